set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r03g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03g_pytest.log
timeout 300 python tools/stage_bench.py 1024 5 > gpurun_out/r03g_stage.log 2>&1
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r03g_bench.json 2> gpurun_out/r03g_bench.err
tail -4 gpurun_out/r03g_pytest.log; cat gpurun_out/r03g_stage.log
