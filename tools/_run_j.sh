set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r03j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03j_pytest.log
for inv in 1 0; do timeout 300 python tools/one_run.py 1024 $inv 2 2>&1 | tail -1 | tr ',' '\n' | grep "ms_paint\|ms_total" | tr '\n' ' ' >> gpurun_out/r03j_stage.log; echo >> gpurun_out/r03j_stage.log; done
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r03j_bench.json 2> gpurun_out/r03j_bench.err
tail -3 gpurun_out/r03j_pytest.log; cat gpurun_out/r03j_stage.log
