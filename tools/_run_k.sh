set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest.log
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r02i_bench_n1_1024.json 2> gpurun_out/r02i_bench_n1_1024.err
timeout 900 bash tools/profile_round.sh r02i > gpurun_out/r02i_profile.log 2>&1
tail -3 gpurun_out/r02i_pytest.log; cut -c1-300 gpurun_out/r02i_bench_n1_1024.json
