set -u
mkdir -p gpurun_out
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r02h_bench_n1_1024.json 2> gpurun_out/r02h_bench_n1_1024.err
timeout 900 bash tools/profile_round.sh r02h > gpurun_out/r02h_profile.log 2>&1
tail -2 gpurun_out/r02h_profile.log; cat gpurun_out/r02h_bench_n1_1024.json | cut -c1-600
