set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_sharded.py tests/test_multi_slab.py -m gpu -x -q > gpurun_out/r02h_pytest_2gpus.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest_2gpus.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02h_bench_n2_1024.json 2> gpurun_out/r02h_bench_n2_1024.err
tail -3 gpurun_out/r02h_pytest_2gpus.log; cut -c1-400 gpurun_out/r02h_bench_n2_1024.json; tail -3 gpurun_out/r02h_bench_n2_1024.err
