set -u
mkdir -p gpurun_out
python tools/one_run.py 512 1 1 > gpurun_out/r03f_one_plain.log 2>&1 &&
timeout 500 ncu --set full --clock-control none --import-source on \
    -k regex:"disc_kernel|zfront_kernel|finish_tma_kernel|ridge_tma_kernel" -c 8 \
    -o gpurun_out/r03f_full_512sp -f python tools/one_run.py 512 1 1 > gpurun_out/r03f_full.log 2>&1
tail -3 gpurun_out/r03f_full.log; cat gpurun_out/r03f_one_plain.log
