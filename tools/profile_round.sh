#!/bin/bash
# Round profile, run on the GPU box:  bash tools/profile_round.sh rNN
#  1. bench.py plain (must exit 0), then the ncu launch list of the same command
#  2. one `ncu --set full` capture of every hot kernel on a 512^3 Tb.Sp run (the 1024^3 replay takes minutes per kernel)
#  3. DRAM traffic of the dominant kernel (disc_kernel) at the bench size: two launches of each run
# Outputs land in gpurun_out/; tools/summarise_profiles.py turns them into profiles/.
set -u
R=${1:-r01}
BJT_BENCH_SHORT_WARMUP=1 python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${R}_bench_plain.json 2> gpurun_out/${R}_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_bench_launches.csv \
    env BJT_BENCH_SHORT_WARMUP=1 python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${R}_bench_ncu.log 2>&1
python tools/one_run.py 512 1 1 > gpurun_out/${R}_one_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:"disc_kernel|zfront_kernel|edt_tile|edt_x_kernel|finish_tma_kernel|finish_kernel|ridge_tma_kernel|ridge_kernel" -c 14 \
    -o gpurun_out/${R}_full_512sp python tools/one_run.py 512 1 1 > gpurun_out/${R}_full.log 2>&1
for INV in 0 1; do
  python tools/one_run.py 1024 $INV 1 > gpurun_out/${R}_traffic_plain_$INV.log 2>&1 &&
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"^disc_kernel$" -c 3 --csv \
      --log-file gpurun_out/${R}_disc_traffic_$INV.csv python tools/one_run.py 1024 $INV 1 > gpurun_out/${R}_traffic_ncu_$INV.log 2>&1
done
ls -la gpurun_out/${R}_*
