#!/bin/bash
# Round profile, run on the GPU box:  bash tools/profile_round.sh rNN
#  1. bench.py plain (must exit 0), then the ncu launch list of the same command
#  2. one `ncu --set full` capture of every hot kernel on a 512^3 Tb.Sp run (the 1024^3 replay takes minutes per kernel)
# Outputs land in gpurun_out/; tools/summarise_profiles.py turns them into profiles/.
set -u
R=${1:-r01}
BJT_BENCH_SHORT_WARMUP=1 python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${R}_bench_plain.json 2> gpurun_out/${R}_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_bench_launches.csv \
    env BJT_BENCH_SHORT_WARMUP=1 python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${R}_bench_ncu.log 2>&1
python tools/one_run.py 512 1 1 > gpurun_out/${R}_one_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:"sep_dyn_kernel|sep_stair_kernel|edt_tile_kernel|finish_kernel|ridge_kernel|edt_x_kernel" -c 12 \
    -o gpurun_out/${R}_full_512sp python tools/one_run.py 512 1 1 > gpurun_out/${R}_full.log 2>&1
ls -la gpurun_out/${R}_*
