#!/bin/bash
# DRAM traffic per launch of the dominant kernel (painter y-stage sweep) at the bench size, from one
# `ncu --set full` capture of a Tb.Th and a Tb.Sp run (all sep_dyn_kernel launches; tools/traffic_json.py keeps the y stage):  bash tools/capture_traffic.sh [SIZE]
set -u
N=${1:-1024}
for INV in 0 1; do
  python tools/one_run.py $N $INV 1 > gpurun_out/traffic_plain_$INV.log 2>&1 &&
  ncu --set full --clock-control none -k regex:"^sep_dyn_kernel$" -c 6 -o gpurun_out/traffic_y_$INV \
      python tools/one_run.py $N $INV 1 > gpurun_out/traffic_ncu_$INV.log 2>&1
done
ls -la gpurun_out/traffic_y_*
