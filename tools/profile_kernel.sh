#!/bin/bash
# one `ncu --set full` capture of selected kernels on a 512^3 run:  bash tools/profile_kernel.sh NAME REGEX [COUNT] [INVERSE]
set -u
N=${1:-k}; RX=${2:-sep_dyn_kernel}; C=${3:-4}; INV=${4:-1}
python tools/one_run.py 512 $INV 1 > gpurun_out/${N}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$RX" -c $C -o gpurun_out/${N} \
    python tools/one_run.py 512 $INV 1 > gpurun_out/${N}.log 2>&1
ls -la gpurun_out/${N}*
