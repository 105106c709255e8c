#!/bin/bash
# Painter occupancy experiments, run on the GPU box:  bash tools/tune_painter.sh [N]
# Rebuilds the library with other shared-memory / register budgets for the line kernels (a
# block takes as long with 7 co-resident blocks as with 9, so more resident warps should buy throughput if the
# spills they cost stay out of the merge loop) and times one Tb.Sp run per variant.  The last build restores the defaults.
set -u
N=${1:-1024}
for v in "12 9 12" "10 10 12" "8 11 12" "8 12 12" "12 9 16"; do
    set -- $v
    echo "== S_SM=$1 DYN_BLOCKS=$2 STAIR_BLOCKS=$3"
    BJT_NVCC_EXTRA="-DBJT_S_SM=$1 -DBJT_DYN_BLOCKS=$2 -DBJT_STAIR_BLOCKS=$3 -Xptxas -v" \
        python -c "from bonej2_b200 import _native; _native.build(force=True)" 2>&1 | grep -A1 "sep_dyn_kernelILb1ELi2ELi1\|sep_stair_kernelILb1" | grep "spill\|Used" | sed 's/ptxas info    : //'
    python tools/one_run.py $N 1 2 2>&1 | tail -1 | tr ',' '\n' | grep "ms_paint\|ms_total"
done
python -c "from bonej2_b200 import _native; _native.build(force=True)"
