#!/bin/bash
# What the driver runs at round end, in one GPU call: the GPU tests, smoke(), one short bench.  bash tools/final_check.sh TAG
set -u
T=${1:-final}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -3 gpurun_out/${T}_pytest.log; tail -2 gpurun_out/${T}_smoke.log; cut -c1-260 gpurun_out/${T}_bench.json
