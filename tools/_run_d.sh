set -u
mkdir -p gpurun_out
for v in cur pf pflpt zu8; do
  cp variants/$v.so bonej2_b200/libbonejthickness.so
  echo "== $v" >> gpurun_out/r03d_stage.log
  timeout 300 python tools/stage_bench.py 1024 5 paint >> gpurun_out/r03d_stage.log 2>&1
  timeout 300 python tools/one_run.py 1024 1 2 >> gpurun_out/r03d_stage.log 2>&1
done
cat gpurun_out/r03d_stage.log
