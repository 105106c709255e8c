set -u
mkdir -p gpurun_out
for v in lpt lptpf1 lptpr2 lptpr3; do
  cp variants/$v.so bonej2_b200/libbonejthickness.so
  echo "== $v" >> gpurun_out/r03e_stage.log
  timeout 300 python tools/one_run.py 1024 1 2 2>&1 | tail -1 | tr ',' '\n' | grep "ms_paint\|ms_total" | tr '\n' ' ' >> gpurun_out/r03e_stage.log
  timeout 300 python tools/one_run.py 1024 0 2 2>&1 | tail -1 | tr ',' '\n' | grep "ms_paint\|ms_total" | tr '\n' ' ' >> gpurun_out/r03e_stage.log
  echo >> gpurun_out/r03e_stage.log
done
cat gpurun_out/r03e_stage.log
