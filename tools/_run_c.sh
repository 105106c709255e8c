set -u
mkdir -p gpurun_out
for v in p64t512 p64t640 p64t1024 p96t768 cur; do
  cp variants/$v.so bonej2_b200/libbonejthickness.so
  echo "== $v" >> gpurun_out/r03c_stage.log
  timeout 300 python tools/stage_bench.py 1024 5 edt >> gpurun_out/r03c_stage.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edt_tile16 -c 8 -o gpurun_out/r03c_edt16_1024 -f python tools/stage_bench.py 1024 1 edt > gpurun_out/r03c_ncu.log 2>&1
cat gpurun_out/r03c_stage.log; tail -3 gpurun_out/r03c_ncu.log
