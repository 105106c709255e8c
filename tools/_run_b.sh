set -u
mkdir -p gpurun_out
for v in e16k1 e16k5 e16y3 zf256 e16k3; do
  cp variants/$v.so bonej2_b200/libbonejthickness.so
  st=edt; [ $v = zf256 ] && st=paint
  echo "== $v" >> gpurun_out/r03b_stage.log
  timeout 300 python tools/stage_bench.py 1024 5 $st >> gpurun_out/r03b_stage.log 2>&1
done
echo "== e16k3 512" >> gpurun_out/r03b_stage.log
timeout 300 python tools/stage_bench.py 512 5 edt >> gpurun_out/r03b_stage.log 2>&1
# e16k3 is now the in-tree library
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r03b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03b_pytest.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r03b_bench.json 2> gpurun_out/r03b_bench.err
cat gpurun_out/r03b_stage.log; tail -5 gpurun_out/r03b_pytest.log
