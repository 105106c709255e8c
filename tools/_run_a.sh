set -u
mkdir -p gpurun_out
for v in base k3 k4 k5; do
  cp variants/$v.so bonej2_b200/libbonejthickness.so
  st=edt; [ $v = base ] && st=edt,paint; [ $v = k5 ] && st=edt,paint
  echo "== $v" >> gpurun_out/r03a_stage.log
  timeout 300 python tools/stage_bench.py 1024 5 $st >> gpurun_out/r03a_stage.log 2>&1
done
# k5 is now the in-tree library
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r03a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03a_pytest.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r03a_bench_k5.json 2> gpurun_out/r03a_bench_k5.err
cat gpurun_out/r03a_stage.log; tail -3 gpurun_out/r03a_pytest.log
