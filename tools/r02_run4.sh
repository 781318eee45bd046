mkdir -p gpurun_out
( ./tools/ubench_ntt4_a2 ) > gpurun_out/r02_ubench_arith3.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest3.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest3.txt
rm -f gpurun_out/r02_bench_ab3.txt
for v in "" variants/libhec_nodefer.so; do
  echo "== HEC_LIB=$v" >> gpurun_out/r02_bench_ab3.txt
  if [ -n "$v" ]; then export HEC_LIB=$PWD/$v; else unset HEC_LIB; fi
  python bench.py --steps 20 --warmup 5 --cpu-sample 0 2>&1 | tail -1 >> gpurun_out/r02_bench_ab3.txt
  for w in keyswitch eval_relu bootstrap_ctos mul_relin; do
    python bench.py --workload $w --steps 20 --warmup 3 --cpu-sample 0 2>&1 | tail -1 >> gpurun_out/r02_bench_ab3.txt
  done
done
tail -5 gpurun_out/r02_pytest3.txt; cat gpurun_out/r02_ubench_arith3.txt
