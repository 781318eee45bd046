mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest4.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest4.txt
rm -f gpurun_out/r02_bench_ab4.txt
for v in "" variants/libhec_b5minb2.so; do
  echo "== HEC_LIB=$v" >> gpurun_out/r02_bench_ab4.txt
  if [ -n "$v" ]; then export HEC_LIB=$PWD/$v; else unset HEC_LIB; fi
  python bench.py --steps 20 --warmup 5 --cpu-sample 0 2>&1 | tail -1 >> gpurun_out/r02_bench_ab4.txt
done
unset HEC_LIB
for e in "" HEC_NO_SMALL=1; do
  echo "== env $e" >> gpurun_out/r02_bench_ab4.txt
  for w in keyswitch eval_relu bootstrap_ctos mul_relin; do
    env $e python bench.py --workload $w --steps 20 --warmup 3 --cpu-sample 0 2>&1 | tail -1 >> gpurun_out/r02_bench_ab4.txt
  done
done
tail -5 gpurun_out/r02_pytest4.txt
