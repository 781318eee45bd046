mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest11.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest11.txt
tail -5 gpurun_out/r02_pytest11.txt
rm -f gpurun_out/r02_bench_ab8.txt
for w in keyswitch eval_relu bootstrap_ctos mul_relin; do
  python bench.py --workload $w --steps 20 --warmup 3 --cpu-sample 0 2>&1 | tail -1 >> gpurun_out/r02_bench_ab8.txt
done
