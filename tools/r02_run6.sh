mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest5.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest5.txt
tail -4 gpurun_out/r02_pytest5.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default.txt 2>&1; tail -c 3000 gpurun_out/r02_bench_default.txt
rm -f gpurun_out/r02_bench_ab5.txt
for e in HEC_DOT_BULK=1 HEC_DOT_BULK=0 HEC_DOT_BULK=1 HEC_DOT_BULK=0; do
  echo "== env $e" >> gpurun_out/r02_bench_ab5.txt
  for w in keyswitch eval_relu bootstrap_ctos "conv_bl --batch 64 --ker 7"; do
    env $e python bench.py --workload $w --steps 20 --warmup 3 --cpu-sample 0 2>&1 | tail -1 >> gpurun_out/r02_bench_ab5.txt
  done
done
for e in HEC_DOT_BULK=1 HEC_DOT_BULK=0; do
env $e ncu --set full --clock-control none --import-source on -k regex:"k_dot" -s 4 -c 2 -o gpurun_out/r02b_dot_$e python bench.py --workload keyswitch --steps 2 --warmup 3 --cpu-sample 0 > gpurun_out/r02b_ncu_$e.log 2>&1
done
