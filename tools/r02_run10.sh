mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest8.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest8.txt
tail -4 gpurun_out/r02_pytest8.txt
rm -f gpurun_out/r02_bench_ab6.txt
for e in HEC_ROT_SCATTER=1 HEC_ROT_SCATTER=0; do
  echo "== env $e" >> gpurun_out/r02_bench_ab6.txt
  for w in keyswitch eval_relu bootstrap_ctos mul_relin "conv_bl --batch 64 --ker 3"; do
    env $e python bench.py --workload $w --steps 20 --warmup 3 --cpu-sample 0 2>&1 | tail -1 >> gpurun_out/r02_bench_ab6.txt
  done
done
timeout 900 python bench.py --workload resnet20 --steps 2 --warmup 1 --cpu-sample 0 > gpurun_out/r02_resnet20.txt 2> gpurun_out/r02_resnet20.err
tail -21 gpurun_out/r02_resnet20.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/r02c_keyswitch_launches.csv python bench.py --workload keyswitch --steps 3 --warmup 3 --cpu-sample 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_row_fwd|k_col_fwd|k_modup2|k_row_inv|k_col_inv" -s 20 -c 14 -o gpurun_out/r02c_generic python bench.py --workload eval_relu --steps 1 --warmup 1 --cpu-sample 0 > gpurun_out/r02c_ncu.log 2>&1
