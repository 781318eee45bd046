mkdir -p gpurun_out
HEC_FWD16=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "transforms or keyswitch or mul_relin or relu or rotations or between_layer" > gpurun_out/r02_pytest10.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest10.txt
tail -5 gpurun_out/r02_pytest10.txt
rm -f gpurun_out/r02_fwd16_ab.txt
for e in HEC_FWD16=1 HEC_FWD16=0 HEC_FWD16=1 HEC_FWD16=0; do
  echo "== env $e" >> gpurun_out/r02_fwd16_ab.txt
  for w in keyswitch eval_relu bootstrap_ctos mul_relin "conv_bl --batch 64 --ker 3"; do
    env $e python bench.py --workload $w --steps 20 --warmup 3 --cpu-sample 0 2>&1 | tail -1 >> gpurun_out/r02_fwd16_ab.txt
  done
done
HEC_FWD16=1 ncu --set full --clock-control none --import-source on -k regex:"k_fwd16" -s 6 -c 4 -o gpurun_out/r02d_fwd16 python bench.py --workload eval_relu --steps 1 --warmup 1 --cpu-sample 0 > gpurun_out/r02d_ncu.log 2>&1
