mkdir -p gpurun_out
HEC_PLAN_CHUNK=1 HEC_PLAN_CHAINS=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "plan" > gpurun_out/r02_pytest6.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest6.txt
tail -3 gpurun_out/r02_pytest6.txt
rm -f gpurun_out/r02_chunk_ab.txt
for cfg in "0 1" "8 2" "8 4" "16 2" "16 4" "4 4" "32 2" "0 1"; do
  set -- $cfg
  echo "== HEC_PLAN_CHUNK=$1 HEC_PLAN_CHAINS=$2" >> gpurun_out/r02_chunk_ab.txt
  HEC_PLAN_CHUNK=$1 HEC_PLAN_CHAINS=$2 python bench.py --steps 20 --warmup 5 --cpu-sample 0 --check 2 --config4 0 2>&1 | tail -1 >> gpurun_out/r02_chunk_ab.txt
done
HEC_PLAN_CHUNK=8 HEC_PLAN_CHAINS=4 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_conv -s 184 -c 184 --csv --log-file gpurun_out/r02_chunk_dram.csv python bench.py --steps 2 --warmup 3 --cpu-sample 0 --check 0 --config4 0 > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_conv -s 23 -c 23 --csv --log-file gpurun_out/r02_nochunk_dram.csv python bench.py --steps 2 --warmup 3 --cpu-sample 0 --check 0 --config4 0 > /dev/null 2>&1
