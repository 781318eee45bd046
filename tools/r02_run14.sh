mkdir -p gpurun_out
rm -f gpurun_out/r02_tables.txt
for B in 4 64 256; do
  python bench.py --batch $B --cts 8 --steps 10 --warmup 3 --check 2 --config4 0 2>&1 | tail -1 >> gpurun_out/r02_tables.txt
done
for k in 3 5 7; do
  python bench.py --workload conv_bl --batch 64 --ker $k --steps 10 --warmup 3 --cpu-sample 0 2>&1 | tail -1 >> gpurun_out/r02_tables.txt
done
for w in keyswitch mul_relin eval_relu bootstrap_ctos prep_ker "prep_ker --batch 256" "eval_relu --cts 1"; do
  python bench.py --workload $w --steps 10 --warmup 3 --cpu-sample 0 2>&1 | tail -1 >> gpurun_out/r02_tables.txt
done
python bench.py --workload bootstrap_ctos --diagonals real --log-slots 13 --steps 5 --warmup 3 --cpu-sample 0 2>&1 | tail -1 >> gpurun_out/r02_tables.txt
for cts in 16 32 128; do
  python bench.py --cts $cts --steps 10 --warmup 3 --check 2 --config4 0 --cpu-sample 0 2>&1 | tail -1 >> gpurun_out/r02_tables.txt
done
python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 >> gpurun_out/r02_tables.txt
wc -l gpurun_out/r02_tables.txt
