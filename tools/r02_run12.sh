mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_final.txt 2>&1
tail -c 600 gpurun_out/r02_bench_final.txt
ncu --set full --clock-control none --import-source on -k regex:"k_conv" -s 23 -c 8 -o gpurun_out/r02c_conv python bench.py --steps 2 --warmup 3 --cpu-sample 0 --check 0 --config4 0 > gpurun_out/r02c_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k regex:k_conv -s 23 -c 23 --csv --log-file gpurun_out/r02c_launches.csv python bench.py --steps 2 --warmup 3 --cpu-sample 0 --check 0 --config4 0 > /dev/null 2>&1
compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; tail -3 gpurun_out/r02_sanitizer_memcheck.txt
compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; tail -3 gpurun_out/r02_sanitizer_racecheck.txt
HEC_DOT_BULK=1 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "keyswitch_and_rotations or mul_relin or between_layer" > gpurun_out/r02_sanitizer_generic.txt 2>&1; tail -3 gpurun_out/r02_sanitizer_generic.txt
python bench.py --workload bootstrap_ctos --diagonals real --steps 5 --warmup 3 --cpu-sample 0 2>&1 | tail -1 > gpurun_out/r02_ctos_real.txt
python bench.py --workload bootstrap_ctos --diagonals real --log-slots 13 --steps 5 --warmup 3 --cpu-sample 0 2>&1 | tail -1 >> gpurun_out/r02_ctos_real.txt
cut -c1-500 gpurun_out/r02_ctos_real.txt
