mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_def|k_convB|k_convA" -s 24 -c 10 -o gpurun_out/r02e_conv -f python bench.py --steps 2 --warmup 3 --cpu-sample 0 --config4 0 > gpurun_out/r02e_ncu.log 2>&1
tail -1 gpurun_out/r02e_ncu.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k regex:"k_def|k_conv" -s 24 -c 24 --csv --log-file gpurun_out/r02e_launches.csv python bench.py --steps 2 --warmup 3 --cpu-sample 0 --config4 0 > /dev/null 2>&1
tail -2 gpurun_out/r02e_launches.csv | cut -c1-300
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02e_sanitizer_memcheck.txt 2>&1; tail -4 gpurun_out/r02e_sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02e_sanitizer_racecheck.txt 2>&1; tail -4 gpurun_out/r02e_sanitizer_racecheck.txt
