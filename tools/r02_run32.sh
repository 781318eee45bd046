mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r02f_ctos_launches.csv python bench.py --workload bootstrap_ctos --steps 1 --warmup 1 --cpu-sample 0 > gpurun_out/r02f_ctos.log 2>&1
tail -1 gpurun_out/r02f_ctos.log | cut -c1-300
wc -l gpurun_out/r02f_ctos_launches.csv
