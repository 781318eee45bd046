mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r02f_relu_launches.csv python bench.py --workload eval_relu --steps 1 --warmup 1 --cpu-sample 0 > gpurun_out/r02f_relu.log 2>&1
tail -2 gpurun_out/r02f_relu.log | cut -c1-300
wc -l gpurun_out/r02f_relu_launches.csv
