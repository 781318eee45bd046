mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_conv" -s 23 -c 8 -o gpurun_out/r02a_conv python bench.py --steps 2 --warmup 3 --cpu-sample 0 > gpurun_out/r02a_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 3000 --csv --log-file gpurun_out/r02a_relu_launches.csv python bench.py --workload eval_relu --steps 1 --warmup 1 --cpu-sample 0 > gpurun_out/r02a_relu.log 2>&1
tail -3 gpurun_out/r02a_ncu.log
