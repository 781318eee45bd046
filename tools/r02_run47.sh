mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k regex:"k_def|k_conv" -s 24 -c 24 --csv --log-file gpurun_out/r02g_launches.csv python bench.py --steps 2 --warmup 3 --cpu-sample 0 --config4 0 > /dev/null 2>&1
tail -1 gpurun_out/r02g_launches.csv | cut -c1-200
