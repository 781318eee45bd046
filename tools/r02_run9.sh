mkdir -p gpurun_out
timeout 1500 python bench.py --workload resnet20 --steps 1 --warmup 1 --cpu-sample 0 > gpurun_out/r02_resnet20.txt 2> gpurun_out/r02_resnet20.err
tail -40 gpurun_out/r02_resnet20.err; tail -c 1500 gpurun_out/r02_resnet20.txt
