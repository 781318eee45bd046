mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest12.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest12.txt
tail -6 gpurun_out/r02_pytest12.txt
python bench.py --steps 20 --warmup 5 --cpu-sample 0 --config4 0 2>&1 | tail -1 > gpurun_out/r02_bench_c.txt
timeout 600 python bench.py --workload resnet20 --steps 2 --warmup 1 --cpu-sample 0 > gpurun_out/r02_resnet20.txt 2> gpurun_out/r02_resnet20.err
tail -21 gpurun_out/r02_resnet20.err
