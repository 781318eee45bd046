mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "conv_bn_relu or layer_pipeline or helpers or ext" > gpurun_out/r02_pytest7.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest7.txt
tail -25 gpurun_out/r02_pytest7.txt
