mkdir -p gpurun_out
timeout 300 python tests/defer_debug.py 4 1 > gpurun_out/r02_defer_dbg.txt 2>&1; echo "rc $?" >> gpurun_out/r02_defer_dbg.txt
timeout 300 python tests/defer_debug.py 16 2 >> gpurun_out/r02_defer_dbg.txt 2>&1; echo "rc $?" >> gpurun_out/r02_defer_dbg.txt
cat gpurun_out/r02_defer_dbg.txt | tail -80
