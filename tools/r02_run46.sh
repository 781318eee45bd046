mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest20.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest20.txt
tail -3 gpurun_out/r02_pytest20.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
