mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest19.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest19.txt
tail -3 gpurun_out/r02_pytest19.txt
python bench.py 2>&1 | tail -1 > gpurun_out/r02_bench_final2.txt
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_final2.txt").read())
print(d["value"], d["e2e"]["value"], d["steps"], d["parity"], d["kernels_ms_per_run"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["traffic"], d["roofline_int"]["frac"], d["roofline_conv"]["frac"], d["roofline_group"]["frac"], d["latency_ms_single_conv"], d["latency_ms_single_call"], d["config4"]["value"], d["gpu_launches"], d["clocks"])
PY
