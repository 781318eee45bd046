mkdir -p gpurun_out
timeout 300 python tests/defer_debug.py 16 2 2>&1 | grep -v " ok$" | tail -5
HEC_DEFER=2 timeout 300 python tests/defer_debug.py 4 1 2>&1 | grep -v " ok$" | tail -5
python bench.py --steps 20 --warmup 5 --cpu-sample 0 --config4 0 2>&1 | tail -1 > gpurun_out/r02_bench_var.txt
python - base <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/r02_bench_var.txt").read())
    print(sys.argv[1], round(d["value"],1), d["parity"]["bit_exact_vs_oracle"], d["kernels_ms_per_run"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
