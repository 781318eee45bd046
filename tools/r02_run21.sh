mkdir -p gpurun_out
timeout 300 python tests/defer_debug.py 4 1 2>&1 | grep -v " ok$" | tail -12
timeout 300 python tests/defer_debug.py 16 2 2>&1 | grep -v " ok$" | tail -12
timeout 300 python tests/defer_debug.py 8 8 2>&1 | grep -v " ok$" | tail -12
python bench.py --steps 20 --warmup 5 --cpu-sample 0 --config4 0 2>&1 | tail -1 > gpurun_out/r02_bench_defer3.txt
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_defer3.txt",):
    try:
        d=json.loads(open(f).read())
        print(f, d["value"], d["e2e"]["value"], d["parity"], d["kernels_ms_per_run"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline_int"]["frac"], d["latency_ms_single_conv"], d["latency_ms_single_call"])
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-2000:])
PY
python tools/defer_crossover.py 16 2>&1 | tail -7
HEC_DEFER=0 python tools/defer_crossover.py 16 2>&1 | tail -7
