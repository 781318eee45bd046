mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --cpu-sample 0 --config4 0 2>&1 | tail -1 > gpurun_out/r02_bench_defer1.txt
HEC_DEFER=0 python bench.py --steps 20 --warmup 5 --cpu-sample 0 --config4 0 2>&1 | tail -1 > gpurun_out/r02_bench_defer0.txt
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_defer1.txt","gpurun_out/r02_bench_defer0.txt"):
    try:
        d=json.loads(open(f).read())
        print(f, d["value"], d["e2e"]["value"], d["parity"], d["kernels_ms_per_run"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline_int"]["frac"], d["latency_ms_single_conv"], d["latency_ms_single_call"])
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-2000:])
PY
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest13.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest13.txt
tail -6 gpurun_out/r02_pytest13.txt
