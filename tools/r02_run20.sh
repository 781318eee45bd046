mkdir -p gpurun_out
timeout 300 python tests/defer_debug.py 4 1 2>&1 | grep -v " ok$" | tail -12
python bench.py --steps 20 --warmup 5 --cpu-sample 0 --config4 0 2>&1 | tail -1 > gpurun_out/r02_bench_defer2.txt
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_defer2.txt",):
    try:
        d=json.loads(open(f).read())
        print(f, d["value"], d["e2e"]["value"], d["parity"], d["kernels_ms_per_run"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline_int"]["frac"], d["latency_ms_single_conv"], d["latency_ms_single_call"])
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-2000:])
PY
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest14.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest14.txt
tail -6 gpurun_out/r02_pytest14.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_def|k_convB" -s 24 -c 9 -o gpurun_out/r02e_conv -f python bench.py --steps 2 --warmup 3 --cpu-sample 0 --config4 0 > gpurun_out/r02e_ncu.log 2>&1
tail -3 gpurun_out/r02e_ncu.log
