mkdir -p gpurun_out
python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/r02_bench_f.txt
for b in 4 64 256; do python bench.py --batch $b --cts 8 --steps 20 --warmup 5 --cpu-sample 0 --config4 0 2>&1 | tail -1 > gpurun_out/r02_bench_f_B$b.txt; done
for m in 16 32 128; do python bench.py --cts $m --steps 20 --warmup 5 --cpu-sample 0 --config4 0 2>&1 | tail -1 > gpurun_out/r02_bench_f_M$m.txt; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_bench_f*.txt")):
    try:
        d=json.loads(open(f).read())
        print(f, round(d["value"],1), round(d["e2e"]["value"],1), d["parity"]["bit_exact_vs_oracle"], d["kernels_ms_per_run"], d["roofline"]["kernel"], round(d["roofline"]["frac"],3), round(d["roofline_int"]["frac"],3), round(d["roofline_conv"]["frac"],3), round(d["roofline_group"]["frac"],3), round(d["latency_ms_single_conv"],3), round(d["latency_ms_single_call"],3), d.get("config4",{}) and d["config4"].get("value"))
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-1500:])
PY
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest16.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest16.txt
tail -4 gpurun_out/r02_pytest16.txt
