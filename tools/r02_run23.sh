mkdir -p gpurun_out
for v in "DB4_MINB=2" "DB2_MINB=2" "DB5_MINB=3" base; do
  if [ $v = base ]; then unset HEC_LIB; else export HEC_LIB=$PWD/tools/variants/libhec_$v.so; fi
  python bench.py --steps 20 --warmup 5 --cpu-sample 0 --config4 0 2>&1 | tail -1 > gpurun_out/r02_bench_var.txt
  python - "$v" <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/r02_bench_var.txt").read())
    print(sys.argv[1], round(d["value"],1), d["parity"]["bit_exact_vs_oracle"], d["kernels_ms_per_run"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
