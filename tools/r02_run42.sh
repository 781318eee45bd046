for v in A1_MINB2 DA2_MINB2 B1_MINB2 base; do
  if [ $v = base ]; then unset HEC_LIB; else export HEC_LIB=$PWD/tools/variants/libhec_$v.so; fi
  python bench.py --steps 20 --warmup 5 --cpu-sample 0 --config4 0 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$v', round(d['value'],1), d['parity']['bit_exact_vs_oracle'], d['kernels_ms_per_run'])"
done
