timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_ref_eval_vectors.py -m gpu -x -q -k "poly or cheby or relu or bootstrap or ctos or layer or bn_relu" 2>&1 | tail -3
python bench.py --workload eval_relu --steps 10 --warmup 3 --cpu-sample 0 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('eval_relu',round(d['ms_per_step'],4))"
