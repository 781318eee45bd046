mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "poly or cheby or relu or bootstrap or mul_relin or ctos or stoc or layer" 2>&1 | tail -4
for f in 1 0; do
  for w in eval_relu bootstrap_ctos; do
    HEC_RELIN_RESCALE=$f python bench.py --workload $w --steps 5 --warmup 2 --cpu-sample 0 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('fuse=$f','$w',round(d['ms_per_step'],4))"
  done
done
