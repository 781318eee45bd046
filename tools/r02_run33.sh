mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_ref_eval_vectors.py -m gpu -x -q -k "bootstrap or ctos or stoc or layer or bn_relu or cheby" 2>&1 | tail -4
for v in 1 0; do
  for w in bootstrap_ctos; do
    HEC_SINE_PAIR=$v python bench.py --workload $w --steps 5 --warmup 2 --cpu-sample 0 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('pair=$v','$w',round(d['ms_per_step'],4))"
    HEC_SINE_PAIR=$v python bench.py --workload $w --diagonals real --steps 5 --warmup 2 --cpu-sample 0 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('pair=$v','$w real',round(d['ms_per_step'],4))"
  done
done
