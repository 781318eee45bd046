mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_ref_eval_vectors.py -m gpu -x -q -k "poly or cheby or relu or bootstrap or ctos or stoc or layer or bn_relu" 2>&1 | tail -4
for w in eval_relu bootstrap_ctos; do
    python bench.py --workload $w --steps 5 --warmup 2 --cpu-sample 0 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$w',round(d['ms_per_step'],4))"
done
python bench.py --workload eval_relu --cts 1 --steps 5 --warmup 2 --cpu-sample 0 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('eval_relu single',round(d['ms_per_step'],4))"
