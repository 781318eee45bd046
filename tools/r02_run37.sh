mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_ref_eval_vectors.py -m gpu -x -q -k "poly or cheby or relu or bootstrap or ctos or stoc or layer or bn_relu or resnet" 2>&1 | tail -4
for w in eval_relu bootstrap_ctos; do
    python bench.py --workload $w --steps 10 --warmup 3 --cpu-sample 0 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$w',round(d['ms_per_step'],4), d.get('gpu_launches'))"
done
timeout 900 python bench.py --workload resnet20 --steps 4 --warmup 1 --cpu-sample 0 > gpurun_out/r02_resnet20c.txt 2> gpurun_out/r02_resnet20c.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_resnet20c.txt").read().strip().splitlines()[-1])
print("resnet20", d["value"], d["eval_ms_total"], d["host_prep_ms_total"], [round(l["eval_ms"],1) for l in d["layers"]])
PY
grep -c . gpurun_out/r02_resnet20c.err; awk '{print $7}' gpurun_out/r02_resnet20c.err | sort -n | tail -5
