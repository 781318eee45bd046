mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest17.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest17.txt
tail -4 gpurun_out/r02_pytest17.txt
for w in keyswitch mul_relin eval_relu bootstrap_ctos; do
    python bench.py --workload $w --steps 10 --warmup 3 --cpu-sample 0 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$w',round(d['ms_per_step'],4), d.get('gpu_launches'))"
done
python bench.py --workload bootstrap_ctos --diagonals real --steps 10 --warmup 3 --cpu-sample 0 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ctos real',round(d['ms_per_step'],4))"
python bench.py --workload bootstrap_ctos --diagonals real --log-slots 13 --steps 10 --warmup 3 --cpu-sample 0 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ctos real logslots13',round(d['ms_per_step'],4))"
timeout 600 python bench.py --workload resnet20 --steps 2 --warmup 1 --cpu-sample 0 > gpurun_out/r02_resnet20b.txt 2> gpurun_out/r02_resnet20b.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_resnet20b.txt").read().strip().splitlines()[-1])
print("resnet20", d["value"], d["eval_ms_total"], d["host_prep_ms_total"], [round(l["eval_ms"],1) for l in d["layers"]])
PY
