mkdir -p gpurun_out
timeout 900 python bench.py --workload resnet20 --steps 3 --warmup 1 --cpu-sample 0 > gpurun_out/r02_resnet20e.txt 2> gpurun_out/r02_resnet20e.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_resnet20e.txt").read().strip().splitlines()[-1])
print("resnet20", d["value"], d["eval_ms_total"], d["host_prep_ms_total"], d["pipelined"])
PY
tail -3 gpurun_out/r02_resnet20e.err
