mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "plan or conv_then_pack or deferred or resnet or 1024 or 4096 or bn_relu or golden" 2>&1 | tail -4
timeout 900 python bench.py --workload resnet20 --steps 5 --warmup 1 --cpu-sample 0 > gpurun_out/r02_resnet20d.txt 2> gpurun_out/r02_resnet20d.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_resnet20d.txt").read().strip().splitlines()[-1])
print("resnet20", d["value"], d["eval_ms_total"], d["host_prep_ms_total"], [round(l["eval_ms"],1) for l in d["layers"]], d["hbm_bytes"])
PY
awk '{print $7}' gpurun_out/r02_resnet20d.err | sort -n | tail -8 | tr '\n' ' '
python bench.py --steps 20 --warmup 5 --cpu-sample 0 --config4 0 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('conv',round(d['value'],1), round(d['e2e']['value'],1), d['latency_ms_single_conv'], d['latency_ms_single_call'], d['parity']['bit_exact_vs_oracle'])"
