mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --cpu-sample 0 > gpurun_out/r02_bench_n2f.txt 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/r02_bench_n2f.txt"):
    if l.startswith("{"):
        d=json.loads(l); print(d["n_gpus"], d["value"], d["e2e"]["value"], d.get("config4"), d["parity"])
PY
tail -3 gpurun_out/r02_bench_n2f.txt | cut -c1-300
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-600
