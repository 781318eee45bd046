#!/bin/bash
# A/B of a run-time experiment knob (environment variable $1 = 0 / 1) on the generic-path workloads
K=${1:-HEC_NTT_SORT}
for v in 0 1 0 1; do
  echo "== $K=$v"
  for w in keyswitch eval_relu bootstrap_ctos; do
    env $K=$v python bench.py --workload $w --steps 20 --warmup 3 --cpu-sample 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  %-16s %.4f ms/step' % (d['config']['workload'], d['ms_per_step']))"
  done
done
