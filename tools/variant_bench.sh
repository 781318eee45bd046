#!/bin/bash
# time the generic-path workloads with each build variant under variants/ (A/B of compile-time knobs)
for so in variants/*.so; do
  echo "== $so"
  for w in keyswitch eval_relu bootstrap_ctos; do
    HEC_LIB=$PWD/$so python bench.py --workload $w --steps 20 --warmup 3 --cpu-sample 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  %-16s %.4f ms/step' % (d['config']['workload'], d['ms_per_step']))"
  done
done
