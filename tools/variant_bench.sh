#!/bin/bash
# time the generic-path workloads with each build variant under variants/ (A/B of compile-time knobs such as
# -DHEC_GEN_MINB=n or -DHEC_DOT_U=n; build them with the flags of __graft_entry__.py)
for so in variants/*.so variants/*.so; do
  echo "== $so"
  for w in keyswitch eval_relu bootstrap_ctos "conv_bl --batch 64 --ker 7"; do
    HEC_LIB=$PWD/$so python bench.py --workload $w --steps 20 --warmup 3 --cpu-sample 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  %-16s %.4f ms/step' % (d['config']['workload'], d['ms_per_step']))"
  done
done
