mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "keyswitch or mul_relin or relu or conv_bl or between_layer or rotate" 2>&1 | tail -3
for v in dot_old dot_cap64 base; do
  if [ $v = base ]; then unset HEC_LIB; else export HEC_LIB=$PWD/tools/variants/libhec_$v.so; fi
  for w in keyswitch eval_relu mul_relin; do
    python bench.py --workload $w --steps 10 --warmup 3 --cpu-sample 0 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$v','$w',round(d['ms_per_step'],4))"
  done
done
