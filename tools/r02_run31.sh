mkdir -p gpurun_out
for rep in 1 2; do
for v in prefuse fuse1 fuse0; do
  unset HEC_LIB HEC_RELIN_RESCALE
  if [ $v = prefuse ]; then export HEC_LIB=$PWD/tools/variants/libhec_prefuse.so; fi
  if [ $v = fuse0 ]; then export HEC_RELIN_RESCALE=0; fi
  for w in eval_relu bootstrap_ctos; do
    python bench.py --workload $w --steps 5 --warmup 2 --cpu-sample 0 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$v','$w',round(d['ms_per_step'],4))"
  done
done
done
