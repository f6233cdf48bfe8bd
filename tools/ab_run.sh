# A/B of build variants under tools/ab/: bench value per variant (no gates, no CPU legs)
mkdir -p gpurun_out
for v in "$@"; do
  lib=$PWD/tools/ab/lib_$v.so; [ "$v" = base ] && lib=$PWD/crispresso2_b200/libc2b200.so
  C2B200_LIB=$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gate --no-api --e2e-steps 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', 'value %.2f M/s  %.3f ms/step  kernel_ms %.3f  e2e %.2f M/s' % (d['value']/1e6, d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value']/1e6))"
done
