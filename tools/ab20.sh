mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.2fM  kernel %.2f ms  e2e %.2fM (%.1f ms) gate %s'%(d['value']/1e6, d['roofline']['kernel_ms'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['config']['parity_gate']))
    elif 'rror' in l: print(l.strip()[:300])
"; }
run C2B_X=1
run C2B200_LIB=$PWD/crispresso2_b200/libc2b200_R0.so
run C2B200_LIB=$PWD/crispresso2_b200/libc2b200_R0.so C2B_NO_REF0=1
