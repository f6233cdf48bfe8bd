mkdir -p gpurun_out
C2B200_LIB=$PWD/crispresso2_b200/libc2b200_tma.so python -m pytest tests -m gpu -x -q > gpurun_out/gputest_tma.log 2>&1; tail -3 gpurun_out/gputest_tma.log
for lib in libc2b200_tma.so; do for env in "X=1" "C2B_NO_TMA_STAGE=1"; do echo "== $lib $env"; env $env C2B200_LIB=$PWD/crispresso2_b200/$lib python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.2fM  kernel %.2f ms  e2e %.2fM (%.1f ms) gate %s'%(d['value']/1e6, d['roofline']['kernel_ms'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['config']['parity_gate']))
    elif 'rror' in l: print(l.strip()[:300])
"; done; done
