mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gputest.log 2>&1; tail -3 gpurun_out/gputest.log
for lib in libc2b200.so libc2b200_nopf.so libc2b200_w10.so; do for ctas in 2; do echo "== $lib"; C2B200_LIB=$PWD/crispresso2_b200/$lib C2B_VERBOSE=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('[c2b]'): print(l.strip())
    elif l.startswith('{'):
        d=json.loads(l); print('value %.2fM  kernel %.1f ms  e2e %.2fM (%.1f ms) gate %s'%(d['value']/1e6, d['roofline']['kernel_ms'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['config']['parity_gate']))
"; done; done
ncu --set full --clock-control none --import-source on -k regex:c2b_align -s 3 -c 1 -o gpurun_out/prof_r01d python bench.py --steps 1 --warmup 3 --no-cpu-baseline --reads 262144 > gpurun_out/b_ncu4.log 2>&1
