mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gputest.log 2>&1; tail -3 gpurun_out/gputest.log
for cfg in "2 8" "1 8" "2 6" "2 4" "2 5" "2 7"; do set -- $cfg; echo "== ctas $1 wpc $2"; C2B_VERBOSE=1 C2B_CTAS_PER_SM=$1 C2B_WARPS_PER_CTA=$2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('[c2b]'): print(l.strip())
    elif l.startswith('{'):
        d=json.loads(l); print('value %.2fM  kernel %.1f ms  e2e %.2fM (%.1f ms) gate %s'%(d['value']/1e6, d['roofline']['kernel_ms'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['config']['parity_gate']))
"; done
