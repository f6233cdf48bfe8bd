mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for env in "C2B_PHASE_WARPS=4" "C2B_PHASE_WARPS=8"; do echo "== $env"; env $env timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.2fM  kernel %.2f ms  e2e %.2fM (%.1f ms) gate %s'%(d['value']/1e6, d['roofline']['kernel_ms'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['config']['parity_gate'])); print({k: d['config'].get(k) for k in ('packed_pair_items','single_items','band_reruns','ring_pairs','ring_fallbacks')})
    elif 'rror' in l: print(l.strip()[:300])
"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:c2b_align -s 3 -c 1 -o gpurun_out/prof_cur python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_cur.log 2>&1
