# bench lines of BASELINE.json configs[2..4] on one GPU (1 Mi reads per step), parity gate on, no CPU legs
tag=${1:-cur}
for c in hdr pooled mixed; do
  python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-api --e2e-steps 3 > gpurun_out/bench_${tag}_$c.json 2> gpurun_out/bench_${tag}_$c.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/bench_${tag}_$c.json')); c=d['config']
print('$c', 'value %.2f M/s  %.2f ms/step  e2e %.2f M/s gate %s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, c['parity_gate']), {k:c[k] for k in ['packed_pair_items','single_items','ring_pairs','ring_fallbacks','aligned_fraction']})" || tail -3 gpurun_out/bench_${tag}_$c.err
done
