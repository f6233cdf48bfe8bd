# host-side layer on one B200: GPU tests, then the bench line with the api block (stage timings of process_fastq, ingest legs)
tag=${1:-cur}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gputest_$tag.log 2>&1; tail -3 gpurun_out/gputest_$tag.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -2 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$tag.json'))
print('value %.1f M  e2e %.1f M  gate %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['config']['parity_gate']))
for k,v in d.get('api',{}).items():
    if 'stages_s' in v: print(' ',k, round(v['reads_per_s']), round(v['unique_per_s']), round(v['seconds'],3), v['stages_s'])
    elif v and all(isinstance(x, dict) for x in v.values()):
        for kk,vv in v.items(): print('   ',kk, {a:(round(b,3) if isinstance(b,float) else b) for a,b in vv.items()})
    else: print(' ',k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in (v or {}).items() if a != 'call'})
PY
