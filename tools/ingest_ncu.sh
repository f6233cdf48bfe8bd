# kernel-level evidence for the GPU FASTQ front end: per-launch duration and DRAM bytes of its kernels (1 Mi-read FASTQ, 541 MB)
tag=${1:-cur}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/ingest_launches_$tag.csv python tools/ingest_profile.py > gpurun_out/ingest_ncu_$tag.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open('gpurun_out/ingest_launches_$tag.csv')))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[h]; ix = {n: i for i, n in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) != len(hdr): continue
    k = r[ix['Kernel Name']].split('(')[0][:60]; m = r[ix['Metric Name']]; u = r[ix['Metric Unit']]; v = float(r[ix['Metric Value']].replace(',', ''))
    a = agg.setdefault(k, {'n': 0, 't': 0.0, 'rd': 0.0, 'wr': 0.0})
    scale = {'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0, 'second': 1e3, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1.0)
    if m == 'gpu__time_duration.sum': a['t'] += v * scale; a['n'] += 1
    elif m == 'dram__bytes_read.sum': a['rd'] += v * scale
    elif m == 'dram__bytes_write.sum': a['wr'] += v * scale
print('%-62s %5s %10s %12s %12s' % ('kernel', 'n', 'ms/launch', 'MB read', 'MB written'))
for k, a in agg.items():
    n = max(1, a['n']); print('%-62s %5d %10.3f %12.1f %12.1f' % (k, a['n'], a['t'] / n, a['rd'] / n / 1e6, a['wr'] / n / 1e6))
PY
