# per-kernel launch times (ncu launch list) of the hdr / pooled / mixed bench configs
tag=${1:-cur}
for c in hdr pooled mixed; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${tag}_$c.csv python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --no-gate --no-api --e2e-steps 2 > gpurun_out/b_ncu_launches_${tag}_$c.log 2>&1
  echo "== $c"; python tools/launch_table.py gpurun_out/launches_${tag}_$c.csv | head -5
done
