# quick GPU iteration: parity tests + bench line (no CPU baselines) + per-kernel launch times
tag=${1:-cur}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gputest_$tag.log 2>&1; tail -3 gpurun_out/gputest_$tag.log
C2B_VERBOSE=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; cut -c1-300 gpurun_out/bench_$tag.json; tail -3 gpurun_out/bench_$tag.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gate --no-api --e2e-steps 2 > gpurun_out/b_ncu_launches_$tag.log 2>&1
python tools/launch_table.py gpurun_out/launches_$tag.csv | head -12
