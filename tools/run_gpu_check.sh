mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gputest.log 2>&1; tail -3 gpurun_out/gputest.log
C2B_VERBOSE=1 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cur.json 2> gpurun_out/bench_cur.err; cat gpurun_out/bench_cur.json; tail -2 gpurun_out/bench_cur.err
ncu --set full --clock-control none --import-source on -k regex:c2b_align -s 3 -c 1 -o gpurun_out/prof_cur python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_cur.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cur.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_launches.log 2>&1
