mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gputest_r01k.log 2>&1; tail -2 gpurun_out/gputest_r01k.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01k.json 2> gpurun_out/bench_r01k.err; cut -c1-300 gpurun_out/bench_r01k.json; tail -2 gpurun_out/bench_r01k.err
CUDA_LAUNCH_BLOCKING=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | cut -c1-200 | tail -2
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01k.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_launches_r01k.log 2>&1; grep -c c2b_align gpurun_out/launches_r01k.csv; tail -2 gpurun_out/b_ncu_launches_r01k.log | cut -c1-300
