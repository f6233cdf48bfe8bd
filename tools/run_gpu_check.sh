# Round-end evidence on one B200: GPU parity tests, bench line, ncu full capture of the align kernel, ncu launch list.
tag=${1:-cur}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gputest_$tag.log 2>&1; tail -3 gpurun_out/gputest_$tag.log
C2B_VERBOSE=1 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; cat gpurun_out/bench_$tag.json; tail -2 gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_reference_arm.json 2>/dev/null; cat gpurun_out/bench_${tag}_reference_arm.json | cut -c1-400
timeout 500 ncu --set full --clock-control none --import-source on -k regex:c2b_align -s 3 -c 1 -o gpurun_out/prof_$tag python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_$tag.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_launches_$tag.log 2>&1
ls -la gpurun_out | tail -8
