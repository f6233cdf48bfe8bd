# Round-end evidence on one B200: GPU parity tests, bench line (all legs), reference arm, ncu full capture of one step's
# kernels, ncu launch list of the same command, the other BASELINE configs.
tag=${1:-cur}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gputest_$tag.log 2>&1; tail -3 gpurun_out/gputest_$tag.log
C2B_VERBOSE=1 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; cut -c1-600 gpurun_out/bench_$tag.json; tail -2 gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_reference_arm.json 2>/dev/null; cut -c1-400 gpurun_out/bench_${tag}_reference_arm.json
bash tools/ncu_two.sh $tag
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gate --no-api --e2e-steps 2 > gpurun_out/b_ncu_launches_$tag.log 2>&1
python tools/launch_table.py gpurun_out/launches_$tag.csv | head -6
bash tools/configs_gpu.sh $tag
ls -la gpurun_out | tail -8
