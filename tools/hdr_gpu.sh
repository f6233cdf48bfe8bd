# after an ALIGN change: GPU tests, the single-amplicon bench line, the other configs
tag=${1:-cur}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gputest_$tag.log 2>&1; tail -3 gpurun_out/gputest_$tag.log
C2B_VERBOSE=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; cut -c1-260 gpurun_out/bench_$tag.json; grep counters gpurun_out/bench_$tag.err | tail -1
bash tools/configs_gpu.sh $tag
