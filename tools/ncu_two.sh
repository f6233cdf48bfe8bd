# ncu --set full captures of one step's kernels (after warm-up): ALIGN tier 1, ALIGN tier 2, CLASSIFY, general kernel
tag=${1:-cur}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:c2b_ -s 12 -c 4 -o gpurun_out/prof_${tag}_step -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gate --no-api --e2e-steps 2 > gpurun_out/b_ncu_${tag}.log 2>&1
ls -la gpurun_out | grep prof_${tag}
