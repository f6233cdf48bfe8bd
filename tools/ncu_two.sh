# ncu --set full captures of the ALIGN and CLASSIFY kernels (one 1 Mi-read launch each, after warm-up) + the general kernel
tag=${1:-cur}
mkdir -p gpurun_out
for k in c2b_align_kernel c2b_classify_kernel c2b_align_classify_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/prof_${tag}_$k -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gate --e2e-steps 2 > gpurun_out/b_ncu_${tag}_$k.log 2>&1
done
ls -la gpurun_out | tail -5
