# compute-sanitizer memcheck over every kernel path incl. the GPU FASTQ front end (small batches)
tag=${1:-cur}
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_check.py > gpurun_out/sanitize_memcheck_$tag.log 2>&1; echo "memcheck rc=$?"; tail -12 gpurun_out/sanitize_memcheck_$tag.log
