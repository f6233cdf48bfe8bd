# GPU FASTQ front end on one B200: parity tests against the reference loop and the host front end, then phase timings
tag=${1:-cur}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fastq_ingest.py -m gpu -x -q > gpurun_out/ingest_test_$tag.log 2>&1; tail -5 gpurun_out/ingest_test_$tag.log
C2B_FASTQ_VERBOSE=1 timeout 600 python tools/ingest_profile.py > gpurun_out/ingest_prof_$tag.log 2>&1; grep -v "^\[" gpurun_out/ingest_prof_$tag.log | tail -14; grep "gpu" gpurun_out/ingest_prof_$tag.log | grep "^\[" | tail -12
