# the wait watchdog (c2b_engine.cu): a healthy run is untouched; with an absurdly small limit the 20-step wait ends the process (exit 70)
tag=${1:-cur}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gate > gpurun_out/wd_ok_$tag.json 2> gpurun_out/wd_ok_$tag.err; echo "normal run rc=$?"; cut -c1-200 gpurun_out/wd_ok_$tag.json
C2B_WATCHDOG_S=0.01 python bench.py --reads 4194304 --steps 6 --warmup 3 --no-api --no-cpu-baseline --no-gate > gpurun_out/wd_trip_$tag.json 2> gpurun_out/wd_trip_$tag.err; echo "tripped run rc=$? (expected 70)"; tail -2 gpurun_out/wd_trip_$tag.err
