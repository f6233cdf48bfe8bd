# weak-scaling check on one multi-GPU box: bench.py at N = 1 and N = 8 (or the Ns given), same build, back to back
tag=${1:-cur}; shift
Ns=${@:-1 8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$tag.txt 2>&1
for n in $Ns; do
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-api > gpurun_out/scale_${tag}_n$n.json 2> gpurun_out/scale_${tag}_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline --no-api > gpurun_out/scale_${tag}_n$n.json 2> gpurun_out/scale_${tag}_n$n.err
  fi
  python -c "
import json
d=json.load(open('gpurun_out/scale_${tag}_n$n.json'))
print('N=$n value %.1f M/s (%.2f ms/step)  e2e %.1f M/s (%.2f ms/step)  gate %s clocks %s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['config']['parity_gate'], d.get('clocks')))" || tail -5 gpurun_out/scale_${tag}_n$n.err
done
python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -3
