#!/bin/bash
# usage: gpurun --gpus N -- bash tools/gpu_multi.sh "<N list>" [test] "<workloads>"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ "$2" = "test" ]; then echo "== distributed parity test =="; timeout 600 python -m pytest tests/test_distributed_gpu.py -q -m gpu 2>&1 | tail -5; fi
for WL in ${3:-rcca}; do
for n in $1; do
  echo "== bench $WL N=$n =="
  if [ "$n" = "1" ]; then timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --workload $WL 2>gpurun_out/scale_${WL}_n$n.err | tail -1 > gpurun_out/scale_${WL}_n$n.json
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 10 --warmup 3 --workload $WL 2>gpurun_out/scale_${WL}_n$n.err | tail -1 > gpurun_out/scale_${WL}_n$n.json; fi
  python -c "import json; d=json.load(open('gpurun_out/scale_${WL}_n$n.json')); print({k:d[k] for k in ['n_gpus','value','ms_per_step']}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d.get('fit_route'))" || tail -5 gpurun_out/scale_${WL}_n$n.err
done; done
