#!/bin/bash
# usage: gpurun --gpus N -- bash tools/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "== distributed parity test =="; timeout 600 python -m pytest tests/test_distributed_gpu.py -q -m gpu 2>&1 | tail -5
for n in 1 $N; do
  echo "== bench N=$n =="
  if [ "$n" = "1" ]; then timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/scale_n$n.json
  else NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 5 --warmup 3 2>&1 | tail -3 | tee gpurun_out/scale_n$n.json; fi
done
