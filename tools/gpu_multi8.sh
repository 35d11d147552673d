#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for WL in rcca mcca4; do
  n=8
  echo "== bench $WL N=$n =="
  NCCL_DEBUG=WARN timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 10 --warmup 3 --workload $WL 2>gpurun_out/scale_${WL}_n$n.err | tail -1 > gpurun_out/scale_${WL}_n$n.json
  python -c "import json; d=json.load(open('gpurun_out/scale_${WL}_n$n.json')); print({k:d[k] for k in ['n_gpus','value','ms_per_step']}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d.get('fit_route'))" || tail -5 gpurun_out/scale_${WL}_n$n.err
done
echo "== config 5: GCCA 8 x 2048 float64, 62500 rows per GPU, 8 ranks =="
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 tools/config5_shard.py 2>gpurun_out/config5_n8.err | tail -1 > gpurun_out/config5_n8.json
python -c "import json; d=json.load(open('gpurun_out/config5_n8.json')); print({k:d[k] for k in ['n_gpus','fit_ms','k1_ms','k1_tflops_fp64_per_gpu','after_k1_ms','allreduce_bytes','peak_mem_gb']}); print(d['properties'])" || tail -8 gpurun_out/config5_n8.err
