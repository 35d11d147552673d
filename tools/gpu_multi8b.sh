#!/bin/bash
mkdir -p gpurun_out
for WL in rcca mcca4; do
  n=8
  echo "== bench $WL N=$n =="
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 10 --warmup 3 --workload $WL 2>gpurun_out/scale_${WL}_n$n.err | tail -1 > gpurun_out/scale_${WL}_n$n.json
  python -c "import json; d=json.load(open('gpurun_out/scale_${WL}_n$n.json')); print({k:d[k] for k in ['n_gpus','value','ms_per_step']}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d.get('fit_route'))" || tail -5 gpurun_out/scale_${WL}_n$n.err
done
echo "== rcca N=1 on the same box =="
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | tail -1 > gpurun_out/scale_rcca_n1_8box.json
python -c "import json; d=json.load(open('gpurun_out/scale_rcca_n1_8box.json')); print({k:d[k] for k in ['n_gpus','value','ms_per_step']})"
