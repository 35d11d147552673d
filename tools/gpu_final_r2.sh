#!/bin/bash
# round-2 closing pass (1 GPU): full GPU test suite, smoke(), every bench workload (FULL-size CPU legs), launch lists
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
echo "== pytest -m gpu =="
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_gpu_tests_final.txt; tail -3 gpurun_out/r2_gpu_tests_final.txt
echo "== smoke =="
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for w in rcca mcca4 ccaloss64 ccaloss512; do
  echo "== bench $w =="
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 2>gpurun_out/bench_final_$w.err | tail -1 > gpurun_out/bench_final_$w.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_final_$w.json'))
print('$w', round(d['ms_per_step'],3), 'ms | e2e', round(d['e2e']['ms_per_step'],3), 'ms | launches', d['gpu_launches'], '| roofline', round(d['roofline']['frac'],3), d['roofline'].get('kernel_ms'), '| cpu', (d.get('cpu_baseline') or {}).get('seconds_per_step'), '| parity', (d.get('parity') or {}).get('max_weight_rel_err'), '| clocks', d.get('clocks'))" || tail -3 gpurun_out/bench_final_$w.err
done
bash tools/gpu_launchlists.sh "rcca mcca4 ccaloss64 ccaloss512"
