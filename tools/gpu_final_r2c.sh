#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_gpu_tests_final.txt; tail -2 gpurun_out/r2_gpu_tests_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 200 python tools/probe_f64k1.py 62500 | tail -1 | tee gpurun_out/f64k1.txt
timeout 600 python bench.py 2>gpurun_out/bench_final_rcca.err | tail -1 > gpurun_out/bench_final_rcca.json
python -c "
import json; d=json.load(open('gpurun_out/bench_final_rcca.json'))
print('rcca', round(d['ms_per_step'],3), 'ms | e2e', round(d['e2e']['ms_per_step'],3), 'ms | launches', d['gpu_launches'], '| roofline', round(d['roofline']['frac'],3), d['roofline'].get('kernel_ms'), '| cpu', (d.get('cpu_baseline') or {}).get('seconds_per_step'), '| parity', (d.get('parity') or {}).get('max_weight_rel_err'), '| clocks', d.get('clocks'))" || tail -3 gpurun_out/bench_final_rcca.err
