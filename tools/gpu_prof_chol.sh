#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_linear_gpu.py -q -x 2>&1 | tail -3
timeout 300 python tools/probe_fitplan.py 16,6,1 16,6,1 2>&1 | tail -1
NCU="ncu --clock-control none"
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e"
timeout 400 $NCU --set full --import-source on --sampling-interval 0 -k regex:chol_diag_inv -s 20 -c 1 -f -o gpurun_out/r2_chol_diag_v2 $B --workload rcca > gpurun_out/ncu_full_chol_v2.log 2>&1; tail -1 gpurun_out/ncu_full_chol_v2.log | cut -c1-100
timeout 400 $NCU --set full --import-source on --sampling-interval 0 -k regex:syevj_small -s 1 -c 1 -f -o gpurun_out/r2_syevj_small_v2 $B --workload rcca > gpurun_out/ncu_full_syevj_v2.log 2>&1; tail -1 gpurun_out/ncu_full_syevj_v2.log | cut -c1-100
ls -la gpurun_out/r2_chol_diag_v2.ncu-rep gpurun_out/r2_syevj_small_v2.ncu-rep | awk '{print $5, $9}'
