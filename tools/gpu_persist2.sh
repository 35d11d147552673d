#!/bin/bash
# fit-plan sweep + ncu capture of the persistent tf32x3b kernel and its pre-pass / reductions
mkdir -p gpurun_out
timeout 300 python tools/probe_fitplan.py 2>&1 | tail -8
NCU="ncu --clock-control none"
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e"
timeout 400 $NCU --set full --import-source on -k regex:moments_x3b_persist -s 1 -c 1 -f -o gpurun_out/r2_k1_persist $B --workload rcca > gpurun_out/ncu_full_k1_persist.log 2>&1; tail -1 gpurun_out/ncu_full_k1_persist.log | cut -c1-100
timeout 400 $NCU --set full -k "regex:split_sums|reduce_colsums|reduce_partials" -s 4 -c 4 -f -o gpurun_out/r2_k1_glue $B --workload rcca > gpurun_out/ncu_full_k1_glue.log 2>&1; tail -1 gpurun_out/ncu_full_k1_glue.log | cut -c1-100
ls -la gpurun_out/r2_k1_persist.ncu-rep gpurun_out/r2_k1_glue.ncu-rep | awk '{print $5, $9}'
