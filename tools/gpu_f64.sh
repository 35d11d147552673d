#!/bin/bash
mkdir -p gpurun_out
./tools/peaks/dmma_peak | tee gpurun_out/dmma_peak.txt
timeout 200 python tools/probe_f64k1.py 62500 | tail -1 | tee gpurun_out/f64k1.txt
timeout 300 ncu --clock-control none --set full --import-source on -k regex:moments_dmma -s 2 -c 1 -f -o gpurun_out/r2_k1_dmma python tools/probe_f64k1.py 20000 > gpurun_out/ncu_full_k1_dmma.log 2>&1; tail -1 gpurun_out/ncu_full_k1_dmma.log | cut -c1-120
