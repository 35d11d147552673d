#!/bin/bash
# first contact with the hardware: descriptor probe, kernel tests, timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
echo "== probe default ==" ; timeout 120 python tools/probe_tf32.py 2>&1 | tail -20
echo "== probe swapped lbo/sbo (1024, 4096) ==" ; timeout 120 python tools/probe_tf32.py 1024 4096 2>&1 | tail -12
echo "== pytest gpu kernels ==" ; timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -m gpu 2>&1 | tail -25
echo "== perf ==" ; timeout 600 python tools/probe_perf.py 2>&1 | tail -20
