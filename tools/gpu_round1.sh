#!/bin/bash
mkdir -p gpurun_out
echo "== probe default ==" ; timeout 120 python tools/probe_tf32.py 2>&1 | tail -20
echo "== pytest gpu kernels ==" ; timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu 2>&1 | tail -40
echo "== perf ==" ; timeout 600 python tools/probe_perf.py 2>&1 | tail -20
