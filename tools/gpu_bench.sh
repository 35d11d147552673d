#!/bin/bash
mkdir -p gpurun_out
echo "== smoke =="; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench tf32x3 =="; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_tf32x3.json
echo "== bench tf32 =="; timeout 600 python bench.py --steps 5 --warmup 3 --precision tf32 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_tf32.json
echo "== ncu launch list =="; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
echo "== ncu full K1 =="; timeout 900 ncu --set full --clock-control none --import-source on -k regex:moments_tf32 -s 1 -c 2 -o gpurun_out/k1_r1 python bench.py --steps 1 --warmup 1 --no-cpu --precision tf32 > gpurun_out/ncu_k1.log 2>&1; tail -2 gpurun_out/ncu_k1.log | cut -c1-200
ls -la gpurun_out
