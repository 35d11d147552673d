#!/bin/bash
mkdir -p gpurun_out
echo "== bench default =="; timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r1_final.json; cut -c1-400 gpurun_out/bench_r1_final.json
echo "== bench tf32 =="; timeout 600 python bench.py --steps 10 --warmup 3 --precision tf32 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_r1_final_tf32.json
echo "== reference arm =="; timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tail -1 > gpurun_out/bench_r1_reference.json; cut -c1-300 gpurun_out/bench_r1_reference.json
echo "== ncu launch list =="; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "== ncu full K1 x3 =="; timeout 900 ncu --set full --clock-control none --import-source on -k regex:moments_tf32 -s 1 -c 1 -o gpurun_out/k1_r1_x3 python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_k1.log 2>&1
echo "== ncu full K1 tf32 =="; timeout 900 ncu --set full --clock-control none --import-source on -k regex:moments_tf32 -s 1 -c 1 -o gpurun_out/k1_r1_tf32 python bench.py --steps 1 --warmup 1 --no-cpu --precision tf32 > gpurun_out/ncu_k1b.log 2>&1
echo "== ncu full jacobi fused =="; timeout 600 ncu --set full --clock-control none -k regex:jacobi_round_fused -s 40 -c 1 -o gpurun_out/jac_r1 python bench.py --steps 1 --warmup 0 --no-cpu > /dev/null 2>&1
ls -la gpurun_out | tail -12
