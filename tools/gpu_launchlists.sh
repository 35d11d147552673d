#!/bin/bash
# launch lists (ncu gpu__time_duration) of every bench workload
mkdir -p gpurun_out
NCU="ncu --clock-control none"
for w in ${1:-rcca mcca4 ccaloss64 ccaloss512}; do
  timeout 400 $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file gpurun_out/r2_launches_$w.csv python bench.py --workload $w --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_$w.log 2>&1; tail -1 gpurun_out/ncu_$w.log | cut -c1-100
done
