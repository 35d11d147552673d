#!/bin/bash
# round-2 evidence pass (1 GPU): launch lists of every bench workload + ncu --set full captures of every kernel family
mkdir -p gpurun_out
NCU="ncu --clock-control none"
for w in rcca mcca4 ccaloss64 ccaloss512; do
  echo "== launch list $w =="
  timeout 400 $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file gpurun_out/r2_launches_$w.csv python bench.py --workload $w --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_$w.log 2>&1; tail -1 gpurun_out/ncu_$w.log | cut -c1-100
done
cap() {  # name regex skip count command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  echo "== full $name =="
  timeout 500 $NCU --set full --import-source on -k regex:$rx -s $skip -c $cnt -f -o gpurun_out/r2_$name "$@" > gpurun_out/ncu_full_$name.log 2>&1; tail -1 gpurun_out/ncu_full_$name.log | cut -c1-100
}
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e"
cap k1_x3b moments_x3b_2cta 1 1 $B --workload rcca
cap k1_x3 moments_tf32_2cta 1 1 $B --workload rcca --precision tf32x3
cap tgemm tgemm_kernel 40 3 $B --workload rcca
cap chol_diag chol_diag_inv 8 1 $B --workload rcca
cap syevj_small syevj_small 1 1 $B --workload rcca
cap dgemm_mma dgemm_mma 60 2 $B --workload mcca4
cap loss_small "ccaloss_small" 2 2 $B --workload ccaloss64
cap jacobi_fused jacobi_round_fused 30 1 python tools/probe_eigen_route.py
cap reduce_cov "reduce_partials|cov_ridge|tf32_bf16_split" 3 3 $B --workload rcca
ls -la gpurun_out/r2_*.ncu-rep | awk '{print $5, $9}'
