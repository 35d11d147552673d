#!/bin/bash
# round-2 evidence pass (1 GPU): launch lists of every bench workload + ncu --set full captures of every kernel family
mkdir -p gpurun_out
NCU="ncu --clock-control none"
echo "== parts =="; timeout 200 python tools/probe_parts.py 2>&1 | grep -E "syev|chol\+inv torch.float32 n=1024|gemm 1024" 
for w in rcca mcca4 ccaloss64 ccaloss512; do
  echo "== launch list $w =="
  timeout 400 $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file gpurun_out/r2_launches_$w.csv python bench.py --workload $w --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_$w.log 2>&1; tail -1 gpurun_out/ncu_$w.log | cut -c1-120
done
cap() {  # name regex skip count workload
  echo "== full $1 =="
  timeout 500 $NCU --set full --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/r2_$1 python bench.py --workload $5 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full_$1.log 2>&1; tail -1 gpurun_out/ncu_full_$1.log | cut -c1-100
}
cap k1_x3 moments_tf32_2cta 1 1 rcca
cap tgemm tgemm_kernel 40 3 rcca
cap chol_diag chol_diag_inv 8 1 rcca
cap syevj_small syevj_small 1 1 rcca
cap dgemm_mma dgemm_mma 30 2 mcca4
cap chol_diag_f64 chol_diag_inv 4 1 mcca4
cap syevj_small_f64 syevj_small 1 1 mcca4
cap loss512 "tgemm_kernel|center_scale|loss_dot|cov_ridge" 30 6 ccaloss512
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
