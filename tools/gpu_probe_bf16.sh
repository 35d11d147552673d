#!/bin/bash
mkdir -p gpurun_out
nvcc -std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a -I cca_zoo_b200/csrc tools/next/umma_bf16_mn_probe.cu -o gpurun_out/umma_bf16 || exit 1
{
for args in "2048 1024 2 1 1 3 16 2048" "4096 1024 2 1 1 3 32 2048" "2048 512 2 1 1 3 16 2048" "2048 1024 2 1 1 3 16 1024" "128 1024 2 1 1 3 16 2048" "2048 1024 1 1 1 3 16 2048"; do
  echo "--- $args"; timeout 20 gpurun_out/umma_bf16 $args 2>&1 | tail -2
done
} > gpurun_out/bf16_probe.txt 2>&1
cat gpurun_out/bf16_probe.txt
