// FP64 tensor-pipe peak on this GPU: register-only mma.sync.m8n8k4.f64 chains (no memory traffic).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/peaks/dmma_peak tools/peaks/dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int ACC>
__global__ void __launch_bounds__(256) k(double* out, int iters, double a0, double b0) {
  double c[ACC][2];
#pragma unroll
  for (int i = 0; i < ACC; ++i) c[i][0] = c[i][1] = 0.0;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}
template <int ACC>
void run(int ctas_per_sm, int sms) {
  double* out; cudaMalloc(&out, 8);
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ACC><<<sms * ctas_per_sm, 256>>>(out, 100, 1.0, 1.0);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<ACC><<<sms * ctas_per_sm, 256>>>(out, iters, 1.0, 1.0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double flop = (double)sms * ctas_per_sm * 8 * (double)iters * ACC * 512.0;
  printf("DMMA m8n8k4: %d independent chains/warp, %d CTAs/SM x 8 warps: %.2f TFLOP/s (%.3f ms)\n", ACC, ctas_per_sm, flop / ms / 1e9, ms);
  cudaFree(out);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  run<4>(1, p.multiProcessorCount); run<8>(1, p.multiProcessorCount); run<8>(2, p.multiProcessorCount);
  run<16>(2, p.multiProcessorCount); run<8>(4, p.multiProcessorCount);
  return 0;
}
