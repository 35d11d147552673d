// PROTOTYPE probe for the next round: single 128 x 128 tile, TMA -> smem -> tcgen05.mma kind::f16 with BF16 operands
// read MN-major (the layout K1 needs to run the two cross terms of 3xTF32 as bf16 MMAs, tools/next/README.md).
// Same structure as tools/umma_unit.cu (which found the TF32 MN-major layout); every layout parameter is an argument
// so that one gpurun call can sweep the candidates:
//
//   nvcc -std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a -I cca_zoo_b200/csrc tools/next/umma_bf16_mn_probe.cu \
//        -o gpurun_out/umma_bf16
//   for layout in 2 1 4 6; do for sbo in 1024 512 256; do gpurun_out/umma_bf16 $((KC*128)) $sbo $layout 1 1 3 16 2048; done; done
//     args: lbo_bytes sbo_bytes desc_layout a_major b_major tma_swizzle kc(rows per box) kstep_bytes
//
// Expected first candidate (by analogy with the documented K-major case): TMA SWIZZLE_128B (3), 64-column boxes,
// descriptor layout 2 (SWIZZLE_128B), LBO = kc*128 (next 64-column atom), SBO = 1024 (8-row group), 2048 B per k-step.
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
namespace ccab { void set_error(const char*, ...) {} int cuda_fail(cudaError_t e, const char*) { return (int)e; } }
using namespace ccab;

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct P { CUtensorMap map; float* d_out; uint32_t lbo, sbo, layout, idesc, kstep; int kc; int nk; };

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ P p) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(smem + 65536);
  uint64_t* mbar = bar + 1;
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(mbar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(slot, 128);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = *slot;
  const int atom = p.kc * 128;                       // 64 bf16 columns x kc rows
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 2 * atom);
    for (int a = 0; a < 2; ++a) tma_load_2d(smem + a * atom, &p.map, bar, 64 * a, 0);
  }
  mbar_wait(bar, 0);
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    for (int kk = 0; kk < p.nk; ++kk) {
      uint64_t d = umma_smem_desc(smem_u32(smem) + kk * p.kstep, p.lbo, p.sbo, p.layout);
      umma_f16(tb, d, d, p.idesc, kk > 0);
    }
    umma_commit(mbar);
  }
  mbar_wait(mbar, 0);
  tc_fence_after();
  for (int cc = 0; cc < 4; ++cc) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(tb + ((uint32_t)(warp * 32) << 16) + cc * 32, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) p.d_out[(warp * 32 + lane) * 128 + cc * 32 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 128); }
}

int main(int argc, char** argv) {
  const int kc = argc > 7 ? atoi(argv[7]) : 16;
  const int nk = kc / 16;                                         // kind::f16: K = 16 per MMA
  const uint32_t lbo = argc > 1 ? atoi(argv[1]) : kc * 128, sbo = argc > 2 ? atoi(argv[2]) : 1024;
  const uint32_t layout = argc > 3 ? atoi(argv[3]) : 2;
  const int amaj = argc > 4 ? atoi(argv[4]) : 1, bmaj = argc > 5 ? atoi(argv[5]) : 1;
  const int tsw = argc > 6 ? atoi(argv[6]) : 3;
  const uint32_t kstep = argc > 8 ? atoi(argv[8]) : 2048;
  const int rows = kc;
  std::vector<__nv_bfloat16> X(rows * 128);
  std::vector<float> Xf(rows * 128);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < 128; ++c) {
      Xf[r * 128 + c] = (float)((r * 7 + c * 3) % 11 - 5);        // small integers: exact in bf16, sums exact in fp32
      X[r * 128 + c] = __float2bfloat16(Xf[r * 128 + c]);
    }
  __nv_bfloat16* dX;
  float* dD;
  cudaMalloc(&dX, X.size() * 2); cudaMemcpy(dX, X.data(), X.size() * 2, cudaMemcpyHostToDevice);
  cudaMalloc(&dD, 128 * 128 * 4); cudaMemset(dD, 0xff, 128 * 128 * 4);
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  P p; memset(&p, 0, sizeof(p));
  cuuint64_t gd[2] = {128, (cuuint64_t)rows}; cuuint64_t gs[1] = {128 * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)kc}; cuuint32_t es[2] = {1, 1};
  CUresult r = ((Enc)f)(&p.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dX, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        (CUtensorMapSwizzle)tsw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)r);
  p.d_out = dD; p.lbo = lbo; p.sbo = sbo; p.layout = layout; p.kc = kc; p.nk = nk; p.kstep = kstep;
  // c_format F32 (1 << 4), a/b format BF16 (1 << 7, 1 << 10), a/b major bits 15/16, N >> 3 at 17, M >> 4 at 24
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)amaj << 15) | ((uint32_t)bmaj << 16) |
            ((128u >> 3) << 17) | ((128u >> 4) << 24);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  probe<<<1, 128, 70000>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  printf("tsw=%d kc=%d kstep=%u kernel: %s  lbo=%u sbo=%u layout=%u amaj=%d bmaj=%d idesc=0x%08x\n", tsw, kc, kstep,
         cudaGetErrorString(e), lbo, sbo, layout, amaj, bmaj, p.idesc);
  std::vector<float> D(128 * 128);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  int dbad = 0; double maxerr = 0;
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < 128; ++j) {
      double ref = 0;
      for (int r2 = 0; r2 < 16 * nk && r2 < rows; ++r2) ref += (double)Xf[r2 * 128 + i] * Xf[r2 * 128 + j];
      const double err = fabs(ref - D[i * 128 + j]);
      if (err > 1e-3) ++dbad;
      if (err > maxerr) maxerr = err;
    }
  printf("D mismatches: %d / 16384 maxerr=%g ; D[0][0..7] = %g %g %g %g %g %g %g %g\n", dbad, maxerr, D[0], D[1], D[2],
         D[3], D[4], D[5], D[6], D[7]);
  return dbad ? 1 : 0;
}
