// Standalone single-tile probe: TMA (128B swizzle) -> smem -> tcgen05.mma kind::tf32 (MN-major) -> TMEM -> dump.
// nvcc -std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a -I cca_zoo_b200/csrc tools/umma_unit.cu -o gpurun_out/umma_unit
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"
namespace ccab { void set_error(const char*, ...) {} int cuda_fail(cudaError_t e, const char*) { return (int)e; } }
using namespace ccab;

struct P { CUtensorMap map; float* smem_dump; float* d_out; uint32_t lbo, sbo, layout, idesc; int kc; int nk; };

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
  const int atom = p.kc * 128;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 4 * atom);
    for (int a = 0; a < 4; ++a) tma_load_2d(smem + a * atom, &p.map, bar, 32 * a, 0);
  }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < 4 * atom / 4; i += 128) p.smem_dump[i] = ((float*)smem)[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    for (int kk = 0; kk < p.nk; ++kk) {
      uint64_t d = umma_smem_desc(smem_u32(smem) + kk * 1024, p.lbo, p.sbo, p.layout);
      umma_tf32(tb, d, d, p.idesc, kk > 0);
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
  int kc = argc > 7 ? atoi(argv[7]) : 8; int nk = kc / 8; int tsw = argc > 6 ? atoi(argv[6]) : 4;
  uint32_t lbo = argc > 1 ? atoi(argv[1]) : kc * 128, sbo = argc > 2 ? atoi(argv[2]) : 512;
  uint32_t layout = argc > 3 ? atoi(argv[3]) : 1;
  int amaj = argc > 4 ? atoi(argv[4]) : 1, bmaj = argc > 5 ? atoi(argv[5]) : 1;
  int rows = kc;
  std::vector<float> X(rows * 128);
  for (int r = 0; r < rows; ++r) for (int c = 0; c < 128; ++c) X[r * 128 + c] = (float)((r * 7 + c * 3) % 11 - 5);
  float *dX, *dS, *dD;
  cudaMalloc(&dX, X.size() * 4); cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&dS, 65536); cudaMalloc(&dD, 128 * 128 * 4); cudaMemset(dD, 0xff, 128 * 128 * 4);
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  P p; memset(&p, 0, sizeof(p));
  cuuint64_t gd[2] = {128, (cuuint64_t)rows}; cuuint64_t gs[1] = {128 * 4}; cuuint32_t box[2] = {32, (cuuint32_t)kc}; cuuint32_t es[2] = {1, 1};
  CUresult r = ((Enc)f)(&p.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dX, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        (CUtensorMapSwizzle)tsw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)r);
  p.smem_dump = dS; p.d_out = dD; p.lbo = lbo; p.sbo = sbo; p.layout = layout; p.kc = kc; p.nk = nk;
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)amaj << 15) | ((uint32_t)bmaj << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  probe<<<1, 128, 70000>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  printf("tsw=%d kc=%d ", tsw, kc); printf("kernel: %s  lbo=%u sbo=%u layout=%u amaj=%d bmaj=%d idesc=0x%08x\n", cudaGetErrorString(e), lbo, sbo, layout, amaj, bmaj, p.idesc);
  std::vector<float> S(4 * kc * 32), D(128 * 128);
  cudaMemcpy(S.data(), dS, S.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  // smem check: expected element (atom a, row r, col c) at a*kc*32 + r*32 + ((c/4) ^ (r%8))*4 + c%4
  int sbad = 0;
  for (int a = 0; a < 4; ++a) for (int r2 = 0; r2 < kc; ++r2) for (int c = 0; c < 32; ++c) {
    float exp = X[r2 * 128 + a * 32 + c];
    float got = tsw == 3 ? S[a * kc * 32 + r2 * 32 + (((c / 4) ^ (r2 % 8)) * 4) + c % 4]
                         : S[a * kc * 32 + r2 * 32 + (((c / 8) ^ (r2 % 4)) * 8) + c % 8];
    if (exp != got) ++sbad;
  }
  printf("smem swizzle-layout mismatches: %d / %d ; smem[0..7] = %g %g %g %g %g %g %g %g\n", sbad, 4 * kc * 32, S[0], S[1], S[2], S[3], S[4], S[5], S[6], S[7]);
  int dbad = 0; double maxerr = 0;
  for (int i = 0; i < 128; ++i) for (int j = 0; j < 128; ++j) {
    double ref = 0; for (int r2 = 0; r2 < kc * nk && r2 < rows; ++r2) ref += (double)X[r2 * 128 + i] * X[r2 * 128 + j];
    double err = fabs(ref - D[i * 128 + j]); if (err > 1e-3) ++dbad; if (err > maxerr) maxerr = err;
  }
  printf("D mismatches: %d / 16384 maxerr=%g ; D[0][0..7] = %g %g %g %g %g %g %g %g\n", dbad, maxerr, D[0], D[1], D[2], D[3], D[4], D[5], D[6], D[7]);
  printf("ref[0][0..3] = "); for (int j = 0; j < 4; ++j) { double ref = 0; for (int r2 = 0; r2 < kc; ++r2) ref += (double)X[r2*128]*X[r2*128+j]; printf("%g ", ref);} printf("\n");
  return 0;
}
