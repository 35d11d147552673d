// PROTOTYPE (not part of libccab200): GEMM shapes of the subspace iteration that the 64 x 64-tile kernel of
// csrc/dense.cu runs on a handful of CTAs (profiles/r1_launches_final.txt: 49-73 us for 0.2 GFLOP):
//   Y  = T   Z      1024 x 1024 times 1024 x p      (p = k + oversampling ~ 80..96)   -> 32 CTAs today
//   Z' = T^T Y      same, transposed operand
//   G  = Y^T Y      p x p x 1024   (Gram matrix of CholQR / Rayleigh-Ritz)              -> 4 CTAs today
//   Y  = Y  R       1024 x p times p x p
// Two remedies, both deterministic (no atomics): (a) 32 x 32 output tiles when the 64 x 64 grid would leave most SMs
// idle, (b) split-K into a partial buffer + a fixed-order reduction when even that is too few (Gram matrices).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o /tmp/thin_gemm tools/next/thin_gemm.cu
//   /tmp/thin_gemm        # self-test against a host double-precision reference + CUDA-event timings (needs a GPU)
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                                 \
    }                                                                                          \
  } while (0)

// C[z] (M x N, partial over the k-range of split z) = alpha * op(A) op(B); BM x BN tile, 16-deep chunks, 256 threads.
// TA: op(A) = A^T (A stored K x M); TB: op(B) = B^T (B stored N x K).  gridDim.z = number of K splits; split z covers
// k in [z * kchunk, min(K, (z+1) * kchunk)) and writes to C + z * split_stride (so the caller reduces in fixed order).
template <typename T, int BM, int BN, int TA, int TB>
__global__ void __launch_bounds__(256) gemm_tile_kernel(int M, int N, int K, int kchunk, T alpha,
                                                        const T* __restrict__ A, int64_t lda,
                                                        const T* __restrict__ B, int64_t ldb, T beta,
                                                        T* __restrict__ C, int64_t ldc, int64_t split_stride) {
  constexpr int KC = 16;
  constexpr int TM = BM / 16, TN = BN / 16;       // per-thread register tile (16 x 16 thread grid)
  __shared__ T As[KC][BM + 4];
  __shared__ T Bs[KC][BN + 4];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * kchunk, k_end = min(K, k_begin + kchunk);
  C += (size_t)blockIdx.z * split_stride;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  T acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = T(0);
  for (int k0 = k_begin; k0 < k_end; k0 += KC) {
    for (int e = tid; e < KC * BM; e += 256) {
      int kk, mm;
      if (TA) { mm = e % BM; kk = e / BM; } else { kk = e % KC; mm = e / KC; }   // fastest index = contiguous one
      const int gk = k0 + kk, gm = m0 + mm;
      T v = T(0);
      if (gk < k_end && gm < M) v = TA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
      As[kk][mm] = v;
    }
    for (int e = tid; e < KC * BN; e += 256) {
      int kk, nn;
      if (TB) { kk = e % KC; nn = e / KC; } else { nn = e % BN; kk = e / BN; }
      const int gk = k0 + kk, gn = n0 + nn;
      T v = T(0);
      if (gk < k_end && gn < N) v = TB ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn];
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      T a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int r = m0 + ty * TM + i;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int c = n0 + tx * TN + j;
      if (c >= N) continue;
      T v = alpha * acc[i][j];
      if (beta != T(0)) v += beta * C[(size_t)r * ldc + c];
      C[(size_t)r * ldc + c] = v;
    }
  }
}

// out = beta * out + sum over splits (fixed order)
template <typename T>
__global__ void reduce_splits_kernel(const T* __restrict__ part, int splits, int64_t split_stride, int M, int N,
                                     int64_t ldp, T beta, T* __restrict__ out, int64_t ldo) {
  const int64_t total = (int64_t)M * N;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / N), c = (int)(e % N);
    T s = T(0);
    for (int z = 0; z < splits; ++z) s += part[(size_t)z * split_stride + (size_t)r * ldp + c];
    T* o = out + (size_t)r * ldo + c;
    *o = beta == T(0) ? s : s + beta * *o;
  }
}

// Dispatch: 64 x 64 tiles when they fill the machine, else 32 x 32 tiles, else split-K on top of 32 x 32 tiles.
// `scratch` must hold splits * M * N elements when split-K is chosen (query with plan()).
struct Plan {
  int bm, splits, kchunk;
};
static Plan plan(int M, int N, int K, int sms) {
  auto tiles = [&](int b) { return ((M + b - 1) / b) * ((N + b - 1) / b); };
  if (tiles(64) >= sms) return {64, 1, K};
  if (tiles(32) >= sms / 2 || K <= 256) return {32, 1, K};
  int splits = std::min(std::max(1, sms / std::max(1, tiles(32))), std::max(1, K / 128));
  int kchunk = ((K + splits - 1) / splits + 15) / 16 * 16;
  splits = (K + kchunk - 1) / kchunk;
  return {32, splits, kchunk};
}

template <typename T, int BM, int BN>
static void launch_tile(int ta, int tb, dim3 grid, int M, int N, int K, int kchunk, T alpha, const T* A, int64_t lda,
                        const T* B, int64_t ldb, T beta, T* C, int64_t ldc, int64_t ss, cudaStream_t st) {
  if (!ta && !tb) gemm_tile_kernel<T, BM, BN, 0, 0><<<grid, 256, 0, st>>>(M, N, K, kchunk, alpha, A, lda, B, ldb, beta, C, ldc, ss);
  else if (ta && !tb) gemm_tile_kernel<T, BM, BN, 1, 0><<<grid, 256, 0, st>>>(M, N, K, kchunk, alpha, A, lda, B, ldb, beta, C, ldc, ss);
  else if (!ta && tb) gemm_tile_kernel<T, BM, BN, 0, 1><<<grid, 256, 0, st>>>(M, N, K, kchunk, alpha, A, lda, B, ldb, beta, C, ldc, ss);
  else gemm_tile_kernel<T, BM, BN, 1, 1><<<grid, 256, 0, st>>>(M, N, K, kchunk, alpha, A, lda, B, ldb, beta, C, ldc, ss);
}

template <typename T>
static void gemm_auto(int ta, int tb, int M, int N, int K, T alpha, const T* A, int64_t lda, const T* B, int64_t ldb,
                      T beta, T* C, int64_t ldc, T* scratch, int sms, cudaStream_t st) {
  const Plan p = plan(M, N, K, sms);
  if (p.splits == 1) {
    if (p.bm == 64) launch_tile<T, 64, 64>(ta, tb, dim3((N + 63) / 64, (M + 63) / 64, 1), M, N, K, K, alpha, A, lda, B, ldb, beta, C, ldc, 0, st);
    else launch_tile<T, 32, 32>(ta, tb, dim3((N + 31) / 32, (M + 31) / 32, 1), M, N, K, K, alpha, A, lda, B, ldb, beta, C, ldc, 0, st);
  } else {
    launch_tile<T, 32, 32>(ta, tb, dim3((N + 31) / 32, (M + 31) / 32, p.splits), M, N, K, p.kchunk, alpha, A, lda, B,
                           ldb, T(0), scratch, N, (int64_t)M * N, st);
    const int blocks = (int)std::min<int64_t>(((int64_t)M * N + 255) / 256, 4 * sms);
    reduce_splits_kernel<T><<<blocks, 256, 0, st>>>(scratch, p.splits, (int64_t)M * N, M, N, N, beta, C, ldc);
  }
  CK(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------------
static double frand() { return (double)rand() / RAND_MAX - 0.5; }

template <typename T>
static bool run_case(const char* what, int ta, int tb, int M, int N, int K, int sms) {
  const int64_t ar = ta ? K : M, ac = ta ? M : K, br = tb ? N : K, bc = tb ? K : N;
  std::vector<T> A((size_t)ar * ac), B((size_t)br * bc), C((size_t)M * N), C0((size_t)M * N);
  for (auto& x : A) x = (T)frand();
  for (auto& x : B) x = (T)frand();
  for (auto& x : C0) x = (T)frand();
  T *dA, *dB, *dC, *dS;
  const Plan p = plan(M, N, K, sms);
  CK(cudaMalloc(&dA, sizeof(T) * A.size()));
  CK(cudaMalloc(&dB, sizeof(T) * B.size()));
  CK(cudaMalloc(&dC, sizeof(T) * C.size()));
  CK(cudaMalloc(&dS, sizeof(T) * (size_t)std::max(1, p.splits) * M * N));
  CK(cudaMemcpy(dA, A.data(), sizeof(T) * A.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), sizeof(T) * B.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dC, C0.data(), sizeof(T) * C.size(), cudaMemcpyHostToDevice));
  const T alpha = (T)0.75, beta = (T)-0.5;
  gemm_auto<T>(ta, tb, M, N, K, alpha, dA, ac, dB, bc, beta, dC, N, dS, sms, 0);
  CK(cudaMemcpy(C.data(), dC, sizeof(T) * C.size(), cudaMemcpyDeviceToHost));
  double worst = 0, scale = 0;
  for (int i = 0; i < M; i += std::max(1, M / 61))          // sampled rows: the host reference is O(M N K)
    for (int j = 0; j < N; ++j) {
      double s = 0;
      for (int k = 0; k < K; ++k) {
        const double a = ta ? A[(size_t)k * ac + i] : A[(size_t)i * ac + k];
        const double b = tb ? B[(size_t)j * bc + k] : B[(size_t)k * bc + j];
        s += a * b;
      }
      const double ref = 0.75 * s - 0.5 * (double)C0[(size_t)i * N + j];
      worst = std::max(worst, std::fabs(ref - (double)C[(size_t)i * N + j]));
      scale = std::max(scale, std::fabs(ref));
    }
  const double tol = sizeof(T) == 4 ? 2e-5 : 1e-12;
  const bool ok = worst <= tol * std::max(scale, 1.0) * std::sqrt((double)K);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) gemm_auto<T>(ta, tb, M, N, K, alpha, dA, ac, dB, bc, T(0), dC, N, dS, sms, 0);
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 50; ++i) gemm_auto<T>(ta, tb, M, N, K, alpha, dA, ac, dB, bc, T(0), dC, N, dS, sms, 0);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("%s %-28s M=%5d N=%5d K=%5d  tile %d splits %2d  err %.1e  %7.1f us  %6.2f TFLOP/s  %s\n",
         sizeof(T) == 4 ? "f32" : "f64", what, M, N, K, p.bm, p.splits, worst, ms / 50 * 1e3,
         2.0 * M * N * K / (ms / 50 * 1e-3) / 1e12, ok ? "PASS" : "FAIL");
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dS);
  return ok;
}

int main() {
  int dev = 0, sms = 148;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  bool ok = true;
  ok &= run_case<float>("Y = T Z", 0, 0, 1024, 96, 1024, sms);
  ok &= run_case<float>("Z = T^T Y", 1, 0, 1024, 96, 1024, sms);
  ok &= run_case<float>("G = Y^T Y (Gram)", 1, 0, 96, 96, 1024, sms);
  ok &= run_case<float>("Y = Y R^T", 0, 1, 1024, 96, 96, sms);
  ok &= run_case<float>("T = Linv C12 (square)", 0, 0, 1024, 1024, 1024, sms);
  ok &= run_case<float>("odd shapes", 1, 1, 77, 45, 333, sms);
  ok &= run_case<double>("G u (config 5)", 0, 0, 16384, 256, 2048, sms);
  ok &= run_case<double>("Gram f64", 1, 0, 256, 256, 4096, sms);
  printf(ok ? "ALL PASS\n" : "SOME FAILED\n");
  return ok ? 0 : 1;
}
