// PROTOTYPE (not part of libccab200): the faster Cholesky / explicit-inverse chain planned in DESIGN.md §8.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o /tmp/fast_chol tools/next/fast_chol.cu
//   /tmp/fast_chol            # self-test against a CPU double-precision reference + timings (needs a GPU)
//
// What changes with respect to cca_zoo_b200/csrc/chol.cu:
//   * diag64_kernel   : the 64 x 64 diagonal block is factored as 2 x 2 blocks of 32, each 32 x 32 factorisation and
//                       triangular inverse done by ONE warp in registers with shuffles (no block barriers inside),
//                       and the INVERSE of the diagonal block is kept (dinv).
//   * panel_kernel    : A_panel L_jj^-T becomes a 64-wide GEMM with dinv (no dependent per-row substitution chains).
//   * trtri_doubling  : L^-1 by recursive doubling, X21 = -X_B (C X_A), log2(n/64) levels of batched GEMMs instead of
//                       n/64 sequential block rows of substitution.
// The data flow of every routine here is emulated lane by lane in tools/next/emulate_fast_chol.py.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                                  \
    }                                                                                           \
  } while (0)

constexpr int NB = 64;
constexpr int LDS = NB + 1;   // shared-memory row pitch (bank skew)

template <typename T>
__device__ __forceinline__ T shfl(T v, int src) {
  return __shfl_sync(0xffffffffu, v, src);
}

// ---------------------------------------------------------------------------------------------------------------
// warp-synchronous 32 x 32 routines: lane i owns row i in registers (statically indexed, loops fully unrolled)
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void warp_potrf32(T (&a)[32], int lane, T tol, int& bad_col) {
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    T piv = shfl(a[k], k);
    if (!(piv > tol)) {
      if (bad_col < 0) bad_col = k;
      piv = T(1);
    }
    const T d = sqrt(piv);
    const T dinv = T(1) / d;
    const T lik = lane > k ? a[k] * dinv : (lane == k ? d : T(0));
    if (lane >= k) a[k] = lik;
#pragma unroll
    for (int j = k + 1; j < 32; ++j) {
      const T ljk = shfl(lik, j);
      if (lane >= j) a[j] = fma(-lik, ljk, a[j]);
    }
  }
}

// x <- row `lane` of L^-1, l = row `lane` of L
template <typename T>
__device__ __forceinline__ void warp_trtri32(const T (&l)[32], T (&x)[32], int lane) {
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] = (j == lane) ? T(1) : T(0);
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const T inv = T(1) / shfl(l[k], k);      // L[k][k] is register k of lane k
#pragma unroll
    for (int j = 0; j <= k; ++j) {
      if (lane == k) x[j] *= inv;
      const T xkj = shfl(x[j], k);
      if (lane > k) x[j] = fma(-l[k], xkj, x[j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// diagonal block: factor (<= 64 x 64, in place, lower) + inverse of the factor (always written as 64 x 64, padded
// with the identity).  128 threads, dynamic shared memory 2 * 64 * 65 * sizeof(T).
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) diag64_kernel(T* __restrict__ A, int64_t lda, int nb, int j0, double piv_tol,
                                                     T* __restrict__ dinv, int* __restrict__ info) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T(*S)[LDS] = reinterpret_cast<T(*)[LDS]>(smem_raw);
  T(*X)[LDS] = reinterpret_cast<T(*)[LDS]>(smem_raw + sizeof(T) * NB * LDS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < NB * NB; e += 128) {
    const int i = e / NB, j = e % NB;
    T v = (i == j) ? T(1) : T(0);
    if (i < nb && j < nb) v = (j <= i) ? A[(size_t)i * lda + j] : A[(size_t)j * lda + i];
    S[i][j] = v;
    X[i][j] = T(0);
  }
  __syncthreads();
  int bad = -1;
  T a[32], x[32];
  // ---- (b) first 32 x 32 block
  if (warp == 0) {
#pragma unroll
    for (int j = 0; j < 32; ++j) a[j] = j <= lane ? S[lane][j] : T(0);
    warp_potrf32(a, lane, (T)piv_tol, bad);
    warp_trtri32(a, x, lane);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      S[lane][j] = j <= lane ? a[j] : T(0);
      X[lane][j] = j <= lane ? x[j] : T(0);
    }
    if (bad >= 0 && lane == 0) atomicCAS(info, 0, j0 + bad + 1);
  }
  __syncthreads();
  const int r = tid >> 2, c0 = (tid & 3) * 8;      // 32 rows x 4 column groups of 8
  T acc[8];
  // ---- (c) L21 = A21 X11^T
#pragma unroll
  for (int cc = 0; cc < 8; ++cc) acc[cc] = T(0);
  for (int k = 0; k < 32; ++k) {
    const T av = S[32 + r][k];
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) acc[cc] = fma(av, X[c0 + cc][k], acc[cc]);
  }
  __syncthreads();
#pragma unroll
  for (int cc = 0; cc < 8; ++cc) S[32 + r][c0 + cc] = acc[cc];
  __syncthreads();
  // ---- (d) A22 <- A22 - L21 L21^T   (reads columns 0..31, writes columns 32..63: no hazard)
#pragma unroll
  for (int cc = 0; cc < 8; ++cc) acc[cc] = S[32 + r][32 + c0 + cc];
  for (int k = 0; k < 32; ++k) {
    const T av = S[32 + r][k];
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) acc[cc] = fma(-av, S[32 + c0 + cc][k], acc[cc]);
  }
#pragma unroll
  for (int cc = 0; cc < 8; ++cc) S[32 + r][32 + c0 + cc] = acc[cc];
  __syncthreads();
  // ---- (e) second 32 x 32 block
  if (warp == 0) {
    bad = -1;
#pragma unroll
    for (int j = 0; j < 32; ++j) a[j] = j <= lane ? S[32 + lane][32 + j] : T(0);
    warp_potrf32(a, lane, (T)piv_tol, bad);
    warp_trtri32(a, x, lane);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      S[32 + lane][32 + j] = j <= lane ? a[j] : T(0);
      X[32 + lane][32 + j] = j <= lane ? x[j] : T(0);
    }
    if (bad >= 0 && lane == 0) atomicCAS(info, 0, j0 + 32 + bad + 1);
  }
  __syncthreads();
  // ---- (f) X21 = -X22 (L21 X11): M = L21 X11 parked in the X21 slot first
#pragma unroll
  for (int cc = 0; cc < 8; ++cc) acc[cc] = T(0);
  for (int k = 0; k < 32; ++k) {
    const T av = S[32 + r][k];
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) acc[cc] = fma(av, X[k][c0 + cc], acc[cc]);
  }
#pragma unroll
  for (int cc = 0; cc < 8; ++cc) X[32 + r][c0 + cc] = acc[cc];     // reads were rows 0..31 of X: no hazard
  __syncthreads();
#pragma unroll
  for (int cc = 0; cc < 8; ++cc) acc[cc] = T(0);
  for (int k = 0; k < 32; ++k) {
    const T xv = X[32 + r][32 + k];                                // X22[r][k], zero for k > r
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) acc[cc] = fma(-xv, X[32 + k][c0 + cc], acc[cc]);
  }
  __syncthreads();
#pragma unroll
  for (int cc = 0; cc < 8; ++cc) X[32 + r][c0 + cc] = acc[cc];
  __syncthreads();
  // ---- (g) write back
  for (int e = tid; e < NB * NB; e += 128) {
    const int i = e / NB, j = e % NB;
    if (i < nb && j <= i) A[(size_t)i * lda + j] = S[i][j];
    dinv[e] = X[i][j];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// panel: rows of B (rows x nb, leading dimension ldb) <- B dinv^T, in place.  One CTA per 64 rows, 256 threads.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) panel_kernel(T* __restrict__ B, int64_t ldb, int rows, int nb,
                                                    const T* __restrict__ dinv) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T(*Bt)[LDS] = reinterpret_cast<T(*)[LDS]>(smem_raw);
  T(*Dv)[LDS] = reinterpret_cast<T(*)[LDS]>(smem_raw + sizeof(T) * NB * LDS);
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * NB;
  for (int e = tid; e < NB * NB; e += 256) {
    const int i = e / NB, j = e % NB;
    Bt[i][j] = (r0 + i < rows && j < nb) ? B[(size_t)(r0 + i) * ldb + j] : T(0);
    Dv[i][j] = dinv[e];
  }
  __syncthreads();
  const int r = tid >> 2, c0 = (tid & 3) * 16;
  T acc[16];
#pragma unroll
  for (int cc = 0; cc < 16; ++cc) acc[cc] = T(0);
  for (int k = 0; k < NB; ++k) {
    const T bv = Bt[r][k];
#pragma unroll
    for (int cc = 0; cc < 16; ++cc) acc[cc] = fma(bv, Dv[c0 + cc][k], acc[cc]);
  }
  if (r0 + r < rows) {
#pragma unroll
    for (int cc = 0; cc < 16; ++cc)
      if (c0 + cc < nb) B[(size_t)(r0 + r) * ldb + c0 + cc] = acc[cc];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// batched SIMT GEMM tile (64 x 64 x 16, 256 threads, 4 x 4 per thread): C = alpha op(A) op(B) + beta C
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) gemm_batched_kernel(int ta, int tb, int M, int N, int K, T alpha,
                                                           const T* __restrict__ A, int64_t lda, int64_t sa,
                                                           const T* __restrict__ B, int64_t ldb, int64_t sb, T beta,
                                                           T* __restrict__ C, int64_t ldc, int64_t sc) {
  __shared__ T As[16][64 + 1];
  __shared__ T Bs[16][64 + 1];
  A += (size_t)blockIdx.z * sa;
  B += (size_t)blockIdx.z * sb;
  C += (size_t)blockIdx.z * sc;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tid = threadIdx.x, tr = (tid >> 4) * 4, tc = (tid & 15) * 4;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int e = tid; e < 16 * 64; e += 256) {
      const int kk = e / 64, mm = e % 64;          // As[kk][mm] = op(A)[m0+mm][k0+kk]
      const int gm = m0 + mm, gk = k0 + kk;
      T v = T(0);
      if (gm < M && gk < K) v = ta ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
      As[kk][mm] = v;
      const int gn = n0 + mm;                      // Bs[kk][mm] = op(B)[k0+kk][n0+mm]
      T w = T(0);
      if (gn < N && gk < K) w = tb ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn];
      Bs[kk][mm] = w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      T av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][tr + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tc + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gm = m0 + tr + i, gn = n0 + tc + j;
      if (gm < M && gn < N) {
        T* c = C + (size_t)gm * ldc + gn;
        *c = beta == T(0) ? alpha * acc[i][j] : fma(alpha, acc[i][j], beta * *c);
      }
    }
}

template <typename T>
static void gemm_b(int ta, int tb, int M, int N, int K, T alpha, const T* A, int64_t lda, int64_t sa, const T* B,
                   int64_t ldb, int64_t sb, T beta, T* C, int64_t ldc, int64_t sc, int batch, cudaStream_t st) {
  if (M <= 0 || N <= 0 || batch <= 0) return;
  dim3 grid((N + 63) / 64, (M + 63) / 64, batch);
  gemm_batched_kernel<T><<<grid, 256, 0, st>>>(ta, tb, M, N, K, alpha, A, lda, sa, B, ldb, sb, beta, C, ldc, sc);
}

template <typename T>
__global__ void place_diag_kernel(const T* __restrict__ dinv, int n, T* __restrict__ X, int64_t ldx) {
  const int b = blockIdx.x, j0 = b * NB;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    const int i = e / NB, j = e % NB;
    if (j0 + i < n && j0 + j < n) X[(size_t)(j0 + i) * ldx + j0 + j] = dinv[(size_t)b * NB * NB + e];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host drivers
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
static void potrf_fast(int n, T* A, int64_t lda, T* dinv, int* info, cudaStream_t st) {
  const size_t smem = 2 * sizeof(T) * NB * LDS;
  static bool attr_set = false;
  if (!attr_set) {
    CK(cudaFuncSetAttribute(diag64_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(panel_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  CK(cudaMemsetAsync(info, 0, sizeof(int), st));
  int b = 0;
  for (int j0 = 0; j0 < n; j0 += NB, ++b) {
    const int nb = std::min(NB, n - j0);
    T* Ajj = A + (size_t)j0 * lda + j0;
    T* dv = dinv + (size_t)b * NB * NB;
    diag64_kernel<T><<<1, 128, smem, st>>>(Ajj, lda, nb, j0, 0.0, dv, info);
    const int rows = n - j0 - nb;
    if (rows > 0) {
      T* panel = A + (size_t)(j0 + nb) * lda + j0;
      panel_kernel<T><<<(rows + NB - 1) / NB, 256, smem, st>>>(panel, lda, rows, nb, dv);
      T* A22 = A + (size_t)(j0 + nb) * lda + (j0 + nb);
      gemm_b<T>(0, 1, rows, rows, nb, T(-1), panel, lda, 0, panel, lda, 0, T(1), A22, lda, 0, 1, st);
    }
  }
  CK(cudaGetLastError());
}

// X (n x n, zero-initialised by the caller) <- L^-1 ; tmp: n x n scratch with the same leading dimension as X
template <typename T>
static void trtri_doubling(int n, const T* L, int64_t ldl, const T* dinv, T* X, T* tmp, int64_t ldx, cudaStream_t st) {
  const int nblk = (n + NB - 1) / NB;
  place_diag_kernel<T><<<nblk, 256, 0, st>>>(dinv, n, X, ldx);
  for (int b = NB; b < n; b *= 2) {
    const int full = n / (2 * b);                 // pairs whose B part is complete
    if (full > 0) {
      // tmp[B,A] = L[B,A] X[A,A]  ;  X[B,A] = -X[B,B] tmp[B,A]     (A = first b rows of the pair, B = next b)
      gemm_b<T>(0, 0, b, b, b, T(1), L + (size_t)b * ldl, ldl, (int64_t)2 * b * (ldl + 1), X, ldx,
                (int64_t)2 * b * (ldx + 1), T(0), tmp + (size_t)b * ldx, ldx, (int64_t)2 * b * (ldx + 1), full, st);
      gemm_b<T>(0, 0, b, b, b, T(-1), X + (size_t)b * ldx + b, ldx, (int64_t)2 * b * (ldx + 1), tmp + (size_t)b * ldx,
                ldx, (int64_t)2 * b * (ldx + 1), T(0), X + (size_t)b * ldx, ldx, (int64_t)2 * b * (ldx + 1), full, st);
    }
    const int r0 = full * 2 * b, rem = n - r0;    // ragged last pair: A complete, B shorter
    if (rem > b) {
      const int mb = rem - b;
      const T* Lc = L + (size_t)(r0 + b) * ldl + r0;
      T* Xa = X + (size_t)r0 * ldx + r0;
      T* Xb = X + (size_t)(r0 + b) * ldx + (r0 + b);
      T* Xc = X + (size_t)(r0 + b) * ldx + r0;
      T* Tc = tmp + (size_t)(r0 + b) * ldx + r0;
      gemm_b<T>(0, 0, mb, b, b, T(1), Lc, ldl, 0, Xa, ldx, 0, T(0), Tc, ldx, 0, 1, st);
      gemm_b<T>(0, 0, mb, b, mb, T(-1), Xb, ldx, 0, Tc, ldx, 0, T(0), Xc, ldx, 0, 1, st);
    }
  }
  CK(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------------
// self-test
// ---------------------------------------------------------------------------------------------------------------
static double frand() { return (double)rand() / RAND_MAX - 0.5; }

template <typename T>
static bool run_case(int n, bool timing) {
  const double tol = sizeof(T) == 4 ? 2e-4 : 1e-11;
  // SPD matrix A = G^T G / m + 0.1 I with m = 2 n (double on the host)
  const int m = 2 * n;
  std::vector<double> G((size_t)m * n), Ad((size_t)n * n, 0.0);
  for (auto& g : G) g = frand();
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = 0;
      for (int k = 0; k < m; ++k) s += G[(size_t)k * n + i] * G[(size_t)k * n + j];
      Ad[(size_t)i * n + j] = Ad[(size_t)j * n + i] = s / m * 12.0 + (i == j ? 0.1 : 0.0);
    }
  std::vector<T> Ah((size_t)n * n);
  for (size_t e = 0; e < Ah.size(); ++e) Ah[e] = (T)Ad[e];
  T *dA, *dA0, *dX, *dTmp, *dDinv;
  int* dInfo;
  const int nblk = (n + NB - 1) / NB;
  CK(cudaMalloc(&dA, sizeof(T) * n * n));
  CK(cudaMalloc(&dA0, sizeof(T) * n * n));
  CK(cudaMemcpy(dA0, Ah.data(), sizeof(T) * n * n, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dX, sizeof(T) * n * n));
  CK(cudaMalloc(&dTmp, sizeof(T) * n * n));
  CK(cudaMalloc(&dDinv, sizeof(T) * nblk * NB * NB));
  CK(cudaMalloc(&dInfo, sizeof(int)));
  auto run = [&]() {
    CK(cudaMemcpyAsync(dA, dA0, sizeof(T) * n * n, cudaMemcpyDeviceToDevice, 0));
    CK(cudaMemsetAsync(dX, 0, sizeof(T) * n * n, 0));
    potrf_fast<T>(n, dA, n, dDinv, dInfo, 0);
    trtri_doubling<T>(n, dA, n, dDinv, dX, dTmp, n, 0);
  };
  run();
  CK(cudaDeviceSynchronize());
  std::vector<T> Lh((size_t)n * n), Xh((size_t)n * n);
  int info = -1;
  CK(cudaMemcpy(Lh.data(), dA, sizeof(T) * n * n, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(Xh.data(), dX, sizeof(T) * n * n, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&info, dInfo, sizeof(int), cudaMemcpyDeviceToHost));
  // checks with random vectors (O(n^2) on the host): || A v - L (L^T v) || / || A v ||  and  || X (L v) - v || / || v ||
  double worst_f = 0, worst_i = 0;
  for (int rep = 0; rep < 4; ++rep) {
    std::vector<double> v(n), t(n, 0.0), u(n, 0.0), av(n, 0.0), lv(n, 0.0), xv(n, 0.0);
    for (auto& x : v) x = frand();
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) av[i] += Ad[(size_t)i * n + j] * v[j];
    for (int j = 0; j < n; ++j)                       // t = L^T v
      for (int i = j; i < n; ++i) t[j] += (double)Lh[(size_t)i * n + j] * v[i];
    for (int i = 0; i < n; ++i)                       // u = L t
      for (int j = 0; j <= i; ++j) u[i] += (double)Lh[(size_t)i * n + j] * t[j];
    double num = 0, den = 0;
    for (int i = 0; i < n; ++i) num += (u[i] - av[i]) * (u[i] - av[i]), den += av[i] * av[i];
    worst_f = std::max(worst_f, std::sqrt(num / den));
    for (int i = 0; i < n; ++i)                       // lv = L v
      for (int j = 0; j <= i; ++j) lv[i] += (double)Lh[(size_t)i * n + j] * v[j];
    for (int i = 0; i < n; ++i)                       // xv = X lv
      for (int j = 0; j <= i; ++j) xv[i] += (double)Xh[(size_t)i * n + j] * lv[j];
    num = den = 0;
    for (int i = 0; i < n; ++i) num += (xv[i] - v[i]) * (xv[i] - v[i]), den += v[i] * v[i];
    worst_i = std::max(worst_i, std::sqrt(num / den));
  }
  double upper = 0;                                   // the strict upper triangle of X must stay zero
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) upper = std::max(upper, std::fabs((double)Xh[(size_t)i * n + j]));
  const bool ok = info == 0 && worst_f < tol && worst_i < 50 * tol && upper == 0.0;
  printf("%s n=%5d  info=%d  factor resid %.2e  inverse resid %.2e  upper %.1e  %s\n", sizeof(T) == 4 ? "f32" : "f64",
         n, info, worst_f, worst_i, upper, ok ? "PASS" : "FAIL");
  if (timing) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) run();
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 20; ++i) run();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("    potrf + explicit inverse (incl. a device-to-device refill of A): %.3f ms per matrix\n", ms / 20);
  }
  cudaFree(dA); cudaFree(dA0); cudaFree(dX); cudaFree(dTmp); cudaFree(dDinv); cudaFree(dInfo);
  return ok;
}

int main() {
  bool ok = true;
  for (int n : {7, 32, 64, 96, 200, 1000, 1024}) ok &= run_case<float>(n, n >= 1000);
  for (int n : {40, 64, 96, 200, 1024, 2048}) ok &= run_case<double>(n, n >= 1024);
  printf(ok ? "ALL PASS\n" : "SOME FAILED\n");
  return ok ? 0 : 1;
}
