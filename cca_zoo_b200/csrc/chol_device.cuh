// Warp-synchronous building blocks of the small Cholesky kernels (cholinv.cu, ccaloss.cu): a 32 x 32 block factored /
// inverted by ONE warp, rows in registers, columns broadcast by shuffles -- no block barrier inside.
#pragma once
#include "common.cuh"

namespace ccab {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float fast_rsqrt(float x) {
  float r = rsqrtf(x);
  return r * fmaf(-0.5f * x * r, r, 1.5f);   // one Newton step: ~1 ulp
}
__device__ __forceinline__ double fast_rsqrt(double x) { return 1.0 / sqrt(x); }

// Warp-synchronous Cholesky of one 32 x 32 block held in shared memory at S[rb.., rb..] (row stride LD): lane i owns
// row i in registers, column k is broadcast by shuffles.  dinv[rb + k] <- 1 / L_kk.  Pivots of rows >= nb (padding:
// identity) are not tested.
template <typename T, int LD>
__device__ __forceinline__ void warp_chol_32(T* S, T* dinv_s, int rb, int nb, int j0, T piv_tol, int* bad) {
  const int lane = threadIdx.x & 31;
  T a[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) a[j] = S[(rb + lane) * LD + rb + j];
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    T piv = __shfl_sync(kFull, a[k], k);
    if (rb + k < nb && !(piv > piv_tol)) {   // warp-uniform
      if (lane == 0 && *bad == 0) *bad = j0 + rb + k + 1;
      piv = T(1);
    }
    const T dinv = fast_rsqrt(piv);
    const T lik = lane > k ? a[k] * dinv : (lane == k ? piv * dinv : T(0));
    a[k] = lik;
    if (lane == k) dinv_s[rb + k] = dinv;
#pragma unroll
    for (int j = k + 1; j < 32; ++j) {
      const T ljk = __shfl_sync(kFull, lik, j);
      a[j] = fma(-lik, ljk, a[j]);    // entries with j > lane are never read
    }
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) S[(rb + lane) * LD + rb + j] = j <= lane ? a[j] : T(0);
}

// Inverse of the 32 x 32 lower-triangular block S[rb.., rb..]: lane c computes column c by forward substitution
// (x_k = 0 for k < c), L read from shared memory (broadcast), result into X[rb.., rb..].
template <typename T, int LD>
__device__ __forceinline__ void warp_trinv_32(const T* S, const T* dinv_s, T* X, int rb) {
  const int lane = threadIdx.x & 31;
  T x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = (lane == i) ? T(1) : T(0);
  // column-oriented substitution: x_i final -> the later right-hand sides update independently (short chain)
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    x[i] *= dinv_s[rb + i];
#pragma unroll
    for (int j = i + 1; j < 32; ++j) x[j] = fma(-S[(rb + j) * LD + rb + i], x[i], x[j]);
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) X[(rb + i) * LD + rb + lane] = x[i];
}


}  // namespace ccab
