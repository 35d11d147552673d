// Cholesky factorisation and triangular solves (exact FMA, fp32 / fp64), blocked at 64.
//
// They implement the textbook reduction of the generalised symmetric problem A v = lambda B v to a
// standard one (B = L L^T, K = L^-1 A L^-T) -- what scipy.linalg.eigh(A, B) does inside LAPACK *sygvx
// (cca_zoo/_utils/_linalg.py:67-71) -- and the ridge whitening of cca_zoo/_utils/_linalg.py:30-38 in its
// Cholesky form: with R_i = (1-c) C_ii + c I = L_i L_i^T the whitened cross-covariance is
// T = L_1^-1 C_12 L_2^-T and the weights are L_i^-T U_k (same singular values, same weights up to sign).
//
//   potrf : right-looking; per 64-column step  (a) diagonal block factored in shared memory,
//           (b) panel  X L_jj^T = A_panel  solved row-wise, (c) trailing update by gemm_kernel.
//   trsm  : forward / backward block substitution; off-diagonal work is gemm_kernel, the 64 x 64
//           diagonal solves run one thread per right-hand-side column with L_jj in shared memory.
#include "chol.cuh"

#include "dense.cuh"

namespace ccab {

constexpr int kNB = 64;

// ---------------------------------------------------------------------------------------------
// (a) factor one diagonal block (<= 64 x 64) in shared memory; info: first non-positive pivot (1-based)
// ---------------------------------------------------------------------------------------------
// Thread i owns row i of the lower triangle in registers (statically indexed: the k and j loops are fully
// unrolled); per column k: the pivot owner publishes 1/sqrt(piv), every thread scales its entry of column k and
// publishes it, then updates its own row with broadcast shared-memory reads.  Two 64-thread barriers per column.
template <typename T>
__global__ void __launch_bounds__(kNB) potrf_diag_kernel(T* __restrict__ A, int64_t lda, int nb, int j0,
                                                         double piv_tol, int* __restrict__ info) {
  __shared__ T col[kNB];
  __shared__ T dinv_s;
  __shared__ int bad;
  const int i = threadIdx.x;
  if (i == 0) bad = 0;
  T a[kNB];
#pragma unroll
  for (int j = 0; j < kNB; ++j) a[j] = (i < nb && j <= i && j < nb) ? A[(size_t)i * lda + j] : T(0);
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kNB; ++k) {
    if (k < nb) {
      if (i == k) {
        const T piv = a[k];
        if (!(piv > (T)piv_tol)) {
          if (!bad) bad = j0 + k + 1;
          dinv_s = T(1);       // keep going with a harmless value; the caller checks info
          a[k] = T(1);
        } else {
          const T d = sqrt(piv);
          dinv_s = T(1) / d;
          a[k] = d;
        }
      }
      __syncthreads();
      if (i > k) a[k] *= dinv_s;
      col[i] = (i > k) ? a[k] : T(0);
      __syncthreads();
      const T lik = (i > k) ? a[k] : T(0);
#pragma unroll
      for (int j = k + 1; j < kNB; ++j) a[j] = fma(-lik, col[j], a[j]);   // entries with j > i stay unused
    }
  }
#pragma unroll
  for (int j = 0; j < kNB; ++j)
    if (i < nb && j <= i && j < nb) A[(size_t)i * lda + j] = a[j];
  __syncthreads();
  if (i == 0 && bad) atomicCAS(info, 0, bad);
}

// ---------------------------------------------------------------------------------------------
// (b) rows of B <- rows of B * L^-T   (L: nb x nb lower, row-major).  One thread per row.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) trsm_right_lt_kernel(const T* __restrict__ L, int64_t ldl, int nb,
                                                            T* __restrict__ B, int64_t ldb, int rows) {
  __shared__ T Ls[kNB][kNB + 1];
  for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) {
    const int i = e / nb, j = e % nb;
    Ls[i][j] = (j <= i) ? L[(size_t)i * ldl + j] : T(0);
  }
  __syncthreads();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  T* b = B + (size_t)r * ldb;
  T x[kNB];
#pragma unroll
  for (int j = 0; j < kNB; ++j) x[j] = j < nb ? b[j] : T(0);
#pragma unroll
  for (int j = 0; j < kNB; ++j) {
    if (j < nb) {
      T acc = x[j];
#pragma unroll
      for (int k = 0; k < j; ++k) acc = fma(-x[k], Ls[j][k], acc);
      x[j] = acc / Ls[j][j];
    }
  }
#pragma unroll
  for (int j = 0; j < kNB; ++j)
    if (j < nb) b[j] = x[j];
}

// ---------------------------------------------------------------------------------------------
// diagonal solves of the left TRSM: B (nb x m) <- L^-1 B  or  L^-T B.  One thread per column.
// ---------------------------------------------------------------------------------------------
template <typename T, int TRANS>
__global__ void __launch_bounds__(128) trsm_left_diag_kernel(const T* __restrict__ L, int64_t ldl, int nb,
                                                             T* __restrict__ B, int64_t ldb, int m) {
  __shared__ T Ls[kNB][kNB + 1];
  for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) {
    const int i = e / nb, j = e % nb;
    Ls[i][j] = (j <= i) ? L[(size_t)i * ldl + j] : T(0);
  }
  __syncthreads();
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= m) return;
  T x[kNB];
#pragma unroll
  for (int i = 0; i < kNB; ++i) x[i] = i < nb ? B[(size_t)i * ldb + col] : T(0);
  if (TRANS == 0) {  // forward: L x = b
#pragma unroll
    for (int i = 0; i < kNB; ++i) {
      if (i < nb) {
        T acc = x[i];
#pragma unroll
        for (int k = 0; k < i; ++k) acc = fma(-Ls[i][k], x[k], acc);
        x[i] = acc / Ls[i][i];
      }
    }
  } else {  // backward: L^T x = b
#pragma unroll
    for (int ii = 0; ii < kNB; ++ii) {
      const int i = kNB - 1 - ii;
      if (i < nb) {
        T acc = x[i];
#pragma unroll
        for (int k = i + 1; k < kNB; ++k)
          if (k < nb) acc = fma(-Ls[k][i], x[k], acc);
        x[i] = acc / Ls[i][i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kNB; ++i)
    if (i < nb) B[(size_t)i * ldb + col] = x[i];
}

// ---------------------------------------------------------------------------------------------
// host drivers
// ---------------------------------------------------------------------------------------------
template <typename T>
int potrf(int n, T* A, int64_t lda, double piv_tol, int* info_dev, cudaStream_t stream) {
  CCAB_CHECK_ARG(n >= 1 && lda >= n, "bad potrf shape");
  CCAB_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), stream));
  for (int j0 = 0; j0 < n; j0 += kNB) {
    const int nb = std::min(kNB, n - j0);
    T* Ajj = A + (size_t)j0 * lda + j0;
    potrf_diag_kernel<T><<<1, kNB, 0, stream>>>(Ajj, lda, nb, j0, piv_tol, info_dev);
    count_launches(1);
    const int rows = n - j0 - nb;
    if (rows > 0) {
      T* panel = A + (size_t)(j0 + nb) * lda + j0;
      trsm_right_lt_kernel<T><<<(unsigned)ceil_div(rows, 128), 128, 0, stream>>>(Ajj, lda, nb, panel, lda, rows);
      count_launches(1);
      // trailing: A22 -= P P^T  (full square update; only the lower part is used afterwards)
      T* A22 = A + (size_t)(j0 + nb) * lda + (j0 + nb);
      int rc = gemm<T>(0, 1, rows, rows, nb, T(-1), panel, lda, panel, lda, T(1), A22, lda, stream);
      if (rc) return rc;
    }
  }
  CCAB_CUDA(cudaGetLastError());
  return 0;
}

template <typename T>
int trsm_left(int trans, int n, int m, const T* L, int64_t ldl, T* B, int64_t ldb, cudaStream_t stream) {
  CCAB_CHECK_ARG(n >= 1 && m >= 1 && ldl >= n && ldb >= m, "bad trsm shape");
  const int nblk = (int)ceil_div(n, kNB);
  if (!trans) {
    for (int bi = 0; bi < nblk; ++bi) {
      const int i0 = bi * kNB, nb = std::min(kNB, n - i0);
      if (i0 > 0) {  // B_i -= L[i, 0:i0] X[0:i0]
        int rc = gemm<T>(0, 0, nb, m, i0, T(-1), L + (size_t)i0 * ldl, ldl, B, ldb, T(1), B + (size_t)i0 * ldb, ldb,
                         stream);
        if (rc) return rc;
      }
      trsm_left_diag_kernel<T, 0><<<(unsigned)ceil_div(m, 128), 128, 0, stream>>>(L + (size_t)i0 * ldl + i0, ldl, nb,
                                                                                  B + (size_t)i0 * ldb, ldb, m);
      count_launches(1);
    }
  } else {
    for (int bi = nblk - 1; bi >= 0; --bi) {
      const int i0 = bi * kNB, nb = std::min(kNB, n - i0);
      const int below = n - i0 - nb;
      if (below > 0) {  // B_i -= L[i0+nb:, i]^T X[i0+nb:]
        int rc = gemm<T>(1, 0, nb, m, below, T(-1), L + (size_t)(i0 + nb) * ldl + i0, ldl,
                         B + (size_t)(i0 + nb) * ldb, ldb, T(1), B + (size_t)i0 * ldb, ldb, stream);
        if (rc) return rc;
      }
      trsm_left_diag_kernel<T, 1><<<(unsigned)ceil_div(m, 128), 128, 0, stream>>>(L + (size_t)i0 * ldl + i0, ldl, nb,
                                                                                  B + (size_t)i0 * ldb, ldb, m);
      count_launches(1);
    }
  }
  CCAB_CUDA(cudaGetLastError());
  return 0;
}

// B (rows x n) <- B L^-T, L n x n lower: block columns left to right
template <typename T>
int trsm_right_lt(int n, int rows, const T* L, int64_t ldl, T* B, int64_t ldb, cudaStream_t stream) {
  CCAB_CHECK_ARG(n >= 1 && rows >= 1 && ldl >= n && ldb >= n, "bad trsm shape");
  for (int j0 = 0; j0 < n; j0 += kNB) {
    const int nb = std::min(kNB, n - j0);
    if (j0 > 0) {  // B_j -= B[:, 0:j0] L[j, 0:j0]^T
      int rc = gemm<T>(0, 1, rows, nb, j0, T(-1), B, ldb, L + (size_t)j0 * ldl, ldl, T(1), B + j0, ldb, stream);
      if (rc) return rc;
    }
    trsm_right_lt_kernel<T><<<(unsigned)ceil_div(rows, 128), 128, 0, stream>>>(L + (size_t)j0 * ldl + j0, ldl, nb,
                                                                               B + j0, ldb, rows);
    count_launches(1);
  }
  CCAB_CUDA(cudaGetLastError());
  return 0;
}

template int potrf<float>(int, float*, int64_t, double, int*, cudaStream_t);
template int potrf<double>(int, double*, int64_t, double, int*, cudaStream_t);
template int trsm_left<float>(int, int, int, const float*, int64_t, float*, int64_t, cudaStream_t);
template int trsm_left<double>(int, int, int, const double*, int64_t, double*, int64_t, cudaStream_t);
template int trsm_right_lt<float>(int, int, const float*, int64_t, float*, int64_t, cudaStream_t);
template int trsm_right_lt<double>(int, int, const double*, int64_t, double*, int64_t, cudaStream_t);

}  // namespace ccab
