// Batched blocked Cholesky with explicit inverse of the factor (see cholinv.cu).
#pragma once
#include "common.cuh"

namespace ccab {

// For each of `batch` SPD matrices A_b (n x n row-major, lda, A + b * strideA; lower triangle referenced):
//   lower triangle of A_b <- L_b (A_b = L_b L_b^T; the strict upper triangle of the diagonal blocks is zeroed, the
//   rest of the upper triangle is left as it was), Linv_b (n x n, ldi, Linv + b * strideLinv) <- L_b^-1 (full
//   matrix, exact zeros above the diagonal).
// info[b] (device) = 0, or the 1-based index of the first pivot <= piv_tol (the results are then meaningless).
template <typename T>
size_t potrf_inv_workspace_bytes(int n, int batch);
template <typename T>
int potrf_inv(int n, int batch, T* A, int64_t lda, int64_t strideA, T* Linv, int64_t ldi, int64_t strideLinv,
              double piv_tol, const double* piv_tol_dev /* device, [batch], overrides piv_tol when non-NULL */, int* info,
              void* ws, size_t ws_bytes, cudaStream_t stream);

// One diagonal block (nb <= potrf_inv_block_size<T>()) per matrix, single launch: factor in place and write the
// dense NB x NB inverse (leading dimension NB, zero padded) to Dinv + b * strideDinv.  info is NOT cleared here.
template <typename T>
int potrf_inv_block_size();
template <typename T>
int potrf_inv_block(T* A, int64_t lda, int64_t strideA, int nb, int j0, T* Dinv, int64_t strideDinv, double piv_tol,
                    const double* piv_tol_dev, int* info, int batch, cudaStream_t stream);

}  // namespace ccab
