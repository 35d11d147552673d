// Cholesky factorisation and triangular solves (see chol.cu).
#pragma once
#include "common.cuh"

namespace ccab {

// In place: lower triangle of A (n x n row-major) <- L with A = L L^T; the strict upper triangle is not
// referenced.  *info_dev (device int) = 0 on success, else 1-based index of the first pivot <= piv_tol.
template <typename T>
int potrf(int n, T* A, int64_t lda, double piv_tol, int* info_dev, cudaStream_t stream);

// B (n x m, row-major) <- L^-1 B (trans = 0) or L^-T B (trans = 1), L n x n lower triangular.
template <typename T>
int trsm_left(int trans, int n, int m, const T* L, int64_t ldl, T* B, int64_t ldb, cudaStream_t stream);

// B (rows x n, row-major) <- B L^-T.
template <typename T>
int trsm_right_lt(int n, int rows, const T* L, int64_t ldl, T* B, int64_t ldb, cudaStream_t stream);

}  // namespace ccab
