// Fused small-matrix stage of the deep-CCA objective (see ccaloss.cu).
#pragma once
#include "common.cuh"

namespace ccab {

// C: (d1+d2) x (d1+d2) block covariance of [z1 z2] (row-major).  Outputs (device): loss[1], G11 (d1 x d1),
// P (d1 x d2), G22 (d2 x d2), min_pivot[1] (smallest elimination pivot of S11 / S22 = squared Cholesky pivot).
template <typename T>
int ccaloss_small(const T* C, int64_t ldc, int d1, int d2, double eps, T* loss, T* G11, T* P, T* G22, T* min_pivot,
                  cudaStream_t stream);

}  // namespace ccab
