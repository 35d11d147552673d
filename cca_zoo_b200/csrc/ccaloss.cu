// K6: fused small-matrix stage of the deep-CCA objective (cca_zoo/deep/objectives.py:86-102) for encoder
// widths d1, d2 <= 64: ONE single-CTA launch turns the block covariance of [z1 z2] into the loss and the
// three small matrices its analytic gradient needs.
//
//   S11 = C11 + eps I, S22 = C22 + eps I, S12 = C12
//   I1 = S11^-1, I2 = S22^-1      in-place Gauss-Jordan without pivoting (safe for SPD); its pivots are the
//                                 squared Cholesky pivots, whose minimum is returned so the caller can certify
//                                 that the reference's clamp(min=eps) of the eigenvalues is inactive
//   P  = I1 S12 I2                loss = -<P, S12> = -|| S11^-1/2 S12 S22^-1/2 ||_F^2
//   G11 = P S21 I1, G22 = I2 S21 P                      (dL/dS11, dL/dS22; dL/dS12 = -2 P)
//
// Everything lives in shared memory (6 matrices of 64 x 65), 1024 threads, two block barriers per
// elimination step.
#include "ccaloss.cuh"

namespace ccab {

constexpr int kLD = 64;
constexpr int kLP = kLD + 1;

template <typename T>
__device__ void spd_inverse_inplace(T* A, int d, T* rowk, T* colk, T* minpiv_s) {
  for (int k = 0; k < d; ++k) {
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
      rowk[i] = A[k * kLP + i];
      colk[i] = A[i * kLP + k];
    }
    __syncthreads();
    const T p = rowk[k];
    if (threadIdx.x == 0) *minpiv_s = fmin(*minpiv_s, p);
    const T ip = T(1) / p;
    for (int e = threadIdx.x; e < d * d; e += blockDim.x) {
      const int i = e / d, j = e % d;
      const T rkj = (j == k ? T(1) : rowk[j]) * ip;
      T v;
      if (i == k) v = rkj;
      else v = (j == k ? T(0) : A[i * kLP + j]) - colk[i] * rkj;
      A[i * kLP + j] = v;
    }
    __syncthreads();
  }
}

// Cm (m x n) = op(A) op(B) with k the contraction length; all operands in shared memory (stride kLP)
template <typename T, int TA, int TB>
__device__ void smem_matmul(const T* A, const T* B, T* Cm, int m, int n, int k) {
  for (int e = threadIdx.x; e < m * n; e += blockDim.x) {
    const int i = e / n, j = e % n;
    T acc = 0;
    for (int t = 0; t < k; ++t) {
      const T a = TA ? A[t * kLP + i] : A[i * kLP + t];
      const T b = TB ? B[j * kLP + t] : B[t * kLP + j];
      acc = fma(a, b, acc);
    }
    Cm[i * kLP + j] = acc;
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(1024) ccaloss_small_kernel(const T* __restrict__ C, int64_t ldc, int d1, int d2,
                                                             T eps, T* __restrict__ loss, T* __restrict__ G11,
                                                             T* __restrict__ Pout, T* __restrict__ G22,
                                                             T* __restrict__ min_pivot) {
  extern __shared__ __align__(16) unsigned char ccl_smem[];
  T* I1 = reinterpret_cast<T*>(ccl_smem);   // S11 -> S11^-1
  T* I2 = I1 + kLD * kLP;                    // S22 -> S22^-1
  T* S12 = I2 + kLD * kLP;
  T* Tm = S12 + kLD * kLP;
  T* Pm = Tm + kLD * kLP;
  T* Gm = Pm + kLD * kLP;
  T* rowk = Gm + kLD * kLP;                  // [64]
  T* colk = rowk + kLD;                      // [64]
  T* red = colk + kLD;                       // [32]
  T* minpiv = red + 32;                      // [1]
  for (int e = threadIdx.x; e < d1 * d1; e += blockDim.x) {
    const int i = e / d1, j = e % d1;
    I1[i * kLP + j] = C[(size_t)i * ldc + j] + (i == j ? eps : T(0));
  }
  for (int e = threadIdx.x; e < d2 * d2; e += blockDim.x) {
    const int i = e / d2, j = e % d2;
    I2[i * kLP + j] = C[(size_t)(d1 + i) * ldc + d1 + j] + (i == j ? eps : T(0));
  }
  for (int e = threadIdx.x; e < d1 * d2; e += blockDim.x) {
    const int i = e / d2, j = e % d2;
    S12[i * kLP + j] = C[(size_t)i * ldc + d1 + j];
  }
  if (threadIdx.x == 0) *minpiv = T(3.0e38);
  __syncthreads();
  spd_inverse_inplace(I1, d1, rowk, colk, minpiv);
  spd_inverse_inplace(I2, d2, rowk, colk, minpiv);
  smem_matmul<T, 0, 0>(I1, S12, Tm, d1, d2, d1);      // Tm = I1 S12
  smem_matmul<T, 0, 0>(Tm, I2, Pm, d1, d2, d2);       // P  = I1 S12 I2
  T acc = 0;
  for (int e = threadIdx.x; e < d1 * d2; e += blockDim.x) {
    const int i = e / d2, j = e % d2;
    acc = fma(Pm[i * kLP + j], S12[i * kLP + j], acc);
    Pout[(size_t)i * d2 + j] = Pm[i * kLP + j];
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : T(0);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) {
      loss[0] = -acc;
      min_pivot[0] = *minpiv;
    }
  }
  __syncthreads();
  smem_matmul<T, 0, 1>(Pm, S12, Tm, d1, d1, d2);      // Tm = P S12^T            (d1 x d1)
  smem_matmul<T, 0, 0>(Tm, I1, Gm, d1, d1, d1);       // G11 = P S21 I1
  for (int e = threadIdx.x; e < d1 * d1; e += blockDim.x) G11[e] = Gm[(e / d1) * kLP + e % d1];
  __syncthreads();
  smem_matmul<T, 1, 0>(S12, Pm, Tm, d2, d2, d1);      // Tm = S12^T P            (d2 x d2)
  smem_matmul<T, 0, 0>(I2, Tm, Gm, d2, d2, d2);       // G22 = I2 S21 P
  for (int e = threadIdx.x; e < d2 * d2; e += blockDim.x) G22[e] = Gm[(e / d2) * kLP + e % d2];
}

template <typename T>
int ccaloss_small(const T* C, int64_t ldc, int d1, int d2, double eps, T* loss, T* G11, T* P, T* G22, T* min_pivot,
                  cudaStream_t stream) {
  CCAB_CHECK_ARG(d1 >= 1 && d2 >= 1 && d1 <= kLD && d2 <= kLD, "ccaloss_small supports widths 1..64, got %d, %d", d1,
                 d2);
  CCAB_CHECK_ARG(ldc >= d1 + d2, "ldc too small");
  const size_t smem = sizeof(T) * (6 * kLD * kLP + 2 * kLD + 33);
  static bool attr = false;
  if (!attr) {
    CCAB_CUDA(cudaFuncSetAttribute(ccaloss_small_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  ccaloss_small_kernel<T><<<1, 1024, smem, stream>>>(C, ldc, d1, d2, (T)eps, loss, G11, P, G22, min_pivot);
  count_launches(1);
  CCAB_CUDA(cudaGetLastError());
  return 0;
}

template int ccaloss_small<float>(const float*, int64_t, int, int, double, float*, float*, float*, float*, float*,
                                  cudaStream_t);
template int ccaloss_small<double>(const double*, int64_t, int, int, double, double*, double*, double*, double*,
                                   double*, cudaStream_t);

}  // namespace ccab
