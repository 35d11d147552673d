// Device-side fit assembly (see fit.cu): moments -> weights without a host round trip.
#pragma once
#include "common.cuh"
#include "moments.cuh"

namespace ccab {

// Result block = [ header: double[kFitHeaderDoubles] | mean: double[D] | sigma: T[k] | W_1: T[d_1 x k] | W_2 ... ],
// every section 256-byte aligned (offsets from *_result_layout).  Header:
//   [0] status bit mask (0 = the weights are valid)   [1] n_total   [2] residual ||T^T U - V diag(sigma)||_F
//   [3] sigma_1   [4] 1-based index of the first failed Cholesky (0 = none)   [5] sweeps of the Ritz eigensolve
constexpr int kFitHeaderDoubles = 32;
constexpr int kFitNotPositiveDefinite = 1;  // a ridge block (or a CholQR Gram matrix) failed the pivot test
constexpr int kFitNotConverged = 2;         // the subspace iteration / Ritz solve missed the tolerance
constexpr int kFitNonFinite = 4;            // NaN / inf in the moments (i.e. in the input)
constexpr int kFitTooFewSamples = 8;        // n <= max(d_i): the covariance blocks are rank deficient by construction

template <typename T>
size_t rcca_fit_workspace_bytes(int d1, int d2, int k, int p);
template <typename T>
void rcca_fit_result_layout(int d1, int d2, int k, int p, int64_t* offsets /* mean, sigma, W1, W2, total */);
// c: host double[2].  n_dev (device, may be NULL) overrides n_host.  p = width of the iterated block (k + oversampling).
template <typename T>
int rcca_fit(const ColumnLayout& L, const double* moments, const double* n_dev, double n_host, int center,
             const double* c, int k, int p, int iters, void* result, size_t result_bytes, void* ws, size_t ws_bytes,
             cudaStream_t stream);

}  // namespace ccab
