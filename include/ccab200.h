/* libccab200 -- C ABI of the B200-native CCA hot path.
 *
 * The reference (jameschapman19/cca_zoo) is pure Python and has NO FFI boundary for this path
 * (SURVEY.md §8b); this header is the boundary a maintainer would bind with ctypes (see
 * INTEGRATION.md).  Each entry point names the reference code it replaces.
 *
 * Conventions
 *   - every pointer marked "device" is a CUDA device pointer owned by the caller (PyTorch's caching
 *     allocator in the Python binding); the library never allocates device memory, the caller passes
 *     a workspace sized by the matching *_workspace_bytes call;
 *   - matrices are row-major with an explicit leading dimension unless stated otherwise;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it, except ccab_syevj /
 *     ccab_gesvj which synchronise it once per Jacobi sweep to read the convergence flag;
 *   - return 0 = OK, <0 = bad argument / unsupported, >0 = cudaError_t.  ccab_last_error() gives the
 *     message of the last failure on the calling thread.  No C++ exception crosses the ABI.
 *   - dtype: CCAB_F32 / CCAB_F64.  There is no CPU fallback: without a sm_100a device every compute
 *     entry point fails.
 */
#ifndef CCAB200_H
#define CCAB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CCAB_F32 0
#define CCAB_F64 1

/* arithmetic of the moment kernel (ccab_moments `precision`) */
#define CCAB_PREC_TF32 0   /* fp32 in, one tcgen05 kind::tf32 pass, fp32 accumulate               */
#define CCAB_PREC_TF32X3 1 /* fp32 in, hi/lo split + 3 tcgen05 passes: fp32-grade accuracy          */
#define CCAB_PREC_EXACT 2  /* FMA in the input dtype on CUDA cores (the only choice for CCAB_F64)  */
#define CCAB_PREC_TF32X3B 3 /* 3xTF32 with the two cross terms as bf16 MMAs (kind::f16): fp32-grade, 2/3 the tensor work */

#define CCAB_MAX_VIEWS 8

int ccab_version(void);
const char* ccab_last_error(void);
/* kernels launched by this library since it was loaded (bench.py's gpu_launches) */
int64_t ccab_launch_count(void);

/* ---- K1: block moments --------------------------------------------------------------------------
 * M = [X_1 .. X_m]^T [X_1 .. X_m] and s = 1^T [X_1 .. X_m] over the n_rows samples this process
 * holds.  Output `moments` (device, double) has ccab_moments_size() entries: a Dp x Dp padded matrix
 * (each view padded to a multiple of 128 columns; only the upper block triangle is meaningful, the rest
 * is zero) followed by the Dp column sums.  The buffer is additive over row shards: all-reduce(sum) it
 * across ranks before ccab_covariance.
 * Replaces: np.linalg.svd(X) cca_zoo/_utils/_linalg.py:28, X1_w.T @ X2_w cca_zoo/linear/_rcca.py:96,
 * np.cov(...) cca_zoo/linear/_mcca.py:150-152,166 and cca_zoo/linear/_gcca.py:101,
 * z.T @ z cca_zoo/deep/objectives.py:86-92; the column sums replace v.mean(axis=0) cca_zoo/_base.py:97.
 * views[v]: device pointer to an n_rows x dims[v] row-major array with leading dimension lds[v]
 * (TF32 paths need 16-byte aligned pointers and lds[v] % 4 == 0). */
int64_t ccab_moments_size(int n_views, const int64_t* dims);
int64_t ccab_moments_padded_dim(int n_views, const int64_t* dims);
size_t ccab_moments_workspace_bytes(int dtype, int precision, int n_views, const int64_t* dims, int64_t n_rows);
int ccab_moments(int dtype, int precision, int n_views, const void* const* views, const int64_t* dims,
                 const int64_t* lds, int64_t n_rows, double* moments, void* workspace, size_t workspace_bytes,
                 void* stream);

/* ---- shifted accumulation (numerical safety of the one-pass covariance) -----------------------------------------------
 * (M - s s^T / n) cancels catastrophically when a column's mean dominates its spread: the relative error of the
 * covariance is eps_prod * (mean / std)^2, eps_prod ~ 1e-7 .. 1e-6 for float32 inputs.  The reference centres the data
 * first (cca_zoo/_base.py:96-99); covariance being shift invariant, the same is had in one pass by accumulating the
 * moments of X - x0 and rebuilding the raw moments in float64:
 *   ccab_column_pilot   x0[j] = mean of column j over the first `rows` rows (device, in the view's dtype);
 *                       ratio_max_dev[0] = max(ratio_max_dev[0], max_j mean_j^2 / var_j)   (zero it first)
 *   ccab_shift_rows     Xs = X - x0   (one extra pass, only taken when the pilot says it matters)
 *   ccab_moments        ... on Xs ...
 *   ccab_moments_unshift  M += x0 s^T + s x0^T + n x0 x0^T, s += n x0 in float64 (x0: per view device pointers or NULL)
 * after which the buffer holds the raw moments again: additive over row shards, all-reducible, whatever x0 each rank
 * chose. */
int ccab_column_pilot(int dtype, const void* X, int64_t rows, int d, int64_t ld, void* x0, float* ratio_max_dev,
                      void* stream);
int ccab_shift_rows(int dtype, const void* X, int64_t n, int d, int64_t ldx, const void* x0, void* Xs, int64_t lds,
                    void* stream);
int ccab_moments_unshift(int dtype, int n_views, const int64_t* dims, double* moments, const void* const* x0,
                         double n_rows, void* stream);

/* ---- exchange step of a sample-sharded fit ---------------------------------------------------------
 * ccab_moments_pack gathers what the all-reduce has to carry into ONE contiguous float64 message of
 * ccab_moments_packed_size() entries: the upper triangle of 128 x 128 blocks of M (row-major over block pairs),
 * the column sums, the local sample count n and one reserved slot.  Sum it over the ranks (NCCL all-reduce over
 * NVLink), then ccab_moments_unpack restores the moment buffer of ccab_moments (zero below the block diagonal);
 * packed[size - 2] is the total sample count, which ccab_rcca_fit / ccab_mcca_fit read on the device (n_total_dev). */
int64_t ccab_moments_packed_size(int n_views, const int64_t* dims);
int ccab_moments_pack(int n_views, const int64_t* dims, const double* moments, double n_local, double* packed,
                      void* stream);
int ccab_moments_unpack(int n_views, const int64_t* dims, const double* packed, double* moments, void* stream);

/* Fused exchange step over NVLink / NVSwitch: pack -> in-switch all-reduce -> unpack in ONE kernel, no NCCL call.
 * Needs a symmetric-memory buffer of `sym_doubles` float64 (>= world * ceil(ccab_moments_packed_size / world),
 * rounded up to even) mapped on every rank with a multicast (NVLS) address, and the ranks' signal pads (uint32 flags,
 * `pad_slots` entries each, zeroed once) -- torch.distributed._symmetric_memory provides both (cca_zoo_b200/parallel.py).
 * Each CTA packs its column of the message, meets the peers on its own flag row (st.release.sys / ld.acquire.sys),
 * reduces this rank's slice with multimem.ld_reduce.add.f64 (the switch adds the ranks' copies), broadcasts it with
 * multimem.st, meets the peers again and scatters the totals into `moments`; n_total_out (device, may be NULL)
 * receives the summed sample count.  `epoch` must be the same on every rank and grow by 2 per call.
 * The result is bit-identical on all ranks (every element is reduced once, in the switch). */
int ccab_moments_exchange_nvls(int n_views, const int64_t* dims, double* moments, double n_local, double* sym_local,
                               double* sym_multicast, void* const* signal_pads_dev, int rank, int world,
                               int pad_slots, int64_t sym_doubles, unsigned epoch, double* n_total_out, void* stream);

/* ---- K2: covariance from (all-reduced) moments ---------------------------------------------------
 * C = (M - s s^T / n_total) / (n_total - 1) (center != 0) or M / (n_total - 1), compact D x D
 * (D = sum dims, hstack order), full symmetric, in out_dtype; mean = s / n_total (or 0).
 * Replaces: the centring of cca_zoo/_base.py:96-99 plus the 1/(n-1) scalings listed above. */
int ccab_covariance(int out_dtype, int n_views, const int64_t* dims, const double* moments, double n_total,
                    int center, void* C, int64_t ldc, void* mean, void* stream);

/* ---- K3: batched symmetric eigensolver (one-sided block Jacobi) ----------------------------------
 * For each of `batch` symmetric n x n matrices A_b (device, row-major == column-major, lda, stride
 * batch_stride elements): eigenvalues descending into evals[b*n ..], eigenvectors as ROWS of
 * evecs_t[b] (n x n, ldv): evecs_t[b][j,:] is the unit eigenvector of the j-th largest eigenvalue.
 * `shift` is added to the diagonal before solving and removed from the eigenvalues afterwards; pass a
 * value >= -lambda_min for indefinite matrices (the one-sided method needs A + shift*I to be PSD to
 * separate +/- pairs).  info[0] (host, may be NULL) = sweeps used, NEGATED when the iteration ran out of sweeps before
 * reaching the tolerance (with info == NULL such a call fails instead), info_offdiag (host, may be NULL) =
 * last normalised off-diagonal.
 * Replaces: scipy.linalg.eigh cca_zoo/_utils/_linalg.py:64-71, np.linalg.eigvalsh
 * cca_zoo/linear/_mcca.py:170,194 cca_zoo/linear/_gcca.py:102, sklearn PCA cca_zoo/linear/_mcca.py:117,
 * torch.linalg.eigh cca_zoo/deep/objectives.py:19, and (via covariance form) the tall SVD of
 * cca_zoo/_utils/_linalg.py:28. */
size_t ccab_syevj_workspace_bytes(int dtype, int n, int batch);
int ccab_syevj(int dtype, int n, int batch, const void* A, int64_t lda, int64_t batch_stride, double shift,
               void* evals, void* evecs_t, int64_t ldv, int* info, float* info_offdiag, void* workspace,
               size_t workspace_bytes, void* stream);

/* Small symmetric eigenproblems (n <= 128 float / 96 double), batched, ONE single-CTA launch per matrix, no host
 * synchronisation: two-sided Jacobi with tournament ordering in shared memory.  Same outputs as ccab_syevj
 * (descending eigenvalues, eigenvectors as rows; evals / evecs_t contiguous per matrix: n and n x ldv); absolute
 * accuracy eps * ||A|| (use ccab_syevj when tiny eigenvalues matter relatively).  info_dev[b] (device int[batch],
 * may be NULL) = sweeps used, negated if the tolerance was not reached.
 * Replaces the Rayleigh-Ritz eigensolve of the top-k route (scipy.linalg.eigh subset_by_index,
 * cca_zoo/_utils/_linalg.py:64-73). */
int ccab_syevj_small(int dtype, int n, int batch, const void* A, int64_t lda, int64_t stride_a, void* evals,
                     void* evecs_t, int64_t ldv, int* info_dev, void* stream);

/* ---- K4: singular value decomposition (one-sided Jacobi) -----------------------------------------
 * G is m x n given by COLUMNS: column j is the contiguous array A + j*lda (length m) -- i.e. a
 * row-major n x m buffer holds G^T.  Outputs (descending): sigma[n]; right_t (n x n, ldr): row j =
 * j-th right singular vector; left_t (n x m, ldl): row j = j-th left singular vector (unit, length m;
 * zero for sigma_j = 0).  Any output may be NULL.
 * Replaces: np.linalg.svd(cross_cov) cca_zoo/linear/_rcca.py:97. */
size_t ccab_gesvj_workspace_bytes(int dtype, int m, int n);
int ccab_gesvj(int dtype, int m, int n, const void* A, int64_t lda, void* sigma, void* right_t, int64_t ldr,
               void* left_t, int64_t ldl, int* info, float* info_offdiag, void* workspace, size_t workspace_bytes,
               void* stream);

/* ---- dense glue ----------------------------------------------------------------------------------
 * C (m x n) = alpha * op(A) * op(B) + beta * C, row-major; transX != 0 means op(X) = X^T.
 * float32 operands that TMA can address (16-byte aligned, leading dimensions % 4 == 0) run on the tensor pipe
 * (ccab_gemm_tc: 3xTF32, fp32-grade); everything else as exact FMA tiles.
 * Replaces the small products W1^T C12 W2, W @ U (cca_zoo/linear/_rcca.py:96,100), components_.T @ w
 * (cca_zoo/linear/_mcca.py:131) and the S11^-1/2 S12 S22^-1/2 chain (cca_zoo/deep/objectives.py:97). */
int ccab_gemm(int dtype, int transa, int transb, int m, int n, int k, double alpha, const void* A, int64_t lda,
              const void* B, int64_t ldb, double beta, void* C, int64_t ldc, void* stream);

/* Tensor-core variant of ccab_gemm for float32 (tcgen05.mma kind::tf32, 3xTF32 split formed in shared memory:
 * fp32-grade products, fp32 accumulation in TMEM), batched: matrix b of the batch lives at X + b * stride_x.
 * Optionally also (or only: C may be NULL when beta == 0) writes the transpose Ct (n x m, row-major, ldct).
 * lower_only != 0 skips the 128-row output tiles that lie strictly above the diagonal (SYRK-type updates).
 * Needs 16-byte aligned A / B with lda, ldb (and the batch strides) multiples of 4; otherwise returns < 0 and the
 * caller uses ccab_gemm.  Replaces the same reference products as ccab_gemm, on the tensor pipe. */
int ccab_gemm_tc(int transa, int transb, int m, int n, int k, double alpha, const void* A, int64_t lda,
                 int64_t stride_a, const void* B, int64_t ldb, int64_t stride_b, double beta, void* C, int64_t ldc,
                 int64_t stride_c, void* Ct, int64_t ldct, int64_t stride_ct, int batch, int lower_only,
                 void* stream);

/* Whitening rows from an eigendecomposition (covariance form of svd_whiten,
 * cca_zoo/_utils/_linalg.py:30-38; also B^-1/2 of cca_zoo/linear/_mcca.py:163-173 and R_i of
 * cca_zoo/linear/_gcca.py:101-105):
 *   keep_j  = lam[j] > rank_tol * max(lam[0],0)  &&  j < max_rank
 *   g_j     = keep_j ? ((1-c)*max(lam[j],lam_floor) + c + floor_add + (floor_dev ? *floor_dev : 0))^-1/2
 *                      * scale^-1/2 : 0
 *   Wt[j,:] = g_j * Vt[j,:]            g_out[j] = g_j (may be NULL)      *rank_out = #kept (device int)
 * lam_floor = 0 with c = 0, floor_add = eps reproduces clamp(eigh(S + eps I), min=eps)
 * (cca_zoo/deep/objectives.py:19-21); pass -inf-like (e.g. -1e300) to disable the clamp. */
int ccab_whiten_rows(int dtype, int d, const void* lam, const void* Vt, int64_t ldv, double c, double floor_add,
                     const void* floor_dev, double scale, double rank_tol, int max_rank, double lam_floor, void* Wt,
                     int64_t ldw, void* g_out, int* rank_out, void* stream);

/* ---- K6: fused small-matrix stage of the deep-CCA objective (widths d1, d2 <= 64) -------------------
 * From the (d1+d2)^2 block covariance C of [z1 z2] (device, row-major): S_ii = C_ii + eps I,
 * P = S11^-1 S12 S22^-1, loss[0] = -<P, S12> = -||S11^-1/2 S12 S22^-1/2||_F^2, G11 = P S21 S11^-1 (d1 x d1),
 * G22 = S22^-1 S21 P (d2 x d2), P (d1 x d2), and min_pivot[0] = the smallest elimination pivot of S11/S22:
 * min_pivot > 4 eps certifies that the eigenvalue clamp of the reference is inactive (otherwise callers take
 * the eigen route).  One single-CTA launch.
 * Replaces: _inv_sqrtm x2, the T / T^T T products and eigvalsh of cca_zoo/deep/objectives.py:94-102 (forward)
 * and supplies the matrices of the analytic backward. */
int ccab_ccaloss_small(int dtype, int d1, int d2, const void* C, int64_t ldc, double eps, void* loss, void* G11,
                       void* P, void* G22, void* min_pivot, void* stream);

/* ---- Cholesky route of the generalised problem ----------------------------------------------------
 * ccab_potrf: lower triangle of A (n x n, row-major, device) <- L with A = L L^T, in place (the strict
 * upper triangle is not referenced).  *info_dev (device int): 0, or the 1-based index of the first pivot
 * <= pivot_tol (matrix not numerically positive definite: callers fall back to the eigen route).
 * ccab_trsm: side 0: B (n x m) <- L^-1 B (trans 0) or L^-T B (trans 1); side 1 (trans must be 1):
 * B (m x n) <- B L^-T.  L is n x n lower triangular.
 * Replaces the Cholesky + back-substitution inside scipy.linalg.eigh(A, B) (LAPACK *sygvx,
 * cca_zoo/_utils/_linalg.py:67-71) and, in Cholesky form, the whitening of _linalg.py:30-38:
 * with R_i = (1-c) C_ii + c I = L_i L_i^T,  T = L_1^-1 C_12 L_2^-T and weights_i = L_i^-T U_k. */
int ccab_potrf(int dtype, int n, void* A, int64_t lda, double pivot_tol, int* info_dev, void* stream);
int ccab_trsm(int dtype, int side, int trans, int n, int m, const void* L, int64_t ldl, void* B, int64_t ldb,
              void* stream);

/* Batched blocked Cholesky WITH the explicit inverse of the factor (the GEMM-friendly form of the whitening):
 * for each of `batch` SPD matrices A_b = A + b * stride_a (n x n row-major, lower triangle referenced)
 *   lower triangle of A_b <- L_b,   Linv_b = Linv + b * stride_i (n x n, ldi) <- L_b^-1 (zeros above the diagonal).
 * info_dev[b] (device int[batch]) = 0 or the 1-based index of the first pivot <= pivot_tol.
 * Diagonal blocks (128 wide for float, 64 for double) are factored AND inverted by one single-CTA launch each
 * (warp-synchronous 32 x 32 sub-blocks); panels, trailing updates and the assembly of L^-1 by recursive doubling are
 * GEMMs (tcgen05 for float).  With Linv,  T = L1^-1 C12 L2^-T  and the weights  L_i^-T U_k  are plain products.
 * Replaces LAPACK potrf / trsm inside scipy.linalg.eigh(A, B) (cca_zoo/_utils/_linalg.py:67-71) and, in Cholesky
 * form, the whitening of cca_zoo/_utils/_linalg.py:30-38 and _inv_sqrtm of cca_zoo/deep/objectives.py:9-21. */
size_t ccab_potrf_inv_workspace_bytes(int dtype, int n, int batch);
int ccab_potrf_inv(int dtype, int n, int batch, void* A, int64_t lda, int64_t stride_a, void* Linv, int64_t ldi,
                   int64_t stride_i, double pivot_tol, int* info_dev, void* workspace, size_t workspace_bytes,
                   void* stream);

/* ---- the fit behind the ABI: rCCA / CCA / PLS ----------------------------------------------------------------------
 * From the (all-reduced) moment buffer of ccab_moments to the weights in ONE asynchronous call: covariance, ridge
 * blocks R_i = (1-c_i) C_ii + c_i I, batched Cholesky + inverse, T = L1^-1 C12 L2^-T, the leading k singular triplets
 * of T by blocked subspace iteration (block width p >= k, `iters` products with T^T T, CholQR, Rayleigh-Ritz) and
 * weights_i = L_i^-T U_k / V_k.  Nothing is read back by the library: every decision that needs a host (pivot
 * failures, rank loss, convergence, NaN / inf in the input, n <= d) is reported in the header of the result block.
 *   result (device, 256-byte aligned, ccab_rcca_fit_result_layout offsets[4] bytes):
 *     double header[32] | double mean[D] | T sigma[k] | T W1[d1 x k] | T W2[d2 x k]      (offsets[0..3] = byte offsets)
 *     header[0] = status bits (0 = valid): 1 a block is not positive definite, 2 not converged, 4 non-finite input,
 *                 8 too few samples (n <= max d_i);  header[1] = n_total;  [2] residual;  [3] sigma_1;
 *                 [4] index of the first failed factorisation;  [5] sweeps of the Ritz eigensolve
 *   n_total_dev (device double, may be NULL) overrides n_total: the sample count can ride in the all-reduced message.
 *   p must satisfy k <= p <= min(d1, d2, 128).  dtype = arithmetic of the whole solve (CCAB_F32 uses tcgen05 GEMMs).
 * Replaces cca_zoo/linear/_rcca.py:83-101 (via cca_zoo/_utils/_linalg.py:9-41) after the moment pass. */
size_t ccab_rcca_fit_workspace_bytes(int dtype, const int64_t* dims, int k, int p);
int ccab_rcca_fit_result_layout(int dtype, const int64_t* dims, int k, int p, int64_t* offsets /* [5] */);
int ccab_rcca_fit(int dtype, const int64_t* dims, const double* moments, const double* n_total_dev, double n_total,
                  int center, const double* c, int k, int p, int iters, void* result, size_t result_bytes,
                  void* workspace, size_t workspace_bytes, void* stream);

/* ---- the fit behind the ABI: MCCA -------------------------------------------------------------------------------------
 * Same contract as ccab_rcca_fit for m >= 2 views: B_i = (1-c_i) C_ii + c_i I = L_i L_i^T (batched Cholesky + inverse
 * when the views share one width), K_ij = L_i^-1 C_ij L_j^-T, the k largest eigenpairs of K by blocked subspace
 * iteration on K + I / (1 - max c) with a Rayleigh-Ritz step, v_i = sqrt(m) L_i^-T y_i (v^T (B/m) v = 1 as scipy).
 * Result block: double header[32] | double mean[D] | T eigenvalues[k] | T W_1[d_1 x k] | .. | T W_m; offsets has m + 3
 * entries (mean, eigenvalues, W_1 .. W_m, total).  `eps` is the reference's floor on lambda_min(B): a block whose pivots
 * fall below it fails the factorisation (status bit 1) and the caller takes the eigen route, which applies the floor.
 * Needs max c <= 0.9 and k <= p <= 128.
 * Replaces cca_zoo/linear/_mcca.py:113-173 (_build_A, _build_B, gevp of cca_zoo/_utils/_linalg.py:44-73). */
size_t ccab_mcca_fit_workspace_bytes(int dtype, int n_views, const int64_t* dims, int k, int p);
int ccab_mcca_fit_result_layout(int dtype, int n_views, const int64_t* dims, int k, int p, int64_t* offsets);
int ccab_mcca_fit(int dtype, int n_views, const int64_t* dims, const double* moments, const double* n_total_dev,
                  double n_total, int center, const double* c, double eps, int k, int p, int iters, void* result,
                  size_t result_bytes, void* workspace, size_t workspace_bytes, void* stream);

/* ---- the deep-CCA objective behind the ABI (any widths) ---------------------------------------------------------
 * ccab_ccaloss_fwd: loss[0] = -|| S11^-1/2 S12 S22^-1/2 ||_F^2 with S_ii = cov(z_i) + eps I, from the moment pass over
 * [z1 z2] (precision as in ccab_moments), a batched Cholesky + inverse and 7 GEMMs; `saved`
 * (T[d1*d1 + d1*d2 + d2*d2 + d1 + d2]) receives G11 = P S21 S11^-1 | P = S11^-1 S12 S22^-1 | G22 = S22^-1 S21 P | the
 * column means of z1, z2 (written by the fused path for widths <= 64) for the backward.
 * Nothing is read back: flags_dev (device int[3]) = Cholesky status of S11, S22 (a pivot^2 <= eps / 4 counts as a failure:
 * rounding destroyed the ridge; the caller re-runs through the eigen route, which clamps like the reference) and a
 * non-finite-input flag -- check lazily.
 * ccab_ccaloss_bwd: g1 = 2/(n-1) center(z1 G11 - z2 P^T) * grad_out[0], g2 = 2/(n-1) center(z2 G22 - z1 P) * grad_out[0]
 * (grad_out: device scalar, may be NULL = 1).  4 tall GEMMs (tcgen05 for float) + the centring (slab partial sums in a
 * 2 MB per-device scratch that this entry point allocates on first use -- the one exception to "the library never
 * allocates": it takes no workspace argument); widths <= 64 run as ONE fused launch instead.
 * Replaces cca_zoo/deep/objectives.py:9-21,79-102 and torch autograd through two eigh + eigvalsh. */
size_t ccab_ccaloss_workspace_bytes(int dtype, int precision, int d1, int d2, int64_t n);
int ccab_ccaloss_fwd(int dtype, int precision, const void* z1, int64_t ld1, const void* z2, int64_t ld2, int64_t n,
                     int d1, int d2, double eps, void* loss, void* saved, int* flags_dev, void* workspace,
                     size_t workspace_bytes, void* stream);
int ccab_ccaloss_bwd(int dtype, const void* z1, int64_t ld1, const void* z2, int64_t ld2, int64_t n, int d1, int d2,
                     const void* saved, const void* grad_out, void* g1, int64_t ldg1, void* g2, int64_t ldg2,
                     void* stream);

/* B[i,j] = A[i,j] * f(r[i]) * f(c[j]); r / c may be NULL; *_pow: 0 -> x, 1 -> 1/x, 2 -> 1/sqrt(x).
 * (column scalings such as diag(sigma)^-1/2 in the GCCA back-substitution, cca_zoo/linear/_gcca.py:109) */
int ccab_scale(int dtype, int m, int n, const void* A, int64_t lda, const void* r, int r_pow, const void* c,
               int c_pow, void* B, int64_t ldb, void* stream);

/* A[:, j] -= mean_i A[i, j] in place (the centring Jacobian of cca_zoo/deep/objectives.py:83-84) */
int ccab_center_columns(int dtype, int m, int n, void* A, int64_t lda, void* stream);

/* out[0] (device) = ||A||_F of an m x n row-major matrix */
int ccab_frobenius_norm(int dtype, int m, int n, const void* A, int64_t lda, void* out, void* stream);

/* Measurement hook: when enabled, CUDA events are recorded on the caller's stream immediately around the
 * tcgen05 moment-kernel launch of ccab_moments (TF32 paths); ccab_profile_moments_last_ms() waits for the
 * last pair and returns the kernel's duration in ms (-1 if none). */
int ccab_profile_moments(int enable);
double ccab_profile_moments_last_ms(void);

/* Debug/tuning knobs ("lbo_bytes", "sbo_bytes", "tma_dtype", "force_splits", "tc_variant", "tc_kc", "x3_split",
 * "x3b_oneshot" = 1: one (tile, split) unit per CTA pair instead of the persistent moment kernel, "f64_simt",
 * "gemm_force_fma", "gemm_split" = 0: never split thin float32 products over k, "jacobi_inner_sweeps",
 * "jacobi_force_unfused"); value < 0 restores the default.  Not part of the stable surface. */
int ccab_debug_set(const char* key, int value);

#ifdef __cplusplus
}
#endif
#endif /* CCAB200_H */
