// Small symmetric eigensolver: ONE single-CTA launch per matrix (n <= 128 float / 96 double), everything in shared
// memory, no host involvement -- the Rayleigh-Ritz step of the subspace iteration (p x p, p = k + oversampling)
// and any other small, well-conditioned symmetric problem of the solver stage.
//
// Classical two-sided Jacobi with the round-robin (tournament) parallel ordering: per step n/2 disjoint index pairs
// are rotated at once.  Thread i < n/2 computes the rotation of pair i from the current 2 x 2 pivot block; then the
// (n/2)^2 2 x 2 blocks  H[{p_i,q_i}, {p_j,q_j}]  are each updated by ONE thread (row rotation of pair i, column
// rotation of pair j -- the blocks partition H, so there is no hazard) and the eigenvector rows are rotated
// alongside.  Two block barriers per step, n - 1 steps per sweep; the off-diagonal mass seen during a sweep decides
// convergence on the device.  Eigenvalues are returned in descending order with the eigenvectors as rows.
//
// Two-sided Jacobi resolves eigenvalues to eps * ||H|| (absolute): right for the Ritz problem, whose block is well
// conditioned; the whitening eigenproblems (tiny eigenvalues matter relatively) stay on the one-sided solver of
// syevj.cu.  Replaces, with the subspace iteration around it, scipy.linalg.eigh / np.linalg.svd of
// cca_zoo/_utils/_linalg.py:64-73 and cca_zoo/linear/_rcca.py:97 when only the leading k pairs are wanted.
#include "syevj_small.cuh"

#include <cmath>

namespace ccab {

namespace {

constexpr int kSmallThreads = 1024;
constexpr int kSmallMaxN = 128;                      // (n/2)^2 2 x 2 blocks over 1024 threads: 4 per thread
constexpr int kHB = (kSmallMaxN / 2) * (kSmallMaxN / 2) / kSmallThreads;   // H blocks per thread
constexpr int kVB = 2;                                // V items per thread: (n/2) * ceil(n/4) <= 2048

template <typename T>
struct SmallEps;
template <>
struct SmallEps<float> {
  static constexpr float v = 1.1920929e-7f;
};
template <>
struct SmallEps<double> {
  static constexpr double v = 2.220446049250313e-16;
};

// rotation of one index pair, as every thread needs it: the indices and Rutishauser's (s, tau = s / (1 + c))
template <typename T>
struct alignas(16) Rot {
  int p, q;
  T s, tau;
};
template <>
struct alignas(32) Rot<double> {
  int p, q;
  double s, tau;
  double pad;
};

template <typename T>
__global__ void __launch_bounds__(kSmallThreads, 1)
syevj_small_kernel(const T* __restrict__ A, int64_t lda, int64_t strideA, int n, T* __restrict__ evals,
                   int64_t strideE, T* __restrict__ evt, int64_t ldv, int64_t strideV, int max_sweeps, T tol,
                   int* __restrict__ info) {
  // grid = (nsplit, batch): every CTA of a matrix runs the SAME rotations on its own copy of H (identical
  // arithmetic, hence identical bits) and accumulates the eigenvector columns [c0, c1) only -- the V update is
  // as large as the H update and splits without any communication.
  extern __shared__ __align__(16) unsigned char sv_smem[];
  const int N = (n + 1) & ~1;      // even number of players; index n (if any) is a bye
  const int LD = N + 1;
  const int m2 = N / 2;
  const int nsplit = gridDim.x;
  const int cw = (n + nsplit - 1) / nsplit;              // eigenvector columns of this CTA
  const int c0 = blockIdx.x * cw;
  const int c1 = min(n, c0 + cw);
  const int ncol = max(0, c1 - c0);
  const int LDV = cw + 1;
  T* H = reinterpret_cast<T*>(sv_smem);
  T* V = H + (size_t)N * LD;                             // [N][LDV]
  T* red = V + (size_t)N * LDV;                          // [32]
  Rot<T>* rot = reinterpret_cast<Rot<T>*>((reinterpret_cast<uintptr_t>(red + 32) + 31) & ~uintptr_t(31));   // [m2]
  int* rank = reinterpret_cast<int*>(rot + m2);          // [N]
  __shared__ int done;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const T* Ab = A + (size_t)blockIdx.y * strideA;

  // ---- load (symmetrised), V = I (own columns), ||H||_F^2 ----
  T fro_local = T(0);
  for (int e = tid; e < N * N; e += kSmallThreads) {
    const int r = e / N, c = e % N;
    T v = T(0);
    if (r < n && c < n) v = T(0.5) * (Ab[(size_t)r * lda + c] + Ab[(size_t)c * lda + r]);
    H[r * LD + c] = v;
    fro_local = fma(v, v, fro_local);
  }
  for (int e = tid; e < N * cw; e += kSmallThreads) {
    const int r = e / cw, c = e % cw;
    V[r * LDV + c] = (r == c0 + c) ? T(1) : T(0);
  }
  for (int o = 16; o > 0; o >>= 1) fro_local += __shfl_xor_sync(0xffffffffu, fro_local, o);
  if (lane == 0) red[warp] = fro_local;
  if (tid == 0) done = 0;
  __syncthreads();
  T fro2 = T(0);
  for (int w = 0; w < kSmallThreads / 32; ++w) fro2 += red[w];
  __syncthreads();

  // static work assignment: H blocks (i, j) = blk / m2, blk % m2 for blk = tid + u * threads; V items likewise
  constexpr int CH = sizeof(T) == 4 ? 2 : 1;   // blocks in flight per thread (64 registers at 1024 threads)
  int hi[kHB], hj[kHB], vi[kVB], vc[kVB];
#pragma unroll
  for (int u = 0; u < kHB; ++u) {
    const int blk = tid + u * kSmallThreads;
    hi[u] = blk < m2 * m2 ? blk / m2 : -1;
    hj[u] = blk < m2 * m2 ? blk % m2 : 0;
  }
#pragma unroll
  for (int u = 0; u < kVB; ++u) {
    const int e = tid + u * kSmallThreads;
    vi[u] = (ncol > 0 && e < m2 * ncol) ? e / ncol : -1;
    vc[u] = (ncol > 0 && e < m2 * ncol) ? e % ncol : 0;
  }

  int sweeps = 0;
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    T off_local = T(0);
    for (int t = 0; t < N - 1; ++t) {
      if (tid < m2) {
        int p, q;
        if (tid == 0) { p = N - 1; q = t; }
        else {
          p = t + tid; if (p >= N - 1) p -= N - 1;
          q = t - tid; if (q < 0) q += N - 1;
        }
        T s = T(0), tau = T(0);
        if (p < n && q < n) {
          const T hpp = H[p * LD + p], hqq = H[q * LD + q], hpq = H[p * LD + q];
          off_local = fma(T(2) * hpq, hpq, off_local);
          if (fabs(hpq) > SmallEps<T>::v * T(0.01) * sqrt(fabs(hpp * hqq)) && hpq != T(0)) {
            const T th = (hqq - hpp) / (T(2) * hpq);
            const T tt = (th >= T(0) ? T(1) : T(-1)) / (fabs(th) + sqrt(T(1) + th * th));
            const T c = T(1) / sqrt(T(1) + tt * tt);
            s = tt * c;
            tau = s / (T(1) + c);
          }
        }
        Rot<T> r;
        r.p = p; r.q = q; r.s = s; r.tau = tau;
        rot[tid] = r;
      }
      __syncthreads();
      // H <- J^T H J : a thread owns the SAME 2 x 2 blocks (rows {p_i, q_i} x columns {p_j, q_j}) in every step;
      // all operands are loaded before anything is stored so that the loads overlap (the blocks partition H: no
      // hazard).  Rutishauser's update  x' = x - s (y + tau x),  y' = y + s (x - tau y)  keeps the error
      // proportional to the rotation.
      if (sizeof(T) == 8) {
        // float64: 64 registers per thread do not hold the preloaded operands; walk the blocks warp by warp instead
        for (int i = warp; i < m2; i += kSmallThreads / 32) {
          const Rot<T> ri = rot[i];
          T* Hp = H + ri.p * LD;
          T* Hq = H + ri.q * LD;
          for (int j = lane; j < m2; j += 32) {
            const Rot<T> rj = rot[j];
            const T a = Hp[rj.p], b = Hp[rj.q], c_ = Hq[rj.p], d = Hq[rj.q];
            const T a1 = a - ri.s * (c_ + ri.tau * a), c1 = c_ + ri.s * (a - ri.tau * c_);
            const T b1 = b - ri.s * (d + ri.tau * b), d1 = d + ri.s * (b - ri.tau * d);
            T a2 = a1 - rj.s * (b1 + rj.tau * a1), b2 = b1 + rj.s * (a1 - rj.tau * b1);
            T c2 = c1 - rj.s * (d1 + rj.tau * c1), d2 = d1 + rj.s * (c1 - rj.tau * d1);
            if (i == j && ri.s != T(0)) { b2 = T(0); c2 = T(0); }
            Hp[rj.p] = a2; Hp[rj.q] = b2; Hq[rj.p] = c2; Hq[rj.q] = d2;
          }
        }
        for (int e = tid; e < m2 * ncol; e += kSmallThreads) {
          const int i = e / ncol, col = e - i * ncol;
          const Rot<T> ri = rot[i];
          const T vp = V[ri.p * LDV + col], vq = V[ri.q * LDV + col];
          V[ri.p * LDV + col] = vp - ri.s * (vq + ri.tau * vp);
          V[ri.q * LDV + col] = vq + ri.s * (vp - ri.tau * vq);
        }
      } else {
#pragma unroll
      for (int u0 = 0; u0 < kHB; u0 += CH) {
        Rot<T> ri[CH], rj[CH];
        T a[CH], b[CH], c_[CH], d[CH];
#pragma unroll
        for (int u = 0; u < CH; ++u) {
          if (hi[u0 + u] >= 0) {
            ri[u] = rot[hi[u0 + u]];
            rj[u] = rot[hj[u0 + u]];
          }
        }
#pragma unroll
        for (int u = 0; u < CH; ++u) {
          if (hi[u0 + u] >= 0) {
            a[u] = H[ri[u].p * LD + rj[u].p]; b[u] = H[ri[u].p * LD + rj[u].q];
            c_[u] = H[ri[u].q * LD + rj[u].p]; d[u] = H[ri[u].q * LD + rj[u].q];
          }
        }
#pragma unroll
        for (int u = 0; u < CH; ++u) {
          if (hi[u0 + u] >= 0) {
            const T si = ri[u].s, ti = ri[u].tau, sj = rj[u].s, tj = rj[u].tau;
            const T a1 = a[u] - si * (c_[u] + ti * a[u]), c1 = c_[u] + si * (a[u] - ti * c_[u]);
            const T b1 = b[u] - si * (d[u] + ti * b[u]), d1 = d[u] + si * (b[u] - ti * d[u]);
            T a2 = a1 - sj * (b1 + tj * a1), b2 = b1 + sj * (a1 - tj * b1);
            T c2 = c1 - sj * (d1 + tj * c1), d2 = d1 + sj * (c1 - tj * d1);
            if (hi[u0 + u] == hj[u0 + u] && si != T(0)) { b2 = T(0); c2 = T(0); }   // the annihilated pivot, exactly
            H[ri[u].p * LD + rj[u].p] = a2; H[ri[u].p * LD + rj[u].q] = b2;
            H[ri[u].q * LD + rj[u].p] = c2; H[ri[u].q * LD + rj[u].q] = d2;
          }
        }
      }
      // eigenvector rows p_i, q_i (own columns), same static assignment
      {
        Rot<T> rv[kVB];
        T vp[kVB], vq[kVB];
#pragma unroll
        for (int u = 0; u < kVB; ++u)
          if (vi[u] >= 0) rv[u] = rot[vi[u]];
#pragma unroll
        for (int u = 0; u < kVB; ++u)
          if (vi[u] >= 0) { vp[u] = V[rv[u].p * LDV + vc[u]]; vq[u] = V[rv[u].q * LDV + vc[u]]; }
#pragma unroll
        for (int u = 0; u < kVB; ++u)
          if (vi[u] >= 0) {
            V[rv[u].p * LDV + vc[u]] = vp[u] - rv[u].s * (vq[u] + rv[u].tau * vp[u]);
            V[rv[u].q * LDV + vc[u]] = vq[u] + rv[u].s * (vp[u] - rv[u].tau * vq[u]);
          }
      }
      }
      __syncthreads();
    }
    ++sweeps;
    // off-diagonal mass met during this sweep (threads >= m2 contribute 0)
    for (int o = 16; o > 0; o >>= 1) off_local += __shfl_xor_sync(0xffffffffu, off_local, o);
    if (lane == 0) red[warp] = off_local;
    __syncthreads();
    if (tid == 0) {
      T off2 = T(0);
      for (int w = 0; w < kSmallThreads / 32; ++w) off2 += red[w];
      done = !(off2 > tol * tol * fro2);
    }
    __syncthreads();
    if (done) break;
  }

  // ---- sort (rank by counting, descending; ties by index) and write ----
  if (tid < n) {
    const T li = H[tid * LD + tid];
    int rk = 0;
    for (int j = 0; j < n; ++j) {
      const T lj = H[j * LD + j];
      rk += (lj > li || (lj == li && j < tid)) ? 1 : 0;
    }
    rank[tid] = rk;
    if (evals && blockIdx.x == 0) evals[(size_t)blockIdx.y * strideE + rk] = li;
  }
  __syncthreads();
  if (evt) {
    T* Eb = evt + (size_t)blockIdx.y * strideV;
    for (int e = tid; e < n * ncol; e += kSmallThreads) {
      const int r = e / ncol, c = e - r * ncol;
      Eb[(size_t)rank[r] * ldv + c0 + c] = V[r * LDV + c];
    }
  }
  if (tid == 0 && info && blockIdx.x == 0) info[blockIdx.y] = done ? sweeps : -sweeps;
}

inline int small_nsplit(int n) { return n >= 64 ? 4 : (n >= 24 ? 2 : 1); }

template <typename T>
size_t small_smem_bytes(int n) {
  const int N = (n + 1) & ~1;
  const int cw = (n + small_nsplit(n) - 1) / small_nsplit(n);
  return sizeof(T) * ((size_t)N * (N + 1) + (size_t)N * (cw + 1) + 32) + sizeof(Rot<T>) * (size_t)(N / 2) +
         sizeof(int) * (size_t)N + 128;
}

}  // namespace

template <typename T>
bool syevj_small_supported(int n) {
  return n >= 1 && n <= kSmallMaxN && small_smem_bytes<T>(n) <= 220 * 1024;
}

template <typename T>
int syevj_small(int n, int batch, const T* A, int64_t lda, int64_t strideA, T* evals, int64_t strideE, T* evt,
                int64_t ldv, int64_t strideV, int* info_dev, cudaStream_t stream) {
  CCAB_CHECK_ARG(syevj_small_supported<T>(n), "syevj_small: n = %d does not fit one CTA's shared memory", n);
  CCAB_CHECK_ARG(batch >= 1 && lda >= n && (!evt || ldv >= n), "bad syevj_small shape");
  const size_t smem = small_smem_bytes<T>(n);
  CCAB_CUDA(cudaFuncSetAttribute(syevj_small_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int max_sweeps = sizeof(T) == 4 ? 14 : 18;
  const T tol = T(4) * SmallEps<T>::v * (T)std::sqrt((double)n);
  syevj_small_kernel<T><<<dim3(small_nsplit(n), batch), kSmallThreads, smem, stream>>>(A, lda, strideA, n, evals, strideE, evt, ldv, strideV,
                                                               max_sweeps, tol, info_dev);
  count_launches(1);
  CCAB_CUDA(cudaGetLastError());
  return 0;
}

template bool syevj_small_supported<float>(int);
template bool syevj_small_supported<double>(int);
template int syevj_small<float>(int, int, const float*, int64_t, int64_t, float*, int64_t, float*, int64_t, int64_t,
                                int*, cudaStream_t);
template int syevj_small<double>(int, int, const double*, int64_t, int64_t, double*, int64_t, double*, int64_t, int64_t,
                                 int*, cudaStream_t);

}  // namespace ccab
