// tgemm: C = alpha * op(A) * op(B) + beta * C in fp32-grade arithmetic on the tcgen05 tensor cores.
//
// The solver stage after K1 (Cholesky panels and trailing updates, L^-1 assembly, T = L1^-1 C12 L2^-T, the
// subspace iteration, the back-substitution of the weights, the tall products of the deep-CCA backward) is a
// chain of small and medium GEMMs.  Round 1 ran them as 64x64 FMA tiles on the CUDA cores; this kernel puts
// them on the tensor pipe with the machinery K1 established:
//
//   * operands straight from row-major storage by TMA, both majors: an operand whose reduction index is
//     contiguous in memory lands K-major (32-float = 128-byte rows, SWIZZLE_128B), one whose reduction index is
//     the strided one lands MN-major (32-column atoms, SWIZZLE_128B with 32-byte chunks -- the only legal
//     MN-major TF32 layout, see moments.cu); the instruction descriptor carries one major bit per operand, so
//     all four op() combinations of ccab_gemm map to the same mainloop without any transposed copy;
//   * 3xTF32 with the split formed IN SHARED MEMORY: the tensor core truncates fp32 operands to TF32, so the raw
//     tile is its own "hi" part; four converter warps derive lo = rna_tf32(x - trunc(x)) from the landed tile
//     into a second buffer of the stage (generic-proxy stores, fence.proxy.async, mbarrier) while the MMA warp
//     works on the previous stage -- no pre-pass over global memory, no second TMA stream;
//   * fp32 accumulator in TMEM (128 lanes x BN columns), warp-specialised: warp 0 TMA producer, warp 1
//     single-thread MMA issuer, warps 2-5 converters and, after the mainloop, the TMEM -> register -> global
//     epilogue (alpha / beta, optional transposed copy, optional lower-triangle-only tiles);
//   * batched through the third TMA coordinate (blockIdx.z).
//
// Out-of-range rows / columns / reduction indices are zero-filled by TMA, so no shape needs padding.
// Replaces the FMA products behind cca_zoo/linear/_rcca.py:96,100, cca_zoo/linear/_mcca.py:131 and
// cca_zoo/deep/objectives.py:97 (and the LAPACK level-3 calls inside scipy.linalg.eigh / numpy.linalg.svd).
#include "tgemm.cuh"

#include <mutex>

namespace ccab {
namespace {

constexpr int kTgThreads = 192;
constexpr int kBM = 128;
constexpr int kKC = 32;       // reduction indices per stage = one 128-byte swizzle row of a K-major tile
constexpr int kAtom = kKC * 128;  // bytes of one 32-column MN-major atom

struct alignas(64) TgParams {
  CUtensorMap mapA, mapB;
  float* C;
  long long ldc, strideC, strideC2;
  float* Ct;
  long long ldct, strideCt, strideCt2;
  int batch1;  // blockIdx.z = b2 * batch1 + b1
  int M, N, K;
  float alpha, beta;
  int lower_only;
  int vec_c;  // rows of C can be read / written as float4
};

template <int BN>
struct TgCfg {
  static constexpr int kABytes = kBM * kKC * 4;
  static constexpr int kBBytes = BN * kKC * 4;
  static constexpr int kRaw = kABytes + kBBytes;
  static constexpr int kStage = 2 * kRaw;  // raw (= hi) tiles, then the lo tiles at the same offsets
  static constexpr int kStages = BN == 128 ? 3 : 4;
  // tcgen05 accumulates with round-toward-zero: one accumulator drifts by ~0.5 ulp per MMA (1e-5 relative after
  // 1024 reduction indices, measured).  The reduction is therefore dealt round-robin, one 32-index chunk at a time,
  // over kSegs independent TMEM accumulators that the epilogue adds in registers (round-to-nearest).
  static constexpr int kSegs = 512 / BN;
  static constexpr int kSmem = kStages * kStage + 1024 + 256;
};

template <bool AK, bool BK, int BN>
__global__ void __launch_bounds__(kTgThreads, 1) tgemm_kernel(const __grid_constant__ TgParams p) {
  using Cfg = TgCfg<BN>;
  constexpr int NS = Cfg::kStages;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * BN;
  const int bz = (int)blockIdx.z % p.batch1, bz2 = (int)blockIdx.z / p.batch1;
  if (p.lower_only && n0 >= m0 + kBM) return;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NS * Cfg::kStage);
  uint64_t* conv_bar = full_bar + NS;
  uint64_t* empty_bar = conv_bar + NS;
  uint64_t* tmem_full_bar = empty_bar + NS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nchunks = (p.K + kKC - 1) / kKC;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&conv_bar[s], 128);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.mapA);
    tma_prefetch_desc(&p.mapB);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::kRaw);
        uint8_t* sA = smem + stage * Cfg::kStage;
        uint8_t* sB = sA + Cfg::kABytes;
        const int k0 = c * kKC;
        if (AK) {
          tma_load_4d(sA, &p.mapA, &full_bar[stage], k0, m0, bz, bz2);
        } else {
#pragma unroll
          for (int a = 0; a < kBM / 32; ++a) tma_load_4d(sA + a * kAtom, &p.mapA, &full_bar[stage], m0 + 32 * a, k0, bz, bz2);
        }
        if (BK) {
          tma_load_4d(sB, &p.mapB, &full_bar[stage], k0, n0, bz, bz2);
        } else {
#pragma unroll
          for (int a = 0; a < BN / 32; ++a) tma_load_4d(sB + a * kAtom, &p.mapB, &full_bar[stage], n0 + 32 * a, k0, bz, bz2);
        }
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = umma_idesc_tf32(kBM, BN, AK, BK);
    // K-major: 8-row groups 1024 B apart (SBO), 128-byte swizzle (layout 2); a k-step of 8 floats = +32 B.
    // MN-major: 32-column atoms kAtom apart (LBO), 4-row groups 512 B apart (SBO), layout 1; a k-step = +1024 B.
    const uint64_t dA0 = AK ? umma_smem_desc(smem_u32(smem), 16, 1024, 2) : umma_smem_desc(smem_u32(smem), kAtom, 512, 1);
    const uint64_t dB0 = BK ? umma_smem_desc(smem_u32(smem) + Cfg::kABytes, 16, 1024, 2)
                            : umma_smem_desc(smem_u32(smem) + Cfg::kABytes, kAtom, 512, 1);
    constexpr uint64_t stepA = AK ? 2 : 64, stepB = BK ? 2 : 64;
    constexpr uint64_t lo_off = (uint64_t)(Cfg::kRaw >> 4);
    int stage = 0;
    uint32_t phase = 0;
    for (int c = 0; c < nchunks; ++c) {
      mbar_wait(&conv_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t so = (uint64_t)((stage * Cfg::kStage) >> 4);
        const uint32_t tacc = tmem_base + (uint32_t)((c % Cfg::kSegs) * BN);
        uint32_t acc = c >= Cfg::kSegs ? 1u : 0u;
#pragma unroll
        for (int kk = 0; kk < kKC / 8; ++kk) {
          const uint64_t a_hi = dA0 + so + kk * stepA, b_hi = dB0 + so + kk * stepB;
          umma_tf32(tacc, a_hi + lo_off, b_hi, idesc, acc);   // small cross terms first
          umma_tf32(tacc, a_hi, b_hi + lo_off, idesc, 1u);
          umma_tf32(tacc, a_hi, b_hi, idesc, 1u);
          acc = 1u;
        }
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == NS) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(tmem_full_bar);
    __syncwarp();
  } else {
    // ================= converters: lo = rna_tf32(x - trunc_tf32(x)) of the landed stage =================
    const int tid = threadIdx.x - 64;
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&full_bar[stage], phase);
        const float4* src = reinterpret_cast<const float4*>(smem + stage * Cfg::kStage);
        float4* dst = reinterpret_cast<float4*>(smem + stage * Cfg::kStage + Cfg::kRaw);
#pragma unroll 4
        for (int i = tid; i < Cfg::kRaw / 16; i += 128) {
          const float4 v = src[i];
          dst[i] = make_float4(tf32_residual(v.x), tf32_residual(v.y), tf32_residual(v.z), tf32_residual(v.w));
        }
        fence_proxy_async_smem();
        mbar_arrive(&conv_bar[stage]);
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
    }
    // ================= epilogue =================
    const int g = warp & 3;
    const int row = m0 + g * 32 + lane;
    const int nsegs = nchunks < Cfg::kSegs ? nchunks : Cfg::kSegs;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    float* Cb = p.C ? p.C + (size_t)bz * p.strideC + (size_t)bz2 * p.strideC2 : nullptr;
    float* Ctb = p.Ct ? p.Ct + (size_t)bz * p.strideCt + (size_t)bz2 * p.strideCt2 : nullptr;
#pragma unroll 1
    for (int cc = 0; cc < BN / 32; ++cc) {
      const int col0 = n0 + cc * 32;
      if (col0 >= p.N) break;   // warp-uniform
      float v[32];
      {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(g * 32) << 16) + cc * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
      }
      for (int sgm = 1; sgm < nsegs; ++sgm) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(g * 32) << 16) + sgm * BN + cc * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += __uint_as_float(r[i]);
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] *= p.alpha;
      if (Cb && row < p.M) {
        float* crow = Cb + (size_t)row * p.ldc + col0;
        if (p.vec_c && col0 + 32 <= p.N) {
          float4* c4 = reinterpret_cast<float4*>(crow);
          if (p.beta != 0.f) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 o = c4[q];
              v[4 * q] += p.beta * o.x; v[4 * q + 1] += p.beta * o.y; v[4 * q + 2] += p.beta * o.z; v[4 * q + 3] += p.beta * o.w;
            }
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) c4[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (col0 + i < p.N) {
              if (p.beta != 0.f) v[i] += p.beta * crow[i];
              crow[i] = v[i];
            }
          }
        }
      } else if (!Cb && p.beta != 0.f) {
        // unreachable: the host rejects beta != 0 without C
      }
      if (Ctb && row < p.M) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (col0 + i < p.N) Ctb[(size_t)(col0 + i) * p.ldct + row] = v[i];   // lanes = consecutive rows: coalesced
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

// K-major operand: stored (rows x K) row-major, box = 32 reduction indices x `box_rows` rows.
// MN-major operand: stored (K x cols) row-major, box = 32 columns x 32 reduction indices.
int encode_operand(CUtensorMap* map, const float* ptr, bool kmajor, int64_t mn, int64_t k, int64_t ld, int64_t stride,
                   int batch, int64_t stride2, int batch2, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available (driver too old?)");
    return -2;
  }
  const int64_t outer = kmajor ? mn : k;
  if (batch == 1) stride = ld * outer;     // unused dimension: any valid (16-byte multiple) stride
  if (batch2 == 1) stride2 = ld * outer;
  cuuint64_t gdim[4] = {(cuuint64_t)(kmajor ? k : mn), (cuuint64_t)outer, (cuuint64_t)batch, (cuuint64_t)batch2};
  cuuint64_t gstride[3] = {(cuuint64_t)ld * 4, (cuuint64_t)stride * 4, (cuuint64_t)stride2 * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)(kmajor ? box_rows : kKC), 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   kmajor ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (mn=%lld k=%lld ld=%lld stride=%lld batch=%d)", (int)r,
              (long long)mn, (long long)k, (long long)ld, (long long)stride, batch);
    return -3;
  }
  return 0;
}

template <bool AK, bool BK, int BN>
int launch(const TgParams& prm, dim3 grid, cudaStream_t stream) {
  using Cfg = TgCfg<BN>;
  static bool attr_done[64] = {};   // cudaFuncSetAttribute is per device
  int dev = 0;
  CCAB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    CCAB_CUDA(cudaFuncSetAttribute(tgemm_kernel<AK, BK, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  tgemm_kernel<AK, BK, BN><<<grid, kTgThreads, Cfg::kSmem, stream>>>(prm);
  count_launches(1);
  CCAB_CUDA(cudaGetLastError());
  return 0;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

bool tgemm_supported(const TgemmArgs& a) {
  if (a.m < 1 || a.n < 1 || a.k < 1 || a.batch < 1) return false;
  if (!a.A || !a.B || !aligned16(a.A) || !aligned16(a.B)) return false;
  if (a.lda % 4 || a.ldb % 4) return false;
  if (a.batch > 1 && (a.strideA % 4 || a.strideB % 4 || a.strideA <= 0 || a.strideB <= 0)) return false;
  if (a.batch2 < 1 || (a.batch2 > 1 && (a.strideA2 % 4 || a.strideB2 % 4 || a.strideA2 <= 0 || a.strideB2 <= 0))) return false;
  return true;
}

int tgemm(const TgemmArgs& a, cudaStream_t stream) {
  {
    // cuTensorMapEncodeTiled is a driver entry point and needs a current context on the CALLING thread; a thread that
    // has only inherited the device ordinal (torch's autograd worker) has none until a runtime call binds the primary
    // context -- cudaSetDevice does (CUDA 12)
    int dev = 0;
    CCAB_CUDA(cudaGetDevice(&dev));
    CCAB_CUDA(cudaSetDevice(dev));
  }
  CCAB_CHECK_ARG(tgemm_supported(a), "tgemm: operands must be 16-byte aligned with leading dimensions % 4 == 0");
  CCAB_CHECK_ARG(a.C || a.Ct, "tgemm: no output");
  CCAB_CHECK_ARG(a.C || a.beta == 0.f, "tgemm: beta != 0 needs C");
  const bool AK = !a.transa;  // op(A) = A (m x k row-major): reduction index contiguous
  const bool BK = a.transb != 0;  // op(B) = B^T with B stored n x k: reduction index contiguous
  CCAB_CHECK_ARG(a.lda >= (AK ? a.k : a.m) && a.ldb >= (BK ? a.k : a.n), "tgemm: leading dimension too small");
  const int64_t tiles128 = ceil_div(a.m, kBM) * ceil_div(a.n, 128) * a.batch * a.batch2;
  int BN = (a.n > 64 && tiles128 >= 96) ? 128 : 64;
  if (a.force_bn == 64 || a.force_bn == 128) BN = a.force_bn;
  TgParams prm;
  memset(&prm, 0, sizeof(prm));
  int rc = encode_operand(&prm.mapA, a.A, AK, a.m, a.k, a.lda, a.strideA, a.batch, a.strideA2, a.batch2, kBM);
  if (rc) return rc;
  rc = encode_operand(&prm.mapB, a.B, BK, a.n, a.k, a.ldb, a.strideB, a.batch, a.strideB2, a.batch2, BN);
  if (rc) return rc;
  prm.C = a.C;
  prm.ldc = a.ldc;
  prm.strideC = a.strideC;
  prm.strideC2 = a.strideC2;
  prm.Ct = a.Ct;
  prm.ldct = a.ldct;
  prm.strideCt = a.strideCt;
  prm.strideCt2 = a.strideCt2;
  prm.batch1 = a.batch;
  prm.M = a.m;
  prm.N = a.n;
  prm.K = a.k;
  prm.alpha = a.alpha;
  prm.beta = a.beta;
  prm.lower_only = a.lower_only;
  prm.vec_c = a.C && aligned16(a.C) && a.ldc % 4 == 0 && a.strideC % 4 == 0 && a.strideC2 % 4 == 0;
  dim3 grid((unsigned)ceil_div(a.n, BN), (unsigned)ceil_div(a.m, kBM), (unsigned)(a.batch * a.batch2));
#define CCAB_TG(AK_, BK_)                                                       \
  (BN == 128 ? launch<AK_, BK_, 128>(prm, grid, stream) : launch<AK_, BK_, 64>(prm, grid, stream))
  if (AK && BK) return CCAB_TG(true, true);
  if (AK && !BK) return CCAB_TG(true, false);
  if (!AK && BK) return CCAB_TG(false, true);
  return CCAB_TG(false, false);
#undef CCAB_TG
}

}  // namespace ccab
