// Block-moment stage (kernels K1/K2 of DESIGN.md): M = [X1..Xm]^T [X1..Xm], s = 1^T X.
#pragma once
#include "common.cuh"

namespace ccab {

constexpr int kMaxViews = 8;
constexpr int kBlk = 128;        // column-block width of the padded tile space
constexpr int kMaxBlocks = 128;  // D_padded <= 16384

// Column layout shared by every kernel of the moment stage.  Each view's columns are padded to a
// multiple of 128 in *tile space*; "compact" is the hstacked view order the reference uses
// (cca_zoo/linear/_mcca.py:150: np.hstack(views)).
struct ColumnLayout {
  int n_views;
  int nblocks;               // total 128-column blocks
  int D;                     // compact width  = sum(dims)
  int Dp;                    // padded width   = nblocks * 128
  int dims[kMaxViews];       // view widths
  int coff[kMaxViews + 1];   // compact offset of each view
  int poff[kMaxViews + 1];   // padded  offset of each view
};

int make_layout(int n_views, const int64_t* dims, ColumnLayout* out);

// --- launchers (all asynchronous on `stream`) -------------------------------------------------
// precision: 0 = TF32 single pass (tcgen05), 1 = 3xTF32 split (tcgen05), 2 = exact SIMT FMA,
//            3 = 3xTF32 with the two cross terms as bf16 MMAs (tcgen05 kind::f16).
size_t moments_workspace_bytes(int dtype, int precision, const ColumnLayout& L, int64_t n_rows);

int moments_tf32(const ColumnLayout& L, const void* const* views, const int64_t* lds, int64_t n_rows,
                 int mode /* 0, 1 or 3 */, double* moments_out, void* ws, size_t ws_bytes, cudaStream_t stream);

template <typename T>
int moments_simt(const ColumnLayout& L, const void* const* views, const int64_t* lds, int64_t n_rows,
                 double* moments_out, void* ws, size_t ws_bytes, cudaStream_t stream);

// moments (double[Dp*Dp + Dp], padded, upper block triangle) -> compact covariance + means
template <typename Tout>
int covariance_from_moments(const ColumnLayout& L, const double* moments, double n_total, int center,
                            Tout* C, int64_t ldc, Tout* mean, cudaStream_t stream);

// exchange-step message: upper triangle of 128 x 128 blocks | column sums | n | reserved (all float64)
int64_t moments_packed_size(const ColumnLayout& L);
int moments_pack(const ColumnLayout& L, const double* moments, double n_local, double* packed, cudaStream_t stream);
int moments_unpack(const ColumnLayout& L, const double* packed, double* moments, cudaStream_t stream);

// Shifted accumulation: pilot statistics of the leading rows, Xs = X - x0, and the float64 correction that turns the
// moments of the shifted data back into raw moments (see moments.cu).
template <typename T>
int column_pilot(const T* X, int64_t rows, int d, int64_t ld, T* x0, float* ratio_max, cudaStream_t stream);
template <typename T>
int shift_rows(const T* X, int64_t n, int d, int64_t ldx, const T* x0, T* Xs, int64_t lds, cudaStream_t stream);
int moments_unshift(const ColumnLayout& L, double* moments, const void* const* x0, int is_f64, double n,
                    cudaStream_t stream);

// Fused exchange step on symmetric memory (pack + in-switch reduction with multimem + unpack in ONE kernel).
// sym_local / sym_multicast: local and multicast address of a symmetric buffer of sym_doubles doubles
// (>= world * ceil(packed / world)); pads_dev: device array of the world signal-pad pointers (uint32, pad_slots
// entries each, zero-initialised once); epoch: 2 x the call counter, identical on all ranks, strictly increasing.
int moments_exchange_nvls(const ColumnLayout& L, double* moments, double n_local, double* sym_local,
                          double* sym_multicast, void* const* pads_dev, int rank, int world, int pad_slots,
                          int64_t sym_doubles, unsigned epoch, double* n_total_out, cudaStream_t stream);

// debug knobs for the tcgen05 kernel (descriptor strides / TMA data type), see tools/umma_probe.py
struct TcDebug {
  int lbo_bytes;    // <0: default
  int sbo_bytes;    // <0: default
  int tma_dtype;    // <0: default (FLOAT32); else a CUtensorMapDataType value
  int force_splits; // <=0: heuristic
  int variant;      // 0: CTA-pair kernel (cta_group::2, default), 1: single-CTA kernel
  int kc;           // CTA-pair kernel, 1-pass: rows per stage 16 / 32 (default) / 64
  int dry_run;      // CTA-pair kernel: skip TMA after the first ring fill (MMA-rate experiment; wrong results)
  int x3_split;     // 3xTF32 operand split: 0 = residual only (raw array is hi by truncation), 1 = round-to-nearest hi/lo
  int f64_simt;     // float64 inputs: 0 = DMMA kernel (default), 1 = CUDA-core FMA kernel
  int x3b_oneshot;  // tf32x3b: 0 = persistent kernel, overlapped epilogue (default), 1 = one (tile, split) unit per CTA pair
};
TcDebug& tc_debug();

// CUDA-event timing of the tcgen05 kernel launch alone (for bench.py's roofline line)
void moments_profile_enable(int on);
float moments_profile_last_ms();

}  // namespace ccab
