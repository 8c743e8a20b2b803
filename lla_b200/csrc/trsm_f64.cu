// lla_b200/csrc/trsm_f64.cu -- left-sided triangular solves op(T) X = B (in place) on the device.
//
// Used by GETRS / POTRS (reference call sites src/linear-algebra.lisp:269-276, :296-301) and by
// the panel steps of the blocked LU / Cholesky factorisations.  The solve is recursive: the
// triangle is split in two, the off-diagonal coupling is a DMMA GEMM (gemm_f64.cu), and only
// 32 x 32 diagonal blocks are solved by substitution (column-oriented, the order of netlib xTRSM)
// with one thread per right-hand side holding its 32 unknowns in registers.
#include "internal.h"
#include "tri128.cuh"
#include "../../include/llacuda.h"

namespace llacuda {

constexpr int TL = 32;        // leaf triangle size
constexpr int TL_COLS = 64;   // right-hand sides per CTA

// E = op(T) restricted to the leaf, identity padded, E[i][k] at e[i * 33 + k].
// EFF_LOWER: E is lower triangular (forward substitution), else upper (backward).
template <bool EFF_LOWER, bool TRANS, bool UNIT>
__global__ void __launch_bounds__(TL_COLS) lla_dtrsm_leaf_kernel(int m, int64_t n, const double* __restrict__ T,
                                                             int64_t ldt, double* __restrict__ B, int64_t ldb,
                                                             const int* abort_flag) {
  if (abort_flag != nullptr && *abort_flag != 0) return;
  __shared__ double e[TL * (TL + 1)];
  __shared__ double bs[TL_COLS * (TL + 1)];
  const int tid = threadIdx.x;
  const int64_t c0 = (int64_t)blockIdx.x * TL_COLS;

#pragma unroll
  for (int idx = tid; idx < TL * TL; idx += TL_COLS) {
    int i = idx % TL, k = idx / TL;  // consecutive threads walk down a column of T
    double v = (i == k) ? 1.0 : 0.0;
    if (i < m && k < m && !(UNIT && i == k)) v = T[i + (int64_t)k * ldt];
    if (TRANS) e[k * (TL + 1) + i] = v;  // E[k][i] = T[i][k]
    else e[i * (TL + 1) + k] = v;
  }
  const int lane = tid & 31, w = tid >> 5;
  {
    double v[TL_COLS / 2];  // all loads of the slab in flight before the first store
#pragma unroll
    for (int q = 0; q < TL_COLS / 2; q++) {
      int64_t gc = c0 + w + 2 * q;
      v[q] = (gc < n && lane < m) ? B[lane + gc * ldb] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < TL_COLS / 2; q++) bs[(w + 2 * q) * (TL + 1) + lane] = v[q];
  }
  __syncthreads();

  double x[TL];
#pragma unroll
  for (int i = 0; i < TL; i++) x[i] = bs[tid * (TL + 1) + i];

  if (EFF_LOWER) {
#pragma unroll
    for (int k = 0; k < TL; k++) {
      if (!UNIT) x[k] = x[k] / e[k * (TL + 1) + k];
#pragma unroll
      for (int i = k + 1; i < TL; i++) x[i] = fma(-e[i * (TL + 1) + k], x[k], x[i]);
    }
  } else {
#pragma unroll
    for (int k = TL - 1; k >= 0; k--) {
      if (!UNIT) x[k] = x[k] / e[k * (TL + 1) + k];
#pragma unroll
      for (int i = 0; i < k; i++) x[i] = fma(-e[i * (TL + 1) + k], x[k], x[i]);
    }
  }

#pragma unroll
  for (int i = 0; i < TL; i++) bs[tid * (TL + 1) + i] = x[i];
  __syncthreads();
#pragma unroll
  for (int c = w; c < TL_COLS; c += TL_COLS / 32) {
    int64_t gc = c0 + c;
    if (gc < n && lane < m) B[lane + gc * ldb] = bs[c * (TL + 1) + lane];
  }
}

static int launch_leaf(bool eff_lower, bool trans, bool unit, int m, int64_t n, const double* T, int64_t ldt,
                       double* B, int64_t ldb, cudaStream_t st, const int* abort_flag) {
  unsigned grid = (unsigned)((n + TL_COLS - 1) / TL_COLS);
#define LLA_LEAF(L, TR, U) lla_dtrsm_leaf_kernel<L, TR, U><<<grid, TL_COLS, 0, st>>>(m, n, T, ldt, B, ldb, abort_flag)
  if (eff_lower) {
    if (trans) { if (unit) LLA_LEAF(true, true, true); else LLA_LEAF(true, true, false); }
    else { if (unit) LLA_LEAF(true, false, true); else LLA_LEAF(true, false, false); }
  } else {
    if (trans) { if (unit) LLA_LEAF(false, true, true); else LLA_LEAF(false, true, false); }
    else { if (unit) LLA_LEAF(false, false, true); else LLA_LEAF(false, false, false); }
  }
#undef LLA_LEAF
  count_launch();
  LLA_CUDA(cudaGetLastError());
  return 0;
}

int dtrsm_left_dev(bool upper, int op, bool unit, int64_t m, int64_t n, const double* T, int64_t ldt, double* B,
                   int64_t ldb, cudaStream_t st, const int* abort_flag) {
  if (m <= 0 || n <= 0) return 0;
  const bool trans = op != 0;
  const bool eff_lower = (upper == trans);
  if (m <= TL) return launch_leaf(eff_lower, trans, unit, (int)m, n, T, ldt, B, ldb, st, abort_flag);
  int64_t m1 = ((m / 2 + TL - 1) / TL) * TL;
  int64_t m2 = m - m1;
  const double* T11 = T;
  const double* T22 = T + m1 + m1 * ldt;
  const double* T21 = T + m1;        // rows m1.., cols 0..m1
  const double* T12 = T + m1 * ldt;  // rows 0..m1, cols m1..
  double* B1 = B;
  double* B2 = B + m1;
  if (eff_lower) {
    LLA_TRY(dtrsm_left_dev(upper, op, unit, m1, n, T11, ldt, B1, ldb, st, abort_flag));
    if (!upper) LLA_TRY(dgemm_dev(0, 0, m2, n, m1, -1.0, T21, ldt, B1, ldb, 1.0, B2, ldb, TRI_FULL, st, abort_flag));
    else LLA_TRY(dgemm_dev(1, 0, m2, n, m1, -1.0, T12, ldt, B1, ldb, 1.0, B2, ldb, TRI_FULL, st, abort_flag));
    LLA_TRY(dtrsm_left_dev(upper, op, unit, m2, n, T22, ldt, B2, ldb, st, abort_flag));
  } else {
    LLA_TRY(dtrsm_left_dev(upper, op, unit, m2, n, T22, ldt, B2, ldb, st, abort_flag));
    if (upper) LLA_TRY(dgemm_dev(0, 0, m1, n, m2, -1.0, T12, ldt, B2, ldb, 1.0, B1, ldb, TRI_FULL, st, abort_flag));
    else LLA_TRY(dgemm_dev(1, 0, m1, n, m2, -1.0, T21, ldt, B2, ldb, 1.0, B1, ldb, TRI_FULL, st, abort_flag));
    LLA_TRY(dtrsm_left_dev(upper, op, unit, m1, n, T11, ldt, B1, ldb, st, abort_flag));
  }
  return 0;
}

// ------------------------------------------------------------------ inverse-based solves
// One CTA per IB x IB diagonal block: Tinv_b = inv(T_bb) (same triangle, other triangle zero).
__global__ void __launch_bounds__(T128_THREADS) lla_dtrtri_blocks_kernel(int64_t m, const double* __restrict__ T,
                                                                         int64_t ldt, double* __restrict__ Tinv,
                                                                         bool upper, bool unit, const int* abort_flag) {
  if (abort_flag != nullptr && *abort_flag != 0) return;
  extern __shared__ __align__(16) double sm128[];
  double* S = sm128;
  double* tmp = sm128 + T128 * T128_SP;
  const int64_t b = blockIdx.x;
  const int64_t rem = m - b * IB;
  const int r = (int)(rem < IB ? rem : IB);
  const double* Tb = T + b * IB + b * IB * ldt;
  tri128_load_lower(S, Tb, ldt, r, upper, unit, threadIdx.x);
  __syncthreads();
  trtri128_lower_smem(S, tmp, unit, threadIdx.x);
  tri128_store(S, Tinv + b * IB * IB, IB, IB, upper, true, threadIdx.x);
}

int dtrtri_blocks_dev(bool upper, bool unit, int64_t m, const double* T, int64_t ldt, double* Tinv, cudaStream_t st,
                      const int* abort_flag) {
  if (m <= 0) return 0;
  static DeviceOnce once;
  if (once.first()) {
    LLA_CUDA(cudaFuncSetAttribute(lla_dtrtri_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T128_SMEM));
  }
  unsigned grid = (unsigned)((m + IB - 1) / IB);
  lla_dtrtri_blocks_kernel<<<grid, T128_THREADS, T128_SMEM, st>>>(m, T, ldt, Tinv, upper, unit, abort_flag);
  count_launch();
  LLA_CUDA(cudaGetLastError());
  return 0;
}

// op(T) X = B with the diagonal-block inverses at hand: recursive halving on multiples of IB, the
// leaves are in-place products B_i <- op(Tinv_i) B_i on the DMMA GEMM (one 128-row tile per column
// slab, so the aliasing of B and C is safe), the couplings are ordinary GEMMs.
int dtrsm_left_inv_dev(bool upper, int op, int64_t m, int64_t n, const double* T, int64_t ldt, const double* Tinv,
                       double* B, int64_t ldb, cudaStream_t st, const int* abort_flag) {
  if (m <= 0 || n <= 0) return 0;
  const bool trans = op != 0;
  const bool eff_lower = (upper == trans);
  if (m <= IB)
    return dgemm_dev(trans ? 1 : 0, 0, m, n, m, 1.0, Tinv, IB, B, ldb, 0.0, B, ldb, TRI_FULL, st, abort_flag,
                     GEMM_INPLACE_B);
  int64_t m1 = ((m / 2 + IB - 1) / IB) * IB;
  int64_t m2 = m - m1;
  const double* T22 = T + m1 + m1 * ldt;
  const double* T21 = T + m1;
  const double* T12 = T + m1 * ldt;
  const double* Tinv2 = Tinv + (m1 / IB) * IB * IB;
  double* B1 = B;
  double* B2 = B + m1;
  if (eff_lower) {
    LLA_TRY(dtrsm_left_inv_dev(upper, op, m1, n, T, ldt, Tinv, B1, ldb, st, abort_flag));
    if (!upper) LLA_TRY(dgemm_dev(0, 0, m2, n, m1, -1.0, T21, ldt, B1, ldb, 1.0, B2, ldb, TRI_FULL, st, abort_flag));
    else LLA_TRY(dgemm_dev(1, 0, m2, n, m1, -1.0, T12, ldt, B1, ldb, 1.0, B2, ldb, TRI_FULL, st, abort_flag));
    LLA_TRY(dtrsm_left_inv_dev(upper, op, m2, n, T22, ldt, Tinv2, B2, ldb, st, abort_flag));
  } else {
    LLA_TRY(dtrsm_left_inv_dev(upper, op, m2, n, T22, ldt, Tinv2, B2, ldb, st, abort_flag));
    if (upper) LLA_TRY(dgemm_dev(0, 0, m1, n, m2, -1.0, T12, ldt, B2, ldb, 1.0, B1, ldb, TRI_FULL, st, abort_flag));
    else LLA_TRY(dgemm_dev(1, 0, m1, n, m2, -1.0, T21, ldt, B2, ldb, 1.0, B1, ldb, TRI_FULL, st, abort_flag));
    LLA_TRY(dtrsm_left_inv_dev(upper, op, m1, n, T, ldt, Tinv, B1, ldb, st, abort_flag));
  }
  return 0;
}

}  // namespace llacuda

using namespace llacuda;

// Left-sided triangular solve op(T) X = B in place on device pointers (the DTRSM LLA's trsm% issues,
// reference src/linear-algebra.lisp:333-347, restricted to side = 'L', alpha = 1): the building block the
// block-cyclic Cholesky / LU of lla_b200/dist.py runs on the block row that the diagonal owner broadcast.
extern "C" int llacuda_dtrsm_dev(char uplo, char trans, char diag, int64_t m, int64_t n, const double* T, int64_t ldt,
                                 double* B, int64_t ldb) {
  ApiLock api_lock;
  const bool upper = (uplo == 'U' || uplo == 'u'), lower = (uplo == 'L' || uplo == 'l');
  const bool unit = (diag == 'U' || diag == 'u'), nonunit = (diag == 'N' || diag == 'n');
  int op = -1;
  if (trans == 'N' || trans == 'n') op = 0;
  if (trans == 'T' || trans == 't' || trans == 'C' || trans == 'c') op = 1;
  if (!(upper || lower) || !(unit || nonunit) || op < 0 || m < 0 || n < 0 || ldt < (m > 1 ? m : 1) || ldb < (m > 1 ? m : 1))
    return LLACUDA_ERR_BAD_ARG;
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  if (m == 0 || n == 0) return 0;
  // per-stream scratch, not the shared arena: this entry point runs next to a factorisation on another stream
  double* tinv = (double*)stream_scratch(c->stream, tinv_bytes(m));
  if (!tinv) { set_error(__FILE__, __LINE__, "dtrsm: out of device memory for the block inverses", (int)m); return LLACUDA_ERR_CUDA_BASE; }
  LLA_TRY(dtrtri_blocks_dev(upper, unit, m, T, ldt, tinv, c->stream));
  return dtrsm_left_inv_dev(upper, op, m, n, T, ldt, tinv, B, ldb, c->stream);
}
