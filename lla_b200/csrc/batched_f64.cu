// lla_b200/csrc/batched_f64.cu -- batched small solves: many independent n x n (n <= 32) double
// systems A_i X_i = B_i, the `(lla:solve a b)` call (DGESV, reference src/linear-algebra.lisp:
// 282-289) issued over a batch (BASELINE.json configs[3]: 200 000 systems of 32 x 32).
//
// HBM-bound by design: one warp per system, lane i owns row i of [A | B] in registers, so each
// matrix is read exactly once (8 n^2 + 8 n nrhs bytes in, 8 n nrhs out, + LU / ipiv when asked).
// Rows are never moved: every lane tracks the logical position of its row, the pivot search is a
// REDUX arg-max over an order-preserving key with the logical position as tie-break (netlib
// IDAMAX: first maximum, NaN never beats a number), and the only cross-lane traffic per column is
// the broadcast of the pivot row.  The arithmetic is DGETF2's / DGETRS's operation for operation
// (reciprocal scaling, separately rounded multiply and subtract, same order), so LU, X and the
// LAPACK swap sequence ipiv are bit-identical to the reference algorithm, ties included.
#include "internal.h"
#include "../../include/llacuda.h"

#include <cfloat>

namespace llacuda {

constexpr int BN32 = 32;

// FULL32: n == 32 known at compile time (BASELINE configs[3]): every "inside the matrix" predicate folds away
template <int NR, bool FULL32>
__global__ void __launch_bounds__(128, 5) lla_dgesv_batched_kernel(int n_arg, int nrhs, int64_t batch, double* __restrict__ A,
                                                                int64_t strideA, int* __restrict__ ipiv,
                                                                double* __restrict__ B, int64_t strideB,
                                                                int* __restrict__ info, int keep_lu) {
  const int n = FULL32 ? BN32 : n_arg;
  const int lane = threadIdx.x & 31;
  const int64_t sys = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (sys >= batch) return;
  double* Am = A + sys * strideA;
  double* Bm = B + sys * strideB;
  const bool rowlive = lane < n;

  double a[BN32], b[NR];
#pragma unroll
  for (int c = 0; c < BN32; c++) a[c] = (rowlive && c < n) ? Am[lane + (int64_t)c * n] : 0.0;
#pragma unroll
  for (int q = 0; q < NR; q++) b[q] = (rowlive && q < nrhs) ? Bm[lane + (int64_t)q * n] : 0.0;

  int pos = lane;          // logical row index of the row this lane holds
  int myinfo = 0;
  int mypiv = 0;           // lane j keeps ipiv[j]
  const double sfmin = DBL_MIN;

#pragma unroll
  for (int j = 0; j < BN32; j++) {
    if (j < n) {
      // ---- pivot: first (in logical order) row >= j with the largest |a[j]|
      unsigned long long key = 0ull;
      unsigned tie = 0x7fffffffu;
      if (rowlive && pos >= j) {
        double x = fabs(a[j]);
        bool isn = (x != x);
        if (pos == j && isn) x = INFINITY;
        if (x == x) key = (unsigned long long)__double_as_longlong(x) + 1ull;
        tie = (unsigned)pos;
      }
      unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
      unsigned mh = __reduce_max_sync(0xffffffffu, hi);
      bool c1 = hi == mh;
      unsigned ml = __reduce_max_sync(0xffffffffu, c1 ? lo : 0u);
      bool c2 = c1 && lo == ml;
      const unsigned ppos = __reduce_min_sync(0xffffffffu, c2 ? tie : 0x7fffffffu);  // logical row of the pivot
      const int plane = __ffs(__ballot_sync(0xffffffffu, rowlive && pos == (int)ppos)) - 1;
      const int jlane = __ffs(__ballot_sync(0xffffffffu, rowlive && pos == j)) - 1;
      if (lane == j) mypiv = (int)ppos + 1;
      const double piv = __shfl_sync(0xffffffffu, a[j], plane);
      if (piv == 0.0) {
        if (myinfo == 0) myinfo = j + 1;   // exact zero pivot: DGETF2 records it and goes on, no swap
      } else {
        // logical interchange of rows j and ppos
        if (lane == plane) pos = j;
        else if (lane == jlane) pos = (int)ppos;
      }
      const int prow_lane = (piv == 0.0) ? jlane : plane;   // the row that now sits at logical j
      const bool below = rowlive && pos > j;
      double l = a[j];
      if (below && piv != 0.0) {
        l = (fabs(piv) >= sfmin) ? a[j] * (1.0 / piv) : a[j] / piv;
        a[j] = l;
      }
#pragma unroll
      for (int c = j + 1; c < BN32; c++) {
        const double u = __shfl_sync(0xffffffffu, a[c], prow_lane);
        if (below) a[c] = __dsub_rn(a[c], __dmul_rn(l, u));
      }
#pragma unroll
      for (int q = 0; q < NR; q++) {
        const double u = __shfl_sync(0xffffffffu, b[q], prow_lane);
        if (below) b[q] = __dsub_rn(b[q], __dmul_rn(l, u));
      }
    }
  }

  // ---- back substitution U x = y (skipped when singular, as DGESV does)
  if (myinfo == 0) {
#pragma unroll
    for (int k = BN32 - 1; k >= 0; k--) {
      if (k < n) {
        const int klane = __ffs(__ballot_sync(0xffffffffu, rowlive && pos == k)) - 1;
#pragma unroll
        for (int q = 0; q < NR; q++) {
          double xk = b[q] / a[k];                       // meaningful on klane only
          xk = __shfl_sync(0xffffffffu, xk, klane);
          if (lane == klane) b[q] = xk;
          else if (rowlive && pos < k) b[q] = __dsub_rn(b[q], __dmul_rn(xk, a[k]));
        }
      }
    }
  }

  // ---- results: X (or the untouched B when singular), ipiv, info, optionally the factors
  if (rowlive) {
    if (myinfo == 0) {
#pragma unroll
      for (int q = 0; q < NR; q++)
        if (q < nrhs) Bm[pos + (int64_t)q * n] = b[q];
    }
    if (ipiv) ipiv[sys * n + lane] = mypiv;
    if (keep_lu) {
#pragma unroll
      for (int c = 0; c < BN32; c++)
        if (c < n) Am[pos + (int64_t)c * n] = a[c];
    }
  }
  if (lane == 0) info[sys] = myinfo;
}

int dgesv_batched_dev(int n, int nrhs, int64_t batch, double* A, int64_t strideA, int* ipiv, double* B,
                      int64_t strideB, int* info, int keep_lu, cudaStream_t st) {
  if (batch <= 0 || n <= 0) return 0;
  if (n > BN32 || nrhs < 0 || nrhs > 8) {
    set_error(__FILE__, __LINE__, "dgesv_batched: supported sizes are n <= 32, nrhs <= 8", n);
    return LLACUDA_ERR_BAD_ARG;
  }
  const int warps = 4;  // 128-thread CTAs: 5 fit the register file at 100 registers per thread (20 warps per SM; 256-thread CTAs stop at 16)
  unsigned grid = (unsigned)((batch + warps - 1) / warps);
#define LLA_BATCHED(NR)                                                                                                       \
  do {                                                                                                                      \
    if (n == BN32) lla_dgesv_batched_kernel<NR, true><<<grid, warps * 32, 0, st>>>(n, nrhs, batch, A, strideA, ipiv, B, strideB, info, keep_lu); \
    else lla_dgesv_batched_kernel<NR, false><<<grid, warps * 32, 0, st>>>(n, nrhs, batch, A, strideA, ipiv, B, strideB, info, keep_lu);          \
  } while (0)
  if (nrhs <= 1) LLA_BATCHED(1);
  else if (nrhs <= 2) LLA_BATCHED(2);
  else if (nrhs <= 4) LLA_BATCHED(4);
  else LLA_BATCHED(8);
#undef LLA_BATCHED
  count_launch();
  LLA_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace llacuda

using namespace llacuda;

extern "C" int llacuda_dgesv_batched_dev(int n, int nrhs, int64_t batch, double* A, int64_t strideA, int* ipiv_dev,
                                         double* B, int64_t strideB, int* info_dev, int keep_lu) {
  ApiLock api_lock;
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  if (strideA < (int64_t)n * n || strideB < (int64_t)n * (nrhs > 0 ? nrhs : 1)) return LLACUDA_ERR_BAD_ARG;
  return dgesv_batched_dev(n, nrhs, batch, A, strideA, ipiv_dev, B, strideB, info_dev, keep_lu, c->stream);
}
