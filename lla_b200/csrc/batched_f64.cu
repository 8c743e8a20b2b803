// lla_b200/csrc/batched_f64.cu -- batched small solves: many independent n x n (n <= 32) double
// systems A_i X_i = B_i, the `(lla:solve a b)` call (DGESV, reference src/linear-algebra.lisp:
// 282-289) issued over a batch (BASELINE.json configs[3]: 200 000 systems of 32 x 32).
//
// One warp per system, lane i owns row i of [A | B] in registers, so each matrix is read exactly
// once (8 n^2 + 8 n nrhs bytes in, 8 n nrhs out, + LU / ipiv when asked).
// Rows are never moved: every lane tracks the logical position of its row, the pivot search is a
// REDUX arg-max over an order-preserving key with the logical position as tie-break (netlib
// IDAMAX: first maximum, NaN never beats a number), and the only cross-lane traffic per column is
// the broadcast of the pivot row.  The DRAM traffic is the algorithmic minimum, but the kernel is
// not HBM-bound: DGETF2's separately rounded multiply and subtract make 1 800 FP64-pipe instructions
// per system (2 issue cycles each), which alone is 0.62 ms for 200 000 systems (profiles/r02_batched.md).  The arithmetic is DGETF2's / DGETRS's operation for operation
// (reciprocal scaling, separately rounded multiply and subtract, same order), so LU, X and the
// LAPACK swap sequence ipiv are bit-identical to the reference algorithm, ties included.
#include "internal.h"
#include "../../include/llacuda.h"

#include <cfloat>
#include <cstdlib>

namespace llacuda {

constexpr int BN32 = 32;

// The pivot row travels through shared memory: the lane that owns it stores it with 16-byte stores, everybody reads it
// back as 16-byte broadcasts -- one LSU instruction per two elements where a 64-bit shuffle costs two per element (the
// round-1 kernel spent 1184 of its 7442 instructions per system on shuffles and as many on selects; this one executes
// 4865, profiles/r02_batched.md).  No lane needs to know WHICH lane owns a row: the owner recognises itself by its
// logical position.  FULL32: n == 32 known at compile time (BASELINE configs[3]): every "inside the matrix" predicate
// folds away.  WARPS systems per CTA, MINB CTAs per SM.
constexpr int BROW = BN32 + 8;   // pivot row + up to 8 right-hand sides
// DGETF2's multiplier for a pivot below the safe minimum is a true division: rare, kept out of line so that the 32
// unrolled columns stay small (the kernel is instruction-fetch sensitive: 20 warps per SM at 32 different columns)
__device__ __noinline__ double lla_div_slow(double a, double b) { return a / b; }

template <int NR, bool FULL32, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) lla_dgesv_batched_smem_kernel(int n_arg, int nrhs, int64_t batch, double* __restrict__ A,
                                                                     int64_t strideA, int* __restrict__ ipiv,
                                                                     double* __restrict__ B, int64_t strideB,
                                                                     int* __restrict__ info, int keep_lu) {
  __shared__ __align__(16) double s_row[WARPS][2][BROW];
  const int n = FULL32 ? BN32 : n_arg;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t sys = (int64_t)blockIdx.x * WARPS + w;
  if (sys >= batch) return;
  double* Am = A + sys * strideA;
  double* Bm = B + sys * strideB;
  const bool rowlive = FULL32 ? true : lane < n;

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
      const int ppos = (int)__reduce_min_sync(0xffffffffu, c2 ? tie : 0x7fffffffu);  // logical row of the pivot
      if (lane == j) mypiv = ppos + 1;
      double* buf = s_row[w][j & 1];
      const int c0 = j & ~1;
      const bool owner = rowlive && pos == ppos;
#pragma unroll
      for (int c = c0; c < BN32; c += 2)
        if (owner) *reinterpret_cast<double2*>(buf + c) = make_double2(a[c], a[c + 1]);
#pragma unroll
      for (int q = 0; q < NR; q++)
        if (owner) buf[BN32 + q] = b[q];
      __syncwarp();
      const double piv = buf[j];
      // An exact zero pivot is recorded and the step goes on without interchange or scaling, as in DGETF2.  The largest
      // magnitude can only be zero when the row at logical j holds a zero itself (a NaN there wins the search), and it is
      // then the first maximum: ppos == j, so the row in `buf` is the row DGETF2 applies in that case too.
      const bool nz = piv != 0.0;
      if (!nz && myinfo == 0) myinfo = j + 1;
      if (nz) {   // logical interchange of rows j and ppos
        if (pos == ppos) pos = j;
        else if (pos == j) pos = ppos;
      }
      const bool below = rowlive && pos > j;
      double l = a[j];
      if (nz) {
        if (fabs(piv) >= sfmin) l = a[j] * (1.0 / piv);
        else l = lla_div_slow(a[j], piv);
        if (below) a[j] = l;
      }
      if (j & 1) {   // odd j: the pair (j - 1, j) is behind us, pairs from j + 1 on
#pragma unroll
        for (int c = j + 1; c < BN32; c += 2) {
          const double2 u = *reinterpret_cast<const double2*>(buf + c);
          if (below) { a[c] = __dsub_rn(a[c], __dmul_rn(l, u.x)); a[c + 1] = __dsub_rn(a[c + 1], __dmul_rn(l, u.y)); }
        }
      } else {       // even j: the pair (j, j + 1) holds the pivot and the first element to update
#pragma unroll
        for (int c = j; c < BN32; c += 2) {
          const double2 u = *reinterpret_cast<const double2*>(buf + c);
          if (below) {
            if (c > j) a[c] = __dsub_rn(a[c], __dmul_rn(l, u.x));
            a[c + 1] = __dsub_rn(a[c + 1], __dmul_rn(l, u.y));
          }
        }
      }
#pragma unroll
      for (int q = 0; q < NR; q++) {
        const double u = buf[BN32 + q];
        if (below) b[q] = __dsub_rn(b[q], __dmul_rn(l, u));
      }
    }
  }

  // ---- back substitution U x = y (skipped when singular, as DGESV does)
  if (myinfo == 0) {
#pragma unroll
    for (int k = BN32 - 1; k >= 0; k--) {
      if (k < n) {
        double* buf = s_row[w][k & 1];
        const bool mine = rowlive && pos == k;
#pragma unroll
        for (int q = 0; q < NR; q++) {
          const double xk = b[q] / a[k];                 // meaningful on the owner of row k only
          if (mine) { b[q] = xk; buf[q] = xk; }
        }
        __syncwarp();
        const bool above = rowlive && pos < k;
#pragma unroll
        for (int q = 0; q < NR; q++) {
          const double xk = buf[q];
          if (above) b[q] = __dsub_rn(b[q], __dmul_rn(xk, a[k]));
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
  // 20 systems per CTA, one CTA per SM (96 registers per thread: 20 warps fill the register file either way) measured
  // 1.60 ms for 200 000 systems against 1.69 ms with five 4-warp CTAs; small batches keep the small CTAs
#define LLA_SMEM(NR, W, MB)                                                                                                  \
  do {                                                                                                                      \
    unsigned g = (unsigned)((batch + (W) - 1) / (W));                                                                       \
    if (n == BN32) lla_dgesv_batched_smem_kernel<NR, true, W, MB><<<g, (W) * 32, 0, st>>>(n, nrhs, batch, A, strideA, ipiv, B, strideB, info, keep_lu); \
    else lla_dgesv_batched_smem_kernel<NR, false, W, MB><<<g, (W) * 32, 0, st>>>(n, nrhs, batch, A, strideA, ipiv, B, strideB, info, keep_lu);          \
  } while (0)
#define LLA_BATCHED(NR)                                                                                                       \
  do {                                                                                                                      \
    if (batch >= 20 * 4 * (int64_t)ctx()->sm_count) LLA_SMEM(NR, 20, 1);                                                    \
    else LLA_SMEM(NR, 4, 5);                                                                                                \
  } while (0)
  if (nrhs <= 1) LLA_BATCHED(1);
  else if (nrhs <= 2) LLA_BATCHED(2);
  else if (nrhs <= 4) LLA_BATCHED(4);
  else LLA_BATCHED(8);
#undef LLA_BATCHED
#undef LLA_SMEM
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
