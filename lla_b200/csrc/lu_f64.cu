// lla_b200/csrc/lu_f64.cu -- DGETRF / DGETRS on the device (column-major, partial pivoting).
//
// Replaces LLA's `dgetrf_` (lu, reference src/linear-algebra.lisp:227-234), `dgetrs_`
// (solve on an lu object, :269-276) and `dgesv_` (solve on raw arrays, :282-289).
//
// Structure (the recursion of LAPACK's DGETRF2; the arithmetic per column is DGETF2's):
//   * the matrix is split recursively in column halves; the coupling is TRSM + DMMA GEMM;
//   * a leaf is a tall panel of <= 32 columns, factored by ONE cooperative kernel: every thread
//     owns one row of the panel in registers, a CTA owns 128 rows, and there is exactly one
//     grid-wide barrier per column.  Before the barrier each CTA publishes its pivot candidate
//     (|value|, row index, and the candidate row's 32 values); after it every CTA picks the same
//     winner -- first index of the maximum |a_ij|, NaN never beats a number, the netlib IDAMAX
//     rule, which is what makes ipiv bit-exact -- and applies swap + scale + rank-1 update to its
//     own rows without touching global memory;
//   * the 32 row interchanges of a leaf are applied to all other columns of the matrix right
//     after the leaf (the swap order of classic blocked DGETRF with nb = 32) by a gather/scatter
//     kernel that first composes the 32 swaps into one permutation, so no load depends on a store.
// INFO follows LAPACK: first zero pivot index, factorisation always completes
// (LLA's lu ignores INFO > 0, reference tests/linear-algebra.lisp:106-107).
#include "internal.h"
#include "../../include/llacuda.h"

#include <cooperative_groups.h>
#include <cfloat>

namespace llacuda {

constexpr int PW = 32;     // leaf panel width (columns held in registers per row)
constexpr int PR = 128;    // rows (= threads) per CTA

struct PanelSlot {        // one per CTA per parity
  double absval;          // candidate |a|, -1 if the CTA has no candidate
  int row;                // candidate row, relative to the panel top
  int pad;
  double vals[PW];        // the candidate row
};

struct PanelShared {      // published by the owner of row j
  double rowj[PW];
};

// monotone grid barrier: `count` starts at 0 for every launch
__device__ __forceinline__ void grid_barrier(unsigned* count, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(count, 1u);
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(count) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

// A: top-left of the panel (diagonal element of the full matrix), m rows x w columns.
// ipiv/info are indexed / valued relative to the full matrix through `off` (global row of the top).
__global__ void __launch_bounds__(PR) lla_dgetf2_panel_kernel(double* __restrict__ A, int64_t lda, int64_t m, int w,
                                                          int64_t off, int* __restrict__ ipiv, int* __restrict__ info,
                                                          PanelSlot* __restrict__ slots, PanelShared* __restrict__ shared_rows,
                                                          unsigned* __restrict__ bar) {
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x;
  const int64_t row = (int64_t)blockIdx.x * PR + tid;  // relative to the panel top
  const bool live = row < m;
  const int nsteps = (int)(m < w ? m : w);

  double a[PW];
#pragma unroll
  for (int c = 0; c < PW; c++) a[c] = (live && c < w) ? A[row + (int64_t)c * lda] : 0.0;

  __shared__ double s_val[PR / 32];
  __shared__ int s_row[PR / 32];
  __shared__ double s_prow[PW];
  __shared__ double s_rowj[PW];
  __shared__ int s_piv;

  const double sfmin = DBL_MIN;

#pragma unroll
  for (int j = 0; j < PW; j++) {
    if (j < nsteps) {
      const int par = j & 1;
      // ---- 1. local candidate: first row (>= j) with the largest |a[j]|; NaN only wins as row j
      double v = -1.0;
      int r = 0x7fffffff;
      if (live && row >= j) {
        double x = fabs(a[j]);
        bool isn = (x != x);
        if (row == j && isn) v = INFINITY;  // IDAMAX: dmax = NaN, nothing compares greater
        else v = isn ? -1.0 : x;
        r = (int)row;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, v, o);
        int orow = __shfl_xor_sync(0xffffffffu, r, o);
        if (ov > v || (ov == v && orow < r)) { v = ov; r = orow; }
      }
      if (lane == 0) { s_val[warp] = v; s_row[warp] = r; }
      __syncthreads();
      if (tid == 0) {
        double bv = s_val[0];
        int br = s_row[0];
#pragma unroll
        for (int q = 1; q < PR / 32; q++)
          if (s_val[q] > bv || (s_val[q] == bv && s_row[q] < br)) { bv = s_val[q]; br = s_row[q]; }
        s_piv = br;
        slots[par * G + blockIdx.x].absval = bv;
        slots[par * G + blockIdx.x].row = br;
      }
      __syncthreads();
      // the thread that owns the CTA's candidate row publishes it; the owner of row j publishes row j
      if (live && (int)row == s_piv) {
        double* dst = slots[par * G + blockIdx.x].vals;
#pragma unroll
        for (int c = 0; c < PW; c++) dst[c] = a[c];
      }
      if (live && row == j) {
        double* dst = shared_rows[par].rowj;
#pragma unroll
        for (int c = 0; c < PW; c++) dst[c] = a[c];
      }
      // ---- 2. the one grid-wide barrier of this column
      grid_barrier(bar, (unsigned)G * (unsigned)(j + 1));
      // ---- 3. every CTA picks the same winner
      if (warp == 0) {
        double bv = -2.0;
        int br = 0x7fffffff, bslot = 0;
        for (int q = lane; q < G; q += 32) {
          const PanelSlot* s = &slots[par * G + q];
          double sv = __ldcg(&s->absval);
          int sr = __ldcg(&s->row);
          if (sv > bv || (sv == bv && sr < br)) { bv = sv; br = sr; bslot = q; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          double ov = __shfl_xor_sync(0xffffffffu, bv, o);
          int orow = __shfl_xor_sync(0xffffffffu, br, o);
          int oslot = __shfl_xor_sync(0xffffffffu, bslot, o);
          if (ov > bv || (ov == bv && orow < br)) { bv = ov; br = orow; bslot = oslot; }
        }
        s_prow[lane] = __ldcg(&slots[par * G + bslot].vals[lane]);
        s_rowj[lane] = __ldcg(&shared_rows[par].rowj[lane]);
        if (lane == 0) s_piv = br;
      }
      __syncthreads();
      const int p = s_piv;
      const double piv = s_prow[j];
      if (blockIdx.x == 0 && tid == 0) {
        ipiv[off + j] = (int)(off + p + 1);
        if (piv == 0.0 && *info == 0) *info = (int)(off + j + 1);
      }
      // ---- 4. swap (only when the pivot is non-zero, as DGETF2 does), scale, rank-1 update
      if (piv != 0.0) {
        if (p != j) {
          if (live && row == j) {
#pragma unroll
            for (int c = 0; c < PW; c++) a[c] = s_prow[c];
          } else if (live && (int)row == p) {
#pragma unroll
            for (int c = 0; c < PW; c++) a[c] = s_rowj[c];
          }
        }
        if (live && row > j) {
          if (fabs(piv) >= sfmin) a[j] = a[j] * (1.0 / piv);
          else a[j] = a[j] / piv;
        }
      }
      if (live && row > j) {
        const double l = a[j];
        // s_prow is row j of the panel after the interchange (a zero pivot implies p == j)
#pragma unroll
        for (int c = j + 1; c < PW; c++) a[c] = fma(-l, s_prow[c], a[c]);
      }
      __syncthreads();  // s_prow / s_rowj / s_piv are rewritten in the next column
    }
  }

#pragma unroll
  for (int c = 0; c < PW; c++)
    if (live && c < w) A[row + (int64_t)c * lda] = a[c];
}

// ------------------------------------------------------------------ row interchanges
// Applies the swaps i <-> ipiv[i]-1 for i in [k1, k2) (forward) or in reverse order to the
// columns [0, ncols) of A except [skip0, skip1), 32 swaps at a time: warp 0 composes the chunk
// into one permutation over the <= 64 touched rows, then the CTA gathers and scatters.
constexpr int SW_CHUNK = 32;
constexpr int SW_COLS = 16;

__global__ void __launch_bounds__(2 * SW_CHUNK * SW_COLS) lla_dlaswp_kernel(double* __restrict__ A, int64_t lda, int64_t ncols,
                                                                       int64_t skip0, int64_t skip1, int64_t k1, int64_t k2,
                                                                       const int* __restrict__ ipiv, bool forward,
                                                                       const int* abort_flag) {
  if (abort_flag != nullptr && *abort_flag != 0) return;
  __shared__ int s_rows[2 * SW_CHUNK];  // touched rows (destinations)
  __shared__ int s_src[2 * SW_CHUNK];   // row whose original content ends up there
  __shared__ int s_n;
  const int tid = threadIdx.x;
  const int s = tid % (2 * SW_CHUNK);
  const int cl = tid / (2 * SW_CHUNK);
  int64_t c = (int64_t)blockIdx.x * SW_COLS + cl;
  const int64_t skipw = skip1 - skip0;
  if (c >= skip0) c += skipw;
  const bool col_ok = c < ncols;

  const int64_t nchunks = (k2 - k1 + SW_CHUNK - 1) / SW_CHUNK;
  for (int64_t ch = 0; ch < nchunks; ch++) {
    int64_t lo, hi;  // this chunk covers swaps [lo, hi)
    if (forward) { lo = k1 + ch * SW_CHUNK; hi = lo + SW_CHUNK < k2 ? lo + SW_CHUNK : k2; }
    else { hi = k2 - ch * SW_CHUNK; lo = hi - SW_CHUNK > k1 ? hi - SW_CHUNK : k1; }
    const int cnt = (int)(hi - lo);
    if (tid < 32) {
      // lane k tracks slot k (row lo+k) and slot 32+k (k-th distinct outside row)
      const int lane = tid;
      int row_a = (int)lo + lane, cur_a = row_a;  // rows of the chunk itself
      int row_b = -1, cur_b = -1;                 // extra rows, filled in order of appearance
      int nextra = 0;
      for (int q = 0; q < cnt; q++) {
        const int i = forward ? (int)lo + q : (int)hi - 1 - q;
        const int p = ipiv[i] - 1;
        if (p == i) continue;
        // locate slot of i (always in part a) and of p
        const int ia = i - (int)lo;
        int ip;  // slot index 0..63
        if (p >= lo && p < hi) ip = p - (int)lo;
        else {
          unsigned hit = __ballot_sync(0xffffffffu, row_b == p);
          if (hit) ip = 32 + (__ffs(hit) - 1);
          else {
            ip = 32 + nextra;
            if (lane == nextra) { row_b = p; cur_b = p; }
            nextra++;
          }
        }
        // swap the contents of slots ia and ip
        int ca = __shfl_sync(0xffffffffu, cur_a, ia);
        int cp = (ip < 32) ? __shfl_sync(0xffffffffu, cur_a, ip) : __shfl_sync(0xffffffffu, cur_b, ip - 32);
        if (lane == ia) cur_a = cp;
        if (ip < 32) { if (lane == ip) cur_a = ca; }
        else { if (lane == ip - 32) cur_b = ca; }
      }
      s_rows[lane] = (lane < cnt) ? row_a : -1;
      s_src[lane] = (lane < cnt) ? cur_a : -1;
      s_rows[32 + lane] = (lane < nextra) ? row_b : -1;
      s_src[32 + lane] = (lane < nextra) ? cur_b : -1;
      if (lane == 0) s_n = nextra;
    }
    __syncthreads();
    const int dst = s_rows[s], src = s_src[s];
    const bool act = col_ok && dst >= 0 && dst != src;
    double v = 0.0;
    if (act) v = A[src + c * lda];
    __syncthreads();
    if (act) A[dst + c * lda] = v;
    __syncthreads();
  }
}

int dlaswp_skip_dev(int64_t ncols, int64_t skip0, int64_t skip1, double* A, int64_t lda, int64_t k1, int64_t k2,
                    const int* ipiv, bool forward, cudaStream_t st, const int* abort_flag) {
  int64_t eff = ncols - (skip1 - skip0);
  if (eff <= 0 || k2 <= k1) return 0;
  unsigned grid = (unsigned)((eff + SW_COLS - 1) / SW_COLS);
  lla_dlaswp_kernel<<<grid, 2 * SW_CHUNK * SW_COLS, 0, st>>>(A, lda, ncols, skip0, skip1, k1, k2, ipiv, forward, abort_flag);
  count_launch();
  LLA_CUDA(cudaGetLastError());
  return 0;
}

int dlaswp_dev(int64_t ncols, double* A, int64_t lda, int64_t k1, int64_t k2, const int* ipiv, bool forward,
               cudaStream_t st, const int* abort_flag) {
  return dlaswp_skip_dev(ncols, ncols, ncols, A, lda, k1, k2, ipiv, forward, st, abort_flag);
}

// ------------------------------------------------------------------ recursive driver
struct LuCtx {
  double* A;       // full matrix
  int64_t lda, M, N;
  int* ipiv;
  int* info;
  PanelSlot* slots;
  PanelShared* shared_rows;
  unsigned* bar;
  cudaStream_t st;
};

static int lu_leaf(const LuCtx& L, int64_t j0, int64_t m, int w) {
  // panel = A[j0 : j0+m, j0 : j0+w]
  double* P = L.A + j0 + j0 * L.lda;
  int G = (int)((m + PR - 1) / PR);
  LLA_CUDA(cudaMemsetAsync(L.bar, 0, sizeof(unsigned), L.st));
  int64_t m_ = m, lda_ = L.lda, off_ = j0;
  int w_ = w;
  int* ipiv_ = L.ipiv;
  int* info_ = L.info;
  PanelSlot* slots_ = L.slots;
  PanelShared* sr_ = L.shared_rows;
  unsigned* bar_ = L.bar;
  void* args[] = {&P, &lda_, &m_, &w_, &off_, &ipiv_, &info_, &slots_, &sr_, &bar_};
  LLA_CUDA(cudaLaunchCooperativeKernel((void*)lla_dgetf2_panel_kernel, dim3(G), dim3(PR), args, 0, L.st));
  count_launch();
  // interchange the rest of the rows now: columns [0, j0) and [j0+w, N)
  int64_t npiv = m < w ? m : w;
  LLA_TRY(dlaswp_skip_dev(L.N, j0, j0 + w, L.A, L.lda, j0, j0 + npiv, L.ipiv, true, L.st, nullptr));
  return 0;
}

static int lu_rec(const LuCtx& L, int64_t j0, int64_t m, int64_t n) {
  if (m <= 0 || n <= 0) return 0;
  if (n <= PW) return lu_leaf(L, j0, m, (int)n);
  int64_t mn = m < n ? m : n;
  int64_t n1 = mn > PW ? ((mn / 2) / PW) * PW : mn;
  if (n1 < PW && mn > PW) n1 = PW;
  int64_t n2 = n - n1;
  LLA_TRY(lu_rec(L, j0, m, n1));
  double* A11 = L.A + j0 + j0 * L.lda;
  double* A12 = A11 + n1 * L.lda;
  double* A21 = A11 + n1;
  double* A22 = A12 + n1;
  // A12 <- L11^{-1} A12 ; A22 <- A22 - A21 A12
  LLA_TRY(dtrsm_left_dev(false, 0, true, n1, n2, A11, L.lda, A12, L.lda, L.st));
  if (m > n1) {
    LLA_TRY(dgemm_dev(0, 0, m - n1, n2, n1, -1.0, A21, L.lda, A12, L.lda, 1.0, A22, L.lda, TRI_FULL, L.st));
    LLA_TRY(lu_rec(L, j0 + n1, m - n1, n2));
  }
  return 0;
}

__global__ void lla_iota_ipiv_kernel(int* ipiv, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ipiv[i] = (int)(i + 1);
}

int dgetrf_dev(int64_t m, int64_t n, double* A, int64_t lda, int* ipiv, int* info_dev, cudaStream_t st) {
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  LLA_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), st));
  if (m <= 0 || n <= 0) return 0;
  int maxG = (int)((m + PR - 1) / PR);
  // co-residency bound of the cooperative panel kernel
  int per_sm = 0;
  LLA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lla_dgetf2_panel_kernel, PR, 0));
  if ((int64_t)per_sm * c->sm_count < maxG) {
    set_error(__FILE__, __LINE__, "dgetrf: panel too tall for one cooperative launch", (int)m);
    return LLACUDA_ERR_BAD_ARG;
  }
  size_t need = sizeof(PanelSlot) * 2 * (size_t)maxG + sizeof(PanelShared) * 2 + 256;
  LLA_TRY(ensure_workspace(need));
  char* ws = (char*)c->ws;
  LuCtx L;
  L.A = A; L.lda = lda; L.M = m; L.N = n; L.ipiv = ipiv; L.info = info_dev; L.st = st;
  L.bar = (unsigned*)ws;
  L.shared_rows = (PanelShared*)(ws + 256);
  L.slots = (PanelSlot*)(ws + 256 + sizeof(PanelShared) * 2);
  return lu_rec(L, 0, m, n);
}

int dgetrs_dev(int op, int64_t n, int64_t nrhs, const double* A, int64_t lda, const int* ipiv, double* B,
               int64_t ldb, cudaStream_t st, const int* abort_flag) {
  if (n <= 0 || nrhs <= 0) return 0;
  if (op == 0) {
    LLA_TRY(dlaswp_dev(nrhs, B, ldb, 0, n, ipiv, true, st, abort_flag));
    LLA_TRY(dtrsm_left_dev(false, 0, true, n, nrhs, A, lda, B, ldb, st, abort_flag));
    LLA_TRY(dtrsm_left_dev(true, 0, false, n, nrhs, A, lda, B, ldb, st, abort_flag));
  } else {
    LLA_TRY(dtrsm_left_dev(true, 1, false, n, nrhs, A, lda, B, ldb, st, abort_flag));
    LLA_TRY(dtrsm_left_dev(false, 1, true, n, nrhs, A, lda, B, ldb, st, abort_flag));
    LLA_TRY(dlaswp_dev(nrhs, B, ldb, 0, n, ipiv, false, st, abort_flag));
  }
  return 0;
}

}  // namespace llacuda
