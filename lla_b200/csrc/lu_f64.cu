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
#include <atomic>
#include <cfloat>
#include <cstdio>
#include <cstdlib>

namespace llacuda {

constexpr int PW = 32;     // leaf panel width (one lane per column when a row travels)
// rows (= threads) per CTA: a template parameter of the panel kernel.  128 rows: up to 128 CTAs exchange through L2 every
// column and sit next to the trailing update's GEMM CTAs on as many SMs; 512 rows: a quarter of the CTAs (the per-column
// all-to-all is the chain of the factorisation: tools/mb/xchg.cu, 3.7k cycles for 128 CTAs, 2.1k for 48, 0.9k for one)
// and a quarter of the SMs taken from the update -- but the CTA-internal barriers and reductions over 16 warps cost what
// the smaller exchange saves (profiles/r02_lu_rows_per_cta.md).  LLACUDA_LU_PR = 128 | 256 | 512 selects; default 128.

// Cross-CTA exchange uses self-validating 8-byte words (32 bits of payload + a 32-bit tag, the
// NCCL "LL" idea): an aligned 8-byte store is single-copy atomic, so a reader that sees the
// expected tag in every word holds a consistent record and no fence or atomic is on the critical
// path.  tag = (launch epoch << 6) | (column + 1) is unique per launch and column.
struct __align__(16) LLDouble { unsigned lo, tag0, hi, tag1; };
// One key header per CTA and column parity, 256 bytes apart: all G CTAs poll all G headers every column, and
// consecutive 256-byte chunks map to different L2 slices, so the polling traffic is spread.  (An all-to-all
// push into per-reader inboxes was measured slower on B200: tools/mb/xchg.cu, 4615 vs 3710 cycles at G = 128.)
// The header is the CTA's best |a_ij| as an order-preserving 64-bit key; ties between CTAs resolve to the
// lowest CTA (= lowest row: each CTA reports its first maximum), so the row index travels with the row.
// Candidate rows are pulled: only the winner's row is read (v[PW] = reciprocal of its pivot value).
struct __align__(256) PanelHdr { LLDouble key; };

struct PanelRow { LLDouble v[PW + 1]; unsigned row, tag2, pad, tag3; };

__device__ __forceinline__ void ll_store(LLDouble* p, double x, unsigned tag) {
  unsigned lo = (unsigned)__double2loint(x), hi = (unsigned)__double2hiint(x);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "r"(lo), "r"(tag), "r"(hi), "r"(tag) : "memory");
}
__device__ __forceinline__ void ll_store_u(void* p, unsigned a, unsigned b, unsigned tag) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "r"(a), "r"(tag), "r"(b), "r"(tag) : "memory");
}
__device__ __forceinline__ bool ll_load_u(const void* p, unsigned tag, unsigned& a, unsigned& b) {
  unsigned t0, t1;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(a), "=r"(t0), "=r"(b), "=r"(t1) : "l"(p) : "memory");
  return t0 == tag && t1 == tag;
}
__device__ __forceinline__ bool ll_load(const LLDouble* p, unsigned tag, double& x) {
  unsigned a, c;
  bool ok = ll_load_u(p, tag, a, c);
  x = __hiloint2double((int)c, (int)a);
  return ok;
}

template <int PR> constexpr size_t panel_smem() { return (size_t)PW * (PR + 1) * sizeof(double); }

// multiplier of DGETF2: reciprocal scaling when |pivot| >= sfmin, division otherwise, nothing for a zero pivot
__device__ __forceinline__ double lu_multiplier(double a, double piv, double rinv) {
  if (piv == 0.0) return a;
  return (fabs(piv) >= DBL_MIN) ? a * rinv : a / piv;
}

// A: top-left of the panel (diagonal element of the full matrix), m rows x w columns.
// ipiv/info are indexed / valued relative to the full matrix through `off` (global row of the top).
// The CTA's 128 x 32 slice of the panel lives in shared memory (column c at S + c*PPITCH), thread t
// owns row t; warp 0 runs the cross-CTA exchange of every column.
//
// The column loop is software-pipelined around the exchange, which is the critical path (two L2 round
// trips per column): once pivot q = j-1 is known every thread applies it to column j ONLY, the arg-max
// of column j starts at once, and each warp stages the fully updated copy of its candidate row (one
// row, lane = column).  The rank-1 update of the remaining columns j+1.. runs while the candidates
// travel through L2.  The arithmetic is unchanged: the update rounds the product and the subtraction
// separately, exactly as DGETF2/DGER do, so a matrix that fits one leaf (n <= 32) is factored
// bit-identically to the reference algorithm and exact ties between pivot candidates resolve the same way.
template <int PR, bool DBG>
__global__ void __launch_bounds__(PR, PR >= 512 ? 1 : (PR >= 256 ? 2 : 3)) lla_dgetf2_panel_kernel(double* __restrict__ A, int64_t lda, int64_t m, int w,
                                                                 int64_t off, int* __restrict__ ipiv, int* __restrict__ info,
                                                                 PanelHdr* __restrict__ hdrs, PanelRow* __restrict__ rows,
                                                                 PanelRow* __restrict__ rowj, unsigned epoch,
                                                                 unsigned long long* __restrict__ dbg) {
  constexpr int PPITCH = PR + 1;   // shared-memory pitch of one panel column
  extern __shared__ __align__(16) double S[];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x;
  const int64_t row0 = (int64_t)blockIdx.x * PR;
  const int64_t grow = row0 + tid;  // relative to the panel top
  const bool live = grow < m;
  const int nsteps = (int)(m < w ? m : w);

#pragma unroll
  for (int c8 = 0; c8 < PW; c8 += 8) {  // 8 loads in flight per thread keeps the register count low
    double v[8];
#pragma unroll
    for (int c = 0; c < 8; c++) v[c] = (live && c8 + c < w) ? A[grow + (int64_t)(c8 + c) * lda] : 0.0;
#pragma unroll
    for (int c = 0; c < 8; c++) S[(c8 + c) * PPITCH + tid] = v[c];
  }

  constexpr int NW = PR / 32;
  __shared__ unsigned long long s_key[NW];
  __shared__ int s_row[NW];
  __shared__ double s_cand[NW][PW + 1];  // each warp's candidate row, updated through the pending pivot; [PW] = 1 / its pivot value
  __shared__ double s_rowj_out[PW];      // CTA 0: row j, updated through the pending pivot
  __shared__ double s_prow[2][PW];       // pivot row of column j in s_prow[j & 1]
  __shared__ double s_rowq[PW];          // the row that sat at position j before the interchange
  __shared__ double s_rinv;
  __shared__ int s_piv;
  __syncthreads();

  double lmul = 0.0;  // this row's multiplier for the pending pivot

#pragma unroll 1
  for (int j = 0; j <= nsteps; j++) {
    const int q = j - 1;  // pending pivot: known, applied to nothing yet
    bool upd = false;
    long long tk[7];
    if (DBG) tk[0] = clock64();
    const double* prow = s_prow[q & 1];
    double xj = 0.0;      // this row's element of column j, after the pending update
    if (q >= 0) {
      const int p = s_piv;
      const double piv = prow[q];
      const double rinv = s_rinv;
      if (blockIdx.x == 0 && tid == 0) {
        ipiv[off + q] = (int)(off + p + 1);
        if (piv == 0.0 && *info == 0) *info = (int)(off + q + 1);
      }
      // ---- interchange (only when the pivot is non-zero, as DGETF2 does): whole warps rewrite the two
      // rows (lane = column); row p is rewritten by the warp that owns it, so __syncwarp orders it
      if (piv != 0.0 && p != q) {
        const int pl = p - (int)row0;
        if (pl >= 0 && pl < PR && warp == (pl >> 5)) { S[lane * PPITCH + pl] = s_rowq[lane]; __syncwarp(); }
        if (blockIdx.x == 0 && warp == NW - 1) S[lane * PPITCH + q] = prow[lane];  // final: nobody reads row q again
      }
      // ---- multiplier, and the update of column j only
      if (live && grow > q) {
        upd = true;
        lmul = lu_multiplier(S[q * PPITCH + tid], piv, rinv);
        S[q * PPITCH + tid] = lmul;
        if (j < w) {
          xj = __dsub_rn(S[j * PPITCH + tid], __dmul_rn(lmul, prow[j]));
          S[j * PPITCH + tid] = xj;
        }
      }
      __syncwarp();
    } else if (live) {
      xj = S[tid];
    }
    if (j < nsteps) {
      // ---- candidate of this warp: first row (>= j) with the largest |a[j]|; NaN only wins as row j.
      // The magnitude becomes an order-preserving 64-bit key (0 = no number here) so that the
      // arg-max is three 32-bit REDUX operations instead of five rounds of 64-bit shuffles.
      unsigned long long key = 0ull;
      unsigned r = 0x7fffffffu;
      if (live && grow >= j) {
        double x = fabs(xj);
        bool isn = (x != x);
        if (grow == j && isn) x = INFINITY;  // IDAMAX: dmax = NaN, nothing compares greater
        if (x == x) key = (unsigned long long)__double_as_longlong(x) + 1ull;
        r = (unsigned)grow;
      }
      const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
      const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
      const bool c1 = hi == mh;
      const unsigned ml = __reduce_max_sync(0xffffffffu, c1 ? lo : 0u);
      const bool c2 = c1 && lo == ml;
      const unsigned rm = __reduce_min_sync(0xffffffffu, c2 ? r : 0x7fffffffu);
      if (lane == 0) { s_key[warp] = ((unsigned long long)mh << 32) | ml; s_row[warp] = (int)rm; }
      if (rm != 0x7fffffffu) {  // warp-uniform: stage the candidate row as it will be after the pending update
        const int lr = (int)rm - (int)row0;
        const double cv = __shfl_sync(0xffffffffu, xj, lr & 31);
        double v = S[lane * PPITCH + lr];
        if (q >= 0 && lane > j) v = __dsub_rn(v, __dmul_rn(S[q * PPITCH + lr], prow[lane]));
        s_cand[warp][lane] = v;
        if (lane == 0) s_cand[warp][PW] = 1.0 / cv;  // the division overlaps the barrier and the CTA reduction
      }
      if (blockIdx.x == 0 && warp == 0) {  // row j lives in CTA 0, warp 0
        double v = S[lane * PPITCH + j];
        if (q >= 0 && lane > j) v = __dsub_rn(v, __dmul_rn(S[q * PPITCH + j], prow[lane]));
        s_rowj_out[lane] = v;
      }
    }
    if (DBG) tk[1] = clock64();
    __syncthreads();
    const int par = j & 1;
    const unsigned tag = (epoch << 6) | (unsigned)(j + 1);
    PanelHdr* hd = hdrs + (size_t)par * G;
    if (warp == 0 && j < nsteps) {
      // ---- CTA candidate: publish the key, then the row
      unsigned long long k8 = (lane < NW) ? s_key[lane] : 0ull;
      unsigned r8 = (lane < NW) ? (unsigned)s_row[lane] : 0x7fffffffu;
      unsigned hi = (unsigned)(k8 >> 32), lo = (unsigned)k8;
      unsigned mh = __reduce_max_sync(0xffffffffu, hi);
      bool c1 = hi == mh;
      unsigned ml = __reduce_max_sync(0xffffffffu, c1 ? lo : 0u);
      bool c2 = c1 && lo == ml;
      const unsigned br = __reduce_min_sync(0xffffffffu, c2 ? r8 : 0x7fffffffu);
      if (lane == 0) ll_store_u(&hd[blockIdx.x].key, ml, mh, tag);
      const int bw = ((int)br - (int)row0) >> 5;  // the CTA's candidate row is always one of its own live rows
      PanelRow* myrow = &rows[par * G + blockIdx.x];
      ll_store(&myrow->v[lane], s_cand[bw][lane], tag);
      if (blockIdx.x == 0) ll_store(&rowj[par].v[lane], s_rowj_out[lane], tag);
      if (lane == 0) {
        ll_store(&myrow->v[PW], s_cand[bw][PW], tag);
        ll_store_u(&myrow->row, br, 0u, tag);
      }
    }
    if (DBG) tk[2] = clock64();
    // ---- the rest of the pending rank-1 update (columns j+1 ..), in the shadow of the exchange.
    // Groups of 8 columns: all loads first, then the arithmetic, then the stores (the compiler cannot
    // prove that S and s_prow do not alias, so a plain loop would serialise on shared-memory latency)
    if (upd) {
      for (int c8 = (j + 1) & ~7; c8 < PW; c8 += 8) {
        double av[8], pv[8];
#pragma unroll
        for (int t = 0; t < 8; t++) { av[t] = S[(c8 + t) * PPITCH + tid]; pv[t] = prow[c8 + t]; }
#pragma unroll
        for (int t = 0; t < 8; t++)
          if (c8 + t > j) S[(c8 + t) * PPITCH + tid] = __dsub_rn(av[t], __dmul_rn(lmul, pv[t]));
      }
    }
    if (j == nsteps) break;
    if (DBG) tk[3] = clock64();
    if (warp == 0) {
      // ---- wait for every CTA's key (all of a lane's slots in flight together), pick the winner
      unsigned whi = 0, wlo = 0, wslot = 0x7fffffffu;
      for (int base = 0; base < G; base += 128) {
        unsigned a[4], b[4];
        bool done[4];
#pragma unroll
        for (int t = 0; t < 4; t++) done[t] = base + t * 32 + lane >= G;
        bool all;
        do {
          all = true;
#pragma unroll
          for (int t = 0; t < 4; t++)
            if (!done[t]) {
              done[t] = ll_load_u(&hd[base + t * 32 + lane].key, tag, a[t], b[t]);
              all = all && done[t];
            }
        } while (!all);
#pragma unroll
        for (int t = 0; t < 4; t++) {
          const int slot = base + t * 32 + lane;
          if (slot < G && (b[t] > whi || (b[t] == whi && a[t] > wlo) || wslot == 0x7fffffffu)) {
            whi = b[t]; wlo = a[t]; wslot = (unsigned)slot;   // slots ascend within a lane: strict > keeps the first
          }
        }
      }
      const unsigned mh = __reduce_max_sync(0xffffffffu, whi);
      const bool c1 = whi == mh && wslot != 0x7fffffffu;
      const unsigned ml = __reduce_max_sync(0xffffffffu, c1 ? wlo : 0u);
      const bool c2 = c1 && wlo == ml;
      const unsigned ws = __reduce_min_sync(0xffffffffu, c2 ? wslot : 0x7fffffffu);
      if (DBG) tk[4] = clock64();
      // ---- fetch the winner's row and the old row j, both in flight together
      const PanelRow* win = &rows[par * G + ws];
      double pv, rj, ri = 0.0;
      unsigned wr = 0, dummy;
      bool ok1, ok2, ok3, ok4;
      do {  // every lane issues the same four loads (no divergent round trips): 2 x 512 bytes + two broadcast words
        ok1 = ll_load(&win->v[lane], tag, pv);
        ok2 = ll_load(&rowj[par].v[lane], tag, rj);
        ok3 = ll_load_u(&win->row, tag, wr, dummy);
        ok4 = ll_load(&win->v[PW], tag, ri);
      } while (!(ok1 && ok2 && ok3 && ok4));
      s_prow[par][lane] = pv;
      s_rowq[lane] = rj;
      if (lane == 0) s_piv = (int)wr;
      if (lane == 1) s_rinv = ri;
      if (DBG) tk[5] = clock64();
    }
    __syncthreads();
    if (DBG && tid == 0 && blockIdx.x == gridDim.x / 2) {  // phases of warp 0 of a middle CTA, cycles summed over columns
      tk[6] = clock64();
      for (int t = 0; t < 6; t++) atomicAdd(&dbg[t], (unsigned long long)(tk[t + 1] - tk[t]));
      atomicAdd(&dbg[6], 1ull);
    }
  }
  __syncthreads();

  if (live) {
#pragma unroll
    for (int c = 0; c < PW; c++)
      if (c < w) A[grow + (int64_t)c * lda] = S[c * PPITCH + tid];
  }
}

// ------------------------------------------------------------------ row interchanges
// Applies the swaps i <-> ipiv[i]-1 for i in [k1, k2) (forward) or in reverse order to the
// columns [0, ncols) of A except [skip0, skip1), 32 swaps at a time: warp 0 composes the chunk
// into one permutation over the <= 64 touched rows, then the CTA gathers and scatters.
constexpr int SW_CHUNK = 32;
constexpr int SW_COLS = 8;

// One warp composes the interchanges [lo, hi) (at most 32, applied forward or in reverse) into one permutation
// over the <= 64 touched rows: dst[s] receives the original content of row src[s] (-1 = unused slot).  Lane k
// tracks slot k (row lo+k) and slot 32+k (k-th distinct row outside [lo, hi)).
__device__ __forceinline__ void laswp_compose_chunk(int lane, int64_t lo, int64_t hi, bool forward,
                                                    const int* __restrict__ ipiv, int* dst, int* src) {
  const int cnt = (int)(hi - lo);
  int row_a = (int)lo + lane, cur_a = row_a;  // rows of the chunk itself
  int row_b = -1, cur_b = -1;                 // extra rows, filled in order of appearance
  int nextra = 0;
  for (int q = 0; q < cnt; q++) {
    const int i = forward ? (int)lo + q : (int)hi - 1 - q;
    const int p = ipiv[i] - 1;
    if (p == i) continue;
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
  dst[lane] = (lane < cnt) ? row_a : -1;
  src[lane] = (lane < cnt) ? cur_a : -1;
  dst[32 + lane] = (lane < nextra) ? row_b : -1;
  src[32 + lane] = (lane < nextra) ? cur_b : -1;
}

__global__ void __launch_bounds__(2 * SW_CHUNK * SW_COLS) lla_dlaswp_kernel(double* __restrict__ A, int64_t lda, int64_t ncols,
                                                                       int64_t skip0, int64_t skip1, int64_t k1, int64_t k2,
                                                                       const int* __restrict__ ipiv, bool forward,
                                                                       const int* abort_flag) {
  if (abort_flag != nullptr && *abort_flag != 0) return;
  __shared__ int s_rows[2 * SW_CHUNK];  // touched rows (destinations)
  __shared__ int s_src[2 * SW_CHUNK];   // row whose original content ends up there
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
    if (tid < 32) laswp_compose_chunk(tid, lo, hi, forward, ipiv, s_rows, s_src);
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

// Long interchange sequences (a whole outer panel: up to 512 swaps) on many columns: the rows [k1, k2) of the
// CTA's 16 columns are loaded into shared memory once (coalesced), all chunk permutations are composed up front
// by eight warps in parallel, and the chunks are then applied one after the other with only the rows OUTSIDE
// [k1, k2) touching global memory (a later chunk may revisit such a row: same CTA, ordered by the barriers).
constexpr int LS_R = 512;
constexpr int LS_CB = 16;
constexpr int LS_T = 256;
constexpr size_t LS_SMEM = (size_t)LS_R * LS_CB * sizeof(double);

__global__ void __launch_bounds__(LS_T) lla_dlaswp_dense_kernel(double* __restrict__ A, int64_t lda, int64_t ncols,
                                                                int64_t skip0, int64_t skip1, int64_t k1, int64_t k2,
                                                                const int* __restrict__ ipiv, bool forward,
                                                                const int* abort_flag) {
  if (abort_flag != nullptr && *abort_flag != 0) return;
  extern __shared__ __align__(16) double Dn[];
  __shared__ int s_dst[LS_R / SW_CHUNK][2 * SW_CHUNK];
  __shared__ int s_src[LS_R / SW_CHUNK][2 * SW_CHUNK];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R = (int)(k2 - k1);
  const int nch = (R + SW_CHUNK - 1) / SW_CHUNK;
  const int64_t skipw = skip1 - skip0;
  const int64_t cbase = (int64_t)blockIdx.x * LS_CB;
  auto column = [&](int cl) -> int64_t {  // global column of local column cl, or -1
    int64_t c = cbase + cl;
    if (c >= skip0) c += skipw;
    return c < ncols ? c : -1;
  };
  for (int ch = warp; ch < nch; ch += LS_T / 32) {
    int64_t lo, hi;
    if (forward) { lo = k1 + (int64_t)ch * SW_CHUNK; hi = lo + SW_CHUNK < k2 ? lo + SW_CHUNK : k2; }
    else { hi = k2 - (int64_t)ch * SW_CHUNK; lo = hi - SW_CHUNK > k1 ? hi - SW_CHUNK : k1; }
    laswp_compose_chunk(lane, lo, hi, forward, ipiv, s_dst[ch], s_src[ch]);
  }
  for (int idx = tid; idx < R * LS_CB; idx += LS_T) {
    const int r = idx % R, cl = idx / R;
    const int64_t c = column(cl);
    if (c >= 0) Dn[cl * LS_R + r] = A[k1 + r + c * lda];
  }
  __syncthreads();
  const int s = tid % (2 * SW_CHUNK), cl0 = tid / (2 * SW_CHUNK);  // 4 column groups of 64 slots
  for (int ch = 0; ch < nch; ch++) {
    const int dst = s_dst[ch][s], src = s_src[ch][s];
    const bool moves = dst >= 0 && dst != src;
    double v[LS_CB / 4];
#pragma unroll
    for (int i = 0; i < LS_CB / 4; i++) {
      const int cl = cl0 + 4 * i;
      const int64_t c = column(cl);
      v[i] = 0.0;
      if (moves && c >= 0) v[i] = (src >= k1 && src < k2) ? Dn[cl * LS_R + (src - (int)k1)] : A[src + c * lda];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < LS_CB / 4; i++) {
      const int cl = cl0 + 4 * i;
      const int64_t c = column(cl);
      if (moves && c >= 0) {
        if (dst >= k1 && dst < k2) Dn[cl * LS_R + (dst - (int)k1)] = v[i];
        else A[dst + c * lda] = v[i];
      }
    }
    __syncthreads();
  }
  for (int idx = tid; idx < R * LS_CB; idx += LS_T) {
    const int r = idx % R, cl = idx / R;
    const int64_t c = column(cl);
    if (c >= 0) A[k1 + r + c * lda] = Dn[cl * LS_R + r];
  }
}

int dlaswp_skip_dev(int64_t ncols, int64_t skip0, int64_t skip1, double* A, int64_t lda, int64_t k1, int64_t k2,
                    const int* ipiv, bool forward, cudaStream_t st, const int* abort_flag) {
  int64_t eff = ncols - (skip1 - skip0);
  if (eff <= 0 || k2 <= k1) return 0;
  static const bool no_dense = getenv("LLACUDA_LASWP_OLD") != nullptr;
  if (k2 - k1 > 2 * SW_CHUNK && k2 - k1 <= LS_R && !no_dense) {
    static DeviceOnce once;
    if (once.first()) {
      LLA_CUDA(cudaFuncSetAttribute(lla_dlaswp_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LS_SMEM));
    }
    unsigned g = (unsigned)((eff + LS_CB - 1) / LS_CB);
    lla_dlaswp_dense_kernel<<<g, LS_T, LS_SMEM, st>>>(A, lda, ncols, skip0, skip1, k1, k2, ipiv, forward, abort_flag);
    count_launch();
    LLA_CUDA(cudaGetLastError());
    return 0;
  }
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
  PanelHdr* hdrs;
  PanelRow* rows;
  PanelRow* rowj;
  double* linv;    // inverses of the 128 x 128 diagonal blocks of L (unit lower), block b at b*IB*IB
  int64_t swap_c0, swap_c1;  // columns that receive a leaf's interchanges right after the leaf
  cudaStream_t st;
  int per_sm = 1;            // co-resident panel CTAs per SM (occupancy of the leaf kernel)
  struct LuExt* ext = nullptr;   // chain-bound phase: the next panel's columns trail this panel block by block
  int pr = 0;                    // rows per CTA of the panel kernel for this panel (0 = panel_rows())
  LuRhs* rhs = nullptr;          // right-hand sides riding along (blocked driver only)
};

// The columns [c0, c1) of the NEXT panel, updated on stream `sx` one 128-column block behind the panel chain (see
// lu_blocked).  The update of block b reads the L columns of block b below the diagonal; the panel's next interchange of
// those columns (the first leaf of block b + 1) therefore waits for `ev_y`.
struct LuExt {
  int64_t c0 = 0, c1 = 0;
  cudaStream_t sx = nullptr;
  cudaEvent_t ev_x = nullptr, ev_y = nullptr;
  bool pending = false;      // a block update is in flight on sx: the next interchange of the panel's columns waits for it
};

static std::atomic<unsigned> g_panel_epoch{1};

static int panel_rows() {   // rows per CTA of the panel kernel
  static int pr = 0;
  if (pr == 0) {
    const char* e = getenv("LLACUDA_LU_PR");
    pr = e ? atoi(e) : 128;   // measured on B200, dgetrf n = 16384 / 8192: 128 rows 147.5 / 39.6 ms, 256: 146.9 / 41.1, 512: 151.9 / 44.2
    if (pr != 128 && pr != 256 && pr != 512) pr = 128;
  }
  return pr;
}

template <int PR>
static int launch_panel(const LuCtx& L, double* P, int64_t m, int w, int64_t off, unsigned epoch, unsigned long long* dbg) {
  static DeviceOnce once;
  if (once.first()) {
    LLA_CUDA(cudaFuncSetAttribute(lla_dgetf2_panel_kernel<PR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)panel_smem<PR>()));
    LLA_CUDA(cudaFuncSetAttribute(lla_dgetf2_panel_kernel<PR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)panel_smem<PR>()));
  }
  int G = (int)((m + PR - 1) / PR);
  int64_t m_ = m, lda_ = L.lda, off_ = off;
  int w_ = w;
  int* ipiv_ = L.ipiv;
  int* info_ = L.info;
  PanelHdr* h_ = L.hdrs;
  PanelRow* r_ = L.rows;
  PanelRow* rj_ = L.rowj;
  void* args[] = {&P, &lda_, &m_, &w_, &off_, &ipiv_, &info_, &h_, &r_, &rj_, &epoch, &dbg};
  LLA_CUDA(cudaLaunchCooperativeKernel(dbg ? (void*)lla_dgetf2_panel_kernel<PR, true> : (void*)lla_dgetf2_panel_kernel<PR, false>, dim3(G),
                                       dim3(PR), args, panel_smem<PR>(), L.st));
  count_launch();
  return 0;
}

static int lu_leaf(const LuCtx& L, int64_t j0, int64_t m, int w) {
  // panel = A[j0 : j0+m, j0 : j0+w]
  double* P = L.A + j0 + j0 * L.lda;
  unsigned epoch = g_panel_epoch.fetch_add(1) & 0x03ffffffu;
  if (epoch == 0) epoch = g_panel_epoch.fetch_add(1) & 0x03ffffffu;
  // LLACUDA_PANEL_DBG=1: per-column phase breakdown of the exchange (instrumented instantiation of the kernel)
  static unsigned long long* dbg = nullptr;
  static bool dbg_checked = false;
  if (!dbg_checked) {
    dbg_checked = true;
    if (getenv("LLACUDA_PANEL_DBG")) { cudaMalloc(&dbg, 128); cudaMemset(dbg, 0, 128); }
  }
  if (dbg && (g_panel_epoch.load() % 128) == 0) {
    unsigned long long h[16];
    cudaMemcpy(h, dbg, 128, cudaMemcpyDeviceToHost);
    if (h[6]) fprintf(stderr, "panel phases, cycles/column over %llu columns (m=%lld): apply+cand=%llu publish=%llu own-update=%llu poll=%llu fetch=%llu sync=%llu\n",
                      h[6], (long long)m, h[0] / h[6], h[1] / h[6], h[2] / h[6], h[3] / h[6], h[4] / h[6], h[5] / h[6]);
    cudaMemset(dbg, 0, 128);
  }
  static const bool leaf_trace = getenv("LLACUDA_TRACE_LEAF") != nullptr;
  if (leaf_trace) trace_mark(L.st, "lu.leaf>", j0, m);
  int pr = L.pr ? L.pr : panel_rows();
  {
    // co-residency of the cooperative launch (launch bounds: 3 / 2 / 1 CTAs per SM for 128 / 256 / 512 rows; the measured
    // occupancy of the default height is in L.per_sm): taller CTAs when the panel does not fit
    const int64_t sms = ctx()->sm_count;
    auto fits = [&](int rows) {
      const int64_t per = rows == panel_rows() ? L.per_sm : (rows == 128 ? 3 : (rows == 256 ? 2 : 1));
      return (m + rows - 1) / rows <= per * sms;
    };
    while (pr < 512 && !fits(pr)) pr *= 2;
  }
  if (pr == 128) LLA_TRY(launch_panel<128>(L, P, m, w, j0, epoch, dbg));
  else if (pr == 256) LLA_TRY(launch_panel<256>(L, P, m, w, j0, epoch, dbg));
  else LLA_TRY(launch_panel<512>(L, P, m, w, j0, epoch, dbg));
  if (leaf_trace) trace_mark(L.st, "lu.leafK<", j0, m);
  if (L.ext && L.ext->pending) {   // the trailing block update still reads the columns left of this leaf
    LLA_CUDA(cudaStreamWaitEvent(L.st, L.ext->ev_y, 0));
    L.ext->pending = false;
  }
  // interchange the other columns of the swap window now: [swap_c0, j0) and [j0+w, swap_c1)
  int64_t npiv = m < w ? m : w;
  LLA_TRY(dlaswp_skip_dev(L.swap_c1 - L.swap_c0, j0 - L.swap_c0, j0 + w - L.swap_c0, L.A + L.swap_c0 * L.lda, L.lda,
                          j0, j0 + npiv, L.ipiv, true, L.st, nullptr));
  if (leaf_trace) trace_mark(L.st, "lu.leaf<", j0, m);
  return 0;
}

// m x n block whose top-left corner is the diagonal element (j0, j0)
static int lu_rec(const LuCtx& L, int64_t j0, int64_t m, int64_t n) {
  if (m <= 0 || n <= 0) return 0;
  if (n <= PW) return lu_leaf(L, j0, m, (int)n);
  const int64_t mn = m < n ? m : n;
  // split on multiples of 128 above 128 columns (GEMM-only solves with the inverted diagonal
  // blocks of L), on multiples of 32 inside a 128-column block (substitution leaves)
  int64_t n1;
  if (mn > IB) n1 = (mn >= 2 * IB) ? ((mn / 2) / IB) * IB : IB;
  else if (mn > PW) n1 = (mn >= 2 * PW) ? ((mn / 2) / PW) * PW : PW;
  else n1 = mn;
  const int64_t n2 = n - n1;
  LLA_TRY(lu_rec(L, j0, m, n1));
  double* A11 = L.A + j0 + j0 * L.lda;
  double* A12 = A11 + n1 * L.lda;
  double* A21 = A11 + n1;
  double* A22 = A12 + n1;
  if (n2 > 0) {
    // A12 <- L11^{-1} A12 ; A22 <- A22 - A21 A12
    if (n1 % IB == 0 && j0 % IB == 0)
      LLA_TRY(dtrsm_left_inv_dev(false, 0, n1, n2, A11, L.lda, L.linv + (j0 / IB) * IB * IB, A12, L.lda, L.st));
    else
      LLA_TRY(dtrsm_left_dev(false, 0, true, n1, n2, A11, L.lda, A12, L.lda, L.st));
    if (m > n1) {
      LLA_TRY(dgemm_dev(0, 0, m - n1, n2, n1, -1.0, A21, L.lda, A12, L.lda, 1.0, A22, L.lda, TRI_FULL, L.st));
      LLA_TRY(lu_rec(L, j0 + n1, m - n1, n2));
    }
  }
  // a 128-column block of L is final once the node that covers exactly that block returns:
  // invert its unit-lower diagonal block for the GEMM-only solves of the levels above
  if (j0 % IB == 0 && n <= IB && (n == IB || j0 + n == L.N || m <= n)) {
    int64_t r = mn;  // rows/cols of the diagonal block that exist
    LLA_TRY(dtrtri_blocks_dev(false, true, r, A11, L.lda, L.linv + (j0 / IB) * IB * IB, L.st));
    if (L.ext && n == IB && m >= IB) {
      // this block's share of the next panel's columns: interchanges, U = L_bb^-1 A, rows below -= L_b U
      LuExt& X = *L.ext;
      const int64_t nc = X.c1 - X.c0;
      double* E = L.A + j0 + X.c0 * L.lda;
      LLA_CUDA(cudaEventRecord(X.ev_x, L.st));
      LLA_CUDA(cudaStreamWaitEvent(X.sx, X.ev_x, 0));
      LLA_TRY(dlaswp_dev(nc, L.A + X.c0 * L.lda, L.lda, j0, j0 + IB, L.ipiv, true, X.sx));
      LLA_TRY(dtrsm_left_inv_dev(false, 0, IB, nc, A11, L.lda, L.linv + (j0 / IB) * IB * IB, E, L.lda, X.sx));
      if (m > IB) {
        LLA_TRY(dgemm_dev(0, 0, m - IB, nc, IB, -1.0, A11 + IB, L.lda, E, L.lda, 1.0, E + IB, L.lda, TRI_FULL, X.sx));
        LLA_CUDA(cudaEventRecord(X.ev_y, X.sx));
        X.pending = true;
      }
    }
  }
  return 0;
}

// Experiment (LLACUDA_LU_PANEL_RL=1): the panel of a blocked step right-looking over its 32-column leaves -- leaf (+ its
// interchanges on the other panel columns), 32-row solve of the columns to its right, rank-32 update below: 66 launches
// per 512-column panel instead of the recursion's 76, but 15 passes over the panel with k = 32 GEMMs.  Measured slower
// (n = 16384: 150.3 against 146.1 ms, profiles/r02_lu_chain.md); the recursion stays the default.
static int lu_panel_rl(const LuCtx& L, int64_t k, int64_t m, int64_t kb) {
  for (int64_t c = 0; c < kb; c += PW) {
    const int64_t w = (kb - c) < PW ? (kb - c) : PW;
    const int64_t j0 = k + c, R = kb - c - w, mr = m - c;
    if (mr <= 0) break;
    LLA_TRY(lu_leaf(L, j0, mr, (int)w));
    double* A11 = L.A + j0 + j0 * L.lda;
    double* A12 = A11 + w * L.lda;
    if (R > 0) {
      LLA_TRY(dtrsm_left_dev(false, 0, true, w < mr ? w : mr, R, A11, L.lda, A12, L.lda, L.st));
      if (mr > w) LLA_TRY(dgemm_dev(0, 0, mr - w, R, w, -1.0, A11 + w, L.lda, A12, L.lda, 1.0, A12 + w, L.lda, TRI_FULL, L.st));
    }
    const int64_t e = c + w;   // columns of the panel finished so far
    if (e % IB == 0 || e == kb) {   // a 128-column block of L is final: invert its unit-lower diagonal block
      const int64_t bs = ((e - 1) / IB) * IB;
      int64_t r = e - bs;
      if (m - bs < r) r = m - bs;
      if (r > 0) LLA_TRY(dtrtri_blocks_dev(false, true, r, L.A + (k + bs) + (k + bs) * L.lda, L.lda, L.linv + ((k + bs) / IB) * IB * IB, L.st));
    }
  }
  return 0;
}

// columns [c0, c1) right of panel (k, kb): interchanges and U12 = L11^-1 A12 ...
static int lu_update_solve(const LuCtx& L, int64_t k, int64_t kb, int64_t c0, int64_t c1, cudaStream_t st) {
  if (c1 <= c0) return 0;
  const int64_t nc = c1 - c0;
  double* A11 = L.A + k + k * L.lda;
  double* A12 = L.A + k + c0 * L.lda;
  static const bool upd_trace = getenv("LLACUDA_TRACE_LEAF") != nullptr;
  LLA_TRY(dlaswp_dev(nc, L.A + c0 * L.lda, L.lda, k, k + kb, L.ipiv, true, st));
  if (upd_trace) trace_mark(st, "lu.upd.swp<", k, nc);
  LLA_TRY(dtrsm_left_inv_dev(false, 0, kb, nc, A11, L.lda, L.linv + (k / IB) * IB * IB, A12, L.lda, st));
  if (upd_trace) trace_mark(st, "lu.upd.trsm<", k, nc);
  return 0;
}
// ... and A22 -= A21 U12
static int lu_update_product(const LuCtx& L, int64_t k, int64_t kb, int64_t c0, int64_t c1, cudaStream_t st, int hint = GEMM_AUTO) {
  if (c1 <= c0 || L.M <= k + kb) return 0;
  double* A11 = L.A + k + k * L.lda;
  double* A12 = L.A + k + c0 * L.lda;
  return dgemm_dev(0, 0, L.M - k - kb, c1 - c0, kb, -1.0, A11 + kb, L.lda, A12, L.lda, 1.0, A12 + kb, L.lda, TRI_FULL, st, nullptr, hint);
}
static int lu_update_cols(const LuCtx& L, int64_t k, int64_t kb, int64_t c0, int64_t c1, cudaStream_t st, int hint = GEMM_AUTO) {
  LLA_TRY(lu_update_solve(L, k, kb, c0, c1, st));
  return lu_update_product(L, k, kb, c0, c1, st, hint);
}

// Blocked right-looking driver with one panel of look-ahead: panel k+1 (recursive, cooperative
// leaf kernels) runs on the high-priority panel stream while the trailing update of step k is
// still running on the main stream; the interchanges of the columns left of a panel go to a third
// stream, since nothing depends on them until the end.
static int lu_blocked(LuCtx L) {
  Context* c = ctx();
  static int64_t NB_env = -1;
  if (NB_env < 0) { const char* e = getenv("LLACUDA_LU_NB"); NB_env = e ? atoi(e) : 0; }
  const int64_t NB = NB_env > 0 ? NB_env : 512;
  const int64_t mn = L.M < L.N ? L.M : L.N;
  const RangeReadyHook* ready = range_ready_hook();
  if (mn <= 2 * NB) {
    if (ready) LLA_TRY(ready->fn(ready->user, 0, L.N, L.st));
    L.swap_c0 = 0; L.swap_c1 = L.N;
    L.rhs = nullptr;   // no rider here: the caller solves
    return lu_rec(L, 0, L.M, L.N);
  }
  cudaStream_t s1 = L.st, s2 = c->aux, s3 = c->h2d;
  cudaEvent_t ev_start = c->ev[3], ev_panel = c->ev[4], ev_a = c->ev[5], ev_b = c->ev[6], ev_left = c->ev[7];
  LLA_CUDA(cudaEventRecord(ev_start, s1));
  LLA_CUDA(cudaStreamWaitEvent(s2, ev_start, 0));
  if (L.rhs) LLA_CUDA(cudaStreamWaitEvent(c->aux4, ev_start, 0));
  // Chain-bound phase: the panel chain (cooperative leaves, interchanges, small solves and products) moves to its own SM
  // partition, the trailing update to the rest.  Under a concurrent update on shared SMs a panel takes 2-3x as long as
  // alone (profiles/r01b_trace_dgetrf_16384_leaves.log: leaves 155 us against 83), and from the step where the update is
  // shorter than the panel that is the critical path.  The cooperative leaf must fit the partition: G <= per_sm * SMs.
  static int64_t part_m = -1;
  if (part_m < 0) { const char* e = getenv("LLACUDA_LU_PART_M"); part_m = e ? atol(e) : 8192; }
  const SmPartition* part = part_m > 0 ? sm_partition() : nullptr;
  bool part_on = false;
  bool have_b = false;
  // Chain-bound phase, second measure: the update of the NEXT panel's columns (interchanges, solve, product -- "updA", 0.6 ms
  // per step after a panel of 1.8-2.1 ms) trails the panel block by block on a stream of its own, so that only the last
  // 128-column block's share is left between two panels (0.3 ms).  Only in the chain-bound phase (remaining rows <=
  // LLACUDA_LU_EXT_M): the trailing stream needs the previous step's update of those columns, which that update therefore
  // does first and signals separately (ev_b1); earlier in the factorisation the panel would stall behind it.  Trailing
  // leaf by leaf (32 columns) was measured slower: 48 small launches per panel lag the chain's interchanges
  // (profiles/r02_lu_chain.md).
  static int64_t ext_m = -1;
  if (ext_m < 0) { const char* e = getenv("LLACUDA_LU_EXT_M"); ext_m = e ? atol(e) : 9216; }
  LuExt X;
  X.ev_x = c->ev[8]; X.ev_y = c->ev[9];
  cudaEvent_t ev_b1 = c->ev[10];
  bool have_b1 = false;
  static const bool side_a = getenv("LLACUDA_LU_NO_SIDE_A") == nullptr;
  static int64_t chunk_w = -1;   // columns per chunk of the pipelined bulk update (0 = off)
  if (chunk_w < 0) { const char* e = getenv("LLACUDA_LU_CHUNK"); chunk_w = e ? atol(e) : 4096; if (chunk_w % NB) chunk_w = 0; }
  bool chunk_live[32] = {};      // the chunk's product event has been recorded in this call
  bool prev_chunked = false;     // the previous step's update recorded per-chunk product events
  cudaEvent_t ev_t0 = c->ev[11], ev_rhs = c->ev[12];
  for (int64_t k = 0; k < mn; k += NB) {
    const int64_t kb = (mn - k) < NB ? (mn - k) : NB;
    if (part && !part_on && k > 0 && (L.M - k) <= part_m &&
        (L.M - k + panel_rows() - 1) / panel_rows() <= (int64_t)L.per_sm * part->sm_panel) {
      // the update streams of the partition follow everything queued so far; the panel stream only what panel k needs
      // (the previous panel and the update of its own columns), so the switch costs no look-ahead
      cudaEvent_t j1 = c->ev[0], j2 = c->ev[1], j3 = c->ev[2];
      LLA_CUDA(cudaEventRecord(j1, s1));
      LLA_CUDA(cudaEventRecord(j2, s2));
      LLA_CUDA(cudaEventRecord(j3, s3));
      for (cudaStream_t t : {part->bulk, part->bulk2})
        for (cudaEvent_t e : {j1, j2, j3}) LLA_CUDA(cudaStreamWaitEvent(t, e, 0));
      LLA_CUDA(cudaStreamWaitEvent(part->panel, j2, 0));
      LLA_CUDA(cudaStreamWaitEvent(part->panel, ev_a, 0));
      s1 = part->bulk; s2 = part->panel; s3 = part->bulk2;
      part_on = true;
      trace_mark(s1, "lu.partition", k, part->sm_panel);
    }
    // ---- panel k on s2
    LuCtx P = L;
    P.st = s2; P.swap_c0 = k; P.swap_c1 = k + kb;
    const int64_t ext_kb = (k + kb < mn) ? ((mn - k - kb) < NB ? (mn - k - kb) : NB) : 0;
    const bool b1_now = have_b1;   // the previous step updated this step's extension columns first and recorded ev_b1
    have_b1 = false;
    const bool use_ext = k > 0 && ext_kb > 0 && kb % IB == 0 && (L.M - k) <= ext_m && L.M - k - kb >= IB && s2 != s1;
    if (use_ext) {
      X.c0 = k + kb; X.c1 = k + kb + ext_kb; X.pending = false;
      X.sx = part_on ? part->bulk3 : c->aux3;
      // the columns received the previous step's update on s1 (ev_b) and, before that switch, nothing else is queued on sx
      if (b1_now) LLA_CUDA(cudaStreamWaitEvent(X.sx, ev_b1, 0));
      else if (have_b) LLA_CUDA(cudaStreamWaitEvent(X.sx, ev_b, 0));
      P.ext = &X;
    }
    // update-bound phase: 256-row panel CTAs -- half as many CTAs taking register file and shared memory from the
    // update's GEMM CTAs on half as many SMs; the panel itself is hidden there (LLACUDA_LU_PR_EARLY = 0 | 256 | 512;
    // dgetrf n = 16384: 138.3 / 136.3 / 137.4 ms)
    static int pr_early = -1;
    if (pr_early < 0) { const char* e = getenv("LLACUDA_LU_PR_EARLY"); pr_early = e ? atoi(e) : 256; if (pr_early != 256 && pr_early != 512) pr_early = 0; }
    if (pr_early && (L.M - k) > ext_m && !part_on) P.pr = pr_early;
    if (ready && k == 0) LLA_TRY(ready->fn(ready->user, 0, kb, s2));
    trace_mark(s2, "lu.panel>", k, kb);
    static const bool panel_rl = getenv("LLACUDA_LU_PANEL_RL") != nullptr;
    if (panel_rl) LLA_TRY(lu_panel_rl(P, k, L.M - k, kb));
    else LLA_TRY(lu_rec(P, k, L.M - k, kb));
    trace_mark(s2, "lu.panel<", k, kb);
    LLA_CUDA(cudaEventRecord(ev_panel, s2));
    if (L.rhs) {
      // the right-hand sides ride along: interchanges of this panel, solve with its unit-lower triangle, rows below
      LuRhs& Rh = *L.rhs;
      cudaStream_t sB = c->aux4;
      double* A11 = L.A + k + k * L.lda;
      LLA_CUDA(cudaStreamWaitEvent(sB, ev_panel, 0));
      LLA_TRY(dlaswp_dev(Rh.nrhs, Rh.B, Rh.ldb, k, k + kb, L.ipiv, true, sB));
      LLA_TRY(dtrsm_left_inv_dev(false, 0, kb, Rh.nrhs, A11, L.lda, L.linv + (k / IB) * IB * IB, Rh.B + k, Rh.ldb, sB));
      if (L.M > k + kb)
        LLA_TRY(dgemm_dev(0, 0, L.M - k - kb, Rh.nrhs, kb, -1.0, A11 + kb, L.lda, Rh.B + k, Rh.ldb, 1.0, Rh.B + k + kb, Rh.ldb, TRI_FULL, sB));
      LLA_CUDA(cudaEventRecord(ev_rhs, sB));   // the left interchanges of later steps permute the rows this product reads
      Rh.done = true;
    }
    // ---- interchanges of the finished columns [0, k) on s3 (after the update that still reads them)
    if (k > 0) {
      LLA_CUDA(cudaStreamWaitEvent(s3, ev_panel, 0));
      if (have_b) LLA_CUDA(cudaStreamWaitEvent(s3, ev_b, 0));
      if (L.rhs) LLA_CUDA(cudaStreamWaitEvent(s3, ev_rhs, 0));
      LLA_TRY(dlaswp_dev(k, L.A, L.lda, k, k + kb, L.ipiv, true, s3));
      LLA_CUDA(cudaEventRecord(ev_left, s3));
    }
    // ---- trailing update of step k on s1: next panel's columns first, then the rest
    const int64_t c0 = k + kb;
    if (c0 >= L.N) break;
    LLA_CUDA(cudaStreamWaitEvent(s1, ev_panel, 0));
    const int64_t next_kb = (k + kb < mn) ? ((mn - c0) < NB ? (mn - c0) : NB) : 0;
    const int64_t c1 = c0 + next_kb;
    bool join_a = false;   // s1 has to see ev_a (recorded on another stream) before the block row counts as final
    if (use_ext) {
      // the next panel's columns are done on sx (the last block's share was queued when the panel returned)
      trace_mark(X.sx, "lu.ext<", k, c1 - c0);
      LLA_CUDA(cudaEventRecord(ev_a, X.sx));
      LLA_CUDA(cudaStreamWaitEvent(s2, ev_a, 0));
      join_a = true;
    } else if (side_a && c1 > c0) {
      // the next panel's columns on the side stream, as soon as the previous update has done them (its first piece,
      // ev_b1): the next panel starts while the rest of that update is still running, and s1 goes from update to update
      cudaStream_t sx = part_on ? part->bulk3 : c->aux3;
      LLA_CUDA(cudaStreamWaitEvent(sx, ev_panel, 0));
      if (b1_now) LLA_CUDA(cudaStreamWaitEvent(sx, ev_b1, 0));
      else if (have_b) LLA_CUDA(cudaStreamWaitEvent(sx, ev_b, 0));
      else LLA_CUDA(cudaStreamWaitEvent(sx, ev_start, 0));
      trace_mark(sx, "lu.updA>", k, c1 - c0);
      if (ready && k == 0) LLA_TRY(ready->fn(ready->user, c0, c1, sx));
      LLA_TRY(lu_update_cols(L, k, kb, c0, c1, sx));
      trace_mark(sx, "lu.updA<", k, c1 - c0);
      LLA_CUDA(cudaEventRecord(ev_a, sx));
      LLA_CUDA(cudaStreamWaitEvent(s2, ev_a, 0));
      join_a = true;
    } else {
      trace_mark(s1, "lu.updA>", k, c1 - c0);
      if (ready && k == 0 && c1 > c0) LLA_TRY(ready->fn(ready->user, c0, c1, s1));
      LLA_TRY(lu_update_cols(L, k, kb, c0, c1, s1));
      trace_mark(s1, "lu.updA<", k, c1 - c0);
      LLA_CUDA(cudaEventRecord(ev_a, s1));
      LLA_CUDA(cudaStreamWaitEvent(s2, ev_a, 0));
    }
    // the rest of the update runs as a persistent GEMM that leaves `reserve` SMs to the panel stream:
    // the factorisation is bound by the panel chain, so the chain gets SMs of its own
    static int reserve = -1;
    if (reserve < 0) { const char* e = getenv("LLACUDA_LU_RESERVE"); reserve = e ? atoi(e) : 0; }  // measured: no gain on B200 (profiles/r01_lu_tuning.txt), off by default
    const bool overlap = (c1 < L.N) && (k + kb < mn);
    bool step_chunked = false;
    if (ready && k == 0) {
      // first step: the update is separable by column block, so each block starts as soon as its columns are on the device
      for (int64_t cc = c1; cc < L.N; cc += READY_BLOCK) {
        const int64_t ce = cc + READY_BLOCK < L.N ? cc + READY_BLOCK : L.N;
        LLA_TRY(ready->fn(ready->user, cc, ce, s1));
        LLA_TRY(lu_update_cols(L, k, kb, cc, ce, s1));
      }
    } else {
      // when the next step trails its extension, that extension's columns (the first NB of this update) come first and
      // signal on their own, so the trailing stream can start while the rest of this update is still running
      int64_t cs = c1;
      const bool next_ext = s2 != s1 && c1 < L.N && (side_a || (ext_m > 0 && (L.M - c0) <= ext_m));
      // Update-bound phase: the interchanges and the GEMM-only solve of a column chunk (0.95 ms per step over all
      // columns, low occupancy) run on a stream of their own as soon as the PREVIOUS step's product has left that chunk,
      // i.e. under the previous step's remaining products; s1 then runs product after product.  Chunks are fixed
      // absolute column ranges, so one event pair per chunk orders step k - 1's product, step k's solve, step k's product.
      const bool chunked = chunk_w > 0 && s2 != s1 && !part_on && (L.M - k) > ext_m && L.N / chunk_w < 32;
      step_chunked = chunked;
      cudaStream_t sT = c->aux2;
      // what last wrote the columns of chunk `id`: the previous step's product of that chunk, or its whole update, or nothing
      auto prev_event = [&](int id) -> cudaEvent_t {
        if (prev_chunked && chunk_live[id]) return c->ev_chunk[1][id];
        return have_b ? ev_b : ev_start;
      };
      auto solve_chunk = [&](int64_t lo, int64_t hi) -> int {     // on sT, after the previous product of these columns
        const int id = (int)(lo / chunk_w);
        LLA_CUDA(cudaStreamWaitEvent(sT, ev_panel, 0));
        LLA_CUDA(cudaStreamWaitEvent(sT, prev_event(id), 0));
        LLA_TRY(lu_update_solve(L, k, kb, lo, hi, sT));
        LLA_CUDA(cudaEventRecord(c->ev_chunk[0][id], sT));
        return 0;
      };
      auto product_chunk = [&](int64_t lo, int64_t hi) -> int {   // on s1, after the solve of these columns
        const int id = (int)(lo / chunk_w);
        LLA_CUDA(cudaStreamWaitEvent(s1, c->ev_chunk[0][id], 0));
        LLA_TRY(lu_update_product(L, k, kb, lo, hi, s1));
        LLA_CUDA(cudaEventRecord(c->ev_chunk[1][id], s1));
        chunk_live[id] = true;
        return 0;
      };
      if (next_ext) {
        cs = c1 + NB < L.N ? c1 + NB : L.N;
        if (chunked) {
          // the first piece lies inside one chunk; its product is recorded in ev_b1, the chunk's event follows below
          const int id = (int)(c1 / chunk_w);
          LLA_CUDA(cudaStreamWaitEvent(sT, ev_panel, 0));
          LLA_CUDA(cudaStreamWaitEvent(sT, prev_event(id), 0));
          LLA_TRY(lu_update_solve(L, k, kb, c1, cs, sT));
          LLA_CUDA(cudaEventRecord(ev_t0, sT));
          LLA_CUDA(cudaStreamWaitEvent(s1, ev_t0, 0));
          LLA_TRY(lu_update_product(L, k, kb, c1, cs, s1));
        } else {
          LLA_TRY(lu_update_cols(L, k, kb, c1, cs, s1));
        }
        LLA_CUDA(cudaEventRecord(ev_b1, s1));
        have_b1 = true;
      }
      if (chunked) {
        for (int64_t lo = cs; lo < L.N;) {
          const int64_t hi = (lo / chunk_w + 1) * chunk_w < L.N ? (lo / chunk_w + 1) * chunk_w : L.N;
          LLA_TRY(solve_chunk(lo, hi));
          lo = hi;
        }
        for (int64_t lo = cs; lo < L.N;) {
          const int64_t hi = (lo / chunk_w + 1) * chunk_w < L.N ? (lo / chunk_w + 1) * chunk_w : L.N;
          LLA_TRY(product_chunk(lo, hi));
          lo = hi;
        }
        // a chunk that the first piece cut (its columns [c1, cs) were updated outside the loop) is covered by s1's order
        if (cs > c1 && cs % chunk_w != 0 && cs >= L.N) {
          LLA_CUDA(cudaEventRecord(c->ev_chunk[1][(int)(c1 / chunk_w)], s1));
          chunk_live[(int)(c1 / chunk_w)] = true;
        }
      } else {
        LLA_TRY(lu_update_cols(L, k, kb, cs, L.N, s1, (overlap && reserve >= 8) ? GEMM_RESERVE_SMS + (reserve - 8) : GEMM_AUTO));
      }
    }
    trace_mark(s1, "lu.updB<", k, L.N - c1);
    LLA_CUDA(cudaEventRecord(ev_b, s1));
    have_b = true;
    prev_chunked = step_chunked;
    if (join_a) LLA_CUDA(cudaStreamWaitEvent(s1, ev_a, 0));
    // rows [k, k+kb) are final: U12 was solved at the head of this update, the L part received its last
    // interchanges from this panel (later panels only touch rows below)
    if (const RowsFinalHook* h = rows_final_hook()) LLA_TRY(h->fn(h->user, k, k + kb, s1, k > 0 ? s3 : nullptr));
  }
  LLA_CUDA(cudaStreamWaitEvent(s1, ev_panel, 0));
  if (mn > NB) LLA_CUDA(cudaStreamWaitEvent(s1, ev_left, 0));
  if (L.rhs) LLA_CUDA(cudaStreamWaitEvent(s1, ev_rhs, 0));
  if (part_on) {   // back to the caller's stream
    LLA_CUDA(cudaEventRecord(ev_start, s1));
    LLA_CUDA(cudaStreamWaitEvent(L.st, ev_start, 0));
  }
  return 0;
}

int dgetrf_dev(int64_t m, int64_t n, double* A, int64_t lda, int* ipiv, int* info_dev, cudaStream_t st, LuRhs* rhs) {
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  LLA_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), st));
  if (m <= 0 || n <= 0) return 0;
  const int pr = panel_rows();
  int maxG = (int)((m + pr - 1) / pr);
  // co-residency bound of the cooperative panel kernel
  static int per_sm_dev[MAX_DEV] = {};
  int& per_sm = per_sm_dev[c->device];
  if (per_sm == 0) {
    if (pr == 128) {
      LLA_CUDA(cudaFuncSetAttribute(lla_dgetf2_panel_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)panel_smem<128>()));
      LLA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lla_dgetf2_panel_kernel<128, false>, 128, panel_smem<128>()));
    } else if (pr == 256) {
      LLA_CUDA(cudaFuncSetAttribute(lla_dgetf2_panel_kernel<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)panel_smem<256>()));
      LLA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lla_dgetf2_panel_kernel<256, false>, 256, panel_smem<256>()));
    } else {
      LLA_CUDA(cudaFuncSetAttribute(lla_dgetf2_panel_kernel<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)panel_smem<512>()));
      LLA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lla_dgetf2_panel_kernel<512, false>, 512, panel_smem<512>()));
    }
  }
  // a panel taller than the co-residency bound of the default CTA height runs with taller CTAs (lu_leaf escalates);
  // only beyond 512 rows per CTA x one CTA per SM there is no cooperative launch left
  if ((m + 511) / 512 > (int64_t)c->sm_count) {
    set_error(__FILE__, __LINE__, "dgetrf: panel too tall for one cooperative launch", (int)m);
    return LLACUDA_ERR_BAD_ARG;
  }
  const int64_t mn = m < n ? m : n;
  Arena ar;
  LLA_TRY(ar.reserve(sizeof(PanelHdr) * 2 * maxG + sizeof(PanelRow) * (2 * maxG + 2) + tinv_bytes(mn) + 4096, st));
  LuCtx L;
  L.A = A; L.lda = lda; L.M = m; L.N = n; L.ipiv = ipiv; L.info = info_dev; L.st = st;
  L.swap_c0 = 0; L.swap_c1 = n;
  L.per_sm = per_sm;
  L.rhs = (rhs && rhs->B && rhs->nrhs > 0 && m == n) ? rhs : nullptr;
  L.hdrs = (PanelHdr*)ar.take(sizeof(PanelHdr) * 2 * maxG);
  L.rows = (PanelRow*)ar.take(sizeof(PanelRow) * 2 * maxG);
  L.rowj = (PanelRow*)ar.take(sizeof(PanelRow) * 2);
  L.linv = (double*)ar.take(tinv_bytes(mn));
  return lu_blocked(L);
}

int dgetrs_dev(int op, int64_t n, int64_t nrhs, const double* A, int64_t lda, const int* ipiv, double* B,
               int64_t ldb, cudaStream_t st, const int* abort_flag) {
  if (n <= 0 || nrhs <= 0) return 0;
  // inverses of the 128 x 128 diagonal blocks of L and of U (two batched launches); both
  // triangular solves are then GEMMs only
  Arena ar;
  LLA_TRY(ar.reserve(2 * tinv_bytes(n) + 1024, st));
  double* linv = (double*)ar.take(tinv_bytes(n));
  double* uinv = (double*)ar.take(tinv_bytes(n));
  LLA_TRY(dtrtri_blocks_dev(false, true, n, A, lda, linv, st, abort_flag));
  LLA_TRY(dtrtri_blocks_dev(true, false, n, A, lda, uinv, st, abort_flag));
  if (op == 0) {
    LLA_TRY(dlaswp_dev(nrhs, B, ldb, 0, n, ipiv, true, st, abort_flag));
    LLA_TRY(dtrsm_left_inv_dev(false, 0, n, nrhs, A, lda, linv, B, ldb, st, abort_flag));
    LLA_TRY(dtrsm_left_inv_dev(true, 0, n, nrhs, A, lda, uinv, B, ldb, st, abort_flag));
  } else {
    LLA_TRY(dtrsm_left_inv_dev(true, 1, n, nrhs, A, lda, uinv, B, ldb, st, abort_flag));
    LLA_TRY(dtrsm_left_inv_dev(false, 1, n, nrhs, A, lda, linv, B, ldb, st, abort_flag));
    LLA_TRY(dlaswp_dev(nrhs, B, ldb, 0, n, ipiv, false, st, abort_flag));
  }
  return 0;
}

// B <- saved copy when the factorisation reported INFO != 0 (DGESV leaves B alone then)
__global__ void lla_restore_rhs_kernel(const int* __restrict__ info, double* __restrict__ B, int64_t ldb,
                                       const double* __restrict__ save, int64_t n, int64_t nrhs) {
  if (*info == 0) return;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n * nrhs; idx += (int64_t)gridDim.x * blockDim.x)
    B[idx % n + (idx / n) * ldb] = save[idx];
}

int dgesv_dev(int64_t n, int64_t nrhs, double* A, int64_t lda, int* ipiv, double* B, int64_t ldb, int* info_dev, cudaStream_t st) {
  static const bool no_rider = getenv("LLACUDA_LU_NO_RHS_RIDER") != nullptr;
  if (n <= 0) return dgetrf_dev(n, n, A, lda, ipiv, info_dev, st);
  if (nrhs <= 0 || nrhs > 1024 || no_rider) {
    LLA_TRY(dgetrf_dev(n, n, A, lda, ipiv, info_dev, st));
    return dgetrs_dev(0, n, nrhs, A, lda, ipiv, B, ldb, st, info_dev);
  }
  // the forward substitution rides along with the factorisation (one panel behind it, on a side stream); a copy of B
  // is kept for the singular case, in which DGESV does not touch B
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  // keyed by the rider's stream: the factorisation's own arena is the scratch of `st`
  double* save = (double*)stream_scratch(c->aux4, (size_t)n * nrhs * sizeof(double));
  if (!save) { set_error(__FILE__, __LINE__, "dgesv: out of device memory for the right-hand-side copy", (int)n); return LLACUDA_ERR_CUDA_BASE; }
  LLA_CUDA(cudaMemcpy2DAsync(save, (size_t)n * 8, B, (size_t)ldb * 8, (size_t)n * 8, (size_t)nrhs, cudaMemcpyDeviceToDevice, st));
  LuRhs rider;
  rider.B = B; rider.ldb = ldb; rider.nrhs = nrhs;
  LLA_TRY(dgetrf_dev(n, n, A, lda, ipiv, info_dev, st, &rider));
  if (!rider.done) return dgetrs_dev(0, n, nrhs, A, lda, ipiv, B, ldb, st, info_dev);
  {
    Arena ar;
    LLA_TRY(ar.reserve(tinv_bytes(n) + 1024, st));
    double* uinv = (double*)ar.take(tinv_bytes(n));
    LLA_TRY(dtrtri_blocks_dev(true, false, n, A, lda, uinv, st, info_dev));
    LLA_TRY(dtrsm_left_inv_dev(true, 0, n, nrhs, A, lda, uinv, B, ldb, st, info_dev));
  }
  lla_restore_rhs_kernel<<<256, 256, 0, st>>>(info_dev, B, ldb, save, n, nrhs);
  count_launch();
  LLA_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace llacuda

using namespace llacuda;

// Row interchanges i <-> ipiv[i]-1, i in [k1, k2), on the ncols columns of A (DLASWP with incx = +-1): what the
// processes of the block-cyclic LU (lla_b200/dist.py) apply to their own columns after a panel's pivots arrive.
extern "C" int llacuda_dlaswp_dev(int64_t ncols, double* A, int64_t lda, int64_t k1, int64_t k2, const int* ipiv_dev,
                                  int forward) {
  ApiLock api_lock;
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  if (ncols < 0 || k1 < 0 || k2 < k1 || lda < 1) return LLACUDA_ERR_BAD_ARG;
  return dlaswp_dev(ncols, A, lda, k1, k2, ipiv_dev, forward != 0, c->stream);
}
