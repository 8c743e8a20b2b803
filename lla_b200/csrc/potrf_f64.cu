// lla_b200/csrc/potrf_f64.cu -- DPOTRF / DPOTRS on the device (column-major).
//
// Replaces LLA's `dpotrf_` (cholesky, reference src/linear-algebra.lisp:670-674: UPLO='U' on the
// untransposed row-major buffer, which yields the row-major lower factor) and `dpotrs_`
// (solve on a cholesky object, :296-301).
//
// Recursive formulation (as LAPACK's DPOTRF2) on multiples of 128: A11 = U11^T U11 ; U12 = U11^-T A12 ;
// A22 -= U12^T U12 (DMMA GEMM restricted to the upper triangle) ; recurse on A22.  Leaves are
// 128 x 128 blocks factored by one CTA in shared memory, which also leaves inv(U_ii) behind so
// that every triangular solve above the leaves is a DMMA GEMM.  Only the named triangle is ever
// written.  A non-positive or NaN pivot stores INFO = its 1-based index in a
// device flag that every later kernel of the call checks, so, as in LAPACK, nothing after the
// failing minor is modified.  UPLO='L' runs the same code on the in-place transpose.
#include "internal.h"
#include "tri128.cuh"
#include "../../include/llacuda.h"

#include <cstdlib>
#include <cstdio>

namespace llacuda {

// One CTA factors an r x r (r <= 128) diagonal block held in shared memory as L = U^T, writes U
// back over the upper triangle of A and leaves inv(U) in `uinv` (128 x 128, upper, zero below) for
// the GEMM-only triangular solves of the levels above.
__global__ void __launch_bounds__(T128_THREADS) lla_dpotrf128_kernel(int r, double* __restrict__ A, int64_t lda,
                                                                     double* __restrict__ uinv, int off,
                                                                     int* __restrict__ info, long long* dbg) {
  if (*info != 0) return;
  long long tk[6];
  if (dbg) tk[0] = clock64();
  extern __shared__ __align__(16) double sm128[];
  double* S = sm128;
  double* tmp = sm128 + T128 * T128_SP;
  const int tid = threadIdx.x;
  tri128_load_lower(S, A, lda, r, /*src_upper=*/true, /*unit=*/false, tid);
  __syncthreads();
  if (dbg) tk[1] = clock64();
  const int fail = potrf128_lower_smem(S, tmp, r, tid);
  if (fail != 0) {  // leading minor off+fail is not positive definite: flag it, later kernels skip
    if (tid == 0) *info = off + fail;
    return;
  }
  if (dbg) tk[2] = clock64();
  tri128_store(S, A, lda, r, /*dst_upper=*/true, /*zero_other=*/false, tid);
  __syncthreads();
  if (dbg) tk[3] = clock64();
  trtri128_lower_smem(S, tmp, /*unit=*/false, tid, /*have_rdiag=*/true);
  if (dbg) tk[4] = clock64();
  tri128_store(S, uinv, IB, IB, /*dst_upper=*/true, /*zero_other=*/true, tid);
  if (dbg && tid == 0) {
    tk[5] = clock64();
    for (int t = 0; t < 5; t++) dbg[t] = tk[t + 1] - tk[t];
  }
}

static int potrf_rec(int64_t n, double* A, int64_t lda, int64_t off, double* uinv, int* info, cudaStream_t st) {
  if (n <= 0) return 0;
  if (n <= IB) {
    static long long* dbg = nullptr;
    static bool dbg_checked = false;
    if (!dbg_checked) { dbg_checked = true; if (getenv("LLACUDA_LEAF_DBG")) { cudaMalloc(&dbg, 64); cudaMemset(dbg, 0, 64); } }
    lla_dpotrf128_kernel<<<1, T128_THREADS, T128_SMEM, st>>>((int)n, A, lda, uinv, (int)off, info, dbg);
    if (dbg) {
      long long h[5];
      cudaMemcpy(h, dbg, 40, cudaMemcpyDeviceToHost);
      fprintf(stderr, "potrf128 phases (cycles): load=%lld factor=%lld store=%lld trtri=%lld store-inv=%lld\n", h[0], h[1], h[2], h[3], h[4]);
    }
    count_launch();
    LLA_CUDA(cudaGetLastError());
    return 0;
  }
  static int64_t rl_max = -1;   // up to this order: right-looking over the 128-column leaves (3 launches per leaf instead of 4)
  if (rl_max < 0) { const char* e = getenv("LLACUDA_CHOL_RL_MAX"); rl_max = e ? atol(e) : 1024; }
  if (n <= rl_max) {
    for (int64_t j = 0; j < n; j += IB) {
      const int64_t jb = (n - j) < IB ? (n - j) : IB, rest = n - j - jb;
      double* Ajj = A + j + j * lda;
      double* Arow = Ajj + jb * lda;
      double* uj = uinv + (j / IB) * IB * IB;
      LLA_TRY(potrf_rec(jb, Ajj, lda, off + j, uj, info, st));
      if (rest <= 0) break;
      LLA_TRY(dtrsm_left_inv_dev(true, 1, jb, rest, Ajj, lda, uj, Arow, lda, st, info));
      LLA_TRY(dgemm_dev(1, 0, rest, rest, jb, -1.0, Arow, lda, Arow, lda, 1.0, Arow + jb, lda, TRI_UPPER, st, info));
    }
    return 0;
  }
  int64_t n1 = ((n / 2 + IB - 1) / IB) * IB;
  int64_t n2 = n - n1;
  double* A12 = A + n1 * lda;
  double* A22 = A12 + n1;
  LLA_TRY(potrf_rec(n1, A, lda, off, uinv, info, st));
  LLA_TRY(dtrsm_left_inv_dev(true, 1, n1, n2, A, lda, uinv, A12, lda, st, info));
  LLA_TRY(dgemm_dev(1, 0, n2, n2, n1, -1.0, A12, lda, A12, lda, 1.0, A22, lda, TRI_UPPER, st, info));
  LLA_TRY(potrf_rec(n2, A22, lda, off + n1, uinv + (n1 / IB) * IB * IB, info, st));
  return 0;
}

// Diagonal block of a blocked step, right-looking over its 128-column leaves on `sP`, together with the `ext` columns of its
// block row that lie above the NEXT diagonal block: each leaf's share of that solve (one product with the leaf's inverse,
// one rank-128 update of the rows below) trails the chain by one leaf on `sX`, so that when the last leaf is done only one
// product is left between this panel and the update of the next diagonal block.  `ev` signals sP -> sX.
static int potrf_panel_ext(int64_t n, double* A, int64_t lda, int64_t off, double* uinv, int64_t ext, int* info,
                           cudaStream_t sP, cudaStream_t sX, cudaEvent_t ev) {
  double* E = A + n * lda;                       // the extension: rows [0, n), columns [n, n + ext)
  for (int64_t j = 0; j < n; j += IB) {
    const int64_t jb = (n - j) < IB ? (n - j) : IB, rest = n - j - jb;
    double* Ajj = A + j + j * lda;
    double* Arow = Ajj + jb * lda;
    double* uj = uinv + (j / IB) * IB * IB;
    LLA_TRY(potrf_rec(jb, Ajj, lda, off + j, uj, info, sP));
    if (rest > 0) LLA_TRY(dtrsm_left_inv_dev(true, 1, jb, rest, Ajj, lda, uj, Arow, lda, sP, info));
    if (sX != sP) {
      LLA_CUDA(cudaEventRecord(ev, sP));
      LLA_CUDA(cudaStreamWaitEvent(sX, ev, 0));
    }
    if (rest > 0) LLA_TRY(dgemm_dev(1, 0, rest, rest, jb, -1.0, Arow, lda, Arow, lda, 1.0, Arow + jb, lda, TRI_UPPER, sP, info));
    LLA_TRY(dtrsm_left_inv_dev(true, 1, jb, ext, Ajj, lda, uj, E + j, lda, sX, info));
    if (rest > 0) LLA_TRY(dgemm_dev(1, 0, rest, ext, jb, -1.0, Arow, lda, E + j, lda, 1.0, E + j + jb, lda, TRI_FULL, sX, info));
  }
  return 0;
}

// Blocked right-looking driver with one panel of look-ahead.  The critical chain of step k -- factor the diagonal block
// (recursively, down to the shared-memory leaves), solve for the part of its block row above the NEXT diagonal block,
// update that block -- runs on two high-priority streams of its own; the rest of the block row goes to a third
// high-priority stream and the trailing updates to the main stream, which fill the machine meanwhile:
//   sP : [diag(k-1)] potrf(k), leaf by leaf -> pot
//   sA : [row1(k-1)] solve U(k, k+1), one leaf behind sP -> t1 | [b(k-1)] A(k+1, k+1) -= U(k, k+1)^T U(k, k+1) -> diag
//   sR : [pot, row(k-1)] solve U(k, k+2:) | [t1] -> panel, rows final                          (beside b(k-1) of s1)
//   s1 : ... b(k-1) | [panel] A(k+1, k+2) -= ... -> row1 | A(k+1, k+3:) -= ... -> row | b(k): rows >= k+2 -> b
static int potrf_blocked(int64_t n, double* A, int64_t lda, double* uinv, int* info, cudaStream_t s1) {
  Context* c = ctx();
  // measured on B200 (profiles/r01_cholesky_nb.txt): n = 16384: NB 256 / 512 / 1024 / 2048 -> 101 / 78.5 / 68.6 / 68.4 ms;
  // n = 8192: 17.3 / 15.2 / 15.9 / 18.4 ms
  static int64_t NB_env = -1;
  if (NB_env < 0) { const char* e = getenv("LLACUDA_CHOL_NB"); NB_env = e ? atoi(e) : 0; }
  const int64_t NB = NB_env > 0 ? NB_env : (n >= 12288 ? 1024 : 512);
  const RangeReadyHook* ready = range_ready_hook();   // rows [lo, hi) of the upper triangle
  if (n <= 2 * NB) {
    if (ready) LLA_TRY(ready->fn(ready->user, 0, n, s1));
    return potrf_rec(n, A, lda, 0, uinv, info, s1);
  }
  const bool lookahead = getenv("LLACUDA_NO_LOOKAHEAD") == nullptr;
  cudaStream_t sP = lookahead ? c->aux : s1;
  cudaStream_t sA = lookahead ? c->aux3 : s1;
  cudaStream_t sR = lookahead ? c->aux2 : s1;
  cudaStream_t s0 = s1;   // the caller's stream
  cudaEvent_t ev_row1 = c->ev[0], ev_panel = c->ev[1], ev_start = c->ev[2], ev_diag = c->ev[3], ev_pot = c->ev[4],
              ev_t1 = c->ev[5], ev_b = c->ev[6], ev_row = c->ev[7], ev_leaf = c->ev[8];
  LLA_CUDA(cudaEventRecord(ev_start, s1));
  LLA_CUDA(cudaStreamWaitEvent(sP, ev_start, 0));
  if (sA != s1) LLA_CUDA(cudaStreamWaitEvent(sA, ev_start, 0));
  if (sR != s1) LLA_CUDA(cudaStreamWaitEvent(sR, ev_start, 0));
  // Optional (LLACUDA_CHOL_PART_N = remaining order at which to switch): the chain moves to its own SM partition, the
  // updates and the block-row solves to the rest.  Off by default: since the chain trails its block-row solve leaf by
  // leaf it is short enough that giving up 16-24 SMs of update throughput costs more than the isolation gains
  // (n = 16384: 54.3 ms without, 54.8-55.1 ms with; profiles/r02_cholesky_chain.md).
  static int64_t part_n = -1;
  if (part_n < 0) { const char* e = getenv("LLACUDA_CHOL_PART_N"); part_n = e ? atol(e) : 0; }
  const SmPartition* part = (part_n > 0 && lookahead) ? sm_partition() : nullptr;
  bool part_on = false;
  bool row1_pending = false, row_pending = false, b_pending = false;
  for (int64_t k = 0; k < n; k += NB) {
    const int64_t kb = (n - k) < NB ? (n - k) : NB;
    const int64_t n2 = n - k - kb;
    double* Akk = A + k + k * lda;
    double* Arow = Akk + kb * lda;             // block row right of the diagonal block
    double* uk = uinv + (k / IB) * IB * IB;
    if (part && !part_on && k > 0 && (n - k) <= part_n) {
      // only the update stream of the partition follows everything queued so far; the streams of the chain and of the
      // block-row solve follow their predecessors (their inputs from the update stream arrive by the per-step events),
      // so the switch costs no look-ahead
      LLA_CUDA(cudaEventRecord(ev_start, s1));
      LLA_CUDA(cudaEventRecord(ev_pot, sP));
      LLA_CUDA(cudaEventRecord(ev_t1, sR));
      LLA_CUDA(cudaEventRecord(ev_leaf, sA));
      LLA_CUDA(cudaStreamWaitEvent(part->bulk, ev_start, 0));
      for (cudaStream_t t : {part->bulk, part->bulk2, part->bulk3, part->panel}) {
        LLA_CUDA(cudaStreamWaitEvent(t, ev_pot, 0));
        LLA_CUDA(cudaStreamWaitEvent(t, ev_t1, 0));
        LLA_CUDA(cudaStreamWaitEvent(t, ev_leaf, 0));
      }
      s1 = part->bulk; sP = part->panel; sA = part->bulk2; sR = part->bulk3;
      part_on = true;
      trace_mark(s1, "chol.partition", k, part->sm_panel);
    }
    // ---- the chain: diagonal block k and, trailing it leaf by leaf on sA, the block row above the next diagonal block
    trace_mark(sP, "chol.panel>", k, kb);
    if (ready && k == 0) {
      LLA_TRY(ready->fn(ready->user, 0, kb, sP));
      if (sA != sP) LLA_TRY(ready->fn(ready->user, 0, kb, sA));
    }
    if (n2 <= 0) {
      LLA_TRY(potrf_rec(kb, Akk, lda, k, uk, info, sP));
      LLA_CUDA(cudaEventRecord(ev_panel, sP));
      if (const RowsFinalHook* h = rows_final_hook()) LLA_TRY(h->fn(h->user, k, k + kb, sP, nullptr));
      break;
    }
    const int64_t kb_next = n2 < NB ? n2 : NB;
    const int64_t n3 = n2 - kb_next;           // columns right of the next diagonal block
    double* Anext = Arow + kb;                 // A[k+kb, k+kb]
    if (row1_pending) LLA_CUDA(cudaStreamWaitEvent(sA, ev_row1, 0));
    LLA_TRY(potrf_panel_ext(kb, Akk, lda, k, uk, kb_next, info, sP, sA, ev_leaf));
    LLA_CUDA(cudaEventRecord(ev_pot, sP));
    LLA_CUDA(cudaEventRecord(ev_t1, sA));
    // the next diagonal block
    if (b_pending) LLA_CUDA(cudaStreamWaitEvent(sA, ev_b, 0));
    if (ready && k == 0) LLA_TRY(ready->fn(ready->user, kb, kb + kb_next, sA));
    trace_mark(sA, "chol.diag>", k, kb_next);
    LLA_TRY(dgemm_dev(1, 0, kb_next, kb_next, kb, -1.0, Arow, lda, Arow, lda, 1.0, Anext, lda, TRI_UPPER, sA, info));
    LLA_CUDA(cudaEventRecord(ev_diag, sA));
    if (sA != sP) LLA_CUDA(cudaStreamWaitEvent(sP, ev_diag, 0));
    trace_mark(sA, "chol.diag<", k, kb_next);
    // ---- the rest of the block row on sR, beside whatever s1 is still doing for step k - 1
    LLA_CUDA(cudaStreamWaitEvent(sR, ev_pot, 0));
    if (row_pending) LLA_CUDA(cudaStreamWaitEvent(sR, ev_row, 0));
    trace_mark(sR, "chol.trsm2>", k, n3);
    if (n3 > 0) LLA_TRY(dtrsm_left_inv_dev(true, 1, kb, n3, Akk, lda, uk, Arow + kb_next * lda, lda, sR, info));
    LLA_CUDA(cudaStreamWaitEvent(sR, ev_t1, 0));
    LLA_CUDA(cudaEventRecord(ev_panel, sR));
    trace_mark(sR, "chol.panel<", k, kb);
    if (const RowsFinalHook* h = rows_final_hook()) LLA_TRY(h->fn(h->user, k, k + kb, sR, nullptr));  // block row k of U is final
    // ---- the trailing update of step k on s1, from top to bottom
    LLA_CUDA(cudaStreamWaitEvent(s1, ev_panel, 0));
    row1_pending = row_pending = b_pending = false;
    if (n3 > 0) {
      // block row k+1: first the columns above diagonal block k+2 (the chain needs them next), then the rest
      trace_mark(s1, "chol.updA>", k, n3);
      if (ready && k == 0) LLA_TRY(ready->fn(ready->user, kb, kb + kb_next, s1));
      const int64_t w1 = n3 < NB ? n3 : NB;
      LLA_TRY(dgemm_dev(1, 0, kb_next, w1, kb, -1.0, Arow, lda, Arow + kb_next * lda, lda, 1.0, Anext + kb_next * lda, lda,
                        TRI_FULL, s1, info));
      LLA_CUDA(cudaEventRecord(ev_row1, s1));
      row1_pending = true;
      if (n3 > w1) {
        LLA_TRY(dgemm_dev(1, 0, kb_next, n3 - w1, kb, -1.0, Arow, lda, Arow + (kb_next + w1) * lda, lda, 1.0,
                          Anext + (kb_next + w1) * lda, lda, TRI_FULL, s1, info));
        LLA_CUDA(cudaEventRecord(ev_row, s1));
        row_pending = true;
      }
      trace_mark(s1, "chol.updA<", k, n3);
      // rows [k+kb+kb_next, n)
      if (ready && k == 0) {
        // first step: the update is separable by block row, so each block row starts as soon as it is on the device
        const int64_t t0 = kb + kb_next;                     // first row / column of this part of the trailing matrix
        for (int64_t r0 = t0; r0 < n; r0 += READY_BLOCK) {
          const int64_t r1 = r0 + READY_BLOCK < n ? r0 + READY_BLOCK : n;
          LLA_TRY(ready->fn(ready->user, r0, r1, s1));
          const double* P = Arow + (r0 - kb) * lda;          // U(k, r0:)
          LLA_TRY(dgemm_dev(1, 0, r1 - r0, n - r0, kb, -1.0, P, lda, P, lda, 1.0, A + r0 + r0 * lda, lda, TRI_UPPER, s1, info));
        }
      } else {
        const double* P = Arow + kb_next * lda;
        LLA_TRY(dgemm_dev(1, 0, n3, n3, kb, -1.0, P, lda, P, lda, 1.0, Anext + kb_next + kb_next * lda, lda, TRI_UPPER, s1, info));
      }
      LLA_CUDA(cudaEventRecord(ev_b, s1));
      b_pending = true;
      trace_mark(s1, "chol.updB<", k, n3);
    }
  }
  LLA_CUDA(cudaStreamWaitEvent(s1, ev_panel, 0));
  if (part_on) {   // back to the caller's stream
    LLA_CUDA(cudaEventRecord(ev_start, s1));
    LLA_CUDA(cudaStreamWaitEvent(s0, ev_start, 0));
  }
  return 0;
}

int dpotrf_dev(bool upper, int64_t n, double* A, int64_t lda, int* info_dev, cudaStream_t st) {
  if (!ctx()) return LLACUDA_ERR_NO_DEVICE;
  LLA_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), st));
  if (n <= 0) return 0;
  static DeviceOnce once;
  if (once.first()) {
    LLA_CUDA(cudaFuncSetAttribute(lla_dpotrf128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T128_SMEM));
  }
  Arena ar;
  LLA_TRY(ar.reserve(tinv_bytes(n), st));
  double* uinv = (double*)ar.take(tinv_bytes(n));
  if (!upper) LLA_TRY(dtranspose_inplace_dev(n, A, lda, st));
  LLA_TRY(potrf_blocked(n, A, lda, uinv, info_dev, st));
  if (!upper) LLA_TRY(dtranspose_inplace_dev(n, A, lda, st));
  return 0;
}

int dpotrs_dev(bool upper, int64_t n, int64_t nrhs, const double* A, int64_t lda, double* B, int64_t ldb,
               cudaStream_t st, const int* abort_flag) {
  if (n <= 0 || nrhs <= 0) return 0;
  // inverses of the 128 x 128 diagonal blocks of the factor (one batched launch), then both
  // triangular solves are GEMMs only
  Arena ar;
  LLA_TRY(ar.reserve(tinv_bytes(n), st));
  double* tinv = (double*)ar.take(tinv_bytes(n));
  LLA_TRY(dtrtri_blocks_dev(upper, false, n, A, lda, tinv, st, abort_flag));
  if (upper) {  // A = U^T U :  U^T y = b ; U x = y
    LLA_TRY(dtrsm_left_inv_dev(true, 1, n, nrhs, A, lda, tinv, B, ldb, st, abort_flag));
    LLA_TRY(dtrsm_left_inv_dev(true, 0, n, nrhs, A, lda, tinv, B, ldb, st, abort_flag));
  } else {      // A = L L^T :  L y = b ; L^T x = y
    LLA_TRY(dtrsm_left_inv_dev(false, 0, n, nrhs, A, lda, tinv, B, ldb, st, abort_flag));
    LLA_TRY(dtrsm_left_inv_dev(false, 1, n, nrhs, A, lda, tinv, B, ldb, st, abort_flag));
  }
  return 0;
}

}  // namespace llacuda
