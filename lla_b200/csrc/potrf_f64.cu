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

// Blocked right-looking driver with one panel of look-ahead.  Step k factors the diagonal block
// (recursively, down to the shared-memory leaves) and solves for its block row on the high-priority
// panel stream while the previous step's trailing update is still running on the main stream: the
// update of step k-1 is split into (a) the block row that panel k needs and (b) everything below.
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
  cudaStream_t s2 = getenv("LLACUDA_NO_LOOKAHEAD") ? s1 : c->aux;
  cudaStream_t s0 = s1;   // the caller's stream
  cudaEvent_t ev_row = c->ev[0], ev_panel = c->ev[1], ev_start = c->ev[2], ev_diag = c->ev[3], ev_pot = c->ev[4];
  LLA_CUDA(cudaEventRecord(ev_start, s1));
  LLA_CUDA(cudaStreamWaitEvent(s2, ev_start, 0));
  // Chain-bound phase (the update of a step is shorter than the panel): the panel chain moves to its own SM partition, the
  // updates and the block-row solve to the rest, so that no kernel of the chain waits for a GEMM CTA to leave an SM.
  static int64_t part_n = -1;
  if (part_n < 0) { const char* e = getenv("LLACUDA_CHOL_PART_N"); part_n = e ? atol(e) : 8192; }
  const SmPartition* part = (part_n > 0 && s2 != s1) ? sm_partition() : nullptr;
  bool part_on = false;
  cudaStream_t st_trsm = s2;
  bool rest_pending = false;  // part (a2) of the previous step: the block row right of the next diagonal block
  for (int64_t k = 0; k < n; k += NB) {
    const int64_t kb = (n - k) < NB ? (n - k) : NB;
    const int64_t n2 = n - k - kb;
    double* Akk = A + k + k * lda;
    double* Arow = Akk + kb * lda;             // block row right of the diagonal block
    double* uk = uinv + (k / IB) * IB * IB;
    if (part && !part_on && k > 0 && (n - k) <= part_n) {
      // the update streams of the partition follow everything queued so far; the panel stream only what panel k needs
      // (the previous panel and the update of its diagonal block), so the switch costs no look-ahead
      LLA_CUDA(cudaEventRecord(ev_start, s1));
      LLA_CUDA(cudaEventRecord(ev_pot, s2));
      for (cudaStream_t t : {part->bulk, part->bulk2}) {
        LLA_CUDA(cudaStreamWaitEvent(t, ev_start, 0));
        LLA_CUDA(cudaStreamWaitEvent(t, ev_pot, 0));
      }
      LLA_CUDA(cudaStreamWaitEvent(part->panel, ev_pot, 0));
      LLA_CUDA(cudaStreamWaitEvent(part->panel, ev_diag, 0));
      s1 = part->bulk; s2 = part->panel; st_trsm = part->bulk2;
      part_on = true;
      trace_mark(s1, "chol.partition", k, part->sm_panel);
    }
    // ---- panel k on s2: the diagonal block needs only part (a1) of the previous update, the block row also (a2)
    trace_mark(s2, "chol.panel>", k, kb);
    if (ready && k == 0) LLA_TRY(ready->fn(ready->user, 0, kb, s2));
    LLA_TRY(potrf_rec(kb, Akk, lda, k, uk, info, s2));
    trace_mark(s2, "chol.trsm>", k, n2);
    if (part_on) {   // the block-row solve is a wide GEMM-shaped job: on the large partition
      LLA_CUDA(cudaEventRecord(ev_pot, s2));
      LLA_CUDA(cudaStreamWaitEvent(st_trsm, ev_pot, 0));
    }
    if (rest_pending) LLA_CUDA(cudaStreamWaitEvent(st_trsm, ev_row, 0));
    if (n2 > 0) LLA_TRY(dtrsm_left_inv_dev(true, 1, kb, n2, Akk, lda, uk, Arow, lda, st_trsm, info));
    trace_mark(st_trsm, "chol.panel<", k, kb);
    LLA_CUDA(cudaEventRecord(ev_panel, st_trsm));
    if (const RowsFinalHook* h = rows_final_hook()) LLA_TRY(h->fn(h->user, k, k + kb, st_trsm, nullptr));  // block row k of U is final
    if (n2 <= 0) break;
    // ---- trailing update of step k on s1
    LLA_CUDA(cudaStreamWaitEvent(s1, ev_panel, 0));
    const int64_t kb_next = n2 < NB ? n2 : NB;
    double* Anext = Arow + kb;                   // A[k+kb, k+kb]
    // (a1) the next diagonal block, (a2) the rest of the next panel's block row: rows [k+kb, k+kb+kb_next), columns [k+kb, n)
    trace_mark(s1, "chol.updA>", k, n2);
    if (ready && k == 0) LLA_TRY(ready->fn(ready->user, kb, kb + kb_next, s1));
    LLA_TRY(dgemm_dev(1, 0, kb_next, kb_next, kb, -1.0, Arow, lda, Arow, lda, 1.0, Anext, lda, TRI_UPPER, s1, info));
    LLA_CUDA(cudaEventRecord(ev_diag, s1));
    LLA_CUDA(cudaStreamWaitEvent(s2, ev_diag, 0));
    rest_pending = n2 > kb_next;
    if (rest_pending) {
      LLA_TRY(dgemm_dev(1, 0, kb_next, n2 - kb_next, kb, -1.0, Arow, lda, Arow + kb_next * lda, lda, 1.0,
                        Anext + kb_next * lda, lda, TRI_FULL, s1, info));
      LLA_CUDA(cudaEventRecord(ev_row, s1));
    }
    trace_mark(s1, "chol.updA<", k, n2);
    // (b) the rest: rows [k+kb+kb_next, n)
    const int64_t n3 = n2 - kb_next;
    if (n3 > 0) {
      // k-split of the big update: shorter-lived CTAs, so that the kernels of the panel chain (high-priority
      // stream) find a free SM sooner (a 128x128 tile with k = 1024 occupies its SM for ~150 us)
      static int64_t ks_env = -1;
      if (ks_env < 0) { const char* e = getenv("LLACUDA_CHOL_KSPLIT"); ks_env = e ? atoi(e) : 0; }
      const int64_t ks = ks_env > 0 ? ks_env : kb;
      if (ready && k == 0) {
        // first step: the update is separable by block row, so each block row starts as soon as it is on the device
        const int64_t t0 = kb + kb_next;                     // first row / column of this part of the trailing matrix
        for (int64_t r0 = t0; r0 < n; r0 += READY_BLOCK) {
          const int64_t r1 = r0 + READY_BLOCK < n ? r0 + READY_BLOCK : n;
          LLA_TRY(ready->fn(ready->user, r0, r1, s1));
          const double* P = Arow + (r0 - kb) * lda;          // U(k, r0:)
          LLA_TRY(dgemm_dev(1, 0, r1 - r0, n - r0, kb, -1.0, P, lda, P, lda, 1.0, A + r0 + r0 * lda, lda, TRI_UPPER, s1, info));
        }
      } else
      for (int64_t kk = 0; kk < kb; kk += ks) {
        const int64_t kw = (kb - kk) < ks ? (kb - kk) : ks;
        const double* P = Arow + kk + kb_next * lda;
        LLA_TRY(dgemm_dev(1, 0, n3, n3, kw, -1.0, P, lda, P, lda, 1.0, Anext + kb_next + kb_next * lda, lda, TRI_UPPER, s1, info,
                          getenv("LLACUDA_CHOL_RESERVE") ? GEMM_RESERVE_SMS : GEMM_AUTO));
      }
    }
    trace_mark(s1, "chol.updB<", k, n3);
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
