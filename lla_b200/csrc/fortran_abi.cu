// lla_b200/csrc/fortran_abi.cu -- the Fortran-symbol drop-in layer (host pointers in, host
// pointers out).  These are the symbols LLA's blas-call / lapack-call macros resolve
// (reference src/fortran-call.lisp:405-439); argument lists as issued by the call sites listed
// in include/llacuda.h.  Everything here is synchronous: LLA's buffers are only valid until the
// call returns (reference src/pinned-array.lisp:105-141).
#include "internal.h"
#include "staging.h"
#include "../../include/llacuda.h"

#include <cstdio>
#include <vector>

using namespace llacuda;

#define LLA_EXPORT extern "C"

static int opcode(char c) {
  switch (c) {
    case 'N': case 'n': return 0;
    case 'T': case 't': return 1;
    case 'C': case 'c': return 2;
    default: return -1;
  }
}

static void blas_arg_error(const char* routine, int pos) {
  // netlib BLAS would call XERBLA and stop; LLA pre-validates with asserts
  // (reference src/blas.lisp:40-53), so this is unreachable from LLA.  Record and return.
  char msg[96];
  snprintf(msg, sizeof(msg), "%s: parameter number %d had an illegal value", routine, pos);
  set_error(__FILE__, __LINE__, msg, pos);
  fprintf(stderr, " ** llacuda: %s\n", msg);
}

// ------------------------------------------------------------------ GEMM
// C = alpha op(A) op(B) + beta C with HOST operands.  Large products are pipelined over column
// chunks of op(B) / C on three streams: while the DMMA GEMM of chunk j runs, chunk j+1 of B (and of C
// when beta != 0) is uploaded and chunk j-1 of C is downloaded, so only the upload of A and the
// first / last chunk transfers are exposed.
static int dgemm_host(int oa, int ob, int64_t m, int64_t n, int64_t k, double alpha, const double* a,
                      int64_t lda, const double* b, int64_t ldb, double beta, double* c, int64_t ldc) {
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  if (m == 0 || n == 0) return 0;
  cudaStream_t st = cx->stream;
  PhaseTimer tm(st);
  tm.mark(0);
  const bool need_ab = (k > 0 && alpha != 0.0);
  Staged sa, sb, sc;
  int64_t ar = oa ? k : m, ac = oa ? m : k;
  int64_t br = ob ? n : k, bc = ob ? k : n;
  if (need_ab) {
    LLA_TRY(stage_alloc(sa, ar, ac, sizeof(double)));
    LLA_TRY(stage_alloc(sb, br, bc, sizeof(double)));
  }
  LLA_TRY(stage_alloc(sc, m, n, sizeof(double)));
  const double work = 2.0 * (double)m * (double)n * (double)(k > 0 ? k : 1);
  const int chunks = (need_ab && !ob && work > 4e10 && n >= 2048) ? 8 : 1;   // op(B) = B: its columns are C's columns
  if (chunks == 1) {
    if (need_ab) {
      LLA_TRY(stage_upload(sa, a, lda, st));
      LLA_TRY(stage_upload(sb, b, ldb, st));
    }
    if (beta != 0.0) LLA_TRY(stage_upload(sc, c, ldc, st));
    tm.mark(1);
    LLA_TRY(dgemm_dev(oa, ob, m, n, k, alpha, sa.ptr<double>(), sa.ld, sb.ptr<double>(), sb.ld, beta,
                      sc.ptr<double>(), sc.ld, TRI_FULL, st));
    tm.mark(2);
    LLA_TRY(stage_download(sc, c, ldc, st));
    tm.mark(3);
    LLA_TRY(stage_sync(st));
    tm.finish();
    return 0;
  }
  // 2-D pipeline: op(A) = A is cut into RC row blocks, op(B) = B / C into `chunks` column blocks; block (r, c) of C
  // is computed as soon as A_r and B_c (and C_rc when beta != 0) have arrived and goes back at once, so only A_0 + B_0
  // on the way in and the last block on the way out are exposed.  The operands of block i+1 are queued right after
  // the GEMM of block i: with pageable host memory the calling thread stages them through the pinned ring while
  // that GEMM runs.  Each GEMM keeps >= ~7 waves of 128x64 tiles so that the per-launch tail stays around 1 %.
  cudaStream_t sh = cx->h2d, sd = cx->d2h;
  const int RC = (!oa && m >= 8192) ? 2 : 1;
  const int64_t rstep = RC == 1 ? m : ((m + RC - 1) / RC + 127) / 128 * 128;
  const int64_t step = ((n + chunks - 1) / chunks + 63) / 64 * 64;
  constexpr int MAXB = 2 * 8;
  struct Events {
    cudaEvent_t in[MAXB] = {}, done[MAXB] = {}, start = nullptr;
    ~Events() {
      for (auto e : in) if (e) cudaEventDestroy(e);
      for (auto e : done) if (e) cudaEventDestroy(e);
      if (start) cudaEventDestroy(start);
    }
  } ev;
  const int nblk = RC * chunks;
  LLA_CUDA(cudaEventCreateWithFlags(&ev.start, cudaEventDisableTiming));
  for (int j = 0; j < nblk; j++) {
    LLA_CUDA(cudaEventCreateWithFlags(&ev.in[j], cudaEventDisableTiming));
    LLA_CUDA(cudaEventCreateWithFlags(&ev.done[j], cudaEventDisableTiming));
  }
  LLA_CUDA(cudaEventRecord(ev.start, st));
  LLA_CUDA(cudaStreamWaitEvent(sh, ev.start, 0));
  LLA_CUDA(cudaStreamWaitEvent(sd, ev.start, 0));
  auto rows_of = [&](int r, int64_t& r0, int64_t& r1) { r0 = (int64_t)r * rstep; r1 = (r0 + rstep < m) ? r0 + rstep : m; };
  auto cols_of = [&](int j, int64_t& c0, int64_t& c1) { c0 = (int64_t)j * step; c1 = (c0 + step < n) ? c0 + step : n; };
  auto fetch = [&](int idx) -> int {   // operands that block idx needs and no earlier block has brought
    const int r = idx / chunks, j = idx % chunks;
    int64_t r0, r1, c0, c1;
    rows_of(r, r0, r1); cols_of(j, c0, c1);
    if (j == 0 && r0 < m) {
      if (RC == 1) LLA_TRY(stage_upload(sa, a, lda, sh));
      else LLA_TRY(stage_upload_block(sa, a, lda, r0, r1, 0, k, sh));
    }
    if (c0 < n && r0 < m) {
      if (r == 0) LLA_TRY(stage_upload_cols(sb, b, ldb, c0, c1, sh));
      if (beta != 0.0) LLA_TRY(stage_upload_block(sc, c, ldc, r0, r1, c0, c1, sh));
    }
    LLA_CUDA(cudaEventRecord(ev.in[idx], sh));
    return 0;
  };
  LLA_TRY(fetch(0));
  tm.mark(1);
  for (int idx = 0; idx < nblk; idx++) {
    const int r = idx / chunks, j = idx % chunks;
    int64_t r0, r1, c0, c1;
    rows_of(r, r0, r1); cols_of(j, c0, c1);
    const bool live = r0 < m && c0 < n;
    if (live) {
      LLA_CUDA(cudaStreamWaitEvent(st, ev.in[idx], 0));
      LLA_TRY(dgemm_dev(oa, ob, r1 - r0, c1 - c0, k, alpha, sa.ptr<double>() + r0, sa.ld, sb.ptr<double>() + c0 * sb.ld, sb.ld,
                        beta, sc.ptr<double>() + r0 + c0 * sc.ld, sc.ld, TRI_FULL, st));
      LLA_CUDA(cudaEventRecord(ev.done[idx], st));
    }
    if (idx + 1 < nblk) LLA_TRY(fetch(idx + 1));
    if (live) {
      LLA_CUDA(cudaStreamWaitEvent(sd, ev.done[idx], 0));
      LLA_TRY(stage_download_block(sc, c, ldc, r0, r1, c0, c1, sd));
    }
  }
  tm.mark(2);
  LLA_CUDA(cudaEventRecord(ev.start, sd));
  LLA_CUDA(cudaStreamWaitEvent(st, ev.start, 0));
  tm.mark(3);
  LLA_TRY(stage_sync(st));
  LLA_CUDA(cudaStreamSynchronize(sd));
  tm.finish();
  return 0;
}

LLA_EXPORT void dgemm_(const char* transa, const char* transb, const fint* m_, const fint* n_, const fint* k_,
                       const double* alpha, const double* a, const fint* lda_, const double* b, const fint* ldb_,
                       const double* beta, double* c, const fint* ldc_) {
  ApiLock api_lock;
  int oa = opcode(*transa), ob = opcode(*transb);
  int64_t m = *m_, n = *n_, k = *k_, lda = *lda_, ldb = *ldb_, ldc = *ldc_;
  int64_t nrowa = oa ? k : m, nrowb = ob ? n : k;
  int bad = 0;
  if (oa < 0) bad = 1;
  else if (ob < 0) bad = 2;
  else if (m < 0) bad = 3;
  else if (n < 0) bad = 4;
  else if (k < 0) bad = 5;
  else if (lda < (nrowa > 1 ? nrowa : 1)) bad = 8;
  else if (ldb < (nrowb > 1 ? nrowb : 1)) bad = 10;
  else if (ldc < (m > 1 ? m : 1)) bad = 13;
  if (bad) { blas_arg_error("DGEMM ", bad); return; }
  llacuda_clear_error();
  const int64_t small = m < n ? (m < k ? m : k) : (n < k ? n : k);
  int r = dist_route(small) ? dist_dgemm(oa, ob, m, n, k, *alpha, a, lda, b, ldb, *beta, c, ldc)
                            : dgemm_host(oa, ob, m, n, k, *alpha, a, lda, b, ldb, *beta, c, ldc);
  if (r != 0) blas_failure("dgemm_", r, c, ldc, m, n, sizeof(double));
}

LLA_EXPORT void llacuda_dgemm(const char* transa, const char* transb, const fint* m, const fint* n, const fint* k,
                              const double* alpha, const double* a, const fint* lda, const double* b,
                              const fint* ldb, const double* beta, double* c, const fint* ldc) {
  ApiLock api_lock;
  dgemm_(transa, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}

// ------------------------------------------------------------------ LAPACK drop-ins
static inline int64_t max1(int64_t x) { return x > 1 ? x : 1; }

// Integer vectors crossing the Fortran ABI: the device works on int32; with -DLLACUDA_ILP64 the caller's cells are
// 64-bit and are converted through a 32-bit host vector (pivots are row indices < 2^31 for every matrix that fits).
struct IntOut {
  fint* dst; int64_t n; std::vector<int> tmp;
  IntOut(fint* d, int64_t count) : dst(d), n(count) { if (sizeof(fint) != sizeof(int)) tmp.resize((size_t)(count > 0 ? count : 1)); }
  int* host32() { return sizeof(fint) == sizeof(int) ? (int*)dst : tmp.data(); }
  void finish() { if (sizeof(fint) != sizeof(int)) for (int64_t i = 0; i < n; i++) dst[i] = (fint)tmp[(size_t)i]; }
};
struct IntIn {
  std::vector<int> tmp; const int* p;
  IntIn(const fint* src, int64_t count) {
    if (sizeof(fint) == sizeof(int)) { p = (const int*)src; return; }
    tmp.resize((size_t)(count > 0 ? count : 1));
    for (int64_t i = 0; i < count; i++) tmp[(size_t)i] = (int)src[i];
    p = tmp.data();
  }
  const int* host32() const { return p; }
};

// Finished block rows of a factor go back to the caller while the factorisation is still running
// (RowsFinalHook of the blocked drivers).  Page-locked callers: asynchronous copies on the d2h stream.  Pageable
// callers (LLA): the pinned ring's downloader thread (hostpipe.cu) drains them, so the calling thread keeps
// queueing the following block steps.
struct RowSender {
  const Staged* s;
  double* host;
  int64_t ldh;
  cudaStream_t sd;
  bool from_diagonal;     // Cholesky 'U': only columns >= r0 carry the factor
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int64_t sent_upto = 0;  // rows [0, sent_upto) are on their way
  RowsFinalHook hook;
  bool armed = false;
  int arm(const Staged& st, double* h, int64_t ld, cudaStream_t d2h, bool diag) {
    s = &st; host = h; ldh = ld; sd = d2h; from_diagonal = diag;
    for (auto& e : ev) LLA_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    hook.fn = &RowSender::rows_final; hook.user = this;
    set_rows_final_hook(&hook);
    armed = true;
    return 0;
  }
  static int rows_final(void* user, int64_t r0, int64_t r1, cudaStream_t a, cudaStream_t b) {
    RowSender* self = (RowSender*)user;
    // rows are reported in order (and on the 512 grid of the triangle upload); anything else is left to the final copy
    if (r0 != self->sent_upto || (self->from_diagonal && r0 % 512 != 0)) return 0;
    LLA_CUDA(cudaEventRecord(self->ev[0], a));
    LLA_CUDA(cudaStreamWaitEvent(self->sd, self->ev[0], 0));
    if (b) {
      LLA_CUDA(cudaEventRecord(self->ev[1], b));
      LLA_CUDA(cudaStreamWaitEvent(self->sd, self->ev[1], 0));
    }
    if (self->from_diagonal) {
      // the device holds exactly the trapezoids stage_upload_tri sent (512-column chunks, rows up to the chunk's
      // end): walk the block row in the same 512 grid so that no byte outside them is written back
      for (int64_t a = r0; a < r1; a += 512)
        LLA_TRY(stage_download_rows(*self->s, self->host, self->ldh, a, a + 512 < r1 ? a + 512 : r1, a, self->sd));
    } else {
      LLA_TRY(stage_download_rows(*self->s, self->host, self->ldh, r0, r1, 0, self->sd));
    }
    self->sent_upto = r1;
    return 0;
  }
  void disarm() { if (armed) set_rows_final_hook(nullptr); }
  ~RowSender() {
    disarm();
    for (auto& e : ev) if (e) cudaEventDestroy(e);
  }
};

// The way in, block by block behind the drivers' RangeReadyHook: column blocks for LU, block rows of the upper triangle
// for Cholesky.  Page-locked callers: every block is queued at once on the h2d stream (asynchronous), the hook only
// makes the consumer stream wait for the blocks it is about to touch.  Pageable callers (LLA): a block is staged through
// the pinned ring when the driver first asks for it -- the calling thread then stages block i + 1 while the GPU
// already works on block i.
struct BlockUploader {
  Staged* s = nullptr;
  const double* host = nullptr;
  int64_t ldh = 0, chunk = 1024, total = 0;
  bool by_rows = false;            // Cholesky 'U': rows [r0, r1) x columns [r0, n)
  cudaStream_t sh = nullptr;
  std::vector<cudaEvent_t> ev;
  std::vector<char> sent;
  RangeReadyHook hook;
  bool armed = false;
  int send(int64_t i) {
    if (sent[(size_t)i]) return 0;
    const int64_t a = i * chunk, b = a + chunk < total ? a + chunk : total;
    if (by_rows) LLA_TRY(stage_upload_block(*s, host, ldh, a, b, a, s->cols, sh));
    else LLA_TRY(stage_upload_cols(*s, host, ldh, a, b, sh));
    LLA_CUDA(cudaEventRecord(ev[(size_t)i], sh));
    sent[(size_t)i] = 1;
    return 0;
  }
  int arm(Staged& st, const double* h, int64_t ld, bool rows, cudaStream_t h2d, cudaStream_t after) {
    s = &st; host = h; ldh = ld; by_rows = rows; sh = h2d;
    total = rows ? st.rows : st.cols;
    const int64_t nchunk = (total + chunk - 1) / chunk;
    ev.assign((size_t)nchunk, nullptr);
    sent.assign((size_t)nchunk, 0);
    for (auto& e : ev) LLA_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    cudaEvent_t start;
    LLA_CUDA(cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
    LLA_CUDA(cudaEventRecord(start, after));       // the staging buffer may still be in use by earlier work on `after`
    LLA_CUDA(cudaStreamWaitEvent(sh, start, 0));
    cudaEventDestroy(start);
    if (host_is_pinned(h))
      for (int64_t i = 0; i < nchunk; i++) LLA_TRY(send(i));
    hook.fn = &BlockUploader::ready; hook.user = this;
    set_range_ready_hook(&hook);
    armed = true;
    return 0;
  }
  static int ready(void* user, int64_t lo, int64_t hi, cudaStream_t st) {
    BlockUploader* self = (BlockUploader*)user;
    if (hi > self->total) hi = self->total;
    for (int64_t i = lo / self->chunk; i * self->chunk < hi; i++) {
      LLA_TRY(self->send(i));
      LLA_CUDA(cudaStreamWaitEvent(st, self->ev[(size_t)i], 0));
    }
    return 0;
  }
  // after the driver: whatever it did not ask for (nothing, normally) and a join on `st`
  int finish(cudaStream_t st) {
    disarm();
    for (size_t i = 0; i < ev.size(); i++) {
      LLA_TRY(send((int64_t)i));
      LLA_CUDA(cudaStreamWaitEvent(st, ev[i], 0));
    }
    return 0;
  }
  void disarm() { if (armed) { set_range_ready_hook(nullptr); armed = false; } }
  ~BlockUploader() {
    disarm();
    for (auto e : ev) if (e) cudaEventDestroy(e);
  }
};

// reads the device INFO word back; a CUDA failure overrides it with the out-of-band code

static int dgetrf_host(int64_t m, int64_t n, double* a, int64_t lda, fint* ipiv, fint* info) {
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  cudaStream_t st = cx->stream;
  PhaseTimer tm(st);
  tm.mark(0);
  Staged sa;
  PoolBuf piv, inf;
  int64_t mn = m < n ? m : n;
  LLA_TRY(stage_alloc(sa, m, n, sizeof(double)));
  LLA_TRY(piv.alloc(sizeof(int) * (size_t)max1(mn)));
  LLA_TRY(inf.alloc(sizeof(int)));
  BlockUploader up;
  LLA_TRY(up.arm(sa, a, lda, false, cx->h2d, st));
  tm.mark(1);
  RowSender rs;
  LLA_TRY(rs.arm(sa, a, lda, cx->d2h, false));
  int rc = dgetrf_dev(m, n, sa.ptr<double>(), sa.ld, (int*)piv.p, (int*)inf.p, st);
  rs.disarm();
  up.disarm();
  LLA_TRY(rc);
  LLA_TRY(up.finish(st));
  tm.mark(2);
  LLA_TRY(stage_download_rows(sa, a, lda, rs.sent_upto, m, 0, st));
  LLA_CUDA(cudaStreamSynchronize(cx->d2h));
  IntOut pivots(ipiv, mn), inf_out(info, 1);
  if (mn > 0) LLA_CUDA(cudaMemcpyAsync(pivots.host32(), piv.p, sizeof(int) * (size_t)mn, cudaMemcpyDeviceToHost, st));
  LLA_CUDA(cudaMemcpyAsync(inf_out.host32(), inf.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  tm.mark(3);
  LLA_TRY(stage_sync(st));
  pivots.finish(); inf_out.finish();
  tm.finish();
  return 0;
}

/* reference call site: lu, src/linear-algebra.lisp:227-234 */
LLA_EXPORT void dgetrf_(const fint* m_, const fint* n_, double* a, const fint* lda_, fint* ipiv, fint* info) {
  ApiLock api_lock;
  int64_t m = *m_, n = *n_, lda = *lda_;
  *info = 0;
  if (m < 0) { *info = -1; return; }
  if (n < 0) { *info = -2; return; }
  if (lda < max1(m)) { *info = -4; return; }
  if (m == 0 || n == 0) return;
  if (m == n && dist_route(n)) {   // LLACUDA_GPUS=N: one factorisation over the GPUs of the box
    int inf = 0;
    int rc = dist_dgesv(n, -1, a, lda, ipiv, nullptr, 0, &inf);
    *info = (fint)(rc != 0 ? rc : inf);
    return;
  }
  int rc = dgetrf_host(m, n, a, lda, ipiv, info);
  if (rc != 0) { drain_streams(); stage_sync(nullptr); *info = (fint)rc; }
}

static int dgetrs_host(int op, int64_t n, int64_t nrhs, const double* a, int64_t lda, const fint* ipiv, double* b,
                       int64_t ldb) {
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  cudaStream_t st = cx->stream;
  PhaseTimer tm(st);
  tm.mark(0);
  Staged sa, sb;
  PoolBuf piv;
  LLA_TRY(stage_alloc(sa, n, n, sizeof(double)));
  LLA_TRY(stage_alloc(sb, n, nrhs, sizeof(double)));
  LLA_TRY(piv.alloc(sizeof(int) * (size_t)n));
  LLA_TRY(stage_upload(sa, a, lda, st));
  LLA_TRY(stage_upload(sb, b, ldb, st));
  IntIn pivots(ipiv, n);
  LLA_CUDA(cudaMemcpyAsync(piv.p, pivots.host32(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
  tm.mark(1);
  LLA_TRY(dgetrs_dev(op, n, nrhs, sa.ptr<double>(), sa.ld, (const int*)piv.p, sb.ptr<double>(), sb.ld, st));
  tm.mark(2);
  LLA_TRY(stage_download(sb, b, ldb, st));
  tm.mark(3);
  LLA_TRY(stage_sync(st));
  tm.finish();
  return 0;
}

/* reference call site: solve (lu), src/linear-algebra.lisp:269-276 */
LLA_EXPORT void dgetrs_(const char* trans, const fint* n_, const fint* nrhs_, const double* a, const fint* lda_,
                        const fint* ipiv, double* b, const fint* ldb_, fint* info) {
  ApiLock api_lock;
  int op = opcode(*trans);
  int64_t n = *n_, nrhs = *nrhs_, lda = *lda_, ldb = *ldb_;
  *info = 0;
  if (op < 0) { *info = -1; return; }
  if (n < 0) { *info = -2; return; }
  if (nrhs < 0) { *info = -3; return; }
  if (lda < max1(n)) { *info = -5; return; }
  if (ldb < max1(n)) { *info = -8; return; }
  if (n == 0 || nrhs == 0) return;
  int rc = dgetrs_host(op, n, nrhs, a, lda, ipiv, b, ldb);
  if (rc != 0) { drain_streams(); stage_sync(nullptr); *info = (fint)rc; }
}

static int dgesv_host(int64_t n, int64_t nrhs, double* a, int64_t lda, fint* ipiv, double* b, int64_t ldb,
                      fint* info) {
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  cudaStream_t st = cx->stream;
  PhaseTimer tm(st);
  tm.mark(0);
  Staged sa, sb;
  PoolBuf piv, inf;
  LLA_TRY(stage_alloc(sa, n, n, sizeof(double)));
  LLA_TRY(stage_alloc(sb, n, nrhs, sizeof(double)));
  LLA_TRY(piv.alloc(sizeof(int) * (size_t)max1(n)));
  LLA_TRY(inf.alloc(sizeof(int)));
  LLA_TRY(stage_upload(sb, b, ldb, st));
  BlockUploader up;
  LLA_TRY(up.arm(sa, a, lda, false, cx->h2d, st));
  tm.mark(1);
  RowSender rs;
  LLA_TRY(rs.arm(sa, a, lda, cx->d2h, false));
  // DGESV solves only when the factorisation reported INFO = 0: the forward substitution rides along with the
  // factorisation, the back substitution is gated by the INFO word and B is restored when it is not zero
  int rc = dgesv_dev(n, nrhs, sa.ptr<double>(), sa.ld, (int*)piv.p, sb.ptr<double>(), sb.ld, (int*)inf.p, st);
  rs.disarm();
  up.disarm();
  LLA_TRY(rc);
  LLA_TRY(up.finish(st));
  tm.mark(2);
  LLA_TRY(stage_download_rows(sa, a, lda, rs.sent_upto, n, 0, st));
  LLA_CUDA(cudaStreamSynchronize(cx->d2h));
  LLA_TRY(stage_download(sb, b, ldb, st));
  IntOut pivots(ipiv, n), inf_out(info, 1);
  LLA_CUDA(cudaMemcpyAsync(pivots.host32(), piv.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
  LLA_CUDA(cudaMemcpyAsync(inf_out.host32(), inf.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  tm.mark(3);
  LLA_TRY(stage_sync(st));
  pivots.finish(); inf_out.finish();
  tm.finish();
  return 0;
}

/* reference call site: solve (array array), src/linear-algebra.lisp:282-289 */
LLA_EXPORT void dgesv_(const fint* n_, const fint* nrhs_, double* a, const fint* lda_, fint* ipiv, double* b,
                       const fint* ldb_, fint* info) {
  ApiLock api_lock;
  int64_t n = *n_, nrhs = *nrhs_, lda = *lda_, ldb = *ldb_;
  *info = 0;
  if (n < 0) { *info = -1; return; }
  if (nrhs < 0) { *info = -2; return; }
  if (lda < max1(n)) { *info = -4; return; }
  if (ldb < max1(n)) { *info = -7; return; }
  if (n == 0) return;
  if (dist_route(n)) {
    int inf = 0;
    int rc = dist_dgesv(n, nrhs, a, lda, ipiv, b, ldb, &inf);
    *info = (fint)(rc != 0 ? rc : inf);
    return;
  }
  int rc = dgesv_host(n, nrhs, a, lda, ipiv, b, ldb, info);
  if (rc != 0) { drain_streams(); stage_sync(nullptr); *info = (fint)rc; }
}

static int dpotrf_host(bool upper, int64_t n, double* a, int64_t lda, fint* info) {
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  cudaStream_t st = cx->stream;
  PhaseTimer tm(st);
  tm.mark(0);
  Staged sa;
  PoolBuf inf;
  LLA_TRY(stage_alloc(sa, n, n, sizeof(double)));
  LLA_TRY(inf.alloc(sizeof(int)));
  // only the named triangle crosses PCIe, in both directions: the other one is neither read nor written
  // (LAPACK: "not referenced"), so the caller's bytes there stay as they are
  BlockUploader up;
  if (upper) LLA_TRY(up.arm(sa, a, lda, true, cx->h2d, st));   // block rows of the upper triangle, behind the driver's hook
  else LLA_TRY(stage_upload_tri(sa, a, lda, upper, st));
  tm.mark(1);
  RowSender rs;
  if (upper) LLA_TRY(rs.arm(sa, a, lda, cx->d2h, true));
  int rc = dpotrf_dev(upper, n, sa.ptr<double>(), sa.ld, (int*)inf.p, st);
  rs.disarm();
  up.disarm();
  LLA_TRY(rc);
  if (upper) LLA_TRY(up.finish(st));
  tm.mark(2);
  LLA_TRY(stage_download_tri(sa, a, lda, upper, rs.sent_upto, st));
  LLA_CUDA(cudaStreamSynchronize(cx->d2h));
  IntOut inf_out(info, 1);
  LLA_CUDA(cudaMemcpyAsync(inf_out.host32(), inf.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  tm.mark(3);
  LLA_TRY(stage_sync(st));
  inf_out.finish();
  tm.finish();
  return 0;
}

/* reference call site: cholesky, src/linear-algebra.lisp:670-674 (always UPLO = 'U') */
LLA_EXPORT void dpotrf_(const char* uplo, const fint* n_, double* a, const fint* lda_, fint* info) {
  ApiLock api_lock;
  int64_t n = *n_, lda = *lda_;
  bool upper = (*uplo == 'U' || *uplo == 'u');
  bool lower = (*uplo == 'L' || *uplo == 'l');
  *info = 0;
  if (!upper && !lower) { *info = -1; return; }
  if (n < 0) { *info = -2; return; }
  if (lda < max1(n)) { *info = -4; return; }
  if (n == 0) return;
  if (upper && dist_route(n)) {
    int inf = 0;
    int rc = dist_dpotrf(n, a, lda, &inf);
    *info = (fint)(rc != 0 ? rc : inf);
    return;
  }
  int rc = dpotrf_host(upper, n, a, lda, info);
  if (rc != 0) { drain_streams(); stage_sync(nullptr); *info = (fint)rc; }
}

static int dpotrs_host(bool upper, int64_t n, int64_t nrhs, const double* a, int64_t lda, double* b, int64_t ldb) {
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  cudaStream_t st = cx->stream;
  PhaseTimer tm(st);
  tm.mark(0);
  Staged sa, sb;
  LLA_TRY(stage_alloc(sa, n, n, sizeof(double)));
  LLA_TRY(stage_alloc(sb, n, nrhs, sizeof(double)));
  LLA_TRY(stage_upload(sa, a, lda, st));
  LLA_TRY(stage_upload(sb, b, ldb, st));
  tm.mark(1);
  LLA_TRY(dpotrs_dev(upper, n, nrhs, sa.ptr<double>(), sa.ld, sb.ptr<double>(), sb.ld, st));
  tm.mark(2);
  LLA_TRY(stage_download(sb, b, ldb, st));
  tm.mark(3);
  LLA_TRY(stage_sync(st));
  tm.finish();
  return 0;
}

/* reference call site: solve (cholesky), src/linear-algebra.lisp:296-301 */
LLA_EXPORT void dpotrs_(const char* uplo, const fint* n_, const fint* nrhs_, const double* a, const fint* lda_,
                        double* b, const fint* ldb_, fint* info) {
  ApiLock api_lock;
  int64_t n = *n_, nrhs = *nrhs_, lda = *lda_, ldb = *ldb_;
  bool upper = (*uplo == 'U' || *uplo == 'u');
  bool lower = (*uplo == 'L' || *uplo == 'l');
  *info = 0;
  if (!upper && !lower) { *info = -1; return; }
  if (n < 0) { *info = -2; return; }
  if (nrhs < 0) { *info = -3; return; }
  if (lda < max1(n)) { *info = -5; return; }
  if (ldb < max1(n)) { *info = -7; return; }
  if (n == 0 || nrhs == 0) return;
  int rc = dpotrs_host(upper, n, nrhs, a, lda, b, ldb);
  if (rc != 0) { drain_streams(); stage_sync(nullptr); *info = (fint)rc; }
}

// llacuda_<name> aliases (INTEGRATION.md option B: re-pointed blas-lapack-function-name)
LLA_EXPORT void llacuda_dgetrf(const fint* m, const fint* n, double* a, const fint* lda, fint* ipiv, fint* info) {
  ApiLock api_lock;
  dgetrf_(m, n, a, lda, ipiv, info);
}
LLA_EXPORT void llacuda_dgetrs(const char* trans, const fint* n, const fint* nrhs, const double* a, const fint* lda,
                               const fint* ipiv, double* b, const fint* ldb, fint* info) {
  ApiLock api_lock;
  dgetrs_(trans, n, nrhs, a, lda, ipiv, b, ldb, info);
}
LLA_EXPORT void llacuda_dgesv(const fint* n, const fint* nrhs, double* a, const fint* lda, fint* ipiv, double* b,
                              const fint* ldb, fint* info) {
  ApiLock api_lock;
  dgesv_(n, nrhs, a, lda, ipiv, b, ldb, info);
}
LLA_EXPORT void llacuda_dpotrf(const char* uplo, const fint* n, double* a, const fint* lda, fint* info) {
  ApiLock api_lock;
  dpotrf_(uplo, n, a, lda, info);
}
LLA_EXPORT void llacuda_dpotrs(const char* uplo, const fint* n, const fint* nrhs, const double* a, const fint* lda,
                               double* b, const fint* ldb, fint* info) {
  ApiLock api_lock;
  dpotrs_(uplo, n, nrhs, a, lda, b, ldb, info);
}

// ------------------------------------------------------------------ device-pointer API
LLA_EXPORT int llacuda_dgetrf_dev(int64_t m, int64_t n, double* A, int64_t lda, int* ipiv_dev, int* info_dev) {
  ApiLock api_lock;
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  if (m < 0 || n < 0 || lda < max1(m)) return LLACUDA_ERR_BAD_ARG;
  return dgetrf_dev(m, n, A, lda, ipiv_dev, info_dev, cx->stream);
}
LLA_EXPORT int llacuda_dgetrs_dev(char trans, int64_t n, int64_t nrhs, const double* A, int64_t lda,
                                  const int* ipiv_dev, double* B, int64_t ldb) {
  ApiLock api_lock;
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  int op = opcode(trans);
  if (op < 0 || n < 0 || nrhs < 0 || lda < max1(n) || ldb < max1(n)) return LLACUDA_ERR_BAD_ARG;
  return dgetrs_dev(op, n, nrhs, A, lda, ipiv_dev, B, ldb, cx->stream);
}
LLA_EXPORT int llacuda_dgesv_dev(int64_t n, int64_t nrhs, double* A, int64_t lda, int* ipiv_dev, double* B,
                                 int64_t ldb, int* info_dev) {
  ApiLock api_lock;
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  if (n < 0 || nrhs < 0 || lda < max1(n) || ldb < max1(n)) return LLACUDA_ERR_BAD_ARG;
  return dgesv_dev(n, nrhs, A, lda, ipiv_dev, B, ldb, info_dev, cx->stream);
}
LLA_EXPORT int llacuda_dpotrf_dev(char uplo, int64_t n, double* A, int64_t lda, int* info_dev) {
  ApiLock api_lock;
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  bool upper = (uplo == 'U' || uplo == 'u');
  if ((!upper && uplo != 'L' && uplo != 'l') || n < 0 || lda < max1(n)) return LLACUDA_ERR_BAD_ARG;
  return dpotrf_dev(upper, n, A, lda, info_dev, cx->stream);
}
LLA_EXPORT int llacuda_dpotrs_dev(char uplo, int64_t n, int64_t nrhs, const double* A, int64_t lda, double* B,
                                  int64_t ldb) {
  ApiLock api_lock;
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  bool upper = (uplo == 'U' || uplo == 'u');
  if ((!upper && uplo != 'L' && uplo != 'l') || n < 0 || nrhs < 0 || lda < max1(n) || ldb < max1(n))
    return LLACUDA_ERR_BAD_ARG;
  return dpotrs_dev(upper, n, nrhs, A, lda, B, ldb, cx->stream);
}
LLA_EXPORT int llacuda_dtranspose_dev(int64_t rows, int64_t cols, const double* in, int64_t ldin, double* out,
                                      int64_t ldout) {
  ApiLock api_lock;
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  return dtranspose_dev(rows, cols, in, ldin, out, ldout, cx->stream);
}

// ------------------------------------------------------------------ "next" rows: SYRK, POSV, TRSM
// C <- alpha A A^T + beta C ('N') or alpha A^T A + beta C ('T' / 'C'), only the uplo triangle of C referenced: the DMMA
// GEMM restricted to that triangle.  The triangle travels as the 512-column trapezoids of stage_upload_tri in both
// directions (uploaded even when beta = 0, so that the bytes of the other triangle inside the diagonal chunks come
// back as they were).
static int dsyrk_host(bool upper, int trans, int64_t n, int64_t k, double alpha, const double* a, int64_t lda, double beta,
                      double* c, int64_t ldc) {
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  if (n == 0) return 0;
  cudaStream_t st = cx->stream;
  PhaseTimer tm(st);
  tm.mark(0);
  Staged sa, sc;
  const bool need_a = k > 0 && alpha != 0.0;
  if (need_a) {
    LLA_TRY(stage_alloc(sa, trans ? k : n, trans ? n : k, sizeof(double)));
    LLA_TRY(stage_upload(sa, a, lda, st));
  }
  LLA_TRY(stage_alloc(sc, n, n, sizeof(double)));
  LLA_TRY(stage_upload_tri(sc, c, ldc, upper, st));
  tm.mark(1);
  LLA_TRY(dgemm_dev(trans ? 1 : 0, trans ? 0 : 1, n, n, need_a ? k : 0, alpha, sa.ptr<double>(), sa.ld, sa.ptr<double>(), sa.ld,
                    beta, sc.ptr<double>(), sc.ld, upper ? TRI_UPPER : TRI_LOWER, st));
  tm.mark(2);
  LLA_TRY(stage_download_tri(sc, c, ldc, upper, 0, st));
  tm.mark(3);
  LLA_TRY(stage_sync(st));
  tm.finish();
  return 0;
}

/* reference call site: mm-hermitian%, src/linear-algebra.lisp:62-76 ((mm a t), (mm t a)) */
LLA_EXPORT void dsyrk_(const char* uplo, const char* trans, const fint* n_, const fint* k_, const double* alpha,
                       const double* a, const fint* lda_, const double* beta, double* c, const fint* ldc_) {
  ApiLock api_lock;
  const bool upper = (*uplo == 'U' || *uplo == 'u'), lower = (*uplo == 'L' || *uplo == 'l');
  const int op = opcode(*trans);
  const int64_t n = *n_, k = *k_, lda = *lda_, ldc = *ldc_;
  const int64_t nrowa = op ? k : n;
  int bad = 0;
  if (!upper && !lower) bad = 1;
  else if (op < 0) bad = 2;
  else if (n < 0) bad = 3;
  else if (k < 0) bad = 4;
  else if (lda < (nrowa > 1 ? nrowa : 1)) bad = 7;
  else if (ldc < (n > 1 ? n : 1)) bad = 10;
  if (bad) { blas_arg_error("DSYRK ", bad); return; }
  llacuda_clear_error();
  int r = dsyrk_host(upper, op, n, k, *alpha, a, lda, *beta, c, ldc);
  if (r != 0) blas_failure("dsyrk_", r, c, ldc, n, n, sizeof(double));
}

LLA_EXPORT void llacuda_dsyrk(const char* uplo, const char* trans, const fint* n, const fint* k, const double* alpha,
                              const double* a, const fint* lda, const double* beta, double* c, const fint* ldc) {
  dsyrk_(uplo, trans, n, k, alpha, a, lda, beta, c, ldc);
}

// op(A) X = alpha B (side 'L') or X op(A) = alpha B (side 'R'), A triangular.  The device solver is left-sided
// (block inverses + DMMA GEMMs); a right-sided solve runs on the transposed right-hand side with op flipped.
static int dtrsm_host(bool left, bool upper, int op, bool unit, int64_t m, int64_t n, double alpha, const double* a,
                      int64_t lda, double* b, int64_t ldb) {
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  if (m == 0 || n == 0) return 0;
  cudaStream_t st = cx->stream;
  PhaseTimer tm(st);
  tm.mark(0);
  const int64_t na = left ? m : n;
  Staged sa, sb;
  LLA_TRY(stage_alloc(sa, na, na, sizeof(double)));
  LLA_TRY(stage_alloc(sb, m, n, sizeof(double)));
  LLA_TRY(stage_upload(sa, a, lda, st));
  LLA_TRY(stage_upload(sb, b, ldb, st));
  tm.mark(1);
  if (alpha != 1.0)  // B <- alpha B (the k = 0 path of the GEMM driver)
    LLA_TRY(dgemm_dev(0, 0, m, n, 0, 0.0, nullptr, 1, nullptr, 1, alpha, sb.ptr<double>(), sb.ld, TRI_FULL, st));
  Arena ar;
  LLA_TRY(ar.reserve(tinv_bytes(na) + 1024, st));
  double* tinv = (double*)ar.take(tinv_bytes(na));
  LLA_TRY(dtrtri_blocks_dev(upper, unit, na, sa.ptr<double>(), sa.ld, tinv, st));
  if (left) {
    LLA_TRY(dtrsm_left_inv_dev(upper, op ? 1 : 0, m, n, sa.ptr<double>(), sa.ld, tinv, sb.ptr<double>(), sb.ld, st));
  } else {
    Staged st_t;
    LLA_TRY(stage_alloc(st_t, n, m, sizeof(double)));
    LLA_TRY(dtranspose_dev(m, n, sb.ptr<double>(), sb.ld, st_t.ptr<double>(), st_t.ld, st));
    LLA_TRY(dtrsm_left_inv_dev(upper, op ? 0 : 1, n, m, sa.ptr<double>(), sa.ld, tinv, st_t.ptr<double>(), st_t.ld, st));
    LLA_TRY(dtranspose_dev(n, m, st_t.ptr<double>(), st_t.ld, sb.ptr<double>(), sb.ld, st));
    LLA_TRY(stage_sync(st));   // st_t goes back to the pool at the end of this scope
  }
  tm.mark(2);
  LLA_TRY(stage_download(sb, b, ldb, st));
  tm.mark(3);
  LLA_TRY(stage_sync(st));
  tm.finish();
  return 0;
}

/* reference call site: trsm%, src/linear-algebra.lisp:333-347 (solve on triangular matrices): side 'R', 'N', 'N' */
LLA_EXPORT void dtrsm_(const char* side, const char* uplo, const char* transa, const char* diag, const fint* m_,
                       const fint* n_, const double* alpha, const double* a, const fint* lda_, double* b, const fint* ldb_) {
  ApiLock api_lock;
  const bool left = (*side == 'L' || *side == 'l'), right = (*side == 'R' || *side == 'r');
  const bool upper = (*uplo == 'U' || *uplo == 'u'), lower = (*uplo == 'L' || *uplo == 'l');
  const bool unit = (*diag == 'U' || *diag == 'u'), nonunit = (*diag == 'N' || *diag == 'n');
  const int op = opcode(*transa);
  const int64_t m = *m_, n = *n_, lda = *lda_, ldb = *ldb_;
  const int64_t nrowa = left ? m : n;
  int bad = 0;
  if (!left && !right) bad = 1;
  else if (!upper && !lower) bad = 2;
  else if (op < 0) bad = 3;
  else if (!unit && !nonunit) bad = 4;
  else if (m < 0) bad = 5;
  else if (n < 0) bad = 6;
  else if (lda < (nrowa > 1 ? nrowa : 1)) bad = 9;
  else if (ldb < (m > 1 ? m : 1)) bad = 11;
  if (bad) { blas_arg_error("DTRSM ", bad); return; }
  llacuda_clear_error();
  int r = dtrsm_host(left, upper, op, unit, m, n, *alpha, a, lda, b, ldb);
  if (r != 0) blas_failure("dtrsm_", r, b, ldb, m, n, sizeof(double));
}

LLA_EXPORT void llacuda_dtrsm(const char* side, const char* uplo, const char* transa, const char* diag, const fint* m,
                              const fint* n, const double* alpha, const double* a, const fint* lda, double* b, const fint* ldb) {
  dtrsm_(side, uplo, transa, diag, m, n, alpha, a, lda, b, ldb);
}

// DPOSV = DPOTRF + DPOTRS, the solve gated by the device INFO word (no host round trip in between)
static int dposv_host(bool upper, int64_t n, int64_t nrhs, double* a, int64_t lda, double* b, int64_t ldb, fint* info) {
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  cudaStream_t st = cx->stream;
  PhaseTimer tm(st);
  tm.mark(0);
  Staged sa, sb;
  PoolBuf inf;
  LLA_TRY(stage_alloc(sa, n, n, sizeof(double)));
  LLA_TRY(stage_alloc(sb, n, nrhs, sizeof(double)));
  LLA_TRY(inf.alloc(sizeof(int)));
  LLA_TRY(stage_upload_tri(sa, a, lda, upper, st));
  LLA_TRY(stage_upload(sb, b, ldb, st));
  tm.mark(1);
  LLA_TRY(dpotrf_dev(upper, n, sa.ptr<double>(), sa.ld, (int*)inf.p, st));
  LLA_TRY(dpotrs_dev(upper, n, nrhs, sa.ptr<double>(), sa.ld, sb.ptr<double>(), sb.ld, st, (const int*)inf.p));
  tm.mark(2);
  LLA_TRY(stage_download_tri(sa, a, lda, upper, 0, st));
  LLA_TRY(stage_download(sb, b, ldb, st));
  IntOut inf_out(info, 1);
  LLA_CUDA(cudaMemcpyAsync(inf_out.host32(), inf.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  tm.mark(3);
  LLA_TRY(stage_sync(st));
  inf_out.finish();
  tm.finish();
  return 0;
}

/* reference call site: solve (hermitian-matrix), src/linear-algebra.lisp:303-315 */
LLA_EXPORT void dposv_(const char* uplo, const fint* n_, const fint* nrhs_, double* a, const fint* lda_, double* b,
                       const fint* ldb_, fint* info) {
  ApiLock api_lock;
  const bool upper = (*uplo == 'U' || *uplo == 'u'), lower = (*uplo == 'L' || *uplo == 'l');
  const int64_t n = *n_, nrhs = *nrhs_, lda = *lda_, ldb = *ldb_;
  *info = 0;
  if (!upper && !lower) { *info = -1; return; }
  if (n < 0) { *info = -2; return; }
  if (nrhs < 0) { *info = -3; return; }
  if (lda < max1(n)) { *info = -5; return; }
  if (ldb < max1(n)) { *info = -7; return; }
  if (n == 0 || nrhs == 0) { if (n == 0) return; }
  int rc = dposv_host(upper, n, nrhs, a, lda, b, ldb, info);
  if (rc != 0) { drain_streams(); stage_sync(nullptr); *info = (fint)rc; }
}

LLA_EXPORT void llacuda_dposv(const char* uplo, const fint* n, const fint* nrhs, double* a, const fint* lda, double* b,
                              const fint* ldb, fint* info) {
  dposv_(uplo, n, nrhs, a, lda, b, ldb, info);
}

// ------------------------------------------------------------------ the invert family
// GETRI / POTRI / TRTRI as solves against the identity on the GEMM-only triangular solvers: (invert lu),
// (invert cholesky), (invert triangular), reference src/linear-algebra.lisp:352-409.  2 n^3 flop instead of LAPACK's
// 4/3 n^3, all of it on the DMMA GEMM.
__global__ void lla_dlacpy_tri_kernel(int64_t n, const double* __restrict__ src, int64_t lds, double* __restrict__ dst,
                                      int64_t ldd, bool upper, bool skip_diag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= n) return;
  const bool in_tri = upper ? (i <= j) : (i >= j);
  if (!in_tri || (skip_diag && i == j)) return;
  dst[i + j * ldd] = src[i + j * lds];
}

static int lacpy_tri(int64_t n, const Staged& src, Staged& dst, bool upper, bool skip_diag, cudaStream_t st) {
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)n);
  lla_dlacpy_tri_kernel<<<grid, 256, 0, st>>>(n, (const double*)src.buf.p, src.ld, (double*)dst.buf.p, dst.ld, upper, skip_diag);
  count_launch();
  LLA_CUDA(cudaGetLastError());
  return 0;
}

static int set_identity(Staged& s, int64_t n, cudaStream_t st) {
  LLA_CUDA(cudaMemset2DAsync(s.buf.p, (size_t)s.ld * 8, 0, (size_t)n * 8, (size_t)n, st));
  std::vector<double> ones((size_t)n, 1.0);   // pageable source: the copy is staged before the call returns
  LLA_CUDA(cudaMemcpy2DAsync(s.buf.p, (size_t)(s.ld + 1) * 8, ones.data(), 8, 8, (size_t)n, cudaMemcpyHostToDevice, st));
  LLA_TRY(stage_sync(st));
  return 0;
}

static int dgetri_host(int64_t n, double* a, int64_t lda, const fint* ipiv) {
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  cudaStream_t st = cx->stream;
  Staged sa, sb;
  PoolBuf piv;
  LLA_TRY(stage_alloc(sa, n, n, sizeof(double)));
  LLA_TRY(stage_alloc(sb, n, n, sizeof(double)));
  LLA_TRY(piv.alloc(sizeof(int) * (size_t)n));
  IntIn pivots(ipiv, n);
  LLA_CUDA(cudaMemcpyAsync(piv.p, pivots.host32(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
  LLA_TRY(stage_upload(sa, a, lda, st));
  LLA_TRY(set_identity(sb, n, st));
  LLA_TRY(dgetrs_dev(0, n, n, sa.ptr<double>(), sa.ld, (const int*)piv.p, sb.ptr<double>(), sb.ld, st));
  LLA_TRY(stage_download(sb, a, lda, st));
  LLA_TRY(stage_sync(st));
  return 0;
}

/* reference call site: invert (lu), src/linear-algebra.lisp:358-367 (lapack-call-w/query: LWORK = -1 asks for the size) */
LLA_EXPORT void dgetri_(const fint* n_, double* a, const fint* lda_, const fint* ipiv, double* work, const fint* lwork_,
                        fint* info) {
  ApiLock api_lock;
  const int64_t n = *n_, lda = *lda_, lwork = *lwork_;
  *info = 0;
  if (n < 0) { *info = -1; return; }
  if (lda < max1(n)) { *info = -3; return; }
  if (lwork < max1(n) && lwork != -1) { *info = -6; return; }
  work[0] = (double)max1(n);
  if (lwork == -1 || n == 0) return;
  for (int64_t i = 0; i < n; i++)
    if (a[i + i * lda] == 0.0) { *info = (fint)(i + 1); return; }   // U is exactly singular: nothing is computed
  int rc = dgetri_host(n, a, lda, ipiv);
  if (rc != 0) { drain_streams(); stage_sync(nullptr); *info = (fint)rc; }
}
LLA_EXPORT void llacuda_dgetri(const fint* n, double* a, const fint* lda, const fint* ipiv, double* work, const fint* lwork,
                               fint* info) {
  dgetri_(n, a, lda, ipiv, work, lwork, info);
}

// shared body of POTRI / TRTRI: solve against the identity, copy the named triangle of the result over the factor
static int tri_inverse_host(bool potri, bool upper, bool unit, int64_t n, double* a, int64_t lda) {
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  cudaStream_t st = cx->stream;
  Staged sa, sb;
  LLA_TRY(stage_alloc(sa, n, n, sizeof(double)));
  LLA_TRY(stage_alloc(sb, n, n, sizeof(double)));
  LLA_TRY(stage_upload_tri(sa, a, lda, upper, st));
  LLA_TRY(set_identity(sb, n, st));
  if (potri) {
    LLA_TRY(dpotrs_dev(upper, n, n, sa.ptr<double>(), sa.ld, sb.ptr<double>(), sb.ld, st));
  } else {
    Arena ar;
    LLA_TRY(ar.reserve(tinv_bytes(n) + 1024, st));
    double* tinv = (double*)ar.take(tinv_bytes(n));
    LLA_TRY(dtrtri_blocks_dev(upper, unit, n, sa.ptr<double>(), sa.ld, tinv, st));
    LLA_TRY(dtrsm_left_inv_dev(upper, 0, n, n, sa.ptr<double>(), sa.ld, tinv, sb.ptr<double>(), sb.ld, st));
  }
  LLA_TRY(lacpy_tri(n, sb, sa, upper, !potri && unit, st));
  LLA_TRY(stage_download_tri(sa, a, lda, upper, 0, st));
  LLA_TRY(stage_sync(st));
  return 0;
}

/* reference call site: invert (cholesky), src/linear-algebra.lisp:399-406 */
LLA_EXPORT void dpotri_(const char* uplo, const fint* n_, double* a, const fint* lda_, fint* info) {
  ApiLock api_lock;
  const bool upper = (*uplo == 'U' || *uplo == 'u'), lower = (*uplo == 'L' || *uplo == 'l');
  const int64_t n = *n_, lda = *lda_;
  *info = 0;
  if (!upper && !lower) { *info = -1; return; }
  if (n < 0) { *info = -2; return; }
  if (lda < max1(n)) { *info = -4; return; }
  if (n == 0) return;
  for (int64_t i = 0; i < n; i++)
    if (a[i + i * lda] == 0.0) { *info = (fint)(i + 1); return; }
  int rc = tri_inverse_host(true, upper, false, n, a, lda);
  if (rc != 0) { drain_streams(); stage_sync(nullptr); *info = (fint)rc; }
}
LLA_EXPORT void llacuda_dpotri(const char* uplo, const fint* n, double* a, const fint* lda, fint* info) {
  dpotri_(uplo, n, a, lda, info);
}

/* reference call site: invert-triangular%, src/linear-algebra.lisp:381-391 */
LLA_EXPORT void dtrtri_(const char* uplo, const char* diag, const fint* n_, double* a, const fint* lda_, fint* info) {
  ApiLock api_lock;
  const bool upper = (*uplo == 'U' || *uplo == 'u'), lower = (*uplo == 'L' || *uplo == 'l');
  const bool unit = (*diag == 'U' || *diag == 'u'), nonunit = (*diag == 'N' || *diag == 'n');
  const int64_t n = *n_, lda = *lda_;
  *info = 0;
  if (!upper && !lower) { *info = -1; return; }
  if (!unit && !nonunit) { *info = -2; return; }
  if (n < 0) { *info = -3; return; }
  if (lda < max1(n)) { *info = -5; return; }
  if (n == 0) return;
  if (nonunit)
    for (int64_t i = 0; i < n; i++)
      if (a[i + i * lda] == 0.0) { *info = (fint)(i + 1); return; }
  int rc = tri_inverse_host(false, upper, unit, n, a, lda);
  if (rc != 0) { drain_streams(); stage_sync(nullptr); *info = (fint)rc; }
}
LLA_EXPORT void llacuda_dtrtri(const char* uplo, const char* diag, const fint* n, double* a, const fint* lda, fint* info) {
  dtrtri_(uplo, diag, n, a, lda, info);
}

// ------------------------------------------------------------------ row-major-native entry points
// SURVEY section 8b (Option B, "richer entry points") / 8f rank 1.  LLA's arrays are row-major; for lu / solve the
// reference converts them with element-wise Lisp loops on both sides of the FFI call (transpose-matrix-to/from-memory,
// src/foreign-memory.lisp:211-255 -- "today's hidden hot loop").  These entry points take the row-major buffers as they
// are (host pointers): the conversion is an HBM-bound transpose on the device, and operands that LLA marks input-only
// (A of `solve`, the factors of `solve` on an lu / cholesky object) are never sent back.
// A row-major r x c array with row width ld is the column-major c x r matrix with leading dimension ld.
struct RowMajorIn {
  Staged raw, cm;   // raw: the bytes as uploaded (column-major c x r); cm: the r x c column-major matrix
  int upload(int64_t r, int64_t c, const double* host, int64_t ld, cudaStream_t st) {
    LLA_TRY(stage_alloc(raw, c, r, sizeof(double)));
    LLA_TRY(stage_alloc(cm, r, c, sizeof(double)));
    LLA_TRY(stage_upload(raw, host, ld, st));
    return dtranspose_dev(c, r, raw.ptr<double>(), raw.ld, cm.ptr<double>(), cm.ld, st);
  }
  int download(int64_t r, int64_t c, double* host, int64_t ld, cudaStream_t st) {
    LLA_TRY(dtranspose_dev(r, c, cm.ptr<double>(), cm.ld, raw.ptr<double>(), raw.ld, st));
    return stage_download(raw, host, ld, st);
  }
};

/* lu on a row-major array: reference src/linear-algebra.lisp:225-234 without the two Lisp transposes */
LLA_EXPORT void llacuda_dgetrf_rm(const fint* m_, const fint* n_, double* a, const fint* lda_, fint* ipiv, fint* info) {
  ApiLock api_lock;
  const int64_t m = *m_, n = *n_, lda = *lda_;
  *info = 0;
  if (m < 0) { *info = -1; return; }
  if (n < 0) { *info = -2; return; }
  if (lda < max1(n)) { *info = -4; return; }
  if (m == 0 || n == 0) return;
  Context* cx = ctx();
  if (!cx) { *info = LLACUDA_ERR_NO_DEVICE; return; }
  cudaStream_t st = cx->stream;
  auto run = [&]() -> int {
    RowMajorIn A;
    PoolBuf piv, inf;
    const int64_t mn = m < n ? m : n;
    LLA_TRY(piv.alloc(sizeof(int) * (size_t)max1(mn)));
    LLA_TRY(inf.alloc(sizeof(int)));
    LLA_TRY(A.upload(m, n, a, lda, st));
    LLA_TRY(dgetrf_dev(m, n, A.cm.ptr<double>(), A.cm.ld, (int*)piv.p, (int*)inf.p, st));
    LLA_TRY(A.download(m, n, a, lda, st));
    IntOut pivots(ipiv, mn), inf_out(info, 1);
    LLA_CUDA(cudaMemcpyAsync(pivots.host32(), piv.p, sizeof(int) * (size_t)mn, cudaMemcpyDeviceToHost, st));
    LLA_CUDA(cudaMemcpyAsync(inf_out.host32(), inf.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    LLA_TRY(stage_sync(st));
    pivots.finish(); inf_out.finish();
    return 0;
  };
  int rc = run();
  if (rc != 0) { drain_streams(); stage_sync(nullptr); *info = (fint)rc; }
}

/* solve on raw row-major arrays: reference src/linear-algebra.lisp:278-289; A is input-only there (&array-in), the
 * factors and pivots are discarded, so nothing but X travels back */
LLA_EXPORT void llacuda_dgesv_rm(const fint* n_, const fint* nrhs_, const double* a, const fint* lda_, double* b,
                                 const fint* ldb_, fint* info) {
  ApiLock api_lock;
  const int64_t n = *n_, nrhs = *nrhs_, lda = *lda_, ldb = *ldb_;
  *info = 0;
  if (n < 0) { *info = -1; return; }
  if (nrhs < 0) { *info = -2; return; }
  if (lda < max1(n)) { *info = -4; return; }
  if (ldb < max1(nrhs)) { *info = -6; return; }
  if (n == 0 || nrhs == 0) return;
  Context* cx = ctx();
  if (!cx) { *info = LLACUDA_ERR_NO_DEVICE; return; }
  cudaStream_t st = cx->stream;
  auto run = [&]() -> int {
    RowMajorIn A, B;
    PoolBuf piv, inf;
    LLA_TRY(piv.alloc(sizeof(int) * (size_t)n));
    LLA_TRY(inf.alloc(sizeof(int)));
    LLA_TRY(A.upload(n, n, a, lda, st));
    LLA_TRY(B.upload(n, nrhs, b, ldb, st));
    LLA_TRY(dgesv_dev(n, nrhs, A.cm.ptr<double>(), A.cm.ld, (int*)piv.p, B.cm.ptr<double>(), B.cm.ld, (int*)inf.p, st));
    LLA_TRY(B.download(n, nrhs, b, ldb, st));   // untouched B comes back when INFO > 0 (restored / the solve was skipped)
    IntOut inf_out(info, 1);
    LLA_CUDA(cudaMemcpyAsync(inf_out.host32(), inf.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    LLA_TRY(stage_sync(st));
    inf_out.finish();
    return 0;
  };
  int rc = run();
  if (rc != 0) { drain_streams(); stage_sync(nullptr); *info = (fint)rc; }
}

/* solve on an lu object: reference src/linear-algebra.lisp:264-276; LU (row-major, as LLA stores it) and ipiv are inputs */
LLA_EXPORT void llacuda_dgetrs_rm(const fint* n_, const fint* nrhs_, const double* lu, const fint* lda_, const fint* ipiv,
                                  double* b, const fint* ldb_, fint* info) {
  ApiLock api_lock;
  const int64_t n = *n_, nrhs = *nrhs_, lda = *lda_, ldb = *ldb_;
  *info = 0;
  if (n < 0) { *info = -1; return; }
  if (nrhs < 0) { *info = -2; return; }
  if (lda < max1(n)) { *info = -4; return; }
  if (ldb < max1(nrhs)) { *info = -7; return; }
  if (n == 0 || nrhs == 0) return;
  Context* cx = ctx();
  if (!cx) { *info = LLACUDA_ERR_NO_DEVICE; return; }
  cudaStream_t st = cx->stream;
  auto run = [&]() -> int {
    RowMajorIn A, B;
    PoolBuf piv;
    LLA_TRY(piv.alloc(sizeof(int) * (size_t)n));
    IntIn pivots(ipiv, n);
    LLA_CUDA(cudaMemcpyAsync(piv.p, pivots.host32(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
    LLA_TRY(A.upload(n, n, lu, lda, st));
    LLA_TRY(B.upload(n, nrhs, b, ldb, st));
    LLA_TRY(dgetrs_dev(0, n, nrhs, A.cm.ptr<double>(), A.cm.ld, (const int*)piv.p, B.cm.ptr<double>(), B.cm.ld, st));
    LLA_TRY(B.download(n, nrhs, b, ldb, st));
    LLA_TRY(stage_sync(st));
    return 0;
  };
  int rc = run();
  if (rc != 0) { drain_streams(); stage_sync(nullptr); *info = (fint)rc; }
}

/* solve on a cholesky object: reference src/linear-algebra.lisp:291-301; the row-major lower factor is shared as it is
 * (POTRS 'U' of its column-major view), only B needs the layout change */
LLA_EXPORT void llacuda_dpotrs_rm(const fint* n_, const fint* nrhs_, const double* l, const fint* lda_, double* b,
                                  const fint* ldb_, fint* info) {
  ApiLock api_lock;
  const int64_t n = *n_, nrhs = *nrhs_, lda = *lda_, ldb = *ldb_;
  *info = 0;
  if (n < 0) { *info = -1; return; }
  if (nrhs < 0) { *info = -2; return; }
  if (lda < max1(n)) { *info = -4; return; }
  if (ldb < max1(nrhs)) { *info = -6; return; }
  if (n == 0 || nrhs == 0) return;
  Context* cx = ctx();
  if (!cx) { *info = LLACUDA_ERR_NO_DEVICE; return; }
  cudaStream_t st = cx->stream;
  auto run = [&]() -> int {
    Staged sa;
    RowMajorIn B;
    LLA_TRY(stage_alloc(sa, n, n, sizeof(double)));
    LLA_TRY(stage_upload_tri(sa, l, lda, true, st));
    LLA_TRY(B.upload(n, nrhs, b, ldb, st));
    LLA_TRY(dpotrs_dev(true, n, nrhs, sa.ptr<double>(), sa.ld, B.cm.ptr<double>(), B.cm.ld, st));
    LLA_TRY(B.download(n, nrhs, b, ldb, st));
    LLA_TRY(stage_sync(st));
    return 0;
  };
  int rc = run();
  if (rc != 0) { drain_streams(); stage_sync(nullptr); *info = (fint)rc; }
}

// ------------------------------------------------------------------ ZGEMM
static int zgemm_host(int oa, int ob, int64_t m, int64_t n, int64_t k, cuDoubleComplex alpha,
                      const cuDoubleComplex* a, int64_t lda, const cuDoubleComplex* b, int64_t ldb,
                      cuDoubleComplex beta, cuDoubleComplex* c, int64_t ldc) {
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  if (m == 0 || n == 0) return 0;
  cudaStream_t st = cx->stream;
  PhaseTimer tm(st);
  tm.mark(0);
  const bool need_ab = (k > 0 && !(alpha.x == 0.0 && alpha.y == 0.0));
  const bool need_c = !(beta.x == 0.0 && beta.y == 0.0);
  Staged sa, sb, sc;
  int64_t ar = oa ? k : m, ac = oa ? m : k;
  int64_t br = ob ? n : k, bc = ob ? k : n;
  if (need_ab) {
    LLA_TRY(stage_alloc(sa, ar, ac, sizeof(cuDoubleComplex)));
    LLA_TRY(stage_alloc(sb, br, bc, sizeof(cuDoubleComplex)));
    LLA_TRY(stage_upload(sa, a, lda, st));
    LLA_TRY(stage_upload(sb, b, ldb, st));
  }
  LLA_TRY(stage_alloc(sc, m, n, sizeof(cuDoubleComplex)));
  if (need_c) LLA_TRY(stage_upload(sc, c, ldc, st));
  tm.mark(1);
  LLA_TRY(zgemm_dev(oa, ob, m, n, k, alpha, sa.ptr<cuDoubleComplex>(), sa.ld, sb.ptr<cuDoubleComplex>(), sb.ld, beta,
                    sc.ptr<cuDoubleComplex>(), sc.ld, TRI_FULL, st));
  tm.mark(2);
  LLA_TRY(stage_download(sc, c, ldc, st));
  tm.mark(3);
  LLA_TRY(stage_sync(st));
  tm.finish();
  return 0;
}

/* reference call sites: mm src/linear-algebra.lisp:50-54, gemm! src/blas.lisp:56-63 (complex-double arrays) */
LLA_EXPORT void zgemm_(const char* transa, const char* transb, const fint* m_, const fint* n_, const fint* k_,
                       const void* alpha, const void* a, const fint* lda_, const void* b, const fint* ldb_,
                       const void* beta, void* c, const fint* ldc_) {
  ApiLock api_lock;
  int oa = opcode(*transa), ob = opcode(*transb);
  int64_t m = *m_, n = *n_, k = *k_, lda = *lda_, ldb = *ldb_, ldc = *ldc_;
  int64_t nrowa = oa ? k : m, nrowb = ob ? n : k;
  int bad = 0;
  if (oa < 0) bad = 1;
  else if (ob < 0) bad = 2;
  else if (m < 0) bad = 3;
  else if (n < 0) bad = 4;
  else if (k < 0) bad = 5;
  else if (lda < (nrowa > 1 ? nrowa : 1)) bad = 8;
  else if (ldb < (nrowb > 1 ? nrowb : 1)) bad = 10;
  else if (ldc < (m > 1 ? m : 1)) bad = 13;
  if (bad) { blas_arg_error("ZGEMM ", bad); return; }
  llacuda_clear_error();
  int r = zgemm_host(oa, ob, m, n, k, *(const cuDoubleComplex*)alpha, (const cuDoubleComplex*)a, lda,
                     (const cuDoubleComplex*)b, ldb, *(const cuDoubleComplex*)beta, (cuDoubleComplex*)c, ldc);
  if (r != 0) blas_failure("zgemm_", r, c, ldc, m, n, sizeof(cuDoubleComplex));
}
LLA_EXPORT void llacuda_zgemm(const char* transa, const char* transb, const fint* m, const fint* n, const fint* k,
                              const void* alpha, const void* a, const fint* lda, const void* b, const fint* ldb,
                              const void* beta, void* c, const fint* ldc) {
  ApiLock api_lock;
  zgemm_(transa, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
LLA_EXPORT int llacuda_zgemm_dev(char opa, char opb, int64_t m, int64_t n, int64_t k, const double* alpha2,
                                 const void* A, int64_t lda, const void* B, int64_t ldb, const double* beta2,
                                 void* C, int64_t ldc) {
  ApiLock api_lock;
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  int oa = opcode(opa), ob = opcode(opb);
  if (oa < 0 || ob < 0 || m < 0 || n < 0 || k < 0) return LLACUDA_ERR_BAD_ARG;
  return zgemm_dev(oa, ob, m, n, k, make_cuDoubleComplex(alpha2[0], alpha2[1]), (const cuDoubleComplex*)A, lda,
                   (const cuDoubleComplex*)B, ldb, make_cuDoubleComplex(beta2[0], beta2[1]), (cuDoubleComplex*)C, ldc,
                   TRI_FULL, cx->stream);
}

// ------------------------------------------------------------------ SGEMM
static int sgemm_host(int oa, int ob, int64_t m, int64_t n, int64_t k, float alpha, const float* a, int64_t lda,
                      const float* b, int64_t ldb, float beta, float* c, int64_t ldc) {
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  if (m == 0 || n == 0) return 0;
  cudaStream_t st = cx->stream;
  PhaseTimer tm(st);
  tm.mark(0);
  const bool need_ab = (k > 0 && alpha != 0.0f);
  Staged sa, sb, sc;
  int64_t ar = oa ? k : m, ac = oa ? m : k;
  int64_t br = ob ? n : k, bc = ob ? k : n;
  if (need_ab) {
    LLA_TRY(stage_alloc(sa, ar, ac, sizeof(float)));
    LLA_TRY(stage_alloc(sb, br, bc, sizeof(float)));
    LLA_TRY(stage_upload(sa, a, lda, st));
    LLA_TRY(stage_upload(sb, b, ldb, st));
  }
  LLA_TRY(stage_alloc(sc, m, n, sizeof(float)));
  if (beta != 0.0f) LLA_TRY(stage_upload(sc, c, ldc, st));
  tm.mark(1);
  LLA_TRY(sgemm_dev(oa, ob, m, n, k, alpha, sa.ptr<float>(), sa.ld, sb.ptr<float>(), sb.ld, beta, sc.ptr<float>(), sc.ld, st));
  tm.mark(2);
  LLA_TRY(stage_download(sc, c, ldc, st));
  tm.mark(3);
  LLA_TRY(stage_sync(st));
  tm.finish();
  return 0;
}

/* reference call sites: mm src/linear-algebra.lisp:50-54, gemm! src/blas.lisp:56-63 (single-float arrays) */
LLA_EXPORT void sgemm_(const char* transa, const char* transb, const fint* m_, const fint* n_, const fint* k_,
                       const float* alpha, const float* a, const fint* lda_, const float* b, const fint* ldb_,
                       const float* beta, float* c, const fint* ldc_) {
  ApiLock api_lock;
  int oa = opcode(*transa), ob = opcode(*transb);
  int64_t m = *m_, n = *n_, k = *k_, lda = *lda_, ldb = *ldb_, ldc = *ldc_;
  int64_t nrowa = oa ? k : m, nrowb = ob ? n : k;
  int bad = 0;
  if (oa < 0) bad = 1;
  else if (ob < 0) bad = 2;
  else if (m < 0) bad = 3;
  else if (n < 0) bad = 4;
  else if (k < 0) bad = 5;
  else if (lda < (nrowa > 1 ? nrowa : 1)) bad = 8;
  else if (ldb < (nrowb > 1 ? nrowb : 1)) bad = 10;
  else if (ldc < (m > 1 ? m : 1)) bad = 13;
  if (bad) { blas_arg_error("SGEMM ", bad); return; }
  llacuda_clear_error();
  int r = sgemm_host(oa, ob, m, n, k, *alpha, a, lda, b, ldb, *beta, c, ldc);
  if (r != 0) blas_failure("sgemm_", r, c, ldc, m, n, sizeof(float));
}
LLA_EXPORT void llacuda_sgemm(const char* transa, const char* transb, const fint* m, const fint* n, const fint* k,
                              const float* alpha, const float* a, const fint* lda, const float* b, const fint* ldb,
                              const float* beta, float* c, const fint* ldc) {
  ApiLock api_lock;
  sgemm_(transa, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
