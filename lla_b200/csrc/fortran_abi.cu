// lla_b200/csrc/fortran_abi.cu -- the Fortran-symbol drop-in layer (host pointers in, host
// pointers out).  These are the symbols LLA's blas-call / lapack-call macros resolve
// (reference src/fortran-call.lisp:405-439); argument lists as issued by the call sites listed
// in include/llacuda.h.  Everything here is synchronous: LLA's buffers are only valid until the
// call returns (reference src/pinned-array.lisp:105-141).
#include "internal.h"
#include "staging.h"
#include "../../include/llacuda.h"

#include <cstdio>

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
    LLA_TRY(stage_upload(sa, a, lda, st));
    LLA_TRY(stage_upload(sb, b, ldb, st));
  }
  LLA_TRY(stage_alloc(sc, m, n, sizeof(double)));
  if (beta != 0.0) LLA_TRY(stage_upload(sc, c, ldc, st));
  tm.mark(1);
  LLA_TRY(dgemm_dev(oa, ob, m, n, k, alpha, sa.ptr<double>(), sa.ld, sb.ptr<double>(), sb.ld, beta,
                    sc.ptr<double>(), sc.ld, TRI_FULL, st));
  tm.mark(2);
  LLA_TRY(stage_download(sc, c, ldc, st));
  tm.mark(3);
  LLA_CUDA(cudaStreamSynchronize(st));
  tm.finish();
  return 0;
}

LLA_EXPORT void dgemm_(const char* transa, const char* transb, const int* m_, const int* n_, const int* k_,
                       const double* alpha, const double* a, const int* lda_, const double* b, const int* ldb_,
                       const double* beta, double* c, const int* ldc_) {
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
  int r = dgemm_host(oa, ob, m, n, k, *alpha, a, lda, b, ldb, *beta, c, ldc);
  if (r != 0) fprintf(stderr, " ** llacuda: dgemm_ failed: %s\n", llacuda_last_error());
}

LLA_EXPORT void llacuda_dgemm(const char* transa, const char* transb, const int* m, const int* n, const int* k,
                              const double* alpha, const double* a, const int* lda, const double* b,
                              const int* ldb, const double* beta, double* c, const int* ldc) {
  dgemm_(transa, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
