// lla_b200/csrc/gemm_c32.cu -- CGEMM (complex single) for LLA's mm / gemm! on complex-single arrays
// (reference src/linear-algebra.lisp:50-54, src/blas.lisp:56-63; type letter "c", src/fortran-call.lisp:405-416).
//
// Built on the tcgen05 SGEMM (gemm_f32.cu, 3xTF32, FP32-grade accuracy): the interleaved operands are split
// into planar real / imaginary images (conjugation folded into the split, so op = 'C' is a plain transpose
// afterwards), four real products give  Tr = Ar Br - Ai Bi,  Ti = Ar Bi + Ai Br  on the tensor cores, and one
// elementwise pass forms  C = alpha T + beta C  (beta == 0 never reads C).  The split and combine passes are
// HBM-bound (8 bytes in, 8 bytes out per element); the products carry all the flops (8 m n k real flop).
#include "internal.h"
#include "staging.h"
#include "../../include/llacuda.h"

#include <cstdio>

namespace llacuda {

__global__ void lla_csplit_kernel(int64_t rows, int64_t cols, const float2* __restrict__ src, int64_t lds,
                                  float* __restrict__ re, float* __restrict__ im, int64_t ldd, bool conj) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t c = blockIdx.y;
  if (r >= rows) return;
  const float2 v = src[r + c * lds];
  re[r + c * ldd] = v.x;
  im[r + c * ldd] = conj ? -v.y : v.y;
}

__global__ void lla_ccombine_kernel(int64_t m, int64_t n, const float* __restrict__ tr, const float* __restrict__ ti,
                                    int64_t ldt, float2 alpha, float2 beta, float2* __restrict__ C, int64_t ldc,
                                    bool have_t) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t c = blockIdx.y;
  if (r >= m) return;
  float2 out = make_float2(0.f, 0.f);
  if (have_t) {
    const float x = tr[r + c * ldt], y = ti[r + c * ldt];
    out.x = alpha.x * x - alpha.y * y;
    out.y = alpha.x * y + alpha.y * x;
  }
  if (beta.x != 0.f || beta.y != 0.f) {
    const float2 cv = C[r + c * ldc];
    out.x += beta.x * cv.x - beta.y * cv.y;
    out.y += beta.x * cv.y + beta.y * cv.x;
  }
  C[r + c * ldc] = out;
}

static inline int64_t pad4(int64_t x) { return (x + 3) / 4 * 4; }

// op codes: 0 = N, 1 = T, 2 = C (conjugate transpose)
int cgemm_dev(int opa, int opb, int64_t m, int64_t n, int64_t k, cuFloatComplex alpha, const cuFloatComplex* A,
              int64_t lda, const cuFloatComplex* B, int64_t ldb, cuFloatComplex beta, cuFloatComplex* C, int64_t ldc,
              cudaStream_t st) {
  if (m <= 0 || n <= 0) return 0;
  if (!ctx()) return LLACUDA_ERR_NO_DEVICE;
  const bool have_t = k > 0 && !(alpha.x == 0.f && alpha.y == 0.f);
  const dim3 blk(256);
  if (!have_t) {
    if (beta.x == 1.f && beta.y == 0.f) return 0;
    lla_ccombine_kernel<<<dim3((unsigned)((m + 255) / 256), (unsigned)n), blk, 0, st>>>(m, n, nullptr, nullptr, 0, alpha, beta,
                                                                                   (float2*)C, ldc, false);
    count_launch();
    LLA_CUDA(cudaGetLastError());
    return 0;
  }
  const int64_t ar = opa ? k : m, ac = opa ? m : k;   // stored shapes
  const int64_t br = opb ? n : k, bc = opb ? k : n;
  const int64_t la = pad4(ar), lb = pad4(br), lt = pad4(m);
  PoolBuf pa, pb, pt;
  LLA_TRY(pa.alloc(sizeof(float) * 2 * (size_t)la * ac));
  LLA_TRY(pb.alloc(sizeof(float) * 2 * (size_t)lb * bc));
  LLA_TRY(pt.alloc(sizeof(float) * 2 * (size_t)lt * n));
  float *Ar = (float*)pa.p, *Ai = Ar + la * ac, *Br = (float*)pb.p, *Bi = Br + lb * bc, *Tr = (float*)pt.p, *Ti = Tr + lt * n;
  lla_csplit_kernel<<<dim3((unsigned)((ar + 255) / 256), (unsigned)ac), blk, 0, st>>>(ar, ac, (const float2*)A, lda, Ar, Ai, la, opa == 2);
  lla_csplit_kernel<<<dim3((unsigned)((br + 255) / 256), (unsigned)bc), blk, 0, st>>>(br, bc, (const float2*)B, ldb, Br, Bi, lb, opb == 2);
  count_launch(2);
  LLA_CUDA(cudaGetLastError());
  const int ta = opa ? 1 : 0, tb = opb ? 1 : 0;
  LLA_TRY(sgemm_dev(ta, tb, m, n, k, 1.0f, Ar, la, Br, lb, 0.0f, Tr, lt, st));
  LLA_TRY(sgemm_dev(ta, tb, m, n, k, -1.0f, Ai, la, Bi, lb, 1.0f, Tr, lt, st));
  LLA_TRY(sgemm_dev(ta, tb, m, n, k, 1.0f, Ar, la, Bi, lb, 0.0f, Ti, lt, st));
  LLA_TRY(sgemm_dev(ta, tb, m, n, k, 1.0f, Ai, la, Br, lb, 1.0f, Ti, lt, st));
  lla_ccombine_kernel<<<dim3((unsigned)((m + 255) / 256), (unsigned)n), blk, 0, st>>>(m, n, Tr, Ti, lt, alpha, beta, (float2*)C, ldc, true);
  count_launch();
  LLA_CUDA(cudaGetLastError());
  // the pool buffers go back to the pool when this frame ends; the pool is stream-ordered per call site:
  LLA_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // namespace llacuda

using namespace llacuda;

static int c_opcode(char c) {
  switch (c) {
    case 'N': case 'n': return 0;
    case 'T': case 't': return 1;
    case 'C': case 'c': return 2;
    default: return -1;
  }
}

extern "C" int llacuda_cgemm_dev(char opa, char opb, int64_t m, int64_t n, int64_t k, const float* alpha2, const void* A,
                                 int64_t lda, const void* B, int64_t ldb, const float* beta2, void* C, int64_t ldc) {
  ApiLock api_lock;
  int oa = c_opcode(opa), ob = c_opcode(opb);
  if (oa < 0 || ob < 0 || m < 0 || n < 0 || k < 0) return LLACUDA_ERR_BAD_ARG;
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  return cgemm_dev(oa, ob, m, n, k, make_cuFloatComplex(alpha2[0], alpha2[1]), (const cuFloatComplex*)A, lda,
                   (const cuFloatComplex*)B, ldb, make_cuFloatComplex(beta2[0], beta2[1]), (cuFloatComplex*)C, ldc, c->stream);
}

static int cgemm_host(int oa, int ob, int64_t m, int64_t n, int64_t k, cuFloatComplex alpha, const cuFloatComplex* a,
                      int64_t lda, const cuFloatComplex* b, int64_t ldb, cuFloatComplex beta, cuFloatComplex* c, int64_t ldc) {
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  if (m == 0 || n == 0) return 0;
  cudaStream_t st = cx->stream;
  PhaseTimer tm(st);
  tm.mark(0);
  const bool need_ab = (k > 0 && !(alpha.x == 0.f && alpha.y == 0.f));
  const bool need_c = !(beta.x == 0.f && beta.y == 0.f);
  Staged sa, sb, sc;
  const int64_t ar = oa ? k : m, ac = oa ? m : k, br = ob ? n : k, bc = ob ? k : n;
  if (need_ab) {
    LLA_TRY(stage_alloc(sa, ar, ac, sizeof(cuFloatComplex)));
    LLA_TRY(stage_alloc(sb, br, bc, sizeof(cuFloatComplex)));
    LLA_TRY(stage_upload(sa, a, lda, st));
    LLA_TRY(stage_upload(sb, b, ldb, st));
  }
  LLA_TRY(stage_alloc(sc, m, n, sizeof(cuFloatComplex)));
  if (need_c) LLA_TRY(stage_upload(sc, c, ldc, st));
  tm.mark(1);
  LLA_TRY(cgemm_dev(oa, ob, m, n, k, alpha, sa.ptr<cuFloatComplex>(), sa.ld, sb.ptr<cuFloatComplex>(), sb.ld, beta,
                    sc.ptr<cuFloatComplex>(), sc.ld, st));
  tm.mark(2);
  LLA_TRY(stage_download(sc, c, ldc, st));
  tm.mark(3);
  LLA_TRY(stage_sync(st));
  tm.finish();
  return 0;
}

/* reference call sites: mm src/linear-algebra.lisp:50-54, gemm! src/blas.lisp:56-63 (complex-single arrays) */
extern "C" void cgemm_(const char* transa, const char* transb, const fint* m_, const fint* n_, const fint* k_, const void* alpha,
                       const void* a, const fint* lda_, const void* b, const fint* ldb_, const void* beta, void* c,
                       const fint* ldc_) {
  ApiLock api_lock;
  int oa = c_opcode(*transa), ob = c_opcode(*transb);
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
  if (bad) {
    char msg[96];
    snprintf(msg, sizeof(msg), "CGEMM : parameter number %d had an illegal value", bad);
    set_error(__FILE__, __LINE__, msg, bad);
    fprintf(stderr, " ** llacuda: %s\n", msg);
    return;
  }
  llacuda_clear_error();
  int r = cgemm_host(oa, ob, m, n, k, *(const cuFloatComplex*)alpha, (const cuFloatComplex*)a, lda,
                     (const cuFloatComplex*)b, ldb, *(const cuFloatComplex*)beta, (cuFloatComplex*)c, ldc);
  if (r != 0) blas_failure("cgemm_", r, c, ldc, m, n, sizeof(cuFloatComplex));
}

extern "C" void llacuda_cgemm(const char* transa, const char* transb, const fint* m, const fint* n, const fint* k,
                              const void* alpha, const void* a, const fint* lda, const void* b, const fint* ldb,
                              const void* beta, void* c, const fint* ldc) {
  ApiLock api_lock;
  cgemm_(transa, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
