/* include/llacuda.h -- C ABI of libllacuda.so, the B200 (sm_100a) back end for the dense
 * linear-algebra hot path of tpapp/lla (mm / gemm!, lu, solve, cholesky).
 *
 * Three groups of entry points:
 *
 *  1. Fortran-symbol drop-ins (`dgemm_`, `dgetrf_`, ...): exactly the symbols LLA's
 *     `blas-call` / `lapack-call` bind through `cffi:foreign-funcall`
 *     (reference src/fortran-call.lisp:405-439: lowercase type letter + name + "_", every
 *     argument a pointer, no hidden string-length arguments, `void` return).  All pointers are
 *     HOST pointers valid only for the duration of the call (reference src/pinned-array.lisp:
 *     105-141); the library stages them to the device, computes, and writes results back in
 *     place before returning.  Select with
 *         (defvar cl-user::*lla-configuration* '(:libraries ("libllacuda.so" "liblapack.so")))
 *     (reference src/libraries.lisp:10-12, README.md:71-77).
 *
 *  2. `llacuda_<routine>` aliases of group 1 with identical signatures, for the re-pointed name
 *     generator variant of the integration (INTEGRATION.md option B).
 *
 *  3. `llacuda_*_dev` entry points on DEVICE pointers (column-major, 64-bit dimensions), plus
 *     memory / stream / timing helpers.  These are what the new device-backed Lisp methods and
 *     the benchmark's HBM-resident leg call.
 *
 * Error convention (reference src/fortran-call.lisp:336-347): LAPACK-style entry points write
 * *info = 0 on success, -k if argument k is illegal, +k for a zero pivot / non-PD minor.
 * CUDA failures are reported out of band: *info = LLACUDA_ERR_CUDA_BASE - cudaError and a
 * message retrievable through llacuda_last_error().  No entry point throws or longjmps.
 */
#ifndef LLACUDA_H
#define LLACUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LLACUDA_ERR_NO_DEVICE (-10001)
#define LLACUDA_ERR_CUDA_BASE (-20000) /* info = BASE - cudaError_t */
#define LLACUDA_ERR_BAD_ARG   (-10002)
#define LLACUDA_ERR_NCCL_BASE (-30000) /* info = BASE - ncclResult_t (multi-GPU entry points) */

/* ------------------------------------------------------------------ runtime / helpers */
int  llacuda_set_device(int device);           /* before the first call; one device per process */
int  llacuda_init(void);                       /* 0, or LLACUDA_ERR_NO_DEVICE (no CPU fallback) */
/* BLAS-style entry points (dgemm_, dsyrk_, dtrsm_, ...) have no INFO argument.  When one of them fails on the device
 * side (out of memory, launch failure, no device) the library does what netlib's XERBLA does: it prints the reason and
 * stops the program (mode 0, the default) -- LLA's mm passes a zeroed C with beta = 0, so returning quietly would look
 * like an all-zero product.  Mode 1 (or LLACUDA_ON_ERROR=return): the output is filled with NaN, the error is recorded
 * and the call returns; the caller must poll llacuda_last_error_code(), which every such entry point clears on entry.
 * The handler, if any, runs first in both modes. */
void llacuda_set_error_mode(int mode);
int  llacuda_error_mode(void);
void llacuda_set_error_handler(void (*handler)(const char* routine, int code, const char* message));
const char* llacuda_last_error(void);
int  llacuda_last_error_code(void);
void llacuda_clear_error(void);
int  llacuda_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);
void* llacuda_stream(void);                    /* cudaStream_t the *_dev calls are enqueued on */
int  llacuda_set_stream(void* cuda_stream);    /* adopt a caller-owned stream (NULL = legacy default stream) */
int  llacuda_reset_stream(void);               /* back to the library's own stream */
int  llacuda_synchronize(void);
unsigned long long llacuda_launch_count(void); /* llacuda kernels launched so far by this process */
void llacuda_trace_enable(int on);             /* diagnostics: event trace of the blocked drivers (also LLACUDA_TRACE=1) */
void llacuda_trace_dump(void);                 /* print "t_ms stream label a b" per mark to stderr, then clear */
int  llacuda_last_timing(float* ms4);          /* last host-pointer call: h2d, compute, d2h, total */
int  llacuda_malloc(void** p, size_t bytes);
int  llacuda_free(void* p);
int  llacuda_host_alloc(void** p, size_t bytes);   /* pinned host memory */
int  llacuda_host_free(void* p);
int  llacuda_memcpy_h2d(void* dst, const void* src, size_t bytes);
int  llacuda_memcpy_d2h(void* dst, const void* src, size_t bytes);
int  llacuda_memset(void* dst, int byte, size_t bytes);
int  llacuda_trim(void);                        /* release the cached device staging buffers */

/* FP64 roof probes: sustained TFLOP/s of a register-resident DMMA (kind 0) or DFMA (kind 1)
 * loop on every SM; the roofline denominator bench.py reports against (MEASURED_PEAKS.json
 * holds no FP64 figure). */
int  llacuda_probe_fp64_peak(int kind, int iters, double* tflops, float* ms);

/* ------------------------------------------------------------------ device-pointer API */
/* op: 'N', 'T' or 'C'.  tri: 0 full, 1 only tiles/elements with row<=col, 2 only row>=col. */
int llacuda_dgemm_dev(char opa, char opb, int64_t m, int64_t n, int64_t k, double alpha,
                      const double* A, int64_t lda, const double* B, int64_t ldb,
                      double beta, double* C, int64_t ldc);
int llacuda_dsyrk_like_dev(char opa, char opb, int64_t m, int64_t n, int64_t k, double alpha,
                           const double* A, int64_t lda, const double* B, int64_t ldb,
                           double beta, double* C, int64_t ldc, int tri);

/* single precision on tcgen05 (kind::tf32, 3xTF32 split, FP32 accumulation in TMEM), TMA-fed when A, B are 16-byte
 * aligned with lda, ldb multiples of 4; any other operand goes to an FP32-FMA tile kernel (same result class as a reference
 * SGEMM, much slower).  Accuracy of the tcgen05 path (also behind sgemm_ / cgemm_): the hi/lo TF32 split is exact to
 * ~2^-21 and TMEM accumulation truncates, which leaves a relative bias of about 3e-6 of max|C| on same-sign data (about
 * 50 FP32 eps; tests/test_gemm_gpu.py bounds it at 4e-6); an infinite operand element yields NaN in its row / column of
 * the product (inf - inf in the split), where a reference SGEMM would return inf. */
int llacuda_sgemm_dev(char opa, char opb, int64_t m, int64_t n, int64_t k, float alpha,
                      const float* A, int64_t lda, const float* B, int64_t ldb,
                      float beta, float* C, int64_t ldc);
/* complex double (interleaved re,im); alpha2 / beta2 point at {re, im} on the host */
int llacuda_zgemm_dev(char opa, char opb, int64_t m, int64_t n, int64_t k, const double* alpha2,
                      const void* A, int64_t lda, const void* B, int64_t ldb,
                      const double* beta2, void* C, int64_t ldc);
/* complex single (interleaved re,im) on the tcgen05 SGEMM: planar split, four real products, complex combine */
int llacuda_cgemm_dev(char opa, char opb, int64_t m, int64_t n, int64_t k, const float* alpha2,
                      const void* A, int64_t lda, const void* B, int64_t ldb,
                      const float* beta2, void* C, int64_t ldc);
/* factorisations and solves on device-resident column-major operands; ipiv_dev / info_dev are
 * DEVICE pointers (int32), written asynchronously on llacuda_stream() */
int llacuda_dgetrf_dev(int64_t m, int64_t n, double* A, int64_t lda, int* ipiv_dev, int* info_dev);
int llacuda_dgetrs_dev(char trans, int64_t n, int64_t nrhs, const double* A, int64_t lda,
                       const int* ipiv_dev, double* B, int64_t ldb);
int llacuda_dgesv_dev(int64_t n, int64_t nrhs, double* A, int64_t lda, int* ipiv_dev, double* B,
                      int64_t ldb, int* info_dev);
int llacuda_dpotrf_dev(char uplo, int64_t n, double* A, int64_t lda, int* info_dev);
int llacuda_dpotrs_dev(char uplo, int64_t n, int64_t nrhs, const double* A, int64_t lda, double* B,
                       int64_t ldb);
/* op(T) X = B in place, T triangular m x m, B m x n (side = 'L', alpha = 1): LLA's trsm%
 * (reference src/linear-algebra.lisp:333-347); used by the block-cyclic factorisations. */
int llacuda_dtrsm_dev(char uplo, char trans, char diag, int64_t m, int64_t n, const double* T, int64_t ldt,
                      double* B, int64_t ldb);
/* rows i <-> ipiv[i]-1 for i in [k1, k2) (0-based i, 1-based LAPACK pivots) on ncols columns, forward or reverse
 * order: the DLASWP step of GETRF/GETRS (reference call sites src/linear-algebra.lisp:227-234, 269-276). */
int llacuda_dlaswp_dev(int64_t ncols, double* A, int64_t lda, int64_t k1, int64_t k2, const int* ipiv_dev, int forward);
/* batched small solves (BASELINE configs[3]): `batch` independent n x n systems, n <= 32,
 * nrhs <= 8; system i lives at A + i*strideA (column-major, lda = n) and B + i*strideB (ldb = n).
 * X overwrites B; ipiv_dev (batch*n, may be NULL) and info_dev (batch) follow DGESV; A is only
 * overwritten with the LU factors when keep_lu != 0 (LLA's solve discards them). */
int llacuda_dgesv_batched_dev(int n, int nrhs, int64_t batch, double* A, int64_t strideA, int* ipiv_dev,
                              double* B, int64_t strideB, int* info_dev, int keep_lu);
/* out[c + r*ldout] = in[r + c*ldin]: the device-side replacement of LLA's
 * transpose-matrix-to/from-memory (reference src/foreign-memory.lisp:211-255) */
int llacuda_dtranspose_dev(int64_t rows, int64_t cols, const double* in, int64_t ldin, double* out,
                           int64_t ldout);

/* convert (+ transpose) on the device: the replacement of copy-array-to-memory / transpose-matrix-to-memory with their
 * element-wise COERCE (reference src/foreign-memory.lisp:168-255).  Type codes are LLA's (src/types.lisp:22-26: 0 single,
 * 1 double, 2 complex single, 3 complex double) plus 4 = int32 and 5 = int64 sources (integer arrays classify as double,
 * src/types.lisp:69-80).  in: rows x cols column-major (ldin); out[c + r ldout] when transpose, out[r + c ldout]
 * otherwise.  A complex source with a real destination is refused, as the reference does. */
int llacuda_convert_dev(int in_type, int out_type, int transpose, int64_t rows, int64_t cols, const void* in, int64_t ldin,
                        void* out, int64_t ldout);

/* Integer type of the Fortran symbols: 32-bit, or 64-bit when the library is built with -DLLACUDA_ILP64
 * (lla_b200/libllacuda64.so) for an LLA compiled with :int64 (reference src/types.lisp:7-8). */
#ifdef LLACUDA_ILP64
typedef int64_t llacuda_int;
#else
typedef int llacuda_int;
#endif

/* ------------------------------------------------------------------ Fortran drop-ins */
/* reference call sites: mm  src/linear-algebra.lisp:50-54 ; gemm!  src/blas.lisp:56-63 */
void dgemm_(const char* transa, const char* transb, const llacuda_int* m, const llacuda_int* n, const llacuda_int* k,
            const double* alpha, const double* a, const llacuda_int* lda, const double* b, const llacuda_int* ldb,
            const double* beta, double* c, const llacuda_int* ldc);
void llacuda_dgemm(const char* transa, const char* transb, const llacuda_int* m, const llacuda_int* n, const llacuda_int* k,
                   const double* alpha, const double* a, const llacuda_int* lda, const double* b,
                   const llacuda_int* ldb, const double* beta, double* c, const llacuda_int* ldc);

void sgemm_(const char* transa, const char* transb, const llacuda_int* m, const llacuda_int* n, const llacuda_int* k,
            const float* alpha, const float* a, const llacuda_int* lda, const float* b, const llacuda_int* ldb,
            const float* beta, float* c, const llacuda_int* ldc);
void llacuda_sgemm(const char* transa, const char* transb, const llacuda_int* m, const llacuda_int* n, const llacuda_int* k,
                   const float* alpha, const float* a, const llacuda_int* lda, const float* b, const llacuda_int* ldb,
                   const float* beta, float* c, const llacuda_int* ldc);
void cgemm_(const char* transa, const char* transb, const llacuda_int* m, const llacuda_int* n, const llacuda_int* k,
            const void* alpha, const void* a, const llacuda_int* lda, const void* b, const llacuda_int* ldb,
            const void* beta, void* c, const llacuda_int* ldc);
void llacuda_cgemm(const char* transa, const char* transb, const llacuda_int* m, const llacuda_int* n, const llacuda_int* k,
                   const void* alpha, const void* a, const llacuda_int* lda, const void* b, const llacuda_int* ldb,
                   const void* beta, void* c, const llacuda_int* ldc);
void zgemm_(const char* transa, const char* transb, const llacuda_int* m, const llacuda_int* n, const llacuda_int* k,
            const void* alpha, const void* a, const llacuda_int* lda, const void* b, const llacuda_int* ldb,
            const void* beta, void* c, const llacuda_int* ldc);
void llacuda_zgemm(const char* transa, const char* transb, const llacuda_int* m, const llacuda_int* n, const llacuda_int* k,
                   const void* alpha, const void* a, const llacuda_int* lda, const void* b, const llacuda_int* ldb,
                   const void* beta, void* c, const llacuda_int* ldc);
/* lu: src/linear-algebra.lisp:227-234 */
void dgetrf_(const llacuda_int* m, const llacuda_int* n, double* a, const llacuda_int* lda, llacuda_int* ipiv, llacuda_int* info);
/* solve (lu): src/linear-algebra.lisp:269-276 */
void dgetrs_(const char* trans, const llacuda_int* n, const llacuda_int* nrhs, const double* a, const llacuda_int* lda,
             const llacuda_int* ipiv, double* b, const llacuda_int* ldb, llacuda_int* info);
/* solve (array array): src/linear-algebra.lisp:282-289 */
void dgesv_(const llacuda_int* n, const llacuda_int* nrhs, double* a, const llacuda_int* lda, llacuda_int* ipiv, double* b,
            const llacuda_int* ldb, llacuda_int* info);
/* cholesky: src/linear-algebra.lisp:670-674 */
void dpotrf_(const char* uplo, const llacuda_int* n, double* a, const llacuda_int* lda, llacuda_int* info);
/* solve (cholesky): src/linear-algebra.lisp:296-301 */
void dpotrs_(const char* uplo, const llacuda_int* n, const llacuda_int* nrhs, const double* a, const llacuda_int* lda,
             double* b, const llacuda_int* ldb, llacuda_int* info);
void llacuda_dgetrf(const llacuda_int* m, const llacuda_int* n, double* a, const llacuda_int* lda, llacuda_int* ipiv, llacuda_int* info);
void llacuda_dgetrs(const char* trans, const llacuda_int* n, const llacuda_int* nrhs, const double* a, const llacuda_int* lda,
                    const llacuda_int* ipiv, double* b, const llacuda_int* ldb, llacuda_int* info);
void llacuda_dgesv(const llacuda_int* n, const llacuda_int* nrhs, double* a, const llacuda_int* lda, llacuda_int* ipiv, double* b,
                   const llacuda_int* ldb, llacuda_int* info);
void llacuda_dpotrf(const char* uplo, const llacuda_int* n, double* a, const llacuda_int* lda, llacuda_int* info);
void llacuda_dpotrs(const char* uplo, const llacuda_int* n, const llacuda_int* nrhs, const double* a, const llacuda_int* lda,
                    double* b, const llacuda_int* ldb, llacuda_int* info);

/* "next" rows: mm-hermitian% src/linear-algebra.lisp:62-76, solve on a hermitian-matrix
 * :303-315, trsm% :333-347 */
void dsyrk_(const char* uplo, const char* trans, const llacuda_int* n, const llacuda_int* k, const double* alpha,
            const double* a, const llacuda_int* lda, const double* beta, double* c, const llacuda_int* ldc);
void llacuda_dsyrk(const char* uplo, const char* trans, const llacuda_int* n, const llacuda_int* k, const double* alpha,
                   const double* a, const llacuda_int* lda, const double* beta, double* c, const llacuda_int* ldc);
void dtrsm_(const char* side, const char* uplo, const char* transa, const char* diag, const llacuda_int* m,
            const llacuda_int* n, const double* alpha, const double* a, const llacuda_int* lda, double* b,
            const llacuda_int* ldb);
void llacuda_dtrsm(const char* side, const char* uplo, const char* transa, const char* diag, const llacuda_int* m,
                   const llacuda_int* n, const double* alpha, const double* a, const llacuda_int* lda, double* b,
                   const llacuda_int* ldb);
void dposv_(const char* uplo, const llacuda_int* n, const llacuda_int* nrhs, double* a, const llacuda_int* lda, double* b,
            const llacuda_int* ldb, llacuda_int* info);
void llacuda_dposv(const char* uplo, const llacuda_int* n, const llacuda_int* nrhs, double* a, const llacuda_int* lda,
                   double* b, const llacuda_int* ldb, llacuda_int* info);

/* the invert family: invert (lu) src/linear-algebra.lisp:358-367 (LWORK = -1 is the size
 * query of lapack-call-w/query), invert (cholesky) :399-406, invert-triangular% :381-391 */
void dgetri_(const llacuda_int* n, double* a, const llacuda_int* lda, const llacuda_int* ipiv, double* work,
             const llacuda_int* lwork, llacuda_int* info);
void llacuda_dgetri(const llacuda_int* n, double* a, const llacuda_int* lda, const llacuda_int* ipiv, double* work,
                    const llacuda_int* lwork, llacuda_int* info);
void dpotri_(const char* uplo, const llacuda_int* n, double* a, const llacuda_int* lda, llacuda_int* info);
void llacuda_dpotri(const char* uplo, const llacuda_int* n, double* a, const llacuda_int* lda, llacuda_int* info);
void dtrtri_(const char* uplo, const char* diag, const llacuda_int* n, double* a, const llacuda_int* lda, llacuda_int* info);
void llacuda_dtrtri(const char* uplo, const char* diag, const llacuda_int* n, double* a, const llacuda_int* lda,
                    llacuda_int* info);

/* Row-major-native entry points (host pointers; SURVEY section 8b Option B): LLA's row-major arrays as they are, the
 * layout conversion runs on the device (HBM-bound transpose) instead of the element-wise Lisp loops of
 * src/foreign-memory.lisp:211-255, and operands LLA marks input-only are never sent back.  lda / ldb are ROW widths.
 *   llacuda_dgetrf_rm : lu              src/linear-algebra.lisp:225-234   (a: m x n row-major, overwritten by LU)
 *   llacuda_dgesv_rm  : solve (a b)     src/linear-algebra.lisp:278-289   (a input-only, b: n x nrhs row-major -> x)
 *   llacuda_dgetrs_rm : solve (lu b)    src/linear-algebra.lisp:264-276   (lu, ipiv input-only)
 *   llacuda_dpotrs_rm : solve (chol b)  src/linear-algebra.lisp:291-301   (row-major lower factor input-only) */
void llacuda_dgetrf_rm(const llacuda_int* m, const llacuda_int* n, double* a, const llacuda_int* lda, llacuda_int* ipiv,
                       llacuda_int* info);
void llacuda_dgesv_rm(const llacuda_int* n, const llacuda_int* nrhs, const double* a, const llacuda_int* lda, double* b,
                      const llacuda_int* ldb, llacuda_int* info);
void llacuda_dgetrs_rm(const llacuda_int* n, const llacuda_int* nrhs, const double* lu, const llacuda_int* lda,
                       const llacuda_int* ipiv, double* b, const llacuda_int* ldb, llacuda_int* info);
void llacuda_dpotrs_rm(const llacuda_int* n, const llacuda_int* nrhs, const double* l, const llacuda_int* lda, double* b,
                       const llacuda_int* ldb, llacuda_int* info);

/* ------------------------------------------------------------------ single, complex-single and complex-double types
 * LLA's lu / solve / cholesky / mm-hermitian% dispatch on the common float type of their operands through the four-way
 * ECASE of src/fortran-call.lisp:428-439 (call sites src/linear-algebra.lisp:62-76,227-234,269-276,282-289,296-315,670-674):
 * same argument lists as the d-symbols above, complex operands interleaved (re, im).  Pivot rule of the complex types: first
 * maximum of |re| + |im| (IZAMAX / ICAMAX).  HERK takes REAL alpha / beta; LLA hands cells of the matrix type, whose
 * leading real part is what gets read. */
void sgetrf_(const llacuda_int* m, const llacuda_int* n, void* a, const llacuda_int* lda, llacuda_int* ipiv, llacuda_int* info);
void sgetrs_(const char* trans, const llacuda_int* n, const llacuda_int* nrhs, const void* a, const llacuda_int* lda,
             const llacuda_int* ipiv, void* b, const llacuda_int* ldb, llacuda_int* info);
void sgesv_(const llacuda_int* n, const llacuda_int* nrhs, void* a, const llacuda_int* lda, llacuda_int* ipiv, void* b,
            const llacuda_int* ldb, llacuda_int* info);
void spotrf_(const char* uplo, const llacuda_int* n, void* a, const llacuda_int* lda, llacuda_int* info);
void spotrs_(const char* uplo, const llacuda_int* n, const llacuda_int* nrhs, const void* a, const llacuda_int* lda, void* b,
             const llacuda_int* ldb, llacuda_int* info);
void sposv_(const char* uplo, const llacuda_int* n, const llacuda_int* nrhs, void* a, const llacuda_int* lda, void* b,
            const llacuda_int* ldb, llacuda_int* info);
void llacuda_sgetrf(const llacuda_int* m, const llacuda_int* n, void* a, const llacuda_int* lda, llacuda_int* ipiv, llacuda_int* info);
void llacuda_sgetrs(const char* trans, const llacuda_int* n, const llacuda_int* nrhs, const void* a, const llacuda_int* lda,
             const llacuda_int* ipiv, void* b, const llacuda_int* ldb, llacuda_int* info);
void llacuda_sgesv(const llacuda_int* n, const llacuda_int* nrhs, void* a, const llacuda_int* lda, llacuda_int* ipiv, void* b,
            const llacuda_int* ldb, llacuda_int* info);
void llacuda_spotrf(const char* uplo, const llacuda_int* n, void* a, const llacuda_int* lda, llacuda_int* info);
void llacuda_spotrs(const char* uplo, const llacuda_int* n, const llacuda_int* nrhs, const void* a, const llacuda_int* lda, void* b,
             const llacuda_int* ldb, llacuda_int* info);
void llacuda_sposv(const char* uplo, const llacuda_int* n, const llacuda_int* nrhs, void* a, const llacuda_int* lda, void* b,
            const llacuda_int* ldb, llacuda_int* info);
void cgetrf_(const llacuda_int* m, const llacuda_int* n, void* a, const llacuda_int* lda, llacuda_int* ipiv, llacuda_int* info);
void cgetrs_(const char* trans, const llacuda_int* n, const llacuda_int* nrhs, const void* a, const llacuda_int* lda,
             const llacuda_int* ipiv, void* b, const llacuda_int* ldb, llacuda_int* info);
void cgesv_(const llacuda_int* n, const llacuda_int* nrhs, void* a, const llacuda_int* lda, llacuda_int* ipiv, void* b,
            const llacuda_int* ldb, llacuda_int* info);
void cpotrf_(const char* uplo, const llacuda_int* n, void* a, const llacuda_int* lda, llacuda_int* info);
void cpotrs_(const char* uplo, const llacuda_int* n, const llacuda_int* nrhs, const void* a, const llacuda_int* lda, void* b,
             const llacuda_int* ldb, llacuda_int* info);
void cposv_(const char* uplo, const llacuda_int* n, const llacuda_int* nrhs, void* a, const llacuda_int* lda, void* b,
            const llacuda_int* ldb, llacuda_int* info);
void llacuda_cgetrf(const llacuda_int* m, const llacuda_int* n, void* a, const llacuda_int* lda, llacuda_int* ipiv, llacuda_int* info);
void llacuda_cgetrs(const char* trans, const llacuda_int* n, const llacuda_int* nrhs, const void* a, const llacuda_int* lda,
             const llacuda_int* ipiv, void* b, const llacuda_int* ldb, llacuda_int* info);
void llacuda_cgesv(const llacuda_int* n, const llacuda_int* nrhs, void* a, const llacuda_int* lda, llacuda_int* ipiv, void* b,
            const llacuda_int* ldb, llacuda_int* info);
void llacuda_cpotrf(const char* uplo, const llacuda_int* n, void* a, const llacuda_int* lda, llacuda_int* info);
void llacuda_cpotrs(const char* uplo, const llacuda_int* n, const llacuda_int* nrhs, const void* a, const llacuda_int* lda, void* b,
             const llacuda_int* ldb, llacuda_int* info);
void llacuda_cposv(const char* uplo, const llacuda_int* n, const llacuda_int* nrhs, void* a, const llacuda_int* lda, void* b,
            const llacuda_int* ldb, llacuda_int* info);
void zgetrf_(const llacuda_int* m, const llacuda_int* n, void* a, const llacuda_int* lda, llacuda_int* ipiv, llacuda_int* info);
void zgetrs_(const char* trans, const llacuda_int* n, const llacuda_int* nrhs, const void* a, const llacuda_int* lda,
             const llacuda_int* ipiv, void* b, const llacuda_int* ldb, llacuda_int* info);
void zgesv_(const llacuda_int* n, const llacuda_int* nrhs, void* a, const llacuda_int* lda, llacuda_int* ipiv, void* b,
            const llacuda_int* ldb, llacuda_int* info);
void zpotrf_(const char* uplo, const llacuda_int* n, void* a, const llacuda_int* lda, llacuda_int* info);
void zpotrs_(const char* uplo, const llacuda_int* n, const llacuda_int* nrhs, const void* a, const llacuda_int* lda, void* b,
             const llacuda_int* ldb, llacuda_int* info);
void zposv_(const char* uplo, const llacuda_int* n, const llacuda_int* nrhs, void* a, const llacuda_int* lda, void* b,
            const llacuda_int* ldb, llacuda_int* info);
void llacuda_zgetrf(const llacuda_int* m, const llacuda_int* n, void* a, const llacuda_int* lda, llacuda_int* ipiv, llacuda_int* info);
void llacuda_zgetrs(const char* trans, const llacuda_int* n, const llacuda_int* nrhs, const void* a, const llacuda_int* lda,
             const llacuda_int* ipiv, void* b, const llacuda_int* ldb, llacuda_int* info);
void llacuda_zgesv(const llacuda_int* n, const llacuda_int* nrhs, void* a, const llacuda_int* lda, llacuda_int* ipiv, void* b,
            const llacuda_int* ldb, llacuda_int* info);
void llacuda_zpotrf(const char* uplo, const llacuda_int* n, void* a, const llacuda_int* lda, llacuda_int* info);
void llacuda_zpotrs(const char* uplo, const llacuda_int* n, const llacuda_int* nrhs, const void* a, const llacuda_int* lda, void* b,
             const llacuda_int* ldb, llacuda_int* info);
void llacuda_zposv(const char* uplo, const llacuda_int* n, const llacuda_int* nrhs, void* a, const llacuda_int* lda, void* b,
            const llacuda_int* ldb, llacuda_int* info);
void ssyrk_(const char* uplo, const char* trans, const llacuda_int* n, const llacuda_int* k, const float* alpha, const void* a,
            const llacuda_int* lda, const float* beta, void* c, const llacuda_int* ldc);
void cherk_(const char* uplo, const char* trans, const llacuda_int* n, const llacuda_int* k, const float* alpha, const void* a,
            const llacuda_int* lda, const float* beta, void* c, const llacuda_int* ldc);
void zherk_(const char* uplo, const char* trans, const llacuda_int* n, const llacuda_int* k, const double* alpha, const void* a,
            const llacuda_int* lda, const double* beta, void* c, const llacuda_int* ldc);
void llacuda_ssyrk(const char* uplo, const char* trans, const llacuda_int* n, const llacuda_int* k, const float* alpha,
                   const void* a, const llacuda_int* lda, const float* beta, void* c, const llacuda_int* ldc);
void llacuda_cherk(const char* uplo, const char* trans, const llacuda_int* n, const llacuda_int* k, const float* alpha,
                   const void* a, const llacuda_int* lda, const float* beta, void* c, const llacuda_int* ldc);
void llacuda_zherk(const char* uplo, const char* trans, const llacuda_int* n, const llacuda_int* k, const double* alpha,
                   const void* a, const llacuda_int* lda, const double* beta, void* c, const llacuda_int* ldc);

/* ------------------------------------------------------------------ group 4: one problem over the GPUs of one box
 * BASELINE.json north_star: "Large GEMM and blocked LU/Cholesky are partitioned across one 8xB200 box as a 2D block-cyclic
 * layout, broadcasting panels with NCCL over NVLink; batched small solves shard one batch slice per GPU ... The Lisp API
 * stays unchanged".  LLA issues one synchronous foreign-funcall from one process (reference src/fortran-call.lisp:428-439),
 * so the engine runs in that process: ncclCommInitAll over the visible devices, one host thread per GPU.
 *
 *   llacuda_dist_init_all(ndev, P, Q)   ndev <= 0: every visible GPU; P, Q <= 0: most-square grid.  Or set LLACUDA_GPUS=N
 *                                       and the plain Fortran symbols (dgemm_, dgesv_, dgetrf_, dpotrf_, zgemm_) route
 *                                       problems of order >= LLACUDA_DIST_MIN_N (8192) here by themselves.
 *   llacuda_pdgemm ... llacuda_pdgesv   HOST pointers, the argument lists of dgemm_ / dpotrf_ / dposv_ / dpotrs_ /
 *                                       dgetrf_ / dgesv_ / zgemm_ (same INFO and failure conventions).  Grids: GEMM most
 *                                       square, factorisations block columns (1 x N) unless LLACUDA_LU_GRID /
 *                                       LLACUDA_CHOL_GRID / LLACUDA_GEMM_GRID = "PxQ" say otherwise; block size 512.
 *   llacuda_dist_init_rank + *_local    the same rank functions with one rank per PROCESS on block-cyclic operands that
 *                                       already live in HBM (bench.py --gpus N, tests): local part of an Np x Np matrix
 *                                       = column-major Np/P x Np/Q, local block (li, lj) = global block (li P + p, lj Q + q),
 *                                       rank = p Q + q, Np a multiple of nb lcm(P, Q). */
int  llacuda_dist_init_all(int ndev, int P, int Q);
int  llacuda_dist_unique_id(void* id128);                 /* 128 bytes; rank 0 creates, the launcher distributes */
int  llacuda_dist_init_rank(int nranks, int rank, const void* id128, int P, int Q);   /* on the current device */
int  llacuda_dist_set_grid(int P, int Q);                 /* collective over the ranks */
int  llacuda_dist_info(int* nranks, int* P, int* Q, int* single_process, int* rank);
int  llacuda_dist_nccl_version(void);
int  llacuda_dist_finalize(void);
void* llacuda_dist_stream(void);                          /* launcher mode: the stream the rank functions end on */
void llacuda_pdgemm(const char* transa, const char* transb, const llacuda_int* m, const llacuda_int* n, const llacuda_int* k,
                    const double* alpha, const double* a, const llacuda_int* lda, const double* b, const llacuda_int* ldb,
                    const double* beta, double* c, const llacuda_int* ldc);
void llacuda_pzgemm(const char* transa, const char* transb, const llacuda_int* m, const llacuda_int* n, const llacuda_int* k,
                    const void* alpha, const void* a, const llacuda_int* lda, const void* b, const llacuda_int* ldb,
                    const void* beta, void* c, const llacuda_int* ldc);
void llacuda_pdpotrf(const char* uplo, const llacuda_int* n, double* a, const llacuda_int* lda, llacuda_int* info);
void llacuda_pdposv(const char* uplo, const llacuda_int* n, const llacuda_int* nrhs, double* a, const llacuda_int* lda, double* b,
                    const llacuda_int* ldb, llacuda_int* info);
void llacuda_pdpotrs(const char* uplo, const llacuda_int* n, const llacuda_int* nrhs, const double* a, const llacuda_int* lda,
                     double* b, const llacuda_int* ldb, llacuda_int* info);
void llacuda_pdgetrf(const llacuda_int* m, const llacuda_int* n, double* a, const llacuda_int* lda, llacuda_int* ipiv,
                     llacuda_int* info);
void llacuda_pdgesv(const llacuda_int* n, const llacuda_int* nrhs, double* a, const llacuda_int* lda, llacuda_int* ipiv, double* b,
                    const llacuda_int* ldb, llacuda_int* info);
/* BASELINE configs[3]: batch of n x n systems (n <= 32, nrhs <= 8), contiguous [batch][n][n] / [batch][nrhs][n]
 * column-major images on the HOST; one contiguous batch slice per GPU, no collective; ipiv / info int32, may be NULL */
int  llacuda_pdgesv_batched(int n, int nrhs, int64_t batch, double* a, double* b, int* ipiv, int* info, int keep_lu);
int  llacuda_pdgemm_local(int64_t Mp, int64_t Np, int64_t Kp, int64_t nb, double alpha, const double* A, int64_t lda,
                          const double* B, int64_t ldb, double beta, double* C, int64_t ldc);
int  llacuda_pzgemm_local(int64_t Mp, int64_t Np, int64_t Kp, int64_t nb, const double* alpha2, const void* A, int64_t lda,
                          const void* B, int64_t ldb, const double* beta2, void* C, int64_t ldc);
int  llacuda_pdpotrf_local(int64_t Np, int64_t nb, double* A, int64_t lld, int* info);            /* info: host int */
int  llacuda_pdgetrf_local(int64_t Np, int64_t nb, double* A, int64_t lld, int* ipiv_dev, int* info);
int  llacuda_pdgetrs_local(int64_t Np, int64_t nb, const double* LU, int64_t lld, const int* ipiv_dev, double* B, int64_t ldb,
                           int64_t nrhs);                                                         /* B replicated */
/* DGESV in one call (reference call site src/linear-algebra.lisp:282-289): on block columns the forward substitution
   rides along with the factorisation; B (replicated) holds X when *info == 0 and is not meaningful otherwise */
int  llacuda_pdgesv_local(int64_t Np, int64_t nb, double* A, int64_t lld, int* ipiv_dev, double* B, int64_t ldb, int64_t nrhs,
                          int* info);
int  llacuda_pdpotrs_local(int64_t Np, int64_t nb, const double* U, int64_t lld, double* B, int64_t ldb, int64_t nrhs);
/* ownership arithmetic of the layout (pure functions) */
int64_t llacuda_dist_numroc(int64_t n, int64_t nb, int iproc, int nprocs);
int64_t llacuda_dist_padded(int64_t n, int64_t nb, int P, int Q);
int64_t llacuda_dist_l2g(int64_t l, int64_t nb, int iproc, int nprocs);
void llacuda_dist_default_grid(int world, int* P, int* Q);

#ifdef __cplusplus
}
#endif
#endif /* LLACUDA_H */
