// lla_b200/csrc/internal.h -- internal declarations shared by the llacuda translation units.
// Nothing in here crosses the C ABI; the public surface is include/llacuda.h.
#pragma once
#include <cuda_runtime.h>
#include <cuComplex.h>
#include <stdint.h>
#include <stddef.h>

// Integer type of the Fortran-symbol layer: LLA passes 32-bit integers unless it was compiled with :int64
// (reference src/types.lisp:7-8, src/foreign-memory.lisp:35-40); -DLLACUDA_ILP64 builds libllacuda64.so for that mode.
#ifdef LLACUDA_ILP64
typedef int64_t fint;
#else
typedef int fint;
#endif

namespace llacuda {

// ------------------------------------------------------------------ errors
// CUDA failures never masquerade as LAPACK INFO codes:
// they are recorded here and surfaced through llacuda_last_error().
void set_error(const char* file, int line, const char* what, int code);
int  record_cuda_error(cudaError_t e, const char* file, int line);

// A BLAS-style entry point (void, no INFO) failed on the device side: drain the streams, call the installed handler,
// then stop like XERBLA (default) or -- error mode 1 -- poison the rows x cols output with NaN and return.
void blas_failure(const char* routine, int rc, void* out, int64_t ld, int64_t rows, int64_t cols, size_t es);

#define LLA_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t _e = (call);                                                  \
    if (_e != cudaSuccess) return ::llacuda::record_cuda_error(_e, __FILE__, __LINE__); \
  } while (0)

#define LLA_TRY(call)                     \
  do {                                    \
    int _r = (call);                      \
    if (_r != 0) return _r;               \
  } while (0)

// One process-global context and per-call arena: calls entering through the C ABI are serialised (LLA itself is
// single-threaded, but user Lisp threads may call concurrently: SURVEY.md section 8b "Threading / reentrancy").
// Recursive: the llacuda_<name> aliases call the Fortran symbols.
struct ApiLock {
  ApiLock();
  ~ApiLock();
};

// ------------------------------------------------------------------ context
constexpr int MAX_DEV = 16;
int cur_dev();                        // cudaGetDevice of the calling thread (-1 without a device)

struct Context {
  int device = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;     // main compute stream
  cudaStream_t aux = nullptr;        // panel / look-ahead stream
  cudaStream_t aux4 = nullptr;       // normal priority: right-hand sides riding along with a factorisation
  cudaStream_t aux3 = nullptr;       // third high-priority stream: the part of a block row the panel chain needs next
  cudaStream_t aux2 = nullptr;       // second high-priority stream: block-row solves beside the trailing update
  cudaStream_t h2d = nullptr;        // staging streams
  cudaStream_t d2h = nullptr;
  cudaEvent_t ev[16] = {};
  cudaEvent_t ev_chunk[2][32] = {};  // per column chunk of a blocked update: [0] solve done, [1] product done (lu_blocked)
  cudaStream_t own_stream = nullptr; // the library's stream while a caller-owned one is adopted (llacuda_set_stream)
  // timing of the last host-pointer call (ms): [0]=h2d [1]=compute [2]=d2h [3]=total
  float last_ms[4] = {};
};
// Spatial split of the SMs (CUDA green contexts) for the chain-bound phase of the blocked factorisations: the panel chain
// (dozens of short dependent kernels per block step) gets `sm_panel` SMs of its own and the trailing update runs on the
// rest, so a panel kernel never queues behind 75-150 us GEMM CTAs (measured: panels take 2-3x longer under a concurrent
// update, profiles/r01b_trace_*.log).  Streams created in a green context only ever run on its SMs.  Optional: when
// the driver refuses (or LLACUDA_NO_PARTITION is set) the drivers keep their priority streams.
struct SmPartition {
  bool tried = false, ok = false;
  cudaStream_t panel = nullptr;    // on the small partition, high priority
  cudaStream_t bulk = nullptr;     // on the large partition
  cudaStream_t bulk2 = nullptr;    // second stream on the large partition, high priority (block-row solves; interchanges of finished columns)
  cudaStream_t bulk3 = nullptr;    // third stream on the large partition, high priority (block-row solves beside the update)
  int sm_panel = 0, sm_bulk = 0;
};
const SmPartition* sm_partition();   // of the current device; nullptr when unavailable

// Lazily initialised, one per device: the context of the calling thread's current device.  The classic entry
// points run on whatever device is current when they are first called (llacuda_set_device); the multi-GPU engine
// (dist.cu) drives one thread per device, each of which sees its own context here.
Context* ctx();

// Scratch that belongs to one stream (grow-only, zero-filled when it is (re)allocated): every user is work queued on
// that stream, so successive calls may alias it without any event.  Scratch of different streams never aliases,
// which is what lets a panel be factored on one stream while another stream solves or updates.
void* stream_scratch(cudaStream_t st, size_t bytes);

// Bump allocator over the scratch of ONE stream for ONE top-level call: reserve() once with the total (may grow the
// stream's scratch: synchronises that stream), then take() pieces (256-byte aligned).
struct Arena {
  char* base = nullptr;
  size_t cap = 0, used = 0;
  int reserve(size_t bytes, cudaStream_t st);
  void* take(size_t bytes);
};

// cudaFuncSetAttribute is per device: `static DeviceOnce once; if (once.first()) { ...set attributes... }`
struct DeviceOnce {
  bool done[MAX_DEV] = {};
  bool first() {
    int d = cur_dev();
    if (d < 0 || d >= MAX_DEV || done[d]) return false;
    done[d] = true;
    return true;
  }
};

// synchronise every stream of the current device's context (error paths of the host-pointer entry points: nothing
// may still read or write the caller's buffers, or the pooled device blocks, after the call has returned)
void drain_streams();

// number of llacuda kernels launched by this process (bench.py reports it as gpu_launches)
void count_launch(int n = 1);

// Event trace of the blocked drivers (diagnostics; LLACUDA_TRACE=1): trace_mark records a timing event
// on `st`; llacuda_trace_dump() prints every mark's time relative to the first one.
void trace_mark(cudaStream_t st, const char* label, long long a = 0, long long b = 0);
bool trace_on();

// device scratch allocations that live for one call
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int alloc(size_t n);
  ~DevBuf();
};

// ------------------------------------------------------------------ GEMM (gemm_f64.cu)
enum GemmTri { TRI_FULL = 0, TRI_UPPER = 1, TRI_LOWER = 2 };
// op codes: 0 = N, 1 = T, 2 = C (== T for real types)
int dgemm_dev(int opa, int opb, int64_t m, int64_t n, int64_t k, double alpha,
              const double* A, int64_t lda, const double* B, int64_t ldb,
              double beta, double* C, int64_t ldc, int tri, cudaStream_t st,
              const int* abort_flag = nullptr, int hint = 0);
// hint 1: m <= 128 and C aliases B (in-place product): force a single row of 128-row tiles so
// that every column slab of B/C is read and then written by exactly one CTA.
// hint 2 + r: persistent 128x128 tiles on (SM count - 8 - r) CTAs, one per SM, so that the kernels of
// a concurrent high-priority panel stream always find free SMs (look-ahead in the blocked factorisations).
enum GemmHint { GEMM_AUTO = 0, GEMM_INPLACE_B = 1, GEMM_RESERVE_SMS = 2 };
int zgemm_dev(int opa, int opb, int64_t m, int64_t n, int64_t k, cuDoubleComplex alpha,
              const cuDoubleComplex* A, int64_t lda, const cuDoubleComplex* B, int64_t ldb,
              cuDoubleComplex beta, cuDoubleComplex* C, int64_t ldc, int tri, cudaStream_t st);

// SGEMM on tcgen05 / TMEM / TMA (gemm_f32.cu); lda, ldb multiples of 4, operands 16-byte aligned
int sgemm_dev(int opa, int opb, int64_t m, int64_t n, int64_t k, float alpha, const float* A, int64_t lda,
              const float* B, int64_t ldb, float beta, float* C, int64_t ldc, cudaStream_t st);

// FP32-FMA tile kernel (xfactor.cu): any alignment, reference-SGEMM accuracy; the fallback of sgemm_dev for operands TMA cannot take
int sgemm_simt_dev(int opa, int opb, int64_t m, int64_t n, int64_t k, float alpha, const float* A, int64_t lda,
                   const float* B, int64_t ldb, float beta, float* C, int64_t ldc, cudaStream_t st);

// ------------------------------------------------------------------ TRSM (trsm_f64.cu), left side only
// solves op(T) X = alpha B in place; T is m x m triangular, B is m x n
int dtrsm_left_dev(bool upper, int op, bool unit, int64_t m, int64_t n,
                   const double* T, int64_t ldt, double* B, int64_t ldb, cudaStream_t st,
                   const int* abort_flag = nullptr);

// inverse-based solve: Tinv holds the inverses of the IB x IB diagonal blocks of T (block b at
// Tinv + b*IB*IB, leading dimension IB), produced by dtrtri_blocks_dev or by the factorisations
constexpr int IB = 128;
int dtrtri_blocks_dev(bool upper, bool unit, int64_t m, const double* T, int64_t ldt, double* Tinv,
                      cudaStream_t st, const int* abort_flag = nullptr);
int dtrsm_left_inv_dev(bool upper, int op, int64_t m, int64_t n, const double* T, int64_t ldt,
                       const double* Tinv, double* B, int64_t ldb, cudaStream_t st,
                       const int* abort_flag = nullptr);
inline size_t tinv_bytes(int64_t m) { return (size_t)((m + IB - 1) / IB) * IB * IB * sizeof(double); }

// ------------------------------------------------------------------ factorisations
// Optional per-call hook of the blocked drivers: rows [r0, r1) of the factor are final once the work queued so
// far on streams `a` and `b` (b may be null) has completed.  The host-pointer entry points use it to send
// finished block rows back over PCIe while the factorisation is still running.
struct RowsFinalHook {
  int (*fn)(void* user, int64_t r0, int64_t r1, cudaStream_t a, cudaStream_t b);
  void* user;
};
// The mirror image on the way in: columns (dgetrf) or rows of the upper triangle (dpotrf 'U') [lo, hi) are about to be
// touched for the first time by work queued on `st`.  The host-pointer entry points upload the matrix in blocks behind
// this hook, so the transfer of the rest overlaps the first block step (its update is separable by column / row block).
struct RangeReadyHook {
  int (*fn)(void* user, int64_t lo, int64_t hi, cudaStream_t st);
  void* user;
};
void set_range_ready_hook(const RangeReadyHook* h);   // nullptr clears; thread-local, valid for the next dgetrf_dev / dpotrf_dev
const RangeReadyHook* range_ready_hook();
constexpr int64_t READY_BLOCK = 2048;                  // granularity the drivers announce ranges in
void set_rows_final_hook(const RowsFinalHook* h);   // nullptr clears; valid for the next dgetrf_dev / dpotrf_dev only
const RowsFinalHook* rows_final_hook();
// Right-hand sides that ride along with a blocked factorisation: after each panel its interchanges, the solve with the
// panel's unit-lower triangle and the product with the rows below are applied to B on a side stream, so that when
// dgetrf_dev returns B holds L^-1 P B and only the back substitution is left.  `done` tells whether the driver took the
// rider (the unblocked path for small orders does not).
struct LuRhs { double* B = nullptr; int64_t ldb = 0, nrhs = 0; bool done = false; };
int dgetrf_dev(int64_t m, int64_t n, double* A, int64_t lda, int* ipiv, int* info_dev, cudaStream_t st, LuRhs* rhs = nullptr);
// DGESV on device pointers: factorisation with the right-hand sides riding along, back substitution, B restored when INFO > 0
int dgesv_dev(int64_t n, int64_t nrhs, double* A, int64_t lda, int* ipiv, double* B, int64_t ldb, int* info_dev, cudaStream_t st);
// rows i in [k1,k2) swapped with ipiv[i]-1 (1-based LAPACK pivots), forward or backward order
int dlaswp_dev(int64_t ncols, double* A, int64_t lda, int64_t k1, int64_t k2, const int* ipiv,
               bool forward, cudaStream_t st, const int* abort_flag = nullptr);
int dgetrs_dev(int op, int64_t n, int64_t nrhs, const double* A, int64_t lda, const int* ipiv,
               double* B, int64_t ldb, cudaStream_t st, const int* abort_flag = nullptr);
int dpotrf_dev(bool upper, int64_t n, double* A, int64_t lda, int* info_dev, cudaStream_t st);
int dpotrs_dev(bool upper, int64_t n, int64_t nrhs, const double* A, int64_t lda,
               double* B, int64_t ldb, cudaStream_t st, const int* abort_flag = nullptr);

// batched small solves (batched_f64.cu): one warp per n <= 32 system
int dgesv_batched_dev(int n, int nrhs, int64_t batch, double* A, int64_t strideA, int* ipiv, double* B,
                      int64_t strideB, int* info, int keep_lu, cudaStream_t st);

// ------------------------------------------------------------------ multi-GPU routing (dist_abi.cu)
// true when LLACUDA_GPUS=N asked for the engine and the problem is large enough: the plain Fortran symbols then
// partition the call over the GPUs of the box (the Lisp side does not change)
bool dist_route(int64_t n);
int dist_dgemm(int oa, int ob, int64_t m, int64_t n, int64_t k, double alpha, const double* a, int64_t lda, const double* b,
               int64_t ldb, double beta, double* c, int64_t ldc);
int dist_dgesv(int64_t n, int64_t nrhs, double* a, int64_t lda, fint* ipiv, double* b, int64_t ldb, int* info);
int dist_dpotrf(int64_t n, double* a, int64_t lda, int* info);

// ------------------------------------------------------------------ layout kernels (transpose.cu)
int dtranspose_dev(int64_t rows, int64_t cols, const double* in, int64_t ldin, double* out, int64_t ldout,
                   cudaStream_t st);   // out[c + r*ldout] = in[r + c*ldin]
int dtranspose_inplace_dev(int64_t n, double* A, int64_t lda, cudaStream_t st);

}  // namespace llacuda
