// lla_b200/csrc/context.cu -- process-global device context, error channel, scratch memory.
//
// LLA threads no handle through its FFI calls (src/fortran-call.lisp:428-439 issues a bare
// foreign-funcall), so the device state is a lazily-initialised process global guarded by a
// mutex (SURVEY.md section 8b "Threading / reentrancy").
#include "internal.h"
#include "staging.h"
#include "../../include/llacuda.h"

#include <atomic>
#include <cuda.h>
#include <vector>
#include <mutex>
#include <cstdio>
#include <cstring>
#include <cstdlib>

namespace llacuda {

static std::mutex g_mu;
static std::recursive_mutex g_api_mu;
ApiLock::ApiLock() { g_api_mu.lock(); }
ApiLock::~ApiLock() { g_api_mu.unlock(); }
static std::atomic<Context*> g_ctx[MAX_DEV];
static thread_local char g_err[512] = "";
static thread_local int g_err_code = 0;

static std::atomic<unsigned long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

void set_error(const char* file, int line, const char* what, int code) {
  snprintf(g_err, sizeof(g_err), "%s:%d: %s (code %d)", file, line, what, code);
  g_err_code = code ? code : -1;
}

int record_cuda_error(cudaError_t e, const char* file, int line) {
  set_error(file, line, cudaGetErrorString(e), (int)e);
  return LLACUDA_ERR_CUDA_BASE - (int)e;
}

// ---- failures of the BLAS-style entry points (no INFO argument).  netlib's XERBLA prints and STOPs; LLA's mm hands
// a zeroed C with beta = 0, so a failure that merely returned would look like an all-zero product.
static std::atomic<int> g_err_mode{-1};
static std::atomic<void (*)(const char*, int, const char*)> g_err_handler{nullptr};
static int err_mode() {
  int m = g_err_mode.load();
  if (m < 0) {
    const char* e = getenv("LLACUDA_ON_ERROR");
    m = (e && (e[0] == 'r' || e[0] == 'R')) ? 1 : 0;
    g_err_mode.store(m);
  }
  return m;
}
void blas_failure(const char* routine, int rc, void* out, int64_t ld, int64_t rows, int64_t cols, size_t es) {
  drain_streams();
  if (g_err_code == 0) set_error(__FILE__, __LINE__, "failure in a BLAS-style entry point", rc);
  fprintf(stderr, " ** llacuda: %s failed (code %d): %s\n", routine, rc, g_err);
  if (auto h = g_err_handler.load()) h(routine, rc, g_err);
  if (err_mode() == 0) {
    fprintf(stderr, " ** llacuda: stopping, as XERBLA would (llacuda_set_error_mode(1) or LLACUDA_ON_ERROR=return: "
                    "poison the output with NaN, record the error and return)\n");
    abort();
  }
  if (out)   // all-ones bytes are a NaN in every IEEE format: the result cannot be mistaken for a product
    for (int64_t j = 0; j < cols; j++) memset((char*)out + (size_t)j * ld * es, 0xff, (size_t)rows * es);
}

int cur_dev() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
  return dev;
}

Context* ctx() {
  const int dev = cur_dev();
  if (dev < 0 || dev >= MAX_DEV) {
    set_error(__FILE__, __LINE__, "no CUDA device: llacuda has no CPU fallback", -1);
    return nullptr;
  }
  if (Context* c = g_ctx[dev].load(std::memory_order_acquire)) return c;
  std::lock_guard<std::mutex> lk(g_mu);
  if (Context* c = g_ctx[dev].load(std::memory_order_acquire)) return c;
  Context* c = new Context();
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    set_error(__FILE__, __LINE__, "cudaGetDeviceProperties failed", -1);
    delete c;
    return nullptr;
  }
  if (prop.major != 10) {
    set_error(__FILE__, __LINE__, "llacuda is built for sm_100a (B200) only", prop.major * 10 + prop.minor);
    delete c;
    return nullptr;
  }
  c->device = dev;
  c->sm_count = prop.multiProcessorCount;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, lo);
  cudaStreamCreateWithPriority(&c->aux, cudaStreamNonBlocking, hi);  // panels ahead of trailing updates
  cudaStreamCreateWithPriority(&c->aux2, cudaStreamNonBlocking, hi);
  cudaStreamCreateWithPriority(&c->aux3, cudaStreamNonBlocking, hi);
  cudaStreamCreateWithPriority(&c->aux4, cudaStreamNonBlocking, lo);
  cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking);
  for (auto& e : c->ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  for (auto& row : c->ev_chunk)
    for (auto& e : row) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  g_ctx[dev].store(c, std::memory_order_release);
  return c;
}

// ---- SM partition (green contexts), one per device
static SmPartition g_part[MAX_DEV];
const SmPartition* sm_partition() {
  const int dev = cur_dev();
  if (dev < 0 || dev >= MAX_DEV || !ctx()) return nullptr;
  SmPartition& P = g_part[dev];
  std::lock_guard<std::mutex> lk(g_mu);
  if (P.tried) return P.ok ? &P : nullptr;
  P.tried = true;
  if (getenv("LLACUDA_NO_PARTITION")) return nullptr;
  // Nsight Compute cannot prepare a cooperative launch on a green-context stream ("Failed to prepare kernel for
  // profiling", and it then shuts the application down): under its injection library the drivers keep their priority
  // streams (measured cost of that: 1-2 % of dgetrf)
  if (FILE* f = fopen("/proc/self/maps", "r")) {
    char line[512];
    bool profiled = false;
    while (!profiled && fgets(line, sizeof line, f))
      profiled = strstr(line, "libcuda-injection") != nullptr || strstr(line, "nsight-compute") != nullptr ||
                 strstr(line, "nsight_compute") != nullptr;
    fclose(f);
    if (profiled) return nullptr;
  }
  int want = 24;   // 3 panel CTAs of 128 rows per SM: LU panels up to 9216 rows fit (measured: profiles/r02_sm_partition.md)
  if (const char* e = getenv("LLACUDA_PANEL_SMS")) want = atoi(e);
  if (want < 8) return nullptr;
  // the driver API is resolved at run time (cudaGetDriverEntryPoint): the library must load on a machine without libcuda
  typedef CUresult (*DevGet)(CUdevice*, int);
  typedef CUresult (*GetRes)(CUdevice, CUdevResource*, CUdevResourceType);
  typedef CUresult (*Split)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
  typedef CUresult (*GenDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int);
  typedef CUresult (*GreenCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
  typedef CUresult (*GreenStream)(CUstream*, CUgreenCtx, unsigned int, int);
  auto entry = [](const char* name) -> void* {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); return nullptr; }
    return fn;
  };
  DevGet dev_get = (DevGet)entry("cuDeviceGet");
  GetRes get_res = (GetRes)entry("cuDeviceGetDevResource");
  Split split = (Split)entry("cuDevSmResourceSplitByCount");
  GenDesc gen_desc = (GenDesc)entry("cuDevResourceGenerateDesc");
  GreenCreate green_create = (GreenCreate)entry("cuGreenCtxCreate");
  GreenStream green_stream = (GreenStream)entry("cuGreenCtxStreamCreate");
  if (!dev_get || !get_res || !split || !gen_desc || !green_create || !green_stream) return nullptr;
  CUdevice cudev;
  CUdevResource all, grp, rem;
  unsigned int ngroups = 1;
  CUdevResourceDesc dp = nullptr, dg = nullptr;
  CUgreenCtx gp = nullptr, gg = nullptr;
  CUstream sp = nullptr, sg = nullptr, sg2 = nullptr, sg3 = nullptr;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  bool ok = dev_get(&cudev, dev) == CUDA_SUCCESS &&
            get_res(cudev, &all, CU_DEV_RESOURCE_TYPE_SM) == CUDA_SUCCESS &&
            split(&grp, &ngroups, &all, &rem, 0, (unsigned)want) == CUDA_SUCCESS && ngroups == 1 &&
            rem.sm.smCount >= 64 &&
            gen_desc(&dp, &grp, 1) == CUDA_SUCCESS && gen_desc(&dg, &rem, 1) == CUDA_SUCCESS &&
            green_create(&gp, dp, cudev, CU_GREEN_CTX_DEFAULT_STREAM) == CUDA_SUCCESS &&
            green_create(&gg, dg, cudev, CU_GREEN_CTX_DEFAULT_STREAM) == CUDA_SUCCESS &&
            green_stream(&sp, gp, CU_STREAM_NON_BLOCKING, hi) == CUDA_SUCCESS &&
            green_stream(&sg, gg, CU_STREAM_NON_BLOCKING, lo) == CUDA_SUCCESS &&
            green_stream(&sg2, gg, CU_STREAM_NON_BLOCKING, hi) == CUDA_SUCCESS &&
            green_stream(&sg3, gg, CU_STREAM_NON_BLOCKING, hi) == CUDA_SUCCESS;
  cudaGetLastError();
  if (!ok) return nullptr;
  P.panel = (cudaStream_t)sp; P.bulk = (cudaStream_t)sg; P.bulk2 = (cudaStream_t)sg2; P.bulk3 = (cudaStream_t)sg3;
  P.sm_panel = (int)grp.sm.smCount; P.sm_bulk = (int)rem.sm.smCount;
  P.ok = true;
  if (getenv("LLACUDA_TRACE")) fprintf(stderr, "llacuda: SM partition %d (panel chain) + %d (trailing update)\n", P.sm_panel, P.sm_bulk);
  return &P;
}

void drain_streams() {
  const int dev = cur_dev();
  if (dev < 0 || dev >= MAX_DEV) return;
  Context* c = g_ctx[dev].load(std::memory_order_acquire);
  if (!c) return;
  cudaStreamSynchronize(c->stream);
  cudaStreamSynchronize(c->aux);
  cudaStreamSynchronize(c->aux2);
  cudaStreamSynchronize(c->aux3);
  cudaStreamSynchronize(c->aux4);
  cudaStreamSynchronize(c->h2d);
  cudaStreamSynchronize(c->d2h);
  if (g_part[dev].ok) {
    cudaStreamSynchronize(g_part[dev].panel);
    cudaStreamSynchronize(g_part[dev].bulk);
    cudaStreamSynchronize(g_part[dev].bulk2);
    cudaStreamSynchronize(g_part[dev].bulk3);
  }
  cudaGetLastError();
}

int Arena::reserve(size_t bytes, cudaStream_t st) {
  if (!ctx()) return LLACUDA_ERR_NO_DEVICE;
  const size_t want = bytes + 4096;
  base = (char*)stream_scratch(st, want);
  if (!base) {
    set_error(__FILE__, __LINE__, "out of device memory for the per-call scratch", (int)(want >> 20));
    return LLACUDA_ERR_CUDA_BASE - (int)cudaErrorMemoryAllocation;
  }
  cap = want;
  used = 0;
  return 0;
}
void* Arena::take(size_t bytes) {
  size_t a = (used + 255) & ~size_t(255);
  if (a + bytes > cap) return nullptr;
  used = a + bytes;
  return base + a;
}

// ---- event trace (diagnostics only)
struct TraceRec { cudaEvent_t ev; cudaStream_t st; const char* label; long long a, b; };
static std::vector<TraceRec> g_trace;
static int g_trace_on = -1;
bool trace_on() {
  if (g_trace_on < 0) g_trace_on = getenv("LLACUDA_TRACE") ? 1 : 0;
  return g_trace_on == 1;
}
void trace_mark(cudaStream_t st, const char* label, long long a, long long b) {
  if (!trace_on()) return;
  TraceRec r{nullptr, st, label, a, b};
  if (cudaEventCreate(&r.ev) != cudaSuccess) return;
  cudaEventRecord(r.ev, st);
  g_trace.push_back(r);
}

struct StreamBuf { void* p; size_t bytes; };
static std::vector<std::pair<cudaStream_t, StreamBuf>> g_stream_scratch;
static void* scratch_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  // the LU panel's exchange area must not start with stray tags
  if (cudaMemset(p, 0, bytes) != cudaSuccess) { cudaGetLastError(); cudaFree(p); return nullptr; }
  return p;
}
void* stream_scratch(cudaStream_t st, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& e : g_stream_scratch) {
    if (e.first != st) continue;
    if (e.second.bytes >= bytes) return e.second.p;
    cudaStreamSynchronize(st);   // the only user of this buffer is work queued on st
    cudaFree(e.second.p);
    size_t want = bytes + bytes / 2;
    e.second = {scratch_alloc(want), want};
    if (!e.second.p) e.second.bytes = 0;
    return e.second.p;
  }
  size_t want = bytes < (size_t(1) << 20) ? (size_t(1) << 20) : bytes;
  StreamBuf b{scratch_alloc(want), want};
  if (!b.p) return nullptr;
  g_stream_scratch.push_back({st, b});
  return b.p;
}

static thread_local const RowsFinalHook* g_rows_hook = nullptr;
void set_rows_final_hook(const RowsFinalHook* h) { g_rows_hook = h; }
const RowsFinalHook* rows_final_hook() { return g_rows_hook; }
static thread_local const RangeReadyHook* g_ready_hook = nullptr;
void set_range_ready_hook(const RangeReadyHook* h) { g_ready_hook = h; }
const RangeReadyHook* range_ready_hook() { return g_ready_hook; }

int DevBuf::alloc(size_t n) {
  if (n == 0) n = 16;
  cudaError_t e = cudaMalloc(&p, n);
  if (e != cudaSuccess) { p = nullptr; return record_cuda_error(e, __FILE__, __LINE__); }
  bytes = n;
  return 0;
}
DevBuf::~DevBuf() { if (p) cudaFree(p); }

}  // namespace llacuda

using namespace llacuda;

extern "C" {

const char* llacuda_last_error(void) { return g_err; }
int llacuda_last_error_code(void) { return g_err_code; }
void llacuda_clear_error(void) { g_err[0] = 0; g_err_code = 0; }

void llacuda_set_error_mode(int mode) { g_err_mode.store(mode ? 1 : 0); }
int llacuda_error_mode(void) { return err_mode(); }
void llacuda_set_error_handler(void (*handler)(const char* routine, int code, const char* message)) { g_err_handler.store(handler); }

unsigned long long llacuda_launch_count(void) { return g_launches.load(); }

int llacuda_set_device(int device) {
  // the classic (single-GPU) entry points run on the calling thread's current device; contexts are per device
  LLA_CUDA(cudaSetDevice(device));
  return 0;
}

int llacuda_init(void) { return ctx() ? 0 : LLACUDA_ERR_NO_DEVICE; }

int llacuda_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem) {
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  LLA_CUDA(cudaGetDeviceProperties(&prop, c->device));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (total_mem) *total_mem = prop.totalGlobalMem;
  return 0;
}

int llacuda_set_stream(void* cuda_stream) {
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  if (!c->own_stream) c->own_stream = c->stream;
  c->stream = (cudaStream_t)cuda_stream;  // NULL is the legacy default stream, a valid choice
  return 0;
}
int llacuda_reset_stream(void) {
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  if (c->own_stream) c->stream = c->own_stream;
  return 0;
}

void* llacuda_stream(void) { Context* c = ctx(); return c ? (void*)c->stream : nullptr; }

int llacuda_synchronize(void) {
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  LLA_CUDA(cudaStreamSynchronize(c->stream));
  LLA_CUDA(cudaStreamSynchronize(c->aux));
  LLA_CUDA(cudaStreamSynchronize(c->aux2));
  LLA_CUDA(cudaStreamSynchronize(c->aux3));
  LLA_CUDA(cudaStreamSynchronize(c->aux4));
  return 0;
}

void llacuda_trace_enable(int on) { g_trace_on = on ? 1 : 0; }
// prints "t_ms stream label a b" for every mark since the last dump, relative to the first mark
void llacuda_trace_dump(void) {
  cudaDeviceSynchronize();
  for (size_t i = 0; i < g_trace.size(); i++) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g_trace[0].ev, g_trace[i].ev);
    fprintf(stderr, "TRACE %10.3f s%03x %-14s %lld %lld\n", ms, (unsigned)(((uintptr_t)g_trace[i].st >> 4) & 0xfff),
            g_trace[i].label, g_trace[i].a, g_trace[i].b);
  }
  for (auto& r : g_trace) cudaEventDestroy(r.ev);
  g_trace.clear();
}

int llacuda_last_timing(float* ms4) {
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  memcpy(ms4, c->last_ms, sizeof(float) * 4);
  return 0;
}

// ---- memory helpers for callers that want device-resident or pinned operands
int llacuda_malloc(void** p, size_t bytes) {
  if (!ctx()) return LLACUDA_ERR_NO_DEVICE;
  LLA_CUDA(cudaMalloc(p, bytes ? bytes : 16));
  return 0;
}
int llacuda_free(void* p) { LLA_CUDA(cudaFree(p)); return 0; }
int llacuda_host_alloc(void** p, size_t bytes) {
  if (!ctx()) return LLACUDA_ERR_NO_DEVICE;
  LLA_CUDA(cudaHostAlloc(p, bytes ? bytes : 16, cudaHostAllocDefault));
  return 0;
}
int llacuda_trim(void) { return pool_trim(); }
int llacuda_host_free(void* p) { LLA_CUDA(cudaFreeHost(p)); return 0; }
int llacuda_memcpy_h2d(void* dst, const void* src, size_t bytes) {
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  LLA_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  LLA_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
int llacuda_memcpy_d2h(void* dst, const void* src, size_t bytes) {
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  LLA_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  LLA_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
int llacuda_memset(void* dst, int byte, size_t bytes) {
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  LLA_CUDA(cudaMemsetAsync(dst, byte, bytes, c->stream));
  return 0;
}

}  // extern "C"
