// lla_b200/csrc/context.cu -- process-global device context, error channel, scratch memory.
//
// LLA threads no handle through its FFI calls (src/fortran-call.lisp:428-439 issues a bare
// foreign-funcall), so the device state is a lazily-initialised process global guarded by a
// mutex (SURVEY.md section 8b "Threading / reentrancy").
#include "internal.h"
#include "staging.h"
#include "../../include/llacuda.h"

#include <atomic>
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
static Context* g_ctx = nullptr;
static thread_local char g_err[512] = "";
static thread_local int g_err_code = 0;

static std::atomic<unsigned long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

void set_error(const char* file, int line, const char* what, int code) {
  snprintf(g_err, sizeof(g_err), "%s:%d: %s (code %d)", file, line, what, code);
  g_err_code = code;
}

int record_cuda_error(cudaError_t e, const char* file, int line) {
  set_error(file, line, cudaGetErrorString(e), (int)e);
  return LLACUDA_ERR_CUDA_BASE - (int)e;
}

Context* ctx() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_ctx) return g_ctx;
  Context* c = new Context();
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error(__FILE__, __LINE__, "no CUDA device: llacuda has no CPU fallback", -1);
    delete c;
    return nullptr;
  }
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
  cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking);
  for (auto& e : c->ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  g_ctx = c;
  return g_ctx;
}

int ensure_workspace(size_t bytes) {
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  if (c->ws_bytes >= bytes) return 0;
  if (c->ws) {
    LLA_CUDA(cudaDeviceSynchronize());
    LLA_CUDA(cudaFree(c->ws));
    c->ws = nullptr;
    c->ws_bytes = 0;
  }
  size_t want = bytes < (size_t(8) << 20) ? (size_t(8) << 20) : bytes;
  LLA_CUDA(cudaMalloc(&c->ws, want));
  LLA_CUDA(cudaMemset(c->ws, 0, want));  // the panel exchange area must not start with stray tags
  c->ws_bytes = want;
  return 0;
}

int Arena::reserve(size_t bytes) {
  LLA_TRY(ensure_workspace(bytes + 4096));
  base = (char*)ctx()->ws;
  cap = ctx()->ws_bytes;
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
void* stream_scratch(cudaStream_t st, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& e : g_stream_scratch) {
    if (e.first != st) continue;
    if (e.second.bytes >= bytes) return e.second.p;
    cudaStreamSynchronize(st);   // the only user of this buffer is work queued on st
    cudaFree(e.second.p);
    e.second = {nullptr, 0};
    if (cudaMalloc(&e.second.p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    e.second.bytes = bytes;
    return e.second.p;
  }
  StreamBuf b{nullptr, bytes};
  if (cudaMalloc(&b.p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  g_stream_scratch.push_back({st, b});
  return b.p;
}

static const RowsFinalHook* g_rows_hook = nullptr;
void set_rows_final_hook(const RowsFinalHook* h) { g_rows_hook = h; }
const RowsFinalHook* rows_final_hook() { return g_rows_hook; }

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

unsigned long long llacuda_launch_count(void) { return g_launches.load(); }

int llacuda_set_device(int device) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_ctx && g_ctx->device != device) {
      set_error(__FILE__, __LINE__, "llacuda_set_device: context already bound to another device", g_ctx->device);
      return LLACUDA_ERR_BAD_ARG;
    }
  }
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

static cudaStream_t g_own_stream = nullptr;
int llacuda_set_stream(void* cuda_stream) {
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  if (!g_own_stream) g_own_stream = c->stream;
  c->stream = (cudaStream_t)cuda_stream;  // NULL is the legacy default stream, a valid choice
  return 0;
}
int llacuda_reset_stream(void) {
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  if (g_own_stream) c->stream = g_own_stream;
  return 0;
}

void* llacuda_stream(void) { Context* c = ctx(); return c ? (void*)c->stream : nullptr; }

int llacuda_synchronize(void) {
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  LLA_CUDA(cudaStreamSynchronize(c->stream));
  LLA_CUDA(cudaStreamSynchronize(c->aux));
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
