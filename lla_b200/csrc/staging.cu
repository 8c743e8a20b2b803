// lla_b200/csrc/staging.cu -- see staging.h.
#include "staging.h"
#include "hostpipe.h"
#include "../../include/llacuda.h"

#include <cstdlib>
#include <mutex>
#include <vector>

namespace llacuda {

namespace {
struct Block { void* p; size_t bytes; bool used; int dev; };
std::mutex g_pool_mu;
std::vector<Block> g_pool;
}  // namespace

int PoolBuf::alloc(size_t n) {
  if (n == 0) n = 16;
  const int dev = cur_dev();
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    int best = -1;
    for (int i = 0; i < (int)g_pool.size(); i++)
      if (!g_pool[i].used && g_pool[i].dev == dev && g_pool[i].bytes >= n && (best < 0 || g_pool[i].bytes < g_pool[best].bytes)) best = i;
    if (best >= 0 && g_pool[best].bytes <= 2 * n + (size_t(1) << 20)) {
      g_pool[best].used = true;
      p = g_pool[best].p;
      bytes = g_pool[best].bytes;
      return 0;
    }
  }
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, n);
  if (e != cudaSuccess) {   // out of memory: drop the cached blocks and retry once
    pool_trim();
    e = cudaMalloc(&q, n);
    if (e != cudaSuccess) return record_cuda_error(e, __FILE__, __LINE__);
  }
  std::lock_guard<std::mutex> lk(g_pool_mu);
  g_pool.push_back({q, n, true, dev});
  p = q;
  bytes = n;
  return 0;
}

PoolBuf::~PoolBuf() {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_pool_mu);
  for (auto& b : g_pool)
    if (b.p == p) { b.used = false; break; }
}

int pool_trim() {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  const int dev = cur_dev();
  cudaDeviceSynchronize();
  std::vector<Block> keep;
  for (auto& b : g_pool) {
    if (b.used || b.dev != dev) keep.push_back(b);
    else cudaFree(b.p);
  }
  g_pool.swap(keep);
  return 0;
}

int stage_alloc(Staged& s, int64_t rows, int64_t cols, size_t es) {
  s.rows = rows;
  s.cols = cols;
  s.es = es;
  int64_t per16 = 16 / (int64_t)es > 0 ? 16 / (int64_t)es : 1;
  s.ld = ((rows > 0 ? rows : 1) + per16 - 1) / per16 * per16;
  return s.buf.alloc((size_t)s.ld * (size_t)(cols > 0 ? cols : 1) * es);
}

// ---- one 2-D region between host and device.  Page-locked host memory (and small regions): one cudaMemcpy2DAsync on
// `st`.  Pageable host memory -- what LLA passes (reference src/pinned-array.lisp:124-141) -- goes through the pinned
// staging ring of hostpipe.cu: uploads return once every DMA has been issued and make `st` wait for them; downloads
// start when the work queued on `st` so far has finished and are completed by stage_sync().
namespace {
constexpr size_t PIPE_MIN_BYTES = size_t(2) << 20;
struct Pending { Xfer x; cudaEvent_t ev; };
thread_local std::vector<Pending> g_pending;
bool g_pipe_off = getenv("LLACUDA_NO_PIPE") != nullptr;

int region_h2d(void* dev, size_t dpitch, const void* host, size_t hpitch, size_t width, size_t ncols, cudaStream_t st) {
  if (width == 0 || ncols == 0) return 0;
  if (g_pipe_off || width * ncols < PIPE_MIN_BYTES || host_is_pinned(host)) {
    LLA_CUDA(cudaMemcpy2DAsync(dev, dpitch, host, hpitch, width, ncols, cudaMemcpyHostToDevice, st));
    return 0;
  }
  Xfer x = pipe_upload(dev, dpitch, host, hpitch, width, ncols);
  LLA_TRY(x.enqueued());
  LLA_CUDA(cudaStreamWaitEvent(st, x.event(), 0));
  g_pending.push_back({x, nullptr});   // keeps the event alive until the call ends
  return 0;
}

int region_d2h(void* host, size_t hpitch, const void* dev, size_t dpitch, size_t width, size_t ncols, cudaStream_t st) {
  if (width == 0 || ncols == 0) return 0;
  if (g_pipe_off || width * ncols < PIPE_MIN_BYTES || host_is_pinned(host)) {
    LLA_CUDA(cudaMemcpy2DAsync(host, hpitch, dev, dpitch, width, ncols, cudaMemcpyDeviceToHost, st));
    return 0;
  }
  cudaEvent_t ev = pipe_event_get();
  if (!ev) return record_cuda_error(cudaErrorMemoryAllocation, __FILE__, __LINE__);
  LLA_CUDA(cudaEventRecord(ev, st));
  g_pending.push_back({pipe_download(host, hpitch, dev, dpitch, width, ncols, ev), ev});
  return 0;
}
}  // namespace

int stage_region_h2d(void* dev, size_t dpitch, const void* host, size_t hpitch, size_t width, size_t ncols, cudaStream_t st) {
  return region_h2d(dev, dpitch, host, hpitch, width, ncols, st);
}
int stage_region_d2h(void* host, size_t hpitch, const void* dev, size_t dpitch, size_t width, size_t ncols, cudaStream_t st) {
  return region_d2h(host, hpitch, dev, dpitch, width, ncols, st);
}

int stage_sync(cudaStream_t st) {
  int rc = 0;
  for (auto& p : g_pending) {
    int r = p.x.done();
    if (r && !rc) rc = r;
    if (p.ev) pipe_event_put(p.ev);
  }
  g_pending.clear();
  if (rc) return rc;
  LLA_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int stage_upload(Staged& s, const void* host, int64_t ldh, cudaStream_t st) {
  if (s.rows <= 0 || s.cols <= 0) return 0;
  return region_h2d(s.buf.p, (size_t)s.ld * s.es, host, (size_t)ldh * s.es, (size_t)s.rows * s.es, (size_t)s.cols, st);
}

int stage_download(const Staged& s, void* host, int64_t ldh, cudaStream_t st) {
  if (s.rows <= 0 || s.cols <= 0) return 0;
  return region_d2h(host, (size_t)ldh * s.es, s.buf.p, (size_t)s.ld * s.es, (size_t)s.rows * s.es, (size_t)s.cols, st);
}

int stage_upload_cols(Staged& s, const void* host, int64_t ldh, int64_t c0, int64_t c1, cudaStream_t st) {
  if (s.rows <= 0 || c1 <= c0) return 0;
  return region_h2d((char*)s.buf.p + (size_t)c0 * s.ld * s.es, (size_t)s.ld * s.es,
                    (const char*)host + (size_t)c0 * ldh * s.es, (size_t)ldh * s.es, (size_t)s.rows * s.es,
                    (size_t)(c1 - c0), st);
}

int stage_download_cols(const Staged& s, void* host, int64_t ldh, int64_t c0, int64_t c1, cudaStream_t st) {
  if (s.rows <= 0 || c1 <= c0) return 0;
  return region_d2h((char*)host + (size_t)c0 * ldh * s.es, (size_t)ldh * s.es,
                    (const char*)s.buf.p + (size_t)c0 * s.ld * s.es, (size_t)s.ld * s.es, (size_t)s.rows * s.es,
                    (size_t)(c1 - c0), st);
}

int stage_upload_block(Staged& s, const void* host, int64_t ldh, int64_t r0, int64_t r1, int64_t c0, int64_t c1, cudaStream_t st) {
  if (r1 <= r0 || c1 <= c0) return 0;
  return region_h2d((char*)s.buf.p + ((size_t)c0 * s.ld + r0) * s.es, (size_t)s.ld * s.es,
                    (const char*)host + ((size_t)c0 * ldh + r0) * s.es, (size_t)ldh * s.es,
                    (size_t)(r1 - r0) * s.es, (size_t)(c1 - c0), st);
}

int stage_download_block(const Staged& s, void* host, int64_t ldh, int64_t r0, int64_t r1, int64_t c0, int64_t c1, cudaStream_t st) {
  if (r1 <= r0 || c1 <= c0) return 0;
  return region_d2h((char*)host + ((size_t)c0 * ldh + r0) * s.es, (size_t)ldh * s.es,
                    (const char*)s.buf.p + ((size_t)c0 * s.ld + r0) * s.es, (size_t)s.ld * s.es,
                    (size_t)(r1 - r0) * s.es, (size_t)(c1 - c0), st);
}

int stage_upload_tri(Staged& s, const void* host, int64_t ldh, bool upper, cudaStream_t st, int64_t chunk) {
  for (int64_t c0 = 0; c0 < s.cols; c0 += chunk) {
    const int64_t c1 = c0 + chunk < s.cols ? c0 + chunk : s.cols;
    const int64_t r0 = upper ? 0 : c0, r1 = upper ? (c1 < s.rows ? c1 : s.rows) : s.rows;
    LLA_TRY(stage_upload_block(s, host, ldh, r0, r1, c0, c1, st));
  }
  return 0;
}

// triangle restricted to rows >= r0 (the rows above were already sent back by the rows-final hook)
int stage_download_tri(const Staged& s, void* host, int64_t ldh, bool upper, int64_t r0, cudaStream_t st, int64_t chunk) {
  for (int64_t c0 = 0; c0 < s.cols; c0 += chunk) {
    const int64_t c1 = c0 + chunk < s.cols ? c0 + chunk : s.cols;
    int64_t a = upper ? 0 : c0, b = upper ? (c1 < s.rows ? c1 : s.rows) : s.rows;
    if (a < r0) a = r0;
    LLA_TRY(stage_download_block(s, host, ldh, a, b, c0, c1, st));
  }
  return 0;
}

int stage_download_rows(const Staged& s, void* host, int64_t ldh, int64_t r0, int64_t r1, int64_t c0, cudaStream_t st) {
  if (r1 <= r0 || c0 >= s.cols) return 0;
  return stage_download_block(s, host, ldh, r0, r1, c0, s.cols, st);
}

// page-locked (cudaHostAlloc / cudaHostRegister) host memory: only then is a D2H copy asynchronous to the host
bool host_is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

PhaseTimer::PhaseTimer(cudaStream_t s) : st(s) {
  ok = true;
  for (auto& x : e)
    if (cudaEventCreate(&x) != cudaSuccess) ok = false;
}
PhaseTimer::~PhaseTimer() {
  for (auto& x : e) cudaEventDestroy(x);
}
void PhaseTimer::mark(int i) {
  if (ok) cudaEventRecord(e[i], st);
}
void PhaseTimer::finish() {
  Context* c = ctx();
  if (!ok || !c) return;
  cudaEventSynchronize(e[3]);
  for (int i = 0; i < 3; i++) cudaEventElapsedTime(&c->last_ms[i], e[i], e[i + 1]);
  cudaEventElapsedTime(&c->last_ms[3], e[0], e[3]);
}

}  // namespace llacuda
