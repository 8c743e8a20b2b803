// lla_b200/csrc/staging.cu -- see staging.h.
#include "staging.h"
#include "../../include/llacuda.h"

#include <mutex>
#include <vector>

namespace llacuda {

namespace {
struct Block { void* p; size_t bytes; bool used; };
std::mutex g_pool_mu;
std::vector<Block> g_pool;
}  // namespace

int PoolBuf::alloc(size_t n) {
  if (n == 0) n = 16;
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    int best = -1;
    for (int i = 0; i < (int)g_pool.size(); i++)
      if (!g_pool[i].used && g_pool[i].bytes >= n && (best < 0 || g_pool[i].bytes < g_pool[best].bytes)) best = i;
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
  g_pool.push_back({q, n, true});
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
  cudaDeviceSynchronize();
  std::vector<Block> keep;
  for (auto& b : g_pool) {
    if (b.used) keep.push_back(b);
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

int stage_upload(Staged& s, const void* host, int64_t ldh, cudaStream_t st) {
  if (s.rows <= 0 || s.cols <= 0) return 0;
  LLA_CUDA(cudaMemcpy2DAsync(s.buf.p, (size_t)s.ld * s.es, host, (size_t)ldh * s.es, (size_t)s.rows * s.es,
                             (size_t)s.cols, cudaMemcpyHostToDevice, st));
  return 0;
}

int stage_download(const Staged& s, void* host, int64_t ldh, cudaStream_t st) {
  if (s.rows <= 0 || s.cols <= 0) return 0;
  LLA_CUDA(cudaMemcpy2DAsync(host, (size_t)ldh * s.es, s.buf.p, (size_t)s.ld * s.es, (size_t)s.rows * s.es,
                             (size_t)s.cols, cudaMemcpyDeviceToHost, st));
  return 0;
}

int stage_upload_cols(Staged& s, const void* host, int64_t ldh, int64_t c0, int64_t c1, cudaStream_t st) {
  if (s.rows <= 0 || c1 <= c0) return 0;
  LLA_CUDA(cudaMemcpy2DAsync((char*)s.buf.p + (size_t)c0 * s.ld * s.es, (size_t)s.ld * s.es,
                             (const char*)host + (size_t)c0 * ldh * s.es, (size_t)ldh * s.es, (size_t)s.rows * s.es,
                             (size_t)(c1 - c0), cudaMemcpyHostToDevice, st));
  return 0;
}

int stage_download_cols(const Staged& s, void* host, int64_t ldh, int64_t c0, int64_t c1, cudaStream_t st) {
  if (s.rows <= 0 || c1 <= c0) return 0;
  LLA_CUDA(cudaMemcpy2DAsync((char*)host + (size_t)c0 * ldh * s.es, (size_t)ldh * s.es,
                             (const char*)s.buf.p + (size_t)c0 * s.ld * s.es, (size_t)s.ld * s.es, (size_t)s.rows * s.es,
                             (size_t)(c1 - c0), cudaMemcpyDeviceToHost, st));
  return 0;
}

int stage_upload_block(Staged& s, const void* host, int64_t ldh, int64_t r0, int64_t r1, int64_t c0, int64_t c1, cudaStream_t st) {
  if (r1 <= r0 || c1 <= c0) return 0;
  LLA_CUDA(cudaMemcpy2DAsync((char*)s.buf.p + ((size_t)c0 * s.ld + r0) * s.es, (size_t)s.ld * s.es,
                             (const char*)host + ((size_t)c0 * ldh + r0) * s.es, (size_t)ldh * s.es,
                             (size_t)(r1 - r0) * s.es, (size_t)(c1 - c0), cudaMemcpyHostToDevice, st));
  return 0;
}

int stage_download_block(const Staged& s, void* host, int64_t ldh, int64_t r0, int64_t r1, int64_t c0, int64_t c1, cudaStream_t st) {
  if (r1 <= r0 || c1 <= c0) return 0;
  LLA_CUDA(cudaMemcpy2DAsync((char*)host + ((size_t)c0 * ldh + r0) * s.es, (size_t)ldh * s.es,
                             (const char*)s.buf.p + ((size_t)c0 * s.ld + r0) * s.es, (size_t)s.ld * s.es,
                             (size_t)(r1 - r0) * s.es, (size_t)(c1 - c0), cudaMemcpyDeviceToHost, st));
  return 0;
}

int stage_upload_tri(Staged& s, const void* host, int64_t ldh, bool upper, cudaStream_t st, int64_t chunk) {
  for (int64_t c0 = 0; c0 < s.cols; c0 += chunk) {
    const int64_t c1 = c0 + chunk < s.cols ? c0 + chunk : s.cols;
    const int64_t r0 = upper ? 0 : c0, r1 = upper ? (c1 < s.rows ? c1 : s.rows) : s.rows;
    if (r1 <= r0) continue;
    LLA_CUDA(cudaMemcpy2DAsync((char*)s.buf.p + ((size_t)c0 * s.ld + r0) * s.es, (size_t)s.ld * s.es,
                               (const char*)host + ((size_t)c0 * ldh + r0) * s.es, (size_t)ldh * s.es,
                               (size_t)(r1 - r0) * s.es, (size_t)(c1 - c0), cudaMemcpyHostToDevice, st));
  }
  return 0;
}

// triangle restricted to rows >= r0 (the rows above were already sent back by the rows-final hook)
int stage_download_tri(const Staged& s, void* host, int64_t ldh, bool upper, int64_t r0, cudaStream_t st, int64_t chunk) {
  for (int64_t c0 = 0; c0 < s.cols; c0 += chunk) {
    const int64_t c1 = c0 + chunk < s.cols ? c0 + chunk : s.cols;
    int64_t a = upper ? 0 : c0, b = upper ? (c1 < s.rows ? c1 : s.rows) : s.rows;
    if (a < r0) a = r0;
    if (b <= a) continue;
    LLA_CUDA(cudaMemcpy2DAsync((char*)host + ((size_t)c0 * ldh + a) * s.es, (size_t)ldh * s.es,
                               (const char*)s.buf.p + ((size_t)c0 * s.ld + a) * s.es, (size_t)s.ld * s.es,
                               (size_t)(b - a) * s.es, (size_t)(c1 - c0), cudaMemcpyDeviceToHost, st));
  }
  return 0;
}

int stage_download_rows(const Staged& s, void* host, int64_t ldh, int64_t r0, int64_t r1, int64_t c0, cudaStream_t st) {
  if (r1 <= r0 || c0 >= s.cols) return 0;
  LLA_CUDA(cudaMemcpy2DAsync((char*)host + ((size_t)c0 * ldh + r0) * s.es, (size_t)ldh * s.es,
                             (const char*)s.buf.p + ((size_t)c0 * s.ld + r0) * s.es, (size_t)s.ld * s.es,
                             (size_t)(r1 - r0) * s.es, (size_t)(s.cols - c0), cudaMemcpyDeviceToHost, st));
  return 0;
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
