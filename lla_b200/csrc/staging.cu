// lla_b200/csrc/staging.cu -- see staging.h.
#include "staging.h"
#include "../../include/llacuda.h"

namespace llacuda {

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
