// lla_b200/csrc/probe.cu -- FP64 roof probes.
//
// MEASURED_PEAKS.json carries HBM and bf16 figures only, so the FP64 tensor (DMMA) and FP64 FMA
// roofs that bench.py's roofline is quoted against are measured in the same run by these
// register-resident loops.
#include "internal.h"
#include "../../include/llacuda.h"

namespace llacuda {

template <int CHAINS>
__global__ void __launch_bounds__(256) lla_dmma_probe_kernel(double* out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  double c[CHAINS][2];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s += c[i][0] + c[i][1];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void __launch_bounds__(256) lla_dfma_probe_kernel(double* out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
  double c[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s += c[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace llacuda
using namespace llacuda;

extern "C" int llacuda_probe_fp64_peak(int kind, int iters, double* tflops, float* ms_out) {
  ApiLock api_lock;
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  constexpr int CH = 16;
  DevBuf out;
  LLA_TRY(out.alloc(sizeof(double) * 256 * 4 * c->sm_count));
  cudaEvent_t e0, e1;
  LLA_CUDA(cudaEventCreate(&e0));
  LLA_CUDA(cudaEventCreate(&e1));
  int grid = c->sm_count * 2;
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    LLA_CUDA(cudaEventRecord(e0, c->stream));
    if (kind == 0) lla_dmma_probe_kernel<CH><<<grid, 256, 0, c->stream>>>((double*)out.p, iters);
    else lla_dfma_probe_kernel<CH><<<grid, 256, 0, c->stream>>>((double*)out.p, iters);
    count_launch();
    LLA_CUDA(cudaGetLastError());
    LLA_CUDA(cudaEventRecord(e1, c->stream));
    LLA_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    LLA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  double warps = double(grid) * 8.0;
  double flops = (kind == 0) ? warps * double(iters) * CH * (8.0 * 8.0 * 4.0 * 2.0)
                             : warps * 32.0 * double(iters) * CH * 2.0;
  if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
  if (ms_out) *ms_out = best;
  return 0;
}
