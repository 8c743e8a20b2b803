// micro-latency probes for the panel-kernel design (B200): build with
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* buf, long long* out, double seed) {
  unsigned* u = (unsigned*)buf;
  long long t0, t1;
  int tid = threadIdx.x;
  // (a) 34 volatile v4 stores from one thread
  __syncthreads();
  t0 = clock64();
  if (tid == 0) {
    for (int c = 0; c < 34; c++)
      asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};\n" ::"l"(u + 4 * c), "r"(c), "r"(7u), "r"(c), "r"(7u) : "memory");
  }
  t1 = clock64();
  if (tid == 0) out[0] = t1 - t0;
  __syncthreads();
  // (b) volatile load round trip (dependent chain of 8)
  t0 = clock64();
  unsigned a = 0, b, c2, d;
  if (tid == 0) {
    for (int i = 0; i < 8; i++)
      asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(a), "=r"(b), "=r"(c2), "=r"(d) : "l"(u + 4 * (a & 3)) : "memory");
  }
  t1 = clock64();
  if (tid == 0) out[1] = (t1 - t0) / 8 + (a & 0);
  // (b2) store then poll until visible (same thread): store->load visibility latency
  __syncthreads();
  t0 = clock64();
  if (tid == 0) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};\n" ::"l"(u + 1024), "r"(99u), "r"(99u), "r"(99u), "r"(99u) : "memory");
    do { asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(a), "=r"(b), "=r"(c2), "=r"(d) : "l"(u + 1024) : "memory"); } while (a != 99u);
  }
  t1 = clock64();
  if (tid == 0) out[2] = t1 - t0;
  // (c) sqrt + div chain
  double x = seed + tid;
  __syncthreads();
  t0 = clock64();
  double r = sqrt(x);
  double q = x / r;
  t1 = clock64();
  if (q == 12345.678) out[9] = 1;
  if (tid == 0) out[3] = t1 - t0;
  t0 = clock64();
  double rr = 1.0 / q;
  t1 = clock64();
  if (rr == 12345.678) out[9] = 2;
  if (tid == 0) out[4] = t1 - t0;
  // (d) 16 back-to-back __syncthreads
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 16; i++) __syncthreads();
  t1 = clock64();
  if (tid == 0) out[5] = (t1 - t0) / 16;
  // (e) dependent DFMA chain of 64
  double acc = x;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; i++) acc = fma(acc, 1.0000001, 0.5);
  t1 = clock64();
  if (acc == 12345.678) out[9] = 3;
  if (tid == 0) out[6] = (t1 - t0) / 64;
  // (f) plain stores + threadfence
  __syncthreads();
  t0 = clock64();
  if (tid == 0) { for (int c = 0; c < 32; c++) buf[2048 + c] = c; __threadfence(); }
  t1 = clock64();
  if (tid == 0) out[7] = t1 - t0;
  // (g) warp shuffle reduce of (double,int) 5 rounds
  double v = x; int rw = tid;
  t0 = clock64();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { double ov = __shfl_xor_sync(~0u, v, o); int orow = __shfl_xor_sync(~0u, rw, o); if (ov > v || (ov == v && orow < rw)) { v = ov; rw = orow; } }
  t1 = clock64();
  if (v == 12345.678 && rw == 999999) out[9] = 4;
  if (tid == 0) out[8] = t1 - t0;
}
int main() {
  double* buf; long long* out; cudaMalloc(&buf, 1 << 20); cudaMalloc(&out, 80); cudaMemset(buf, 0, 1 << 20);
  for (int rep = 0; rep < 3; rep++) {
    k<<<1, 256>>>(buf, out, 3.0); cudaDeviceSynchronize();
    long long h[10]; cudaMemcpy(h, out, 80, cudaMemcpyDeviceToHost);
    printf("rep%d: 34 vol stores=%lld cyc | vol ld rt=%lld | st->ld visible=%lld | sqrt+div=%lld | recip=%lld | syncthreads=%lld | dfma lat=%lld | 32 st+fence=%lld | shfl reduce=%lld\n", rep, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8]);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
