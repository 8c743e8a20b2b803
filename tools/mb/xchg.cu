// Exchange-latency probes for the LU panel design (B200).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xchg xchg.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void st16(void* p, unsigned a, unsigned b, unsigned tag) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "r"(a), "r"(tag), "r"(b), "r"(tag) : "memory");
}
__device__ __forceinline__ bool ld16(const void* p, unsigned tag, unsigned& a, unsigned& b) {
  unsigned t0, t1;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(a), "=r"(t0), "=r"(b), "=r"(t1) : "l"(p) : "memory");
  return t0 == tag && t1 == tag;
}
__device__ __forceinline__ void st16r(void* p, unsigned a, unsigned b, unsigned tag) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "r"(a), "r"(tag), "r"(b), "r"(tag) : "memory");
}
__device__ __forceinline__ bool ld16r(const void* p, unsigned tag, unsigned& a, unsigned& b) {
  unsigned t0, t1;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(a), "=r"(t0), "=r"(b), "=r"(t1) : "l"(p) : "memory");
  return t0 == tag && t1 == tag;
}
// all-to-all barrier carrying 8 bytes, G CTAs of one warp each, ITER rounds.
// MODE 0: shared headers (256 B apart), everybody polls all G.  MODE 1: inbox push.  MODE 2: as 1 with relaxed.gpu ops
template <int MODE>
__global__ void xchg(uint4* buf, int iters, long long* out) {
  const int G = gridDim.x, b = blockIdx.x, lane = threadIdx.x;
  long long t0 = clock64();
  unsigned acc = 0;
  for (int it = 0; it < iters; it++) {
    const unsigned tag = it + 1;
    const int par = it & 1;
    if (MODE == 0) {
      uint4* hd = buf + (size_t)par * G * 16;
      if (lane == 0) st16(hd + b * 16, b, it, tag);
      for (int s = lane; s < G; s += 32) { unsigned x, y; while (!ld16(hd + s * 16, tag, x, y)) {} acc += x; }
    } else {
      uint4* in = buf + (size_t)par * G * G;
      for (int r = lane; r < G; r += 32) { if (MODE == 1) st16(in + (size_t)r * G + b, b, it, tag); else st16r(in + (size_t)r * G + b, b, it, tag); }
      for (int s = lane; s < G; s += 32) { unsigned x, y; if (MODE == 1) { while (!ld16(in + (size_t)b * G + s, tag, x, y)) {} } else { while (!ld16r(in + (size_t)b * G + s, tag, x, y)) {} } acc += x; }
    }
    __syncwarp();
  }
  long long t1 = clock64();
  if (lane == 0 && b == 0) out[0] = (t1 - t0) / iters;
  if (acc == 0xdeadbeef) out[1] = 1;
}
__global__ void lat(double* buf, long long* out, double seed) {
  int tid = threadIdx.x;
  long long t0, t1;
  unsigned v = tid * 2654435761u;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 16; i++) v = __reduce_max_sync(~0u, v + i) ^ tid;
  t1 = clock64();
  if (tid == 0) out[0] = (t1 - t0) / 16;
  double x = seed + tid;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 32; i++) x = __dsub_rn(x, __dmul_rn(x, 1e-9));
  t1 = clock64();
  if (tid == 0) out[1] = (t1 - t0) / 32;
  t0 = clock64();
  double r = x;
#pragma unroll
  for (int i = 0; i < 8; i++) r = 1.0 / r + 0.5;
  t1 = clock64();
  if (tid == 0) out[2] = (t1 - t0) / 8;
  __shared__ int sm[64];
  sm[tid & 63] = (tid + 1) & 63;
  __syncthreads();
  int idx = tid & 63;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 32; i++) idx = ((volatile int*)sm)[idx];
  t1 = clock64();
  if (tid == 0) out[3] = (t1 - t0) / 32;
  unsigned s = v;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 16; i++) s = __shfl_sync(~0u, s, (s + i) & 31);
  t1 = clock64();
  if (tid == 0) out[4] = (t1 - t0) / 16;
  // same-thread store -> poll visible, volatile and relaxed
  unsigned a, b2;
  uint4* u = (uint4*)buf;
  __syncthreads();
  t0 = clock64();
  if (tid == 0) { st16(u + 100, 1, 2, 77u); while (!ld16(u + 100, 77u, a, b2)) {} }
  t1 = clock64();
  if (tid == 0) out[5] = t1 - t0;
  t0 = clock64();
  if (tid == 0) { st16r(u + 200, 1, 2, 78u); while (!ld16r(u + 200, 78u, a, b2)) {} }
  t1 = clock64();
  if (tid == 0) out[6] = t1 - t0;
  t0 = clock64();
  if (tid == 0) { for (int i = 0; i < 8; i++) ld16(u + 300 + 16 * (a & 1), 0, a, b2); }
  t1 = clock64();
  if (tid == 0) out[7] = (t1 - t0) / 8;
  if (x == 1.2345 || r == 1.2345 || idx == 777 || s == 12345) out[9] = 1;
}
int main() {
  uint4* buf; long long* out; cudaMalloc(&buf, 64 << 20); cudaMalloc(&out, 80);
  long long h[10];
  for (int rep = 0; rep < 2; rep++) {
    cudaMemset(buf, 0, 64 << 20);
    lat<<<1, 128>>>((double*)buf, out, 3.0); cudaDeviceSynchronize(); cudaMemcpy(h, out, 80, cudaMemcpyDeviceToHost);
    printf("redux=%lld dmul+dsub=%lld recip(+add)=%lld lds=%lld shfl=%lld st->ld(vol)=%lld st->ld(relaxed)=%lld volld-rt=%lld\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
  }
  int Gs[] = {1, 2, 8, 16, 48, 128};
  for (int g : Gs) {
    long long r[3];
    cudaMemset(buf, 0, 64 << 20); xchg<0><<<g, 32>>>(buf, 2000, out); cudaDeviceSynchronize(); cudaMemcpy(h, out, 80, cudaMemcpyDeviceToHost); r[0] = h[0];
    cudaMemset(buf, 0, 64 << 20); xchg<1><<<g, 32>>>(buf, 2000, out); cudaDeviceSynchronize(); cudaMemcpy(h, out, 80, cudaMemcpyDeviceToHost); r[1] = h[0];
    cudaMemset(buf, 0, 64 << 20); xchg<2><<<g, 32>>>(buf, 2000, out); cudaDeviceSynchronize(); cudaMemcpy(h, out, 80, cudaMemcpyDeviceToHost); r[2] = h[0];
    printf("G=%3d all-to-all round, cycles: shared-headers=%lld inbox-push=%lld inbox-push-relaxed=%lld\n", g, r[0], r[1], r[2]);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
