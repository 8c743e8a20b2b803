// lla_b200/csrc/staging.h -- host <-> device staging for the Fortran drop-in entry points.
//
// LLA hands the library HOST pointers that are only valid for the duration of the call: either
// SBCL heap vectors pinned in place (pageable memory, reference src/pinned-array.lisp:135-141)
// or foreign scratch (src/pinned-array.lisp:105-109).  A matrix operand (rows x cols, leading
// dimension ldh) is copied into a device buffer whose leading dimension is padded to an even
// number of elements so that every column starts 16-byte aligned for the cp.async loaders.
#pragma once
#include "internal.h"

namespace llacuda {

// Device staging buffers come from a grow-only pool (cudaMalloc / cudaFree of multi-GiB blocks cost
// tens of milliseconds each and cudaFree synchronises the device); llacuda_trim() releases it.
struct PoolBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int alloc(size_t n);
  ~PoolBuf();
};
int pool_trim();

struct Staged {
  PoolBuf buf;
  int64_t ld = 0;
  int64_t rows = 0, cols = 0;
  size_t es = 0;
  template <class T> T* ptr() { return (T*)buf.p; }
};

int stage_alloc(Staged& s, int64_t rows, int64_t cols, size_t es);
int stage_upload(Staged& s, const void* host, int64_t ldh, cudaStream_t st);
int stage_download(const Staged& s, void* host, int64_t ldh, cudaStream_t st);
// rows [r0, r1) of the columns [c0, c1)
int stage_upload_block(Staged& s, const void* host, int64_t ldh, int64_t r0, int64_t r1, int64_t c0, int64_t c1, cudaStream_t st);
int stage_download_block(const Staged& s, void* host, int64_t ldh, int64_t r0, int64_t r1, int64_t c0, int64_t c1, cudaStream_t st);
// only the named triangle (incl. diagonal) of a square matrix, as trapezoids of `chunk` columns
int stage_upload_tri(Staged& s, const void* host, int64_t ldh, bool upper, cudaStream_t st, int64_t chunk = 512);
int stage_download_tri(const Staged& s, void* host, int64_t ldh, bool upper, int64_t r0, cudaStream_t st, int64_t chunk = 512);
// rows [r0, r1) of the columns [c0, cols)
int stage_download_rows(const Staged& s, void* host, int64_t ldh, int64_t r0, int64_t r1, int64_t c0, cudaStream_t st);
bool host_is_pinned(const void* p);
// one 2-D region (ncols pieces of `width` bytes, pitches in bytes) between host and device on `st`; pageable host
// memory goes through the pinned ring (hostpipe.cu), completed by stage_sync()
int stage_region_h2d(void* dev, size_t dpitch, const void* host, size_t hpitch, size_t width, size_t ncols, cudaStream_t st);
int stage_region_d2h(void* host, size_t hpitch, const void* dev, size_t dpitch, size_t width, size_t ncols, cudaStream_t st);
// end of a host-pointer call: completes the pageable transfers this thread started through the pinned ring
// (hostpipe.cu), then synchronises `st`
int stage_sync(cudaStream_t st);
// column range [c0, c1) only
int stage_upload_cols(Staged& s, const void* host, int64_t ldh, int64_t c0, int64_t c1, cudaStream_t st);
int stage_download_cols(const Staged& s, void* host, int64_t ldh, int64_t c0, int64_t c1, cudaStream_t st);

// wall-clock phases of one host-pointer call, written to ctx()->last_ms
struct PhaseTimer {
  cudaEvent_t e[4];
  cudaStream_t st;
  bool ok = false;
  explicit PhaseTimer(cudaStream_t s);
  ~PhaseTimer();
  void mark(int i);   // 0 start, 1 after h2d, 2 after compute, 3 after d2h
  void finish();      // synchronises and stores the four durations
};

}  // namespace llacuda
