// lla_b200/csrc/hostpipe.h -- the pinned staging ring for PAGEABLE host operands.
//
// LLA shares Lisp vectors with the foreign call by pinning them in place (`with-pinned-objects` + `vector-sap`,
// reference src/pinned-array.lisp:124-141): that is ordinary pageable heap memory, not page-locked.  A
// cudaMemcpyAsync from / to such memory is staged by the driver through one small bounce buffer on the calling
// thread (a single-threaded memcpy: 6-12 GB/s, a fraction of the 50+ GB/s of a PCIe Gen5 x16 link) and blocks the
// caller, so nothing overlaps.  This ring does the staging itself:
//   * page-locked slabs (allocated once, on the first pageable transfer);
//   * a pool of helper threads copies between the caller's memory and a slab in parallel;
//   * one uploader and one downloader thread own the slabs and issue ONE cudaMemcpyAsync per slab on their own
//     stream, so the copy of slab i+1 (host side) overlaps the DMA of slab i, and transfers in both directions
//     overlap each other and the kernels;
//   * the calling thread gets a ticket: `enqueued()` returns once every DMA of the transfer has been issued (the
//     consumer stream can then wait for `event()`), `done()` once the bytes are where they belong.
// The threads never call back into the host language and sleep between calls ("foreign helper
// threads are fine as long as they never call into Lisp").
#pragma once
#include "internal.h"

#include <memory>

namespace llacuda {

struct XferState;
struct Xfer {
  std::shared_ptr<XferState> s;
  // uploads: all DMAs issued and event() recorded (host blocks for the staging copies only, not for the DMA)
  int enqueued();
  // the whole transfer has completed (downloads: the caller's memory holds the data)
  int done();
  cudaEvent_t event();
  bool valid() const { return (bool)s; }
};

// Host 2-D region = ncols pieces of `width` bytes, hpitch bytes apart; device region likewise with dpitch.
// Pageable host memory only (page-locked callers use cudaMemcpy2DAsync directly: staging.cu decides).
Xfer pipe_upload(void* dev, size_t dpitch, const void* host, size_t hpitch, size_t width, size_t ncols);
// The download starts once `after` (an event already recorded by the caller, may be null) has completed.
Xfer pipe_download(void* host, size_t hpitch, const void* dev, size_t dpitch, size_t width, size_t ncols,
                   cudaEvent_t after);
// events handed to pipe_download must stay untouched until the transfer is done: take them from here
cudaEvent_t pipe_event_get();
void pipe_event_put(cudaEvent_t e);

}  // namespace llacuda
