// lla_b200/csrc/hostpipe.cu -- see hostpipe.h: pinned staging ring + helper threads for pageable host operands
// (what LLA passes: reference src/pinned-array.lisp:124-141).
#include "hostpipe.h"
#include "../../include/llacuda.h"

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace llacuda {

namespace {

// LLACUDA_PIPE_STATS=1: where the staging threads spend their time (printed at exit): host copies, waits for a slab's
// DMA, bytes -- per direction
struct PipeStats {
  std::atomic<long long> bytes{0}, copy_us{0}, wait_us{0}, busy_us{0}, xfers{0};
};
PipeStats g_stats[2];
bool stats_on() { static const bool on = getenv("LLACUDA_PIPE_STATS") != nullptr; return on; }
long long now_us() {
  timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
  return (long long)ts.tv_sec * 1000000ll + ts.tv_nsec / 1000;
}
struct StatsDump {
  ~StatsDump() {
    if (!stats_on()) return;
    for (int d = 0; d < 2; d++) {
      const PipeStats& t = g_stats[d];
      if (!t.xfers.load()) continue;
      fprintf(stderr, "llacuda pipe %s: %lld transfers, %.1f MB, busy %.1f ms (host copies %.1f ms = %.1f GB/s, slab / DMA waits %.1f ms)\n",
              d == 0 ? "up  " : "down", t.xfers.load(), t.bytes.load() / 1e6, t.busy_us.load() / 1e3, t.copy_us.load() / 1e3,
              t.copy_us.load() ? t.bytes.load() / 1e3 / t.copy_us.load() : 0.0, t.wait_us.load() / 1e3);
    }
  }
} g_stats_dump;

// Slab geometry (per direction and device).  Every staged byte crosses the host's memory three times (source -> slab by
// the copy threads, slab -> device by the DMA engine), so the ring is kept small enough to stay in the last-level cache
// between the copy and the DMA that follows it, and fine-grained enough that copy and DMA of one transfer overlap;
// LLACUDA_SLAB_MB / LLACUDA_NSLAB / LLACUDA_PIECE_KB override.  Measured (profiles/r02_pageable_probe.log, n = 16384,
// dgesv_ / dpotrf_ from pageable memory): 3 x 16 MiB slabs 256 / 132 ms, 8 x 4 MiB slabs with 256 KiB pieces 204 / 89 ms.
constexpr int NSLAB_MAX = 8;
size_t env_size(const char* name, size_t dflt, size_t lo, size_t hi) {
  const char* e = getenv(name);
  if (!e) return dflt;
  long v = atol(e);
  if (v < (long)lo) v = (long)lo;
  if (v > (long)hi) v = (long)hi;
  return (size_t)v;
}
const size_t SLAB_BYTES = env_size("LLACUDA_SLAB_MB", 4, 1, 64) << 20;
const int NSLAB = (int)env_size("LLACUDA_NSLAB", 8, 2, NSLAB_MAX);
const size_t PIECE_BYTES = env_size("LLACUDA_PIECE_KB", 256, 64, 4096) << 10;

// ------------------------------------------------------------------ parallel host memcpy
struct Batch {
  int remaining = 0;   // guarded by mu
  std::mutex mu;
  std::condition_variable cv;
};
// `cnt` runs of `bytes` bytes, dpitch / spitch apart (cnt = 1: one contiguous run)
struct Piece { char* dst; const char* src; size_t bytes; size_t cnt, dpitch, spitch; Batch* batch; };

class CopyPool {
 public:
  static CopyPool& get() { static CopyPool* p = new CopyPool(); return *p; }
  // copies every piece, using the helpers and the calling thread; returns when all are done
  // `first`: uploads -- they gate the computation, downloads only its end -- are served before queued download pieces
  void run(std::vector<Piece>& pieces, bool first) {
    if (pieces.empty()) return;
    Batch b;
    b.remaining = (int)pieces.size();
    {
      std::lock_guard<std::mutex> lk(mu_);
      for (auto& p : pieces) { p.batch = &b; (first ? q_first_ : q_).push_back(p); }
    }
    cv_.notify_all();
    for (;;) {  // help until the queues are empty, then wait for this batch's stragglers
      Piece p;
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (!take(p)) break;
      }
      work(p);
    }
    std::unique_lock<std::mutex> lk(b.mu);
    b.cv.wait(lk, [&] { return b.remaining == 0; });
  }

 private:
  CopyPool() {
    int n = 0;
    if (const char* e = getenv("LLACUDA_COPY_THREADS")) n = atoi(e);
    if (n <= 0) {
      unsigned hw = std::thread::hardware_concurrency();
      n = (int)(hw * 3 / 4);
      if (n < 2) n = 2;
      if (n > 12) n = 12;
    }
    for (int i = 0; i < n - 1; i++) std::thread([this] { loop(); }).detach();
  }
  static void work(const Piece& p) {
    for (size_t i = 0; i < p.cnt; i++) memcpy(p.dst + i * p.dpitch, p.src + i * p.spitch, p.bytes);
    Batch* b = p.batch;   // the batch lives on its submitter's stack: nothing of it is touched after the unlock
    std::lock_guard<std::mutex> lk(b->mu);
    if (--b->remaining == 0) b->cv.notify_all();
  }
  void loop() {
    for (;;) {
      Piece p;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return !q_.empty() || !q_first_.empty(); });
        take(p);
      }
      work(p);
    }
  }
  bool take(Piece& p) {   // under mu_
    std::deque<Piece>& q = !q_first_.empty() ? q_first_ : q_;
    if (q.empty()) return false;
    p = q.front();
    q.pop_front();
    return true;
  }
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<Piece> q_, q_first_;
};

}  // namespace

// ------------------------------------------------------------------ transfers
struct XferState {
  std::mutex mu;
  std::condition_variable cv;
  bool enq = false, fin = false;
  int rc = 0;
  char err[256] = "";
  cudaEvent_t ev = nullptr;    // recorded after the last DMA of the transfer
  int dev = 0;
  bool up = true;
  char* devp = nullptr;
  char* host = nullptr;
  size_t dpitch = 0, hpitch = 0, width = 0, ncols = 0;
  cudaEvent_t after = nullptr;
  ~XferState() { if (ev) { cudaSetDevice(dev); cudaEventDestroy(ev); } }
};

namespace {

// one chunk = `cols` consecutive pieces of `w` bytes starting `off` bytes into column c0 (w < width only when a
// single column is wider than a slab)
struct Chunk { size_t c0, cols, off, w; };

std::vector<Chunk> make_chunks(size_t width, size_t ncols) {
  std::vector<Chunk> v;
  if (width == 0 || ncols == 0) return v;
  if (width > SLAB_BYTES) {
    for (size_t c = 0; c < ncols; c++)
      for (size_t off = 0; off < width; off += SLAB_BYTES) v.push_back({c, 1, off, width - off < SLAB_BYTES ? width - off : SLAB_BYTES});
    return v;
  }
  const size_t per = SLAB_BYTES / width;
  for (size_t c = 0; c < ncols; c += per) v.push_back({c, ncols - c < per ? ncols - c : per, 0, width});
  return v;
}

// pieces of one chunk between the caller's memory and the (densely packed) slab, about PIECE_BYTES each
void chunk_pieces(std::vector<Piece>& out, const Chunk& ch, char* host, size_t hpitch, char* slab, bool to_slab) {
  out.clear();
  auto add = [&](char* h, char* s, size_t bytes, size_t cnt) {
    if (to_slab) out.push_back(Piece{s, h, bytes, cnt, ch.w, hpitch, nullptr});
    else out.push_back(Piece{h, s, bytes, cnt, hpitch, ch.w, nullptr});
  };
  if (ch.w == hpitch) {   // whole columns with no gap between them on the host side: one contiguous range
    const size_t total = ch.cols * ch.w;
    char* h = host + ch.c0 * hpitch;
    for (size_t o = 0; o < total; o += PIECE_BYTES) add(h + o, slab + o, total - o < PIECE_BYTES ? total - o : PIECE_BYTES, 1);
    return;
  }
  if (ch.w >= PIECE_BYTES) {   // long columns: cut each one
    for (size_t c = 0; c < ch.cols; c++)
      for (size_t o = 0; o < ch.w; o += PIECE_BYTES)
        add(host + (ch.c0 + c) * hpitch + ch.off + o, slab + c * ch.w + o, ch.w - o < PIECE_BYTES ? ch.w - o : PIECE_BYTES, 1);
    return;
  }
  const size_t per = PIECE_BYTES / ch.w;   // short columns: several per piece
  for (size_t c = 0; c < ch.cols; c += per)
    add(host + (ch.c0 + c) * hpitch + ch.off, slab + c * ch.w, ch.w, ch.cols - c < per ? ch.cols - c : per);
}

class Pipe {
 public:
  explicit Pipe(bool up) : up_(up) { std::thread([this] { loop(); }).detach(); }
  void submit(const std::shared_ptr<XferState>& s) {
    { std::lock_guard<std::mutex> lk(mu_); q_.push_back(s); }
    cv_.notify_one();
  }

 private:
  bool up_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::shared_ptr<XferState>> q_;
  char* slab_[NSLAB_MAX] = {};
  cudaEvent_t slab_ev_[MAX_DEV][NSLAB_MAX] = {};
  cudaEvent_t last_[NSLAB_MAX] = {};      // event of the DMA that used the slab last
  cudaStream_t stream_[MAX_DEV] = {};

  static void fail(XferState& x, cudaError_t e, int line) {
    x.rc = LLACUDA_ERR_CUDA_BASE - (int)e;
    snprintf(x.err, sizeof(x.err), "%s:%d: %s", __FILE__, line, cudaGetErrorString(e));
  }
#define PIPE_CUDA(call)                                               \
  do {                                                                \
    cudaError_t _e = (call);                                          \
    if (_e != cudaSuccess) { fail(x, _e, __LINE__); return; }         \
  } while (0)

  void prepare(XferState& x) {
    PIPE_CUDA(cudaSetDevice(x.dev));
    for (int i = 0; i < NSLAB; i++)
      if (!slab_[i]) PIPE_CUDA(cudaHostAlloc((void**)&slab_[i], SLAB_BYTES, cudaHostAllocPortable));
    if (!stream_[x.dev]) {
      PIPE_CUDA(cudaStreamCreateWithFlags(&stream_[x.dev], cudaStreamNonBlocking));
      for (int i = 0; i < NSLAB; i++) PIPE_CUDA(cudaEventCreateWithFlags(&slab_ev_[x.dev][i], cudaEventDisableTiming));
    }
  }

  void upload(XferState& x) {
    prepare(x);
    if (x.rc) return;
    cudaStream_t st = stream_[x.dev];
    std::vector<Piece> pieces;
    const std::vector<Chunk> chunks = make_chunks(x.width, x.ncols);
    for (size_t i = 0; i < chunks.size(); i++) {
      const Chunk& ch = chunks[i];
      const int slot = (int)(i % NSLAB);
      const long long t0 = stats_on() ? now_us() : 0;
      if (last_[slot]) PIPE_CUDA(cudaEventSynchronize(last_[slot]));   // the DMA that read this slab has finished
      const long long t1 = stats_on() ? now_us() : 0;
      chunk_pieces(pieces, ch, x.host, x.hpitch, slab_[slot], true);
      CopyPool::get().run(pieces, true);
      if (stats_on()) { g_stats[0].wait_us += t1 - t0; g_stats[0].copy_us += now_us() - t1; g_stats[0].bytes += (long long)(ch.w * ch.cols); }
      PIPE_CUDA(cudaMemcpy2DAsync(x.devp + ch.c0 * x.dpitch + ch.off, x.dpitch, slab_[slot], ch.w, ch.w, ch.cols,
                                  cudaMemcpyHostToDevice, st));
      PIPE_CUDA(cudaEventRecord(slab_ev_[x.dev][slot], st));
      last_[slot] = slab_ev_[x.dev][slot];
    }
    PIPE_CUDA(cudaEventRecord(x.ev, st));
  }

  void download(XferState& x) {
    prepare(x);
    if (x.rc) return;
    cudaStream_t st = stream_[x.dev];
    if (x.after) PIPE_CUDA(cudaStreamWaitEvent(st, x.after, 0));
    std::vector<Piece> pieces;
    const std::vector<Chunk> chunks = make_chunks(x.width, x.ncols);
    const size_t nc = chunks.size();
    auto drain = [&](size_t i) -> bool {
      const int slot = (int)(i % NSLAB);
      const long long t0 = stats_on() ? now_us() : 0;
      cudaError_t e = cudaEventSynchronize(slab_ev_[x.dev][slot]);
      if (e != cudaSuccess) { fail(x, e, __LINE__); return false; }
      const long long t1 = stats_on() ? now_us() : 0;
      chunk_pieces(pieces, chunks[i], x.host, x.hpitch, slab_[slot], false);
      CopyPool::get().run(pieces, false);
      if (stats_on()) { g_stats[1].wait_us += t1 - t0; g_stats[1].copy_us += now_us() - t1; g_stats[1].bytes += (long long)(chunks[i].w * chunks[i].cols); }
      return true;
    };
    for (size_t i = 0; i < nc; i++) {
      const Chunk& ch = chunks[i];
      const int slot = (int)(i % NSLAB);
      if (i >= (size_t)NSLAB && !drain(i - NSLAB)) return;              // the slab must be empty before it is refilled
      PIPE_CUDA(cudaMemcpy2DAsync(slab_[slot], ch.w, x.devp + ch.c0 * x.dpitch + ch.off, x.dpitch, ch.w, ch.cols,
                                  cudaMemcpyDeviceToHost, st));
      PIPE_CUDA(cudaEventRecord(slab_ev_[x.dev][slot], st));
    }
    for (size_t i = nc > (size_t)NSLAB ? nc - NSLAB : 0; i < nc; i++)
      if (!drain(i)) return;
  }
#undef PIPE_CUDA

  void loop() {
    for (;;) {
      std::shared_ptr<XferState> s;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return !q_.empty(); });
        s = q_.front();
        q_.pop_front();
      }
      const long long tb = stats_on() ? now_us() : 0;
      if (up_) upload(*s); else download(*s);
      if (stats_on()) { g_stats[up_ ? 0 : 1].busy_us += now_us() - tb; g_stats[up_ ? 0 : 1].xfers += 1; }
      {
        std::lock_guard<std::mutex> lk(s->mu);
        s->enq = true;
        if (!up_ || s->rc) s->fin = true;   // uploads complete on the device: done() waits for the event
      }
      s->cv.notify_all();
    }
  }
};

// one uploader and one downloader per device: the rank threads of the multi-GPU engine stage concurrently
std::mutex g_pipe_mu;
Pipe* g_up[MAX_DEV] = {};
Pipe* g_down[MAX_DEV] = {};
Pipe& up_pipe(int dev) {
  std::lock_guard<std::mutex> lk(g_pipe_mu);
  if (!g_up[dev]) g_up[dev] = new Pipe(true);
  return *g_up[dev];
}
Pipe& down_pipe(int dev) {
  std::lock_guard<std::mutex> lk(g_pipe_mu);
  if (!g_down[dev]) g_down[dev] = new Pipe(false);
  return *g_down[dev];
}

std::shared_ptr<XferState> make_state(bool up, void* dev, size_t dpitch, void* host, size_t hpitch, size_t width,
                                      size_t ncols, cudaEvent_t after) {
  auto s = std::make_shared<XferState>();
  s->up = up;
  s->dev = cur_dev();
  s->devp = (char*)dev; s->dpitch = dpitch;
  s->host = (char*)host; s->hpitch = hpitch;
  s->width = width; s->ncols = ncols;
  s->after = after;
  if (cudaEventCreateWithFlags(&s->ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); s->ev = nullptr; }
  return s;
}

std::mutex g_ev_mu;
std::vector<cudaEvent_t> g_ev_pool[MAX_DEV];

}  // namespace

int Xfer::enqueued() {
  if (!s) return 0;
  std::unique_lock<std::mutex> lk(s->mu);
  s->cv.wait(lk, [&] { return s->enq; });
  if (s->rc) set_error(__FILE__, __LINE__, s->err, s->rc);
  return s->rc;
}

int Xfer::done() {
  if (!s) return 0;
  {
    std::unique_lock<std::mutex> lk(s->mu);
    s->cv.wait(lk, [&] { return s->enq; });
    if (s->rc) { set_error(__FILE__, __LINE__, s->err, s->rc); return s->rc; }
    if (!s->up) {
      s->cv.wait(lk, [&] { return s->fin; });
      return s->rc;
    }
  }
  LLA_CUDA(cudaEventSynchronize(s->ev));
  return 0;
}

cudaEvent_t Xfer::event() { return s ? s->ev : nullptr; }

Xfer pipe_upload(void* dev, size_t dpitch, const void* host, size_t hpitch, size_t width, size_t ncols) {
  Xfer x;
  x.s = make_state(true, dev, dpitch, const_cast<void*>(host), hpitch, width, ncols, nullptr);
  up_pipe(x.s->dev).submit(x.s);
  return x;
}

Xfer pipe_download(void* host, size_t hpitch, const void* dev, size_t dpitch, size_t width, size_t ncols,
                   cudaEvent_t after) {
  Xfer x;
  x.s = make_state(false, const_cast<void*>(dev), dpitch, host, hpitch, width, ncols, after);
  down_pipe(x.s->dev).submit(x.s);
  return x;
}

cudaEvent_t pipe_event_get() {
  const int d = cur_dev();
  if (d < 0 || d >= MAX_DEV) return nullptr;
  {
    std::lock_guard<std::mutex> lk(g_ev_mu);
    if (!g_ev_pool[d].empty()) {
      cudaEvent_t e = g_ev_pool[d].back();
      g_ev_pool[d].pop_back();
      return e;
    }
  }
  cudaEvent_t e = nullptr;
  if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return e;
}

void pipe_event_put(cudaEvent_t e) {
  const int d = cur_dev();
  if (!e || d < 0 || d >= MAX_DEV) return;
  std::lock_guard<std::mutex> lk(g_ev_mu);
  g_ev_pool[d].push_back(e);
}

}  // namespace llacuda
