// hostio.cu -- host-side staging for the buffers the reference really passes.
//
// commander3/src/sharp.f90:219-224 hands sharp_execute `c_loc` of ordinary (pageable) Fortran arrays.
// DMA needs page-locked memory; registering the caller's arrays (cudaHostRegister) would be the obvious
// route, but Commander allocates and frees its comm_map arrays all the time (cr_matmulA builds temporary
// comm_map objects in every CG iteration, commander3/src/comm_cr_mod.f90:877-918) and a registration that
// outlives a free()/munmap silently points the DMA engine at stale physical pages.  So the library owns a
// pinned staging arena instead and moves the data between the caller's pages and the arena with a small pool
// of copy threads, chunk by chunk inside the same pipeline that overlaps PCIe with the kernels:
//
//   upload   : worker threads copy chunk c+1 caller -> arena while chunk c is in flight over PCIe / in the kernels
//   download : the D2H of a chunk lands in the arena; once its event completes the workers copy it to the caller
//              while the GPU already works on the next chunk
//
// Pinned caller buffers (cudaMallocHost / cudaHostRegister done by the application) bypass the arena.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "kernels.h"

namespace cmdr {

// ---------------------------------------------------------------- pinned arena (grow-only, per tag and device)
struct PinnedBuf { void *ptr = nullptr; size_t bytes = 0; };
static std::map<std::string, PinnedBuf> g_pinned;
static std::mutex g_pin_mu;

void *pinned_get(const char *name, size_t bytes, bool write_combined) {
  int dev = 0;
  CMDR_CUDA_CHECK(cudaGetDevice(&dev));
  std::string key = std::string(name) + (write_combined ? "/wc@" : "@") + std::to_string(dev);
  std::lock_guard<std::mutex> lk(g_pin_mu);
  PinnedBuf &b = g_pinned[key];
  if (b.bytes < bytes) {
    if (b.ptr) { CMDR_CUDA_CHECK(cudaDeviceSynchronize()); CMDR_CUDA_CHECK(cudaFreeHost(b.ptr)); }
    size_t want = bytes + bytes / 16 + 4096;
    CMDR_CUDA_CHECK(cudaHostAlloc(&b.ptr, want, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
    b.bytes = want;
  }
  return b.ptr;
}

void pinned_release() {
  std::lock_guard<std::mutex> lk(g_pin_mu);
  for (auto &kv : g_pinned) if (kv.second.ptr) cudaFreeHost(kv.second.ptr);
  g_pinned.clear();
}

// ---------------------------------------------------------------- copy-thread pool
// One job at a time: memcpy(dst, src, n) cut into equal slices, one per worker plus the calling thread.
class CopyPool {
 public:
  explicit CopyPool(int nworkers) : nw_(nworkers) {
    for (int i = 0; i < nw_; ++i) std::thread([this, i] { run(i); }).detach();   // live until the process exits
  }
  void copy(void *dst, const void *src, size_t n) {
    if (n < (size_t)(256 << 10) || nw_ == 0) { memcpy(dst, src, n); return; }
    std::lock_guard<std::mutex> job(job_mu_);                      // one job at a time (callers: the API thread)
    const int parts = nw_ + 1;
    const size_t slice = ((n / parts) + 4095) & ~(size_t)4095;
    dst_ = static_cast<char *>(dst); src_ = static_cast<const char *>(src); n_ = n; slice_ = slice;
    pending_.store(nw_, std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> lk(mu_);
      gen_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
    do_slice(nw_);                                  // the caller takes the last slice
    while (pending_.load(std::memory_order_acquire) != 0) cpu_relax();
  }
  int workers() const { return nw_; }

 private:
  static void cpu_relax() {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  void do_slice(int i) {
    const size_t b = slice_ * (size_t)i;
    if (b >= n_) return;
    const size_t e = b + slice_ < n_ ? b + slice_ : n_;
    memcpy(dst_ + b, src_ + b, e - b);
  }
  // Workers spin for a short while after a job before they go to sleep: the staged uploads / downloads of one
  // transform arrive as a burst of jobs a few hundred microseconds apart, and a condition-variable wake-up per job
  // would cost as much as a small job itself.
  void run(int i) {
    unsigned long long seen = 0;
    for (;;) {
      bool got = false;
      for (int spin = 0; spin < 20000 && !got; ++spin) {           // ~100-200 us
        if (gen_.load(std::memory_order_acquire) != seen) got = true; else cpu_relax();
      }
      if (!got) {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen; });
      }
      seen = gen_.load(std::memory_order_acquire);
      do_slice(i);
      pending_.fetch_sub(1, std::memory_order_release);
    }
  }
  int nw_;
  std::mutex mu_, job_mu_;
  std::condition_variable cv_;
  char *dst_ = nullptr;
  const char *src_ = nullptr;
  size_t n_ = 0, slice_ = 0;
  std::atomic<int> pending_{0};
  std::atomic<unsigned long long> gen_{0};
};

// Ranks that share this host's cores (set by cmdr_sht_comm_register: one process per GPU, all on one node): the copy
// threads of all of them must fit into the CPUs the node has, or they only take each other's time slices -- 8 ranks x 16
// spinning workers on a 32-core host made the pageable path slower than one GPU.
static std::atomic<int> g_ranks_per_node{1};
void host_copy_set_ranks(int n) { if (n >= 1) g_ranks_per_node.store(n); }

static int copy_threads_wanted() {
  if (const char *e = getenv("CMDR_SHT_COPY_THREADS")) { const int n = atoi(e); if (n > 0) return n; }
  // the threads this process may run on (respects taskset / cgroup / mpirun binding), shared with the other ranks of the
  // node; one is left for the thread that drives the streams; at most 16
  cpu_set_t set;
  int avail = (sched_getaffinity(0, sizeof(set), &set) == 0) ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
  int n = avail / g_ranks_per_node.load() - 1;
  if (n > 16) n = 16;
  if (n < 1) n = 1;
  return n;
}

static CopyPool *copy_pool() {
  static std::mutex mu;
  static CopyPool *pool = nullptr;
  static int pool_threads = 0;
  std::lock_guard<std::mutex> lk(mu);
  const int want = copy_threads_wanted();
  if (!pool || want != pool_threads) {
    // pools are never destroyed (detached workers, no exit-order hazards); a superseded pool's workers just stay asleep
    pool = new CopyPool(want - 1);
    pool_threads = want;
  }
  return pool;
}

void host_copy(void *dst, const void *src, size_t bytes) { copy_pool()->copy(dst, src, bytes); }
int host_copy_threads() { return copy_pool()->workers() + 1; }

// ---------------------------------------------------------------- cache-resident upload ring
// Uploads of pageable arrays do not need an arena as large as the array: the copy threads fill a small ring of pinned
// slots with ordinary (cached) stores, the DMA engine reads each slot while it is still in the CPU's caches, and the
// slot is reused as soon as its copy has left.  Per payload byte DRAM then sees one read (the caller's page) instead of
// read + write + read -- the staging copy had made the host memory system, not PCIe, the limit of the end-to-end path.
// $CMDR_SHT_UP_PIECE_MB (default 4) x $CMDR_SHT_UP_SLOTS (default 6); CMDR_SHT_UP_PIECE_MB=0 stages whole ranges through
// the big arena as before.  Downloads take the mirror image ($CMDR_SHT_DN_PIECE_MB x $CMDR_SHT_DN_SLOTS, same defaults): the
// D2H copies of a finished chunk are issued at drain time, piece by piece into a second ring, and the copy threads move
// piece p to the caller while pieces p + 1 ... are in flight.  Measured at nside 2048 / lmax 4000 on one B200 (16 host
// cores): pageable pair 100.8 (arena both ways) -> 96.3 (upload ring) -> 90.4 ms (both rings); pinned: 83.1.  The distributed pipeline keeps the arena: there a rank that blocks on a ring slot delays its next
// exchange barrier and with it every other rank (2 GPUs: 96.8 ms with the arena, 112.6 ms with the ring).
struct UpRing {
  char *base = nullptr;
  size_t piece = 0;
  int nslot = 0, next = 0;
  std::vector<cudaEvent_t> ev;
  std::vector<char> busy;
};
static UpRing *make_ring(int which) {         // 0: upload ring, 1: download ring
  static std::map<int, UpRing *> rings;
  int dev = 0;
  CMDR_CUDA_CHECK(cudaGetDevice(&dev));
  const int key = dev * 2 + which;
  auto it = rings.find(key);
  if (it != rings.end()) return it->second;
  UpRing *R = new UpRing;
  int mb = 4, ns = 6;
  if (const char *e = getenv(which == 0 ? "CMDR_SHT_UP_PIECE_MB" : "CMDR_SHT_DN_PIECE_MB")) mb = atoi(e);
  if (const char *e = getenv(which == 0 ? "CMDR_SHT_UP_SLOTS" : "CMDR_SHT_DN_SLOTS")) ns = atoi(e);
  if (mb > 0 && ns >= 2) {
    R->piece = (size_t)mb << 20; R->nslot = ns;
    CMDR_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void **>(&R->base), R->piece * ns, cudaHostAllocDefault));
    R->ev.resize(ns); R->busy.assign(ns, 0);
    for (int k = 0; k < ns; ++k) CMDR_CUDA_CHECK(cudaEventCreateWithFlags(&R->ev[k], cudaEventDisableTiming));
  }
  rings[key] = R;
  return R;
}
static UpRing *up_ring() { return make_ring(0); }
static UpRing *dn_ring() { return make_ring(1); }
static cudaStream_t dn_stream() {
  static std::map<int, cudaStream_t> ss;
  int dev = 0;
  CMDR_CUDA_CHECK(cudaGetDevice(&dev));
  auto it = ss.find(dev);
  if (it != ss.end()) return it->second;
  cudaStream_t s;
  CMDR_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  ss[dev] = s;
  return s;
}

// ---------------------------------------------------------------- pointer classes
HostKind host_kind(const void *p) {
  if (!p) return HostKind::Device;                  // null columns (n_local == 0) never get dereferenced
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return HostKind::Pageable; }
  if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) return HostKind::Device;
  if (at.type == cudaMemoryTypeHost) return HostKind::Pinned;
  return HostKind::Pageable;
}

// ---------------------------------------------------------------- HostIO
void HostIO::init(const char *tag, double *const *cols, int ncols, long long count, bool upload_ring) {
  ncols_ = ncols; count_ = count;
  ring_ = upload_ring;
  pageable_ = false;
  for (int c = 0; c < ncols; ++c) { user_[c] = cols[c]; pageable_ = pageable_ || host_kind(cols[c]) == HostKind::Pageable; }
  stage_ = stage_up_ = nullptr;                    // allocated on first use: a transform uploads one array and downloads the other
  tag_ = tag;
}

double *HostIO::stage_dn() {
  if (!stage_) stage_ = static_cast<double *>(pinned_get(tag_, sizeof(double) * (size_t)count_ * ncols_, false));
  return stage_;
}
double *HostIO::stage_up() {
  // write-combined pages for the upload arena were measured and bring nothing here (105.8 ms plain vs 108.5 ms WC per
  // pageable pair; pool copy rate 73.7 vs 74.8 GB/s): plain pinned memory unless CMDR_SHT_STAGE_WC=1
  static const bool wc = getenv("CMDR_SHT_STAGE_WC") && atoi(getenv("CMDR_SHT_STAGE_WC")) != 0;
  if (!stage_up_) stage_up_ = static_cast<double *>(pinned_get(tag_, sizeof(double) * (size_t)count_ * ncols_, wc));
  return stage_up_;
}

void HostIO::h2d(double *dev, int c, long long ofs, long long n, cudaStream_t s) {
  if (n <= 0) return;
  const double *src = user_[c] + ofs;
  if (pageable_) {
    UpRing *R = ring_ ? up_ring() : nullptr;
    if (R && R->nslot) {                             // piece by piece through the cache-resident ring
      const char *from = reinterpret_cast<const char *>(src);
      char *to = reinterpret_cast<char *>(dev);
      const size_t bytes = sizeof(double) * (size_t)n;
      for (size_t off = 0; off < bytes; off += R->piece) {
        const size_t len = bytes - off < R->piece ? bytes - off : R->piece;
        const int k = R->next;
        R->next = (k + 1) % R->nslot;
        if (R->busy[k]) CMDR_CUDA_CHECK(cudaEventSynchronize(R->ev[k]));
        char *slot = R->base + (size_t)k * R->piece;
        host_copy(slot, from + off, len);
        CMDR_CUDA_CHECK(cudaMemcpyAsync(to + off, slot, len, cudaMemcpyHostToDevice, s));
        CMDR_CUDA_CHECK(cudaEventRecord(R->ev[k], s));
        R->busy[k] = 1;
      }
      return;
    }
    double *sp = stage_up() + (size_t)c * count_ + ofs;
    host_copy(sp, src, sizeof(double) * (size_t)n);
    src = sp;
  }
  CMDR_CUDA_CHECK(cudaMemcpyAsync(dev, src, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
}

void HostIO::d2h(const double *dev, int c, long long ofs, long long n, cudaStream_t s) {
  if (n <= 0) return;
  double *dst = user_[c] + ofs;
  if (pageable_ && ring_ && dn_ring()->nslot) {      // fetched piece by piece at drain time through the download ring
    drains_.push_back(Drain{nullptr, dst, dev, sizeof(double) * (size_t)n, false, true});
    return;
  }
  if (pageable_) {
    double *sp = stage_dn() + (size_t)c * count_ + ofs;
    CMDR_CUDA_CHECK(cudaMemcpyAsync(sp, dev, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s));
    drains_.push_back(Drain{nullptr, dst, sp, sizeof(double) * (size_t)n});
    return;
  }
  CMDR_CUDA_CHECK(cudaMemcpyAsync(dst, dev, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s));
}

void HostIO::commit(cudaStream_t s) {
  if (!pageable_) return;
  cudaEvent_t e = nullptr;
  for (Drain &d : drains_)
    if (!d.ev && !d.done) {
      if (!e) {
        CMDR_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CMDR_CUDA_CHECK(cudaEventRecord(e, s));
        events_.push_back(e);
      }
      d.ev = e;
    }
}

void HostIO::drain() {
  // requests deferred to the download ring: D2H of piece p + 1 ... p + nslot - 1 runs while the copy threads move piece p
  // out of its slot (still cache resident) into the caller's pages
  {
    UpRing *R = dn_ring();
    cudaStream_t ds = dn_stream();
    struct Fly { int k; double *dst; size_t len; };
    std::vector<Fly> fly;
    size_t head = 0;
    auto retire = [&]() {
      const Fly &f = fly[head++];
      CMDR_CUDA_CHECK(cudaEventSynchronize(R->ev[f.k]));
      host_copy(f.dst, R->base + (size_t)f.k * R->piece, f.len);
    };
    cudaEvent_t last = nullptr;
    for (Drain &d : drains_) {
      if (!d.deferred || d.done) continue;
      if (d.ev && d.ev != last) { CMDR_CUDA_CHECK(cudaStreamWaitEvent(ds, d.ev, 0)); last = d.ev; }
      else if (!d.ev) CMDR_CUDA_CHECK(cudaDeviceSynchronize());
      const char *from = reinterpret_cast<const char *>(d.src);
      char *to = reinterpret_cast<char *>(d.dst);
      for (size_t off = 0; off < d.bytes; off += R->piece) {
        const size_t len = d.bytes - off < R->piece ? d.bytes - off : R->piece;
        if ((int)(fly.size() - head) == R->nslot) retire();
        const int k = R->next;
        R->next = (k + 1) % R->nslot;
        CMDR_CUDA_CHECK(cudaMemcpyAsync(R->base + (size_t)k * R->piece, from + off, len, cudaMemcpyDeviceToHost, ds));
        CMDR_CUDA_CHECK(cudaEventRecord(R->ev[k], ds));
        fly.push_back(Fly{k, reinterpret_cast<double *>(to + off), len});
      }
      d.done = true;
    }
    while (head < fly.size()) retire();
  }
  for (Drain &d : drains_) {
    if (d.done) continue;
    if (d.ev) CMDR_CUDA_CHECK(cudaEventSynchronize(d.ev));
    else CMDR_CUDA_CHECK(cudaDeviceSynchronize());     // never committed: wait for everything
    host_copy(d.dst, d.src, d.bytes);
    d.done = true;
  }
  drains_.clear();
  for (cudaEvent_t e : events_) cudaEventDestroy(e);
  events_.clear();
}

}  // namespace cmdr

// Throughput of the copy-thread pool, GB/s of payload (best of `reps`): pageable -> pinned arena (direction 0, the
// upload staging: write-combined when wc != 0) or pinned arena -> pageable (direction 1).  Diagnostic for bench.py.
extern "C" double cmdr_sht_measure_host_copy(size_t bytes, int direction, int wc, int reps) {
  using namespace cmdr;
  void *pin = nullptr;
  CMDR_CUDA_CHECK(cudaHostAlloc(&pin, bytes, (wc && direction == 0) ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
  char *pg = static_cast<char *>(malloc(bytes));
  memset(pg, 1, bytes);
  memset(pin, 2, bytes);
  double best = 0.0;
  for (int r = 0; r < reps; ++r) {
    auto t0 = std::chrono::steady_clock::now();
    if (direction == 0) host_copy(pin, pg, bytes); else host_copy(pg, pin, bytes);
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (bytes / s / 1e9 > best) best = bytes / s / 1e9;
  }
  free(pg);
  cudaFreeHost(pin);
  return best;
}
extern "C" int cmdr_sht_host_copy_threads(void) { return cmdr::host_copy_threads(); }
