// calico_b200 — host driver behind the C ABI (include/calico_b200.h).
//
// Mirrors, for ONE path of the reference, calico::BatchOptimizer::Optimize (calico/batch_optimizer.cpp:53-81):
//   problem assembly      world_model.cpp:40-77, bspline.hpp:10-17, camera.cpp:92-153, gyroscope.cpp:10-54,
//                         accelerometer.cpp:10-56  -> one SoA pack + upload (Problem::upload)
//   ceres::Solve          batch_optimizer.cpp:73 with DefaultSolverOptions (batch_optimizer.cpp:10-17): trust-region
//                         Levenberg-Marquardt (Ceres external: trust_region_minimizer.cc, levenberg_marquardt_strategy.cc)
//                         -> Problem::minimize, every numerical step a CUDA kernel, the host reading one small scalar
//                         block per iteration
//   Sensor::UpdateResiduals  camera.cpp:70-80 -> Problem::refresh_residuals
// There is no CPU fallback: without a CUDA device every device entry point returns CB2_INTERNAL.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <memory>
#include <unordered_map>
#include <vector>
#include <map>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#if defined(CB2_EMUL)
#include <condition_variable>
#include <mutex>
#else
#include <dlfcn.h>
#endif

#include "../../include/calico_b200.h"
#include "cb2_eval.cuh"
#include "cb2_lmkernels.cuh"
#include "cb2_normal.cuh"
#include "cb2_schur.cuh"
#include "cb2_cr.cuh"
#include "cb2_fit.cuh"

namespace cb2 {

static int env_int(const char* name, int fallback) { const char* e = std::getenv(name); return e ? std::atoi(e) : fallback; }
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct CudaFail { std::string msg; };
#define CB2_CUDA(expr)                                                                                              \
  do {                                                                                                              \
    cudaError_t e_ = (expr);                                                                                        \
    if (e_ != cudaSuccess) throw CudaFail{std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #expr};      \
  } while (0)

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
  ~DevBuf() { release(); }
  // Device memory comes from the device's default stream-ordered pool with the release threshold lifted (dev_pool_init): freed
  // blocks stay cached in the process, so the second and later problems of a session (re-optimisation after outlier marking,
  // camera.cpp:258-299, creates a new problem per Optimize call) do not pay cudaMalloc / cudaFree again. Allocation and release are
  // ordered on the legacy default stream; the owner synchronises its own streams before releasing and the device after allocating.
  void release() {
#ifdef CB2_EMUL
    if (p) cudaFree(p);
#else
    if (p) {
      const cudaError_t e = cudaFreeAsync(p, nullptr);
      if (e != cudaSuccess) {   // a destructor cannot report through the status code: say it once on stderr and clear the sticky error
        static std::atomic<bool> warned{false};
        if (!warned.exchange(true)) std::fprintf(stderr, "[calico_b200] cudaFreeAsync failed: %s\n", cudaGetErrorString(e));
        cudaGetLastError();
      }
    }
#endif
    p = nullptr; n = 0;
  }
  // Pool blocks come back with the previous owner's contents: every allocation that is not immediately overwritten by an upload is
  // cleared (per-segment partials of (segment, sensor) pairs without observations, padding rows, ... are never written by a kernel).
  void alloc(size_t count, bool clear = true) {
    release();
    n = count;
#ifdef CB2_EMUL
    if (count) CB2_CUDA(cudaMalloc(reinterpret_cast<void**>(&p), std::max<size_t>(count, 1) * sizeof(T)));
    if (count && clear) CB2_CUDA(cudaMemsetAsync(p, 0, count * sizeof(T), nullptr));
#else
    if (count) CB2_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&p), std::max<size_t>(count, 1) * sizeof(T), nullptr));
    if (count && clear) CB2_CUDA(cudaMemsetAsync(p, 0, count * sizeof(T), nullptr));
#endif
  }
  void zero(cudaStream_t s) { if (n) CB2_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
  void upload(const std::vector<T>& h, int64_t* counter = nullptr) {
    alloc(h.size(), false);
    if (!h.empty()) { CB2_CUDA(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice)); if (counter) *counter += int64_t(h.size() * sizeof(T)); }
  }
  void download(std::vector<T>& h, int64_t* counter = nullptr) const {
    h.resize(n);
    if (n) { CB2_CUDA(cudaMemcpy(h.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost)); if (counter) *counter += int64_t(n * sizeof(T)); }
  }
};

// Process-wide pool of pinned (page-locked) host blocks for the staging of uploads and result downloads. Pinning is expensive (of the order
// of 0.3 ms per MB), transfers from / to pinned memory run at the full PCIe / C2C rate and asynchronously; a calibration session creates a
// problem per Optimize call (batch_optimizer.cpp:57), so blocks are kept for the life of the process and handed out by size.
struct PinnedPool {
  std::mutex mu;
  std::multimap<size_t, void*> free_blocks;
  void* get(size_t bytes, size_t* cap) {
    bytes = std::max<size_t>(bytes, 256);
    {
      std::lock_guard<std::mutex> lk(mu);
      auto it = free_blocks.lower_bound(bytes);
      if (it != free_blocks.end() && it->first <= 2 * bytes + (size_t(1) << 20)) { void* p = it->second; *cap = it->first; free_blocks.erase(it); return p; }
    }
    const size_t c = (bytes + (size_t(1) << 20) - 1) >> 20 << 20;
    void* p = nullptr;
    if (cudaMallocHost(&p, c) != cudaSuccess) {
      // No device (problem assembly and the shard plan work without one) or pinning refused: ordinary host memory — copies from it still
      // work, only slower. Blocks are never returned to the system, so the two kinds can share the pool.
      cudaGetLastError();
      p = std::aligned_alloc(256, c);
      if (!p) throw std::bad_alloc();
    }
    *cap = c;
    return p;
  }
  void put(void* p, size_t cap) { if (p) { std::lock_guard<std::mutex> lk(mu); free_blocks.emplace(cap, p); } }
};
static PinnedPool& pinned_pool() { static PinnedPool* pool = new PinnedPool(); return *pool; }   // never destroyed: outlives every handle
struct PinnedBuf {
  void* p = nullptr; size_t cap = 0;
  PinnedBuf() = default;
  PinnedBuf(const PinnedBuf&) = delete;
  PinnedBuf& operator=(const PinnedBuf&) = delete;
  PinnedBuf(PinnedBuf&& o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr; o.cap = 0; }
  PinnedBuf& operator=(PinnedBuf&& o) noexcept { if (this != &o) { release(); p = o.p; cap = o.cap; o.p = nullptr; o.cap = 0; } return *this; }
  ~PinnedBuf() { release(); }
  void acquire(size_t bytes) { if (cap < bytes) { release(); p = pinned_pool().get(bytes, &cap); } }
  void release() { pinned_pool().put(p, cap); p = nullptr; cap = 0; }
};

// Growable host array backed by the pinned pool: the per-observation input copies of a problem (tens of MB) reuse warm, already-mapped
// blocks from one Optimize call of a session to the next instead of paying fresh page faults for every std::vector.
template <class T>
struct PooledArray {
  PinnedBuf buf;
  size_t n = 0;
  PooledArray() = default;
  PooledArray(PooledArray&& o) noexcept : buf(std::move(o.buf)), n(o.n) { o.n = 0; }
  PooledArray& operator=(PooledArray&& o) noexcept { buf = std::move(o.buf); n = o.n; o.n = 0; return *this; }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  T* data() { return static_cast<T*>(buf.p); }
  const T* data() const { return static_cast<const T*>(buf.p); }
  T& operator[](size_t i) { return data()[i]; }
  const T& operator[](size_t i) const { return data()[i]; }
  void reserve(size_t count) {
    if (count * sizeof(T) <= buf.cap) return;
    PinnedBuf nb;
    nb.acquire(std::max(count * sizeof(T), 2 * buf.cap));
    if (n) std::memcpy(nb.p, buf.p, n * sizeof(T));
    std::swap(nb.p, buf.p); std::swap(nb.cap, buf.cap);
  }
  void append(const T* src, size_t count) { reserve(n + count); if (count) std::memcpy(data() + n, src, count * sizeof(T)); n += count; }
  void append_fill(size_t count, T v) { reserve(n + count); for (size_t i = 0; i < count; ++i) data()[n + i] = v; n += count; }
};

struct HostBody {
  int id = 0;
  double q[4] = {0, 0, 0, 1}, t[3] = {0, 0, 0};
  bool pose_const = true, model_const = true;
  std::vector<int> feature_ids;
  std::vector<double> pts;
  std::unordered_map<int, int> slot;
  std::vector<int> dense_slot;   // feature id -> slot for small non-negative ids (the common case), else empty
  int pw0 = 0;   // first index of this body's points in the world-point table
};

struct HostSensor {
  int kind = 0, model = 0;
  std::string name;
  std::vector<double> intr;
  double q[4] = {0, 0, 0, 1}, t[3] = {0, 0, 0};
  double latency = 0, sigma = 1;
  int loss_type = 0;
  double loss_scale = 1;
  bool en_intr = false, en_extr = false, en_lat = false;
  PooledArray<double> stamp, meas;
  PooledArray<int> body_slot, feat_slot;
  PooledArray<uint8_t> outlier;
  const double* residuals = nullptr;          // [n_obs][m] after cb2_optimize: points into the problem's pinned result block
  const uint8_t* residual_valid = nullptr;    // [n_obs]
  int m() const { return kind == kCamera ? 2 : 3; }
  int n_obs() const { return int(stamp.size()); }
  // device-side bookkeeping
  std::vector<int> perm;      // sorted position -> original observation index
  int n_active = 0;
  DevBuf<unsigned char> d_obs;   // ONE block per sensor: [stamp | measurement | segment | point | image | permutation], filled by one H2D copy
  const int* d_perm = nullptr;   // (inside d_obs)
  DevBuf<double> d_r, d_J;
  DevBuf<int> d_seg_start, d_frame_obs, d_seg_frame;
  DevBuf<unsigned char> d_valid;
};

struct ChunkPlan { int a, b; };   // interior control points [a, b)

// CUDA-event phase timing on the library's stream; pairs are resolved after the next stream synchronisation.
enum Phase { kPhJacobian = 0, kPhNormal = 1, kPhSchur = 2, kPhCost = 3, kPhLoop = 4, kPhCamera = 5, kPhCount = 6 };
struct PhaseTimer {
  struct Pair { cudaEvent_t a, b; int phase; };
  std::vector<cudaEvent_t> pool;
  std::vector<Pair> open_pairs;
  cudaEvent_t cur_a[kPhCount] = {};
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  void begin(int ph, cudaStream_t s) { cur_a[ph] = get(); cudaEventRecord(cur_a[ph], s); }
  void end(int ph, cudaStream_t s) { cudaEvent_t b = get(); cudaEventRecord(b, s); open_pairs.push_back(Pair{cur_a[ph], b, ph}); }
  void resolve(double* ms_by_phase) {   // call only after the stream has been synchronised
    for (auto& pr : open_pairs) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, pr.a, pr.b);
      ms_by_phase[pr.phase] += ms;
      pool.push_back(pr.a); pool.push_back(pr.b);
    }
    open_pairs.clear();
  }
  ~PhaseTimer() { for (auto e : pool) cudaEventDestroy(e); for (auto& pr : open_pairs) { cudaEventDestroy(pr.a); cudaEventDestroy(pr.b); } }
};

// Developer aid (CB2_PROFILE=1): CUDA events around every kernel launch, warm and in pipeline order (what ncu's cold-cache,
// serialised launch list cannot show); per-kernel totals are printed to stderr when the problem is destroyed.
struct KernelProfiler {
  struct Rec { const char* name; cudaEvent_t a, b; };
  bool on = std::getenv("CB2_PROFILE") != nullptr;
  std::vector<Rec> open;
  std::map<std::string, std::pair<double, long>> tot;
  void resolve() {
    for (auto& r : open) { float ms = 0.f; cudaEventElapsedTime(&ms, r.a, r.b); auto& t = tot[r.name]; t.first += ms; ++t.second; cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    open.clear();
  }
  void report() {
    if (!on || tot.empty()) return;
    std::vector<std::pair<std::string, std::pair<double, long>>> v(tot.begin(), tot.end());
    std::sort(v.begin(), v.end(), [](const auto& x, const auto& y) { return x.second.first > y.second.first; });
    std::fprintf(stderr, "[cb2 profile] kernel, launches, total ms, avg us\n");
    for (auto& e : v) std::fprintf(stderr, "[cb2 profile] %-44s %6ld %10.3f %9.2f\n", e.first.c_str(), e.second.second, e.second.first, 1e3 * e.second.first / e.second.second);
  }
};

// ----------------------------------------------------------------------------------------------------------------
// Cross-rank collectives (one process per GPU). The product uses NCCL, resolved at run time from the libnccl.so.2 the
// process already has (torch's bundled copy) or the system one, so that single-GPU use has no NCCL dependency.
// The SIMT-emulation test build replaces it by an in-process rendezvous between host threads.
// ----------------------------------------------------------------------------------------------------------------
struct Comm {
  int world = 1, rank = 0;
  std::atomic<bool> dead{false};   // aborted after a failed / timed-out collective: every later use fails fast
  double timeout_s = std::getenv("CB2_COLLECTIVE_TIMEOUT_S") ? std::atof(std::getenv("CB2_COLLECTIVE_TIMEOUT_S")) : 60.0;
  virtual void arm() {}            // a region that may block on a peer begins / ends (watchdog deadline, see NcclComm)
  virtual void disarm() {}
  virtual ~Comm() {}
  virtual void allreduce_sum(double* buf, size_t n, cudaStream_t s) = 0;
  virtual void allreduce_max(double* buf, size_t n, cudaStream_t s) = 0;
  virtual void allreduce_sum_u8(unsigned char* buf, size_t n, cudaStream_t s) = 0;
  virtual bool async_error(std::string*) { return false; }   // an asynchronous failure of the communicator (peer died, ...)
  virtual void abort() { dead = true; }                       // tears the communicator down so that stuck collectives return
};

#if !defined(CB2_EMUL)
struct NcclApi {
  typedef struct { char internal[128]; } UniqueId;
  typedef void* CommT;
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(CommT*, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(CommT) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, CommT, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*CommGetAsyncError)(CommT, int*) = nullptr;
  int (*CommAbort)(CommT) = nullptr;
  void* handle = nullptr;
  std::string error;
  bool load() {
    if (handle) return true;
    const char* names[] = {std::getenv("CB2_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n) continue;
      handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) { error = "cannot load libnccl.so.2 (set CB2_NCCL_LIB)"; return false; }
    GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(dlsym(handle, "ncclGetUniqueId"));
    CommInitRank = reinterpret_cast<decltype(CommInitRank)>(dlsym(handle, "ncclCommInitRank"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(handle, "ncclCommDestroy"));
    AllReduce = reinterpret_cast<decltype(AllReduce)>(dlsym(handle, "ncclAllReduce"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(handle, "ncclGetErrorString"));
    CommGetAsyncError = reinterpret_cast<decltype(CommGetAsyncError)>(dlsym(handle, "ncclCommGetAsyncError"));
    CommAbort = reinterpret_cast<decltype(CommAbort)>(dlsym(handle, "ncclCommAbort"));
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce) { error = "libnccl is missing expected symbols"; return false; }
    return true;
  }
};
static NcclApi& nccl() { static NcclApi api; return api; }
// A collective a peer never joins blocks either inside the NCCL call on the host (the first collective of a communicator sets up its
// connections) or inside the kernel on the stream. Both are covered by a WATCHDOG thread per communicator: regions that may block arm a
// deadline (CB2_COLLECTIVE_TIMEOUT_S, default 60 s); when it passes — or the communicator reports an asynchronous error — the watchdog calls
// ncclCommAbort, which makes the blocked call / kernel return, and the caller reports status 13 instead of hanging.
struct NcclComm : Comm {
  NcclApi::CommT comm = nullptr;
  std::thread watchdog;
  std::mutex wd_mu;
  std::condition_variable wd_cv;
  std::atomic<double> deadline{0.0};    // 0 = disarmed
  std::atomic<bool> aborting{false};
  bool stop = false;
  std::string why;                      // set by the watchdog before `dead`
  void start_watchdog() {
    watchdog = std::thread([this] {
      std::unique_lock<std::mutex> lk(wd_mu);
      while (!stop) {
        wd_cv.wait_for(lk, std::chrono::milliseconds(100));
        if (stop || dead.load()) continue;
        const double dl = deadline.load();
        if (dl == 0.0) continue;
        std::string reason;
        bool fire = false;
        if (now_s() > dl) { fire = true; reason = "Cross-rank collective did not complete within " + std::to_string(timeout_s) + " s (a peer rank is missing or the ranks' call sequences diverged)"; }
        else if (nccl().CommGetAsyncError) {
          int st = 0;
          if (nccl().CommGetAsyncError(comm, &st) != 0 || (st != 0 && st != 7 /* ncclInProgress */)) { fire = true; reason = std::string("NCCL asynchronous error: ") + (nccl().GetErrorString ? nccl().GetErrorString(st) : "?"); }
        }
        if (fire) {
          why = reason + "; the communicator has been aborted.";
          aborting.store(true);            // before the abort: the call it unblocks reports `why`, not NCCL's post-abort error
          if (nccl().CommAbort) nccl().CommAbort(comm);
          dead.store(true);
        }
      }
    });
  }
  ~NcclComm() override {
    { std::lock_guard<std::mutex> lk(wd_mu); stop = true; }
    wd_cv.notify_all();
    if (watchdog.joinable()) watchdog.join();
    if (comm && !dead.load()) nccl().CommDestroy(comm);
  }
  void arm() override { deadline.store(now_s() + timeout_s); }
  void disarm() override { deadline.store(0.0); }
  bool async_error(std::string* w) override { if (dead.load()) { if (w) *w = why; return true; } return false; }
  void abort() override { dead.store(true); }
  void check(int rc, const char* what) {
    if (dead.load() || aborting.load()) throw CudaFail{why.empty() ? std::string("NCCL communicator was aborted after an earlier failure; create a new one (cb2_comm_init).") : why};
    if (rc != 0) throw CudaFail{std::string("NCCL error in ") + what + ": " + (nccl().GetErrorString ? nccl().GetErrorString(rc) : "?")};
  }
  int guarded_allreduce(void* buf, size_t n, int dtype, int op, cudaStream_t s) {
    if (dead.load()) return 0;
    arm();                               // the enqueue itself can block (connection set-up of the first collective)
    const int rc = nccl().AllReduce(buf, buf, n, dtype, op, comm, s);
    disarm();
    return rc;
  }
  // ncclFloat64 = 8; ncclSum = 0, ncclMax = 2
  void allreduce_sum(double* buf, size_t n, cudaStream_t s) override { check(guarded_allreduce(buf, n, 8, 0, s), "allreduce(sum)"); }
  void allreduce_max(double* buf, size_t n, cudaStream_t s) override { check(guarded_allreduce(buf, n, 8, 2, s), "allreduce(max)"); }
  void allreduce_sum_u8(unsigned char* buf, size_t n, cudaStream_t s) override { check(guarded_allreduce(buf, n, 1 /* ncclUint8 */, 0, s), "allreduce(sum, u8)"); }
};
#else
// Test-only rendezvous: `world` host threads of one process, each driving its own handle, meet in every collective.
struct LocalGroup {
  std::mutex mu;
  std::condition_variable cv;
  int world = 0, arrived = 0, generation = 0;
  std::vector<double*> bufs;
  std::vector<double> result;
  std::vector<unsigned char*> bufs8;
  std::vector<unsigned char> result8;
};
static std::mutex g_groups_mu;
static std::map<std::string, LocalGroup*>& local_groups() { static std::map<std::string, LocalGroup*> m; return m; }
struct LocalComm : Comm {
  LocalGroup* grp = nullptr;
  void reduce(double* buf, size_t n, bool is_max) {
    std::unique_lock<std::mutex> lk(grp->mu);
    const int gen = grp->generation;
    if (grp->arrived == 0) grp->bufs.assign(world, nullptr);
    grp->bufs[rank] = buf;
    if (++grp->arrived == world) {
      grp->result.assign(n, 0.0);
      for (size_t i = 0; i < n; ++i) {
        double v = grp->bufs[0][i];
        for (int r = 1; r < world; ++r) v = is_max ? std::max(v, grp->bufs[r][i]) : v + grp->bufs[r][i];
        grp->result[i] = v;
      }
      for (int r = 0; r < world; ++r) std::memcpy(grp->bufs[r], grp->result.data(), n * sizeof(double));
      grp->arrived = 0;
      ++grp->generation;
      grp->cv.notify_all();
    } else {
      grp->cv.wait(lk, [&] { return grp->generation != gen; });
    }
  }
  void allreduce_sum(double* buf, size_t n, cudaStream_t) override { reduce(buf, n, false); }
  void allreduce_max(double* buf, size_t n, cudaStream_t) override { reduce(buf, n, true); }
  void allreduce_sum_u8(unsigned char* buf, size_t n, cudaStream_t) override {
    std::unique_lock<std::mutex> lk(grp->mu);
    const int gen = grp->generation;
    if (grp->arrived == 0) grp->bufs8.assign(world, nullptr);
    grp->bufs8[rank] = buf;
    if (++grp->arrived == world) {
      grp->result8.assign(n, 0);
      for (size_t i = 0; i < n; ++i) { unsigned v = 0; for (int r = 0; r < world; ++r) v += grp->bufs8[r][i]; grp->result8[i] = (unsigned char)v; }
      for (int r = 0; r < world; ++r) std::memcpy(grp->bufs8[r], grp->result8.data(), n);
      grp->arrived = 0;
      ++grp->generation;
      grp->cv.notify_all();
    } else {
      grp->cv.wait(lk, [&] { return grp->generation != gen; });
    }
  }
};
#endif

}  // namespace cb2

using namespace cb2;

struct cb2_problem {
  // ---- host-side description (what the reference keeps in Trajectory / WorldModel / Sensor objects) ----
  int k = 6;
  std::vector<double> knots, ctrl;
  double gravity[3] = {0, 0, -9.80665};   // world_model.h:78
  std::vector<HostBody> bodies;
  std::unordered_map<int, int> body_slot;
  std::vector<HostSensor> sensors;
  std::string error;

  // ---- device state ----
  std::atomic<bool> uploaded{false};   // atomic: cb2_add_*_observations may run concurrently on DIFFERENT sensors (each clears it)
  int device = -1;
  cudaStream_t stream = nullptr, stream_imu = nullptr;   // IMU sweeps overlap the camera sweep on a second stream
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool imu_side_stream = std::getenv("CB2_IMU_STREAM") != nullptr;
  bool imu_pair_streams = std::getenv("CB2_NO_IMU_PAIR") == nullptr;
  bool debug_log = std::getenv("CB2_DEBUG") != nullptr;
  bool capturing = false;       // inside graphed(): stay on one stream
  bool speculative_imu = std::getenv("CB2_SPECULATIVE_IMU") != nullptr;   // measured neutral on C4 (Jacobian-mode IMU blocks cost +53 us over cost mode): off by default
  int imu_jac_point = -1;       // parameter buffer (0 / 1) whose IMU Jacobians, residuals and cost partials are current; -1 = none
  // Speculative sweep: when the previous step was accepted the next one very likely is too, so its trial point is evaluated with the
  // full residual + Jacobian sweep instead of the cost-only pass; on acceptance the sweep of the new x is then already done (saves the
  // cost pass, K0 and a reduction per accepted iteration), on rejection the extra Jacobian write is the price (CB2_NO_SPECULATIVE_SWEEP=1: off).
  bool speculative_sweep = std::getenv("CB2_NO_SPECULATIVE_SWEEP") == nullptr;
  int jac_point = -1;           // parameter buffer whose complete sweep (all sensors: J, r, cost partials) is current; -1 = none
  bool trial_was_sweep = false, sweep_skipped = false, speculate_next = true;
  bool defer_normal_sync = std::getenv("CB2_NO_DEFERRED_SYNC") == nullptr;
  bool last_sweep_imu = true;
  int n_cp = 0, n_seg = 0, N_c = 0, csz = 0, n_tiles = 0;
  long n_a = 0, n_tot = 0;
  std::vector<SensorDesc> h_desc;
  std::vector<SensorState> h_state;
  std::vector<int> tile_off = std::vector<int>(4, 0);   // per kind offsets into the tile table
  DevBuf<SensorDesc> d_desc;
  DevBuf<SensorState> d_state[2], d_state0;
  DevBuf<double> d_ctrl[2], d_ctrl0, d_knots, d_basis;
  // World model on the device: world points per parameter buffer; with freed rigid-body poses / model points (world_model.cpp:52-70) the
  // body states are double-buffered unknowns like everything else and the world points are recomputed from them.
  DevBuf<double> d_pw[2], d_body_q[2], d_body_t[2], d_pm[2], d_body_q0, d_body_t0, d_pm0;
  DevBuf<int> d_pt_body, d_body_u, d_pt_u;
  std::vector<int> h_body_u, h_pt_u;
  bool world_free = false;
  int n_points = 0, n_bodies = 0;
  WorldRefs world_refs() const { return WorldRefs{n_bodies, n_points, d_body_u.p, d_pt_u.p}; }
  DevBuf<EvalTile> d_tiles;
  // camera images (frames): one record per (camera, stamp), shared by the residual blocks of that image
  int n_frames = 0;
  DevBuf<int> d_frame_sensor, d_frame_seg;
  DevBuf<double> d_frame_stamp, d_frames;
  DevBuf<double> d_cost_partial;
  DevBuf<int> d_invalid_partial;
  DevBuf<double> d_scal;
  DevBuf<unsigned char> d_cp_ref;
  // normal equations
  DevBuf<int> d_c2off, d_plain_idx;
  DevBuf<CalibEntry> d_centries;
  DevBuf<double> d_cpartial;
  DevBuf<double> d_segA, d_segG, d_segB, d_segC, d_segGc, d_Aband, d_Bmat, d_Cmat, d_grad, d_diag, d_scaling, d_dtil2, d_ytil;
  // multi-GPU sharding: this rank owns the chunks [chunk_lo, chunk_hi) and the segments [g_lo, g_hi)
  std::shared_ptr<Comm> comm;   // may be shared between handles of one process (cb2_comm_clone)
  int world = 1, rank = 0;
  int chunk_lo = 0, chunk_hi = 0, g_lo = 0, g_hi = 0;
  long total_blocks = 0, total_residuals = 0;    // over ALL ranks (summary counts)
  DevBuf<unsigned char> d_cp_own;
  DevBuf<int> d_shared_idx;
  DevBuf<double> d_shared_buf, d_gradG, d_red, d_normbuf;
  int n_shared = 0;
  // Schur
  std::vector<ChunkPlan> chunks;
  std::vector<BandSys> h_l1;
  BandSys h_l2;
  DevBuf<BandSys> d_l1;
  DevBuf<BandSys> d_l2;
  DevBuf<int> d_chunk_sys, d_rowidx, d_colidx;
  DevBuf<double> d_L1, d_W1, d_T1, d_T2, d_rawdiag, d_Dinv;
  // level 1 by block cyclic reduction (cb2_cr.cuh): the default; CB2_SCHUR=band selects the chunked left-to-right band factor
  bool use_cr = true;
  int cr_nlevels = 0, cr_max_nblk = 0;
  // Border Gram product in two parts (border_gram_dmma_kernel): part = {blk_res, blk_mod, k_off, k_cnt, level it follows (-1: after every
  // level)}. Empty = one launch over all rows after the last level.
  struct GramPart { int res, mod, k_off, k_cnt, after_level; };
  std::vector<GramPart> gram_parts;
  bool gram_prereduce = std::getenv("CB2_GRAM_PREREDUCE") != nullptr;   // opt-in: folding the early partials beside the late levels gained nothing (profiles/r02_variants.md)
  bool calib_fork = std::getenv("CB2_NO_CALIB_FORK") == nullptr;
  bool back_cluster = std::getenv("CB2_BACK_CLUSTER") != nullptr;   // opt-in (measured neutral on C4, profiles/r02_variants.md): back-substitution of the narrow levels in one thread-block-cluster launch
  cudaStream_t stream_gram = nullptr;
  cudaEvent_t ev_gram[4] = {nullptr, nullptr, nullptr, nullptr};
  int cr_zsplit = 1;                // CTAs per block of a cyclic-reduction level (column split of the forward substitution + Schur update)
  bool cr_tma = false;              // non-first levels fetch their state with TMA bulk copies (cb2_cr.cuh)
  DevBuf<double> d_crD, d_crBd, d_crU, d_crWef, d_crL;
  // the separator level between the chunks by the same cyclic-reduction kernels (CB2_SEP_BAND=1: the single-CTA band factor instead)
  bool sep_cr = false, gram_dmma2 = false;
  int cr2_nlevels = 0;
  DevBuf<double> d_cr2D, d_cr2Bd, d_cr2U, d_cr2Wef, d_cr2L;
  DevBuf<unsigned> d_grid_sync;   // arrival counter of the fused back-substitution launch
  DevBuf<double> d_lm_partial;    // grid_reduce_last (lmkernels): per-CTA partial sums and arrival tickets of gradient_norm_kernel / apply_step_kernel
  DevBuf<unsigned> d_lm_ticket;
  int cur = 0;   // which of the two parameter buffers holds x
  size_t smem_eval[3] = {0, 0, 0};
  int max_tilepairs1 = 0, max_ksplit1 = 1;
  bool gram_dmma1 = false;   // level-1 Gram product on the FP64 tensor pipe
  // Cameras with <= 16 calibration unknowns: compact per-image Gram slots written by the sweep, expanded by expand_gram_kernel (cb2_normal.cuh);
  // their Jacobian is never read back. CB2_NO_SWEEP_GRAM=1 sends every sensor through accumulate_kernel (J re-read) instead.
  bool sweep_gram = std::getenv("CB2_NO_SWEEP_GRAM") == nullptr;
  DevBuf<double> d_gslots, d_gcta, d_segA2, d_segG2;
  DevBuf<int> d_ext_tab, d_ext_dst;
  DevBuf<int2> d_gslot_tab, d_gram_meta;   // per (segment, slot-producing camera): {first slot, count}; per such camera: {calib_off, n_calib}
  int n_gram_sensors = 0, n_plain_sensors = 0;
  int n_plain_idx = 0, acc_spc = 1;            // accumulate_kernel: its sensors (d_plain_idx) and the segments per CTA
  double* h_scal = nullptr;   // pinned
  double* h_param = nullptr;  // pinned: {radius, min_lm_diagonal, max_lm_diagonal} of the coming solve
  bool use_graphs = std::getenv("CB2_NO_GRAPHS") == nullptr;
  bool graph_collectives = std::getenv("CB2_NO_NCCL_GRAPHS") == nullptr;
  bool ne_shared_pending = false;   // multi-rank: the shared-row sums of the current normal equations ride in the next solve's collective
  struct GraphSlot { cudaGraphExec_t exec = nullptr; int64_t kernels = 0; };
  GraphSlot g_solve[4], g_trial[2], g_normal[4];   // LM phases as CUDA graphs, one per parameter buffer (cur = 0 / 1) [+ 2: multi-rank variant with the shared-row sums deferred into the solve]
  cb2_stats stats{};
  PhaseTimer timer;
  KernelProfiler kprof;
  double phase_ms[kPhCount] = {0, 0, 0, 0, 0, 0};
  bool scaling_set = false;

  ~cb2_problem() {
    if (stream) sync_stream(false);                   // device buffers are released (stream-ordered, default stream) after this body
    if (stream_imu) cudaStreamSynchronize(stream_imu);
    if (stream_gram) cudaStreamSynchronize(stream_gram);
    kprof.report();
    drop_graphs();
    if (h_scal) cudaFreeHost(h_scal);
    if (h_param) cudaFreeHost(h_param);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    for (auto& e : ev_gram) if (e) cudaEventDestroy(e);
    if (stream_gram) cudaStreamDestroy(stream_gram);
    if (stream_imu) cudaStreamDestroy(stream_imu);
    if (stream) cudaStreamDestroy(stream);
  }

  void drop_graphs() {
#ifndef CB2_EMUL
    for (int i = 0; i < 4; ++i) { for (auto* set : {g_solve, g_normal}) { if (set[i].exec) cudaGraphExecDestroy(set[i].exec); set[i] = GraphSlot{}; } }
    for (int i = 0; i < 2; ++i) { if (g_trial[i].exec) cudaGraphExecDestroy(g_trial[i].exec); g_trial[i] = GraphSlot{}; }
#endif
  }
  // Runs `body` (kernel launches and memsets on `stream` only) as a CUDA graph: captured and instantiated on first use, replayed afterwards.
  // With several ranks the cross-rank sums inside the body are captured too (NCCL collectives are graph-capturable); CB2_NO_NCCL_GRAPHS=1
  // keeps multi-rank runs on plain launches. Not while the per-kernel profiler is on.
  template <class F>
  void graphed(GraphSlot& slot, F&& body) {
#ifndef CB2_EMUL
    const bool graph = use_graphs && !kprof.on && (world == 1 || graph_collectives);
    if (graph && slot.exec) { CB2_CUDA(cudaGraphLaunch(slot.exec, stream)); stats.kernel_launches += slot.kernels; return; }
    const int64_t before = stats.kernel_launches;
    if (graph) { CB2_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal)); capturing = true; }
    body();
    capturing = false;
    if (graph) {
      cudaGraph_t g = nullptr;
      CB2_CUDA(cudaStreamEndCapture(stream, &g));
      CB2_CUDA(cudaGraphInstantiate(&slot.exec, g, 0));
      CB2_CUDA(cudaGraphDestroy(g));
      slot.kernels = stats.kernel_launches - before;
      CB2_CUDA(cudaGraphLaunch(slot.exec, stream));
    }
#else
    body();
#endif
  }

  std::mutex error_mu;                 // concurrent observation uploads (one thread per sensor) may fail at the same time
  int fail(int code, const std::string& msg) { std::lock_guard<std::mutex> lk(error_mu); error = msg; return code; }

  // Waits for the library's stream. With several ranks a collective that a peer never joins (crashed rank, mismatched call sequence)
  // would block forever: the wait then polls with a deadline (CB2_COLLECTIVE_TIMEOUT_S, default 60 s) and the communicator's
  // asynchronous error state (NCCL async error check); on either, the communicator is aborted — which makes the stuck kernel
  // return — and the call fails with CB2_INTERNAL instead of hanging.
  void sync_stream(bool may_throw = true) {
    if (world <= 1 || !comm) {
      const cudaError_t e = cudaStreamSynchronize(stream);
      if (e != cudaSuccess && may_throw) throw CudaFail{std::string("CUDA error: ") + cudaGetErrorString(e) + " at cudaStreamSynchronize"};
      return;
    }
    comm->arm();
    for (;;) {
      const cudaError_t q = cudaStreamQuery(stream);
      if (q == cudaSuccess) { comm->disarm(); return; }
      if (q != cudaErrorNotReady) {
        comm->disarm();
        if (may_throw) throw CudaFail{std::string("CUDA error: ") + cudaGetErrorString(q) + " at cudaStreamQuery"};
        return;
      }
      if (!comm->dead.load()) continue;
      // the watchdog aborted the communicator (deadline passed or asynchronous NCCL error): the stuck collective has been torn down
      comm->disarm();
      cudaStreamSynchronize(stream);
      cudaGetLastError();
      uploaded = false;
      std::string why;
      comm->async_error(&why);
      if (may_throw) throw CudaFail{why.empty() ? std::string("Cross-rank collective failed; the communicator has been aborted.") : why};
      return;
    }
  }

  // ------------------------------------------------------------------------------------------------------------
  // Spline bookkeeping (host): BSpline::GetSplineIndex bspline.hpp:139-151, basis matrices bspline.hpp:192-244.
  // ------------------------------------------------------------------------------------------------------------
  static void basis_matrix(const std::vector<double>& kn, int i, double* M /*36*/) {
    // Qin's recursion, built bottom-up: M_1 = [1]; M_j = [M_{j-1}; 0] A_j + [0; M_{j-1}] B_j.
    std::vector<double> cur(1, 1.0);
    for (int j = 2; j <= kK; ++j) {
      const int n = j - 1;
      std::vector<double> nxt(size_t(j) * j, 0.0);
      for (int idx = 0; idx < n; ++idx) {
        const int jj = i - j + 2 + idx;
        const double den = kn[jj + j - 1] - kn[jj];
        const double d0 = den <= 0.0 ? 0.0 : (kn[i] - kn[jj]) / den;
        const double d1 = den <= 0.0 ? 0.0 : (kn[i + 1] - kn[i]) / den;
        // column idx and idx+1 of the new matrix receive contributions from column idx of the old one
        for (int r = 0; r < n; ++r) {
          const double v = cur[size_t(r) * n + idx];
          nxt[size_t(r) * j + idx] += v * (1.0 - d0);
          nxt[size_t(r) * j + idx + 1] += v * d0;
          nxt[size_t(r + 1) * j + idx] += v * (-d1);
          nxt[size_t(r + 1) * j + idx + 1] += v * d1;
        }
      }
      cur.swap(nxt);
    }
    std::memcpy(M, cur.data(), sizeof(double) * kK * kK);
  }
  int spline_index(double t) const {
    const double* vk = knots.data() + (kK - 1);
    const int nv = int(knots.size()) - 2 * (kK - 1);
    if (t == vk[nv - 1]) return nv - 2;
    if (!(t < vk[nv - 1])) return -1;
    // Same answer as upper_bound(valid knots, t) - 1 (bspline.hpp:139-151); the knots are (nearly always) uniform, so a linear guess
    // corrected against the knot values replaces the binary search.
    if (t >= vk[0] && nv >= 2) {
      int i = int((t - vk[0]) / (vk[nv - 1] - vk[0]) * (nv - 1));
      i = std::min(std::max(i, 0), nv - 2);
      int steps = 0;
      while (i > 0 && vk[i] > t && steps < 4) { --i; ++steps; }
      while (i < nv - 2 && vk[i + 1] <= t && steps < 4) { ++i; ++steps; }
      if (vk[i] <= t && t < vk[i + 1]) return i;
    }
    return int(std::upper_bound(vk, vk + nv, t) - vk) - 1;
  }

  // ------------------------------------------------------------------------------------------------------------
  // Upload: validate, pack SoA (observations sorted by spline segment), allocate every device array.
  // ------------------------------------------------------------------------------------------------------------
  int upload() {
    try {
      return upload_impl();
    } catch (const CudaFail& f) {
      return fail(CB2_INTERNAL, f.msg);
    } catch (const std::bad_alloc&) {
      return fail(CB2_INTERNAL, "Out of host memory while packing the problem.");
    }
  }

  int ensure_device() {
    if (stream) return CB2_OK;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0)
      return fail(CB2_INTERNAL, "No CUDA device available: calico_b200 has no CPU fallback.");
    if (device >= 0) CB2_CUDA(cudaSetDevice(device));
#ifndef CB2_EMUL
    {
      int dev = 0;
      CB2_CUDA(cudaGetDevice(&dev));
      cudaMemPool_t pool;
      CB2_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
      unsigned long long keep = ~0ull;   // never trim: freed blocks are reused by the next problem of this process
      CB2_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
#endif
    CB2_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    {
      int lo = 0, hi = 0;   // the IMU sweep runs at a higher priority so that its long-latency CTAs are placed as soon as a slot frees up
      CB2_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CB2_CUDA(cudaStreamCreateWithPriority(&stream_imu, cudaStreamNonBlocking, hi));
    }
    CB2_CUDA(cudaEventCreate(&ev_fork));
    CB2_CUDA(cudaEventCreate(&ev_join));
    CB2_CUDA(cudaStreamCreateWithFlags(&stream_gram, cudaStreamNonBlocking));
    for (auto& e : ev_gram) CB2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CB2_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h_scal), sizeof(double) * kScCount));
    CB2_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h_param), sizeof(double) * 4));
    return CB2_OK;
  }

  bool timing_log = std::getenv("CB2_TIMING") != nullptr;
  int upload_impl() {
    const double t_upload0 = now_s();
    uploaded = false;
    if (stream) sync_stream();
    drop_graphs();   // every device pointer baked into a captured solve phase is about to change
    if (knots.empty() || ctrl.empty()) return fail(CB2_FAILED_PRECONDITION, "Trajectory has not been set.");
    if (k != kK) return fail(CB2_UNIMPLEMENTED, "Only spline order 6 (calico::Trajectory::kSplineOrder, trajectory.h:28) is supported.");
    world_free = false;
    for (const auto& b : bodies) world_free = world_free || !b.pose_const || !b.model_const;
    int rc = ensure_device();
    if (rc != CB2_OK) return rc;
    n_cp = int(ctrl.size() / 6);
    n_seg = n_cp - (kK - 1);
    if (n_seg < 1) return fail(CB2_INVALID_ARGUMENT, "Trajectory has too few control points.");
    n_a = 6L * n_cp;
    rc = plan_chunks();
    if (rc != CB2_OK) return rc;
    // World points p_w = q_wm * p_m + t_wm; rigid-body states and the body of every point.
    std::vector<double> pw, body_q, body_t, pm;
    std::vector<int> pt_body;
    for (size_t bi = 0; bi < bodies.size(); ++bi) {
      HostBody& b = bodies[bi];
      b.pw0 = int(pw.size() / 3);
      const M3 R = quat_matrix(Q4{b.q[0], b.q[1], b.q[2], b.q[3]});
      for (int j = 0; j < 4; ++j) body_q.push_back(b.q[j]);
      for (int j = 0; j < 3; ++j) body_t.push_back(b.t[j]);
      for (size_t f = 0; f < b.feature_ids.size(); ++f) {
        const V3 p = R * v3(b.pts[3 * f], b.pts[3 * f + 1], b.pts[3 * f + 2]);
        pw.push_back(p.x + b.t[0]); pw.push_back(p.y + b.t[1]); pw.push_back(p.z + b.t[2]);
        pm.push_back(b.pts[3 * f]); pm.push_back(b.pts[3 * f + 1]); pm.push_back(b.pts[3 * f + 2]);
        pt_body.push_back(int(bi));
      }
    }
    n_points = int(pt_body.size()); n_bodies = int(bodies.size());
    std::vector<double> basis(size_t(n_seg) * 36);
    for (int s = 0; s < n_seg; ++s) basis_matrix(knots, s + kK - 1, &basis[size_t(s) * 36]);
    // Sensors.
    const int ns = int(sensors.size());
    h_desc.assign(ns, SensorDesc{});
    h_state.assign(ns, SensorState{});
    std::vector<unsigned char> cp_ref(n_cp, 0);
    total_blocks = total_residuals = 0;
    N_c = 0; csz = 0;
    std::vector<int> c2off(std::max(ns, 1), 0);
    std::vector<EvalTile> tiles;
    std::vector<std::vector<EvalTile>> tiles_by_kind(3);
    std::vector<int> frame_sensor, frame_seg;
    std::vector<double> frame_stamp;
    size_t n_gslots = 0, n_gcta = 0;
    std::vector<char> gram_sensor(std::max(ns, 1), 0);
    std::vector<size_t> gcta_off(std::max(ns, 1), 0);
    std::vector<int> gcta_cnt(std::max(ns, 1), 0);
    int64_t* h2d = &stats.h2d_bytes;
    // Phase A (one host thread per sensor): validation, segment lookup, counting sort by segment, SoA packing.
    struct Packed {
      int rc = CB2_OK; std::string err;
      int want = 0, n_active = 0; long blocks = 0, residuals = 0; bool ref_any = false;
      std::vector<int> seg_start, frame_seg, frame_obs, seg_frame;
      std::vector<double> frame_stamp;
      std::vector<unsigned char> cp_ref, pt_ref;
      // the per-observation arrays, packed straight into ONE pinned staging block (offsets in bytes, 256-byte aligned)
      PinnedBuf pin;
      size_t off_stamp = 0, off_meas = 0, off_seg = 0, off_pt = 0, off_frm = 0, off_perm = 0, bytes = 0;
    };
    std::vector<Packed> packed(ns);
    // A sensor with many observations can be packed by several threads (CB2_PACK_SUB_THREADS=n for sensors of at least CB2_PACK_SPLIT_MIN
    // observations). Off by default: on C4 (8 cameras of 125 k observations, 16 host cores) two threads per camera on top of the one
    // thread per sensor measured slower and erratic (pack 6-8 ms with spikes to 34 ms against 4.3-6.2 ms).
    const int pack_split_min = env_int("CB2_PACK_SPLIT_MIN", 65536);
    const int sub_threads = std::max(1, env_int("CB2_PACK_SUB_THREADS", 1));
    auto pack_sensor = [&](int si) {
      HostSensor& s = sensors[si];
      Packed& P = packed[si];
      auto bad = [&](int code, const char* msg) { P.rc = code; P.err = msg; };
      const int want = s.kind == kCamera ? camera_num_params(s.model) : imu_num_params(s.model);
      P.want = want;
      if (want < 0) return bad(CB2_FAILED_PRECONDITION, "Cannot add sensor parameters. Model is not yet defined.");   // camera.cpp:95
      if (int(s.intr.size()) != want) return bad(CB2_INVALID_ARGUMENT, "Invalid number of intrinsics parameters.");
      const int m = s.m(), n = s.n_obs();
      P.cp_ref.assign(n_cp, 0);
      if (world_free && s.kind == kCamera) P.pt_ref.assign(n_points, 0);
      // Every pass below runs over T disjoint observation (or output) ranges: a large sensor is packed by T threads (sub_threads: the
      // cores the per-sensor threads leave idle), a small one inline (T = 1). The result does not depend on T.
      const int T = n >= pack_split_min ? std::max(1, sub_threads) : 1;
      auto parallel = [&](auto&& fn) {
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) th.emplace_back([&fn, t] { fn(t); });
        fn(0);
        for (auto& x : th) x.join();
      };
      auto lo_of = [&](long total, int t) { return int(total * t / T); };
      // Pass 1: segment of each active observation, per-range histograms (-> stable counting sort by segment), reference marks.
      std::vector<int> seg_of(n, -1);
      std::vector<std::vector<int>> hist(T, std::vector<int>(n_seg + 1, 0));
      std::vector<std::vector<unsigned char>> seg_seen(T, std::vector<unsigned char>(n_seg, 0));
      std::vector<std::vector<unsigned char>> pt_seen(T);
      struct RangeStat { int rc = CB2_OK; const char* err = nullptr; long blocks = 0; int n_active = 0, n_frames = 0; };
      std::vector<RangeStat> rs(T);
      parallel([&](int t) {
        RangeStat& R = rs[t];
        std::vector<int>& count = hist[t];
        unsigned char* seen = seg_seen[t].data();
        if (!P.pt_ref.empty()) pt_seen[t].assign(n_points, 0);
        for (int o = lo_of(n, t), oe = lo_of(n, t + 1); o < oe; ++o) {
          if (s.kind == kCamera && !s.outlier.empty() && s.outlier[o]) continue;   // camera.cpp:121-124
          if (s.kind == kCamera && s.body_slot[o] < 0) {                            // camera.cpp:125-131
            R.rc = CB2_FAILED_PRECONDITION; R.err = "Attempted to create cost function from an observation for a rigidbody that does not exist in the world model."; return;
          }
          const int sg = spline_index(s.stamp[o]);
          if (sg < 0 || sg >= n_seg) { R.rc = CB2_INVALID_ARGUMENT; R.err = "Observation stamp is outside the valid knots of the trajectory."; return; }
          seen[sg] = 1;                                              // referenced by some rank's residual block
          if (!pt_seen[t].empty()) pt_seen[t][bodies[s.body_slot[o]].pw0 + s.feat_slot[o]] = 1;
          ++R.blocks;
          if (sg < g_lo || sg >= g_hi) continue;                     // another rank's time range
          seg_of[o] = sg;
          ++count[sg + 1];
          ++R.n_active;
        }
      });
      int n_active = 0;
      for (int t = 0; t < T; ++t) {
        if (rs[t].rc != CB2_OK) return bad(rs[t].rc, rs[t].err);   // the first failing range in observation order
        P.blocks += rs[t].blocks; n_active += rs[t].n_active;
        for (int g = 0; g < n_seg; ++g) if (seg_seen[t][g]) for (int c = 0; c < kK; ++c) P.cp_ref[g + c] = 1;
        if (!pt_seen[t].empty()) for (int q = 0; q < n_points; ++q) P.pt_ref[q] |= pt_seen[t][q];
      }
      P.residuals = P.blocks * m;
      P.ref_any = P.blocks > 0;
      // seg_start = prefix sums of the summed histograms; hist[t] becomes range t's write cursor per segment (ranges in order: stable)
      P.seg_start.assign(n_seg + 1, 0);
      for (int g = 0; g < n_seg; ++g) {
        int c = 0;
        for (int t = 0; t < T; ++t) c += hist[t][g + 1];
        P.seg_start[g + 1] = P.seg_start[g] + c;
      }
      for (int g = 0; g < n_seg; ++g) {
        int at = P.seg_start[g];
        for (int t = 0; t < T; ++t) { const int c = hist[t][g + 1]; hist[t][g + 1] = at; at += c; }
      }
      const std::vector<int>& seg_start = P.seg_start;
      s.perm.assign(n_active, 0);
      parallel([&](int t) {
        int* cursor = hist[t].data() + 1;
        for (int o = lo_of(n, t), oe = lo_of(n, t + 1); o < oe; ++o) if (seg_of[o] >= 0) s.perm[cursor[seg_of[o]]++] = o;
      });
      if (s.kind == kCamera) {
        // The corners of one image (same stamp) must be adjacent: order each segment's observations by stamp (stable; usually a no-op).
        auto by_stamp = [&](int a, int b) { return s.stamp[a] < s.stamp[b]; };
        parallel([&](int t) {
          for (int g = lo_of(n_seg, t), ge = lo_of(n_seg, t + 1); g < ge; ++g) {
            int* b = s.perm.data() + seg_start[g];
            int* e = s.perm.data() + seg_start[g + 1];
            if (!std::is_sorted(b, e, by_stamp)) std::stable_sort(b, e, by_stamp);
          }
        });
      }
      s.n_active = P.n_active = n_active;
      {
        auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
        size_t off = 0;
        P.off_stamp = off; off = al(off + size_t(n_active) * 8);
        P.off_meas = off; off = al(off + size_t(n_active) * m * 8);
        P.off_seg = off; off = al(off + size_t(n_active) * 4);
        P.off_pt = off; off = al(off + (s.kind == kCamera ? size_t(n_active) * 4 : 0));
        P.off_frm = off; off = al(off + (s.kind == kCamera ? size_t(n_active) * 4 : 0));
        P.off_perm = off; off = al(off + size_t(n_active) * 4);
        P.bytes = off;
        P.pin.acquire(std::max<size_t>(off, 256));
      }
      unsigned char* base = static_cast<unsigned char*>(P.pin.p);
      double* p_stamp = reinterpret_cast<double*>(base + P.off_stamp);
      double* p_meas = reinterpret_cast<double*>(base + P.off_meas);
      int* p_seg = reinterpret_cast<int*>(base + P.off_seg);
      int* p_pt = reinterpret_cast<int*>(base + P.off_pt);
      int* p_frm = reinterpret_cast<int*>(base + P.off_frm);
      std::memcpy(base + P.off_perm, s.perm.data(), size_t(n_active) * 4);
      // Pass 3a: the SoA arrays in packed order; image starts (a new stamp) counted per range.
      parallel([&](int t) {
        int nf = 0;
        for (int i = lo_of(n_active, t), ie = lo_of(n_active, t + 1); i < ie; ++i) {
          const int o = s.perm[i];
          p_stamp[i] = s.stamp[o];
          p_seg[i] = seg_of[o];
          for (int q = 0; q < m; ++q) p_meas[size_t(i) * m + q] = s.meas[size_t(o) * m + q];
          if (s.kind == kCamera) {
            p_pt[i] = bodies[s.body_slot[o]].pw0 + s.feat_slot[o];
            if (i == 0 || s.stamp[o] != s.stamp[s.perm[i - 1]]) ++nf;
          }
        }
        rs[t].n_frames = nf;
      });
      if (s.kind == kCamera) {
        // Pass 3b: per-image records (segment, stamp, first observation) and the sensor-local image index of every observation
        // (SensorDesc::frame_base makes it global); range t's images start at the number of images before it.
        int total_frames = 0;
        std::vector<int> frame0(T);
        for (int t = 0; t < T; ++t) { frame0[t] = total_frames; total_frames += rs[t].n_frames; }
        P.frame_seg.resize(total_frames); P.frame_stamp.resize(total_frames); P.frame_obs.resize(total_frames);
        parallel([&](int t) {
          int f = frame0[t];                                       // images started before the current observation
          for (int i = lo_of(n_active, t), ie = lo_of(n_active, t + 1); i < ie; ++i) {
            if (i == 0 || p_stamp[i] != p_stamp[i - 1]) { P.frame_seg[f] = p_seg[i]; P.frame_stamp[f] = p_stamp[i]; P.frame_obs[f] = i; ++f; }
            p_frm[i] = f - 1;
          }
        });
      }
      if (s.kind == kCamera) {                                     // image CSR for the structured accumulation (cb2_normal.cuh)
        P.frame_obs.push_back(n_active);
        P.seg_frame.assign(n_seg + 1, 0);
        for (int sg : P.frame_seg) ++P.seg_frame[sg + 1];
        for (int g = 0; g < n_seg; ++g) P.seg_frame[g + 1] += P.seg_frame[g];
      }
    };
#ifdef CB2_EMUL
    for (int si = 0; si < ns; ++si) pack_sensor(si);
#else
    {
      const int nth = std::max(1, std::min<int>(ns, int(std::thread::hardware_concurrency())));
      std::atomic<int> next{0};
      std::vector<std::thread> pool;
      for (int th = 0; th < nth; ++th) pool.emplace_back([&] { for (int si = next++; si < ns; si = next++) pack_sensor(si); });
      for (auto& th : pool) th.join();
    }
#endif
    const double t_packed = now_s();
    // Phase B (serial): unknown layout, device buffers, uploads.
    for (int si = 0; si < ns; ++si) {
      HostSensor& s = sensors[si];
      Packed& P = packed[si];
      if (P.rc != CB2_OK) return fail(P.rc, P.err);
      const int want = P.want, m = s.m(), n_active = P.n_active;
      const bool ref_any = P.ref_any;
      total_blocks += P.blocks; total_residuals += P.residuals;
      for (int c = 0; c < n_cp; ++c) cp_ref[c] |= P.cp_ref[c];
      const int frame_base = int(frame_stamp.size());
      frame_sensor.insert(frame_sensor.end(), P.frame_stamp.size(), si);
      frame_seg.insert(frame_seg.end(), P.frame_seg.begin(), P.frame_seg.end());
      frame_stamp.insert(frame_stamp.end(), P.frame_stamp.begin(), P.frame_stamp.end());
      const std::vector<int>& seg_start = P.seg_start;
      // ONE asynchronous host -> device copy per sensor, from the pinned block its packing thread filled.
      s.d_obs.alloc(std::max<size_t>(P.bytes, 256), false);
      if (P.bytes) { CB2_CUDA(cudaMemcpyAsync(s.d_obs.p, P.pin.p, P.bytes, cudaMemcpyHostToDevice, nullptr)); *h2d += int64_t(P.bytes); }
      s.d_perm = reinterpret_cast<const int*>(s.d_obs.p + P.off_perm);
      s.d_seg_start.upload(seg_start, h2d);
      if (s.kind == kCamera) { s.d_frame_obs.upload(P.frame_obs, h2d); s.d_seg_frame.upload(P.seg_frame, h2d); }
      // Unknown layout of this sensor (constant or unreferenced blocks drop out, as in Ceres's reduced program).
      SensorDesc& d = h_desc[si];
      d.kind = s.kind; d.model = s.model; d.ni = want; d.m = m; d.n_obs = n_active; d.frame_base = frame_base;
      const bool ref = ref_any;   // referenced by a residual block on ANY rank: same unknown layout everywhere
      int u = 0;
      d.u_intr = (ref && s.en_intr) ? u : -1; if (d.u_intr >= 0) u += want;
      d.u_rot = (ref && s.en_extr) ? u : -1; if (d.u_rot >= 0) u += 3;
      d.u_trans = (ref && s.en_extr) ? u : -1; if (d.u_trans >= 0) u += 3;
      d.u_lat = (ref && s.en_lat) ? u : -1; if (d.u_lat >= 0) u += 1;
      d.n_calib = u; d.calib_off = N_c; N_c += u;
      c2off[si] = csz; csz += u * u;
      // Stored Jacobian columns: canonical order [intr | rot | trans | lat]; the gyroscope's translation columns are
      // structurally zero (gyroscope_cost_functor.h:59-118 never reads t) and are not stored.
      int nj = 0;
      if (d.u_intr >= 0) for (int j = 0; j < want; ++j) { d.jcanon[nj] = j; d.junk[nj] = d.u_intr + j; ++nj; }
      if (d.u_rot >= 0) for (int j = 0; j < 3; ++j) { d.jcanon[nj] = want + j; d.junk[nj] = d.u_rot + j; ++nj; }
      if (d.u_trans >= 0 && s.kind != kGyroscope) for (int j = 0; j < 3; ++j) { d.jcanon[nj] = want + 3 + j; d.junk[nj] = d.u_trans + j; ++nj; }
      if (d.u_lat >= 0) { d.jcanon[nj] = want + 6; d.junk[nj] = d.u_lat; ++nj; }
      d.n_jcal = nj; d.jw = kCpCols + nj;
      d.gslots = nullptr; d.gslot_base = 0;
      if (sweep_gram && ns <= kAccMaxSensors && s.kind == kCamera && d.n_calib <= 16 && n_active > 0) {
        d.gslot_base = int(n_gslots);
        const size_t n_cta = size_t((n_active + eval_tile(kCamera) - 1) / eval_tile(kCamera));
        n_gslots += P.frame_stamp.size() + n_cta;
        const size_t n_part = n_cta * (eval_tile(kCamera) / 32);      // one calibration partial per (CTA, warp)
        gram_sensor[si] = 1; gcta_off[si] = n_gcta; gcta_cnt[si] = int(n_part);
        int canon_of_unknown[16];
        for (int u = 0; u < 16; ++u) canon_of_unknown[u] = -1;
        for (int j = 0; j < nj; ++j) canon_of_unknown[d.junk[j]] = d.jcanon[j];
        for (int q = 0; q < 2; ++q) for (int col = 0; col < 24; ++col) d.gfield[q][col] = (unsigned char)gram_field(want, q, col, canon_of_unknown, 16);
        n_gcta += n_part;
      }
      s.d_r.alloc(size_t(n_active) * m);
      s.d_J.alloc(size_t(n_active) * m * d.jw);
      s.d_valid.alloc(n_active);
      d.stamp = reinterpret_cast<const double*>(s.d_obs.p + P.off_stamp); d.meas = reinterpret_cast<const double*>(s.d_obs.p + P.off_meas);
      d.seg = reinterpret_cast<const int*>(s.d_obs.p + P.off_seg); d.pt = reinterpret_cast<const int*>(s.d_obs.p + P.off_pt);
      d.frm = reinterpret_cast<const int*>(s.d_obs.p + P.off_frm); d.seg_start = s.d_seg_start.p;
      d.r = s.d_r.p; d.J = s.d_J.p; d.valid = s.d_valid.p;
      d.frame_obs = s.kind == kCamera ? s.d_frame_obs.p : nullptr; d.seg_frame = s.kind == kCamera ? s.d_seg_frame.p : nullptr;
      for (int o0 = 0, T = eval_tile(s.kind); o0 < n_active; o0 += T) tiles_by_kind[s.kind].push_back(EvalTile{si, o0, std::min(T, n_active - o0)});
      // state
      SensorState& st = h_state[si];
      st.kind = s.kind; st.model = s.model; st.ni = want;
      for (int j = 0; j < kMaxIntrinsics; ++j) st.intr[j] = j < want ? s.intr[j] : 0.0;
      st.q = Q4{s.q[0], s.q[1], s.q[2], s.q[3]}; st.t = v3(s.t[0], s.t[1], s.t[2]);
      st.latency = s.latency; st.inv_sigma = 1.0 / s.sigma; st.loss_type = s.loss_type; st.loss_scale = s.loss_scale;
      smem_eval[s.kind] = std::max(smem_eval[s.kind], size_t(rec_size(s.kind, want)) * eval_rec_stride(s.kind) * sizeof(double));
    }
    // Freed world-model blocks join the calibration vector after the sensors' blocks: per rigid body rotation (3, tangent) + translation (3)
    // when its pose is estimated, 3 per model point when its definition is; like Ceres's reduced program, only blocks some residual
    // block (of any rank) references.
    n_sensor_unknowns = N_c;
    h_body_u.assign(2 * std::max(n_bodies, 1), -1);
    h_pt_u.assign(std::max(n_points, 1), -1);
    n_world_pose_blocks = n_world_point_blocks = 0;
    if (world_free) {
      std::vector<unsigned char> pt_ref(n_points, 0);
      for (int si = 0; si < ns; ++si) if (!packed[si].pt_ref.empty()) for (int p = 0; p < n_points; ++p) pt_ref[p] |= packed[si].pt_ref[p];
      for (int bi = 0; bi < n_bodies; ++bi) {
        const HostBody& b = bodies[bi];
        bool ref = false;
        for (size_t f = 0; f < b.feature_ids.size(); ++f) ref = ref || pt_ref[b.pw0 + f];
        if (!ref) continue;
        if (!b.pose_const) { h_body_u[2 * bi] = N_c; h_body_u[2 * bi + 1] = N_c + 3; N_c += 6; ++n_world_pose_blocks; }
        if (!b.model_const) for (size_t f = 0; f < b.feature_ids.size(); ++f) if (pt_ref[b.pw0 + f]) { h_pt_u[b.pw0 + f] = N_c; N_c += 3; ++n_world_point_blocks; }
      }
    }
    n_tot = n_a + N_c;
    if (N_c > kRedThreads) return fail(CB2_UNIMPLEMENTED, "More than 512 calibration unknowns (sensor blocks + freed world-model blocks) are not supported.");
    tile_off[0] = 0;
    for (int kd = 0; kd < 3; ++kd) { tiles.insert(tiles.end(), tiles_by_kind[kd].begin(), tiles_by_kind[kd].end()); tile_off[kd + 1] = int(tiles.size()); }
    n_tiles = int(tiles.size());
    d_tiles.upload(tiles, h2d);
    n_frames = int(frame_stamp.size());
    d_frame_sensor.upload(frame_sensor, h2d); d_frame_seg.upload(frame_seg, h2d); d_frame_stamp.upload(frame_stamp, h2d);
    d_frames.alloc(size_t(std::max(n_frames, 1)) * FrameRec::kSize);
    d_cost_partial.alloc(std::max(n_tiles, 1)); d_invalid_partial.alloc(std::max(n_tiles, 1));
    d_gslots.alloc(n_gslots * kGramSlot, false);
    d_gcta.alloc(n_gcta * kGramCta);
    { std::vector<int> tab(32 * kExpExt), dst(32 * kExpExt); expand_ext_table(tab.data(), dst.data()); d_ext_tab.upload(tab, h2d); d_ext_dst.upload(dst, h2d); }
    {
      // where expand_gram_kernel finds the slots of every (segment, camera): slot = gslot_base + image + CTA-in-sensor (see SensorDesc::gslots)
      std::vector<int> gram_ids;
      for (int si = 0; si < ns; ++si) if (gram_sensor[si]) gram_ids.push_back(si);
      const int ng = int(gram_ids.size()), nsl_t = std::max(g_hi - g_lo, 0);
      std::vector<int2> tab(size_t(std::max(nsl_t, 1)) * std::max(ng, 1), int2{0, 0}), meta(std::max(ng, 1), int2{0, 0});
      for (int j = 0; j < ng; ++j) {
        const int si = gram_ids[j];
        const Packed& P = packed[si];
        meta[j] = int2{h_desc[si].calib_off, h_desc[si].n_calib};
        for (int gl = 0; gl < nsl_t; ++gl) {
          const int g = g_lo + gl;
          const int f_begin = P.seg_frame[g], f_end = P.seg_frame[g + 1];
          if (f_end == f_begin) continue;
          const int o_first = P.seg_start[g], o_last = P.seg_start[g + 1] - 1;
          const int lo = h_desc[si].gslot_base + f_begin + o_first / eval_tile(kCamera);
          tab[size_t(gl) * ng + j] = int2{lo, h_desc[si].gslot_base + f_end - 1 + o_last / eval_tile(kCamera) - lo + 1};
        }
      }
      d_gslot_tab.upload(tab, h2d); d_gram_meta.upload(meta, h2d);
    }
    n_gram_sensors = n_plain_sensors = 0;
    for (int si = 0; si < ns; ++si) {
      h_desc[si].gcta = nullptr;
      if (gram_sensor[si]) { h_desc[si].gslots = d_gslots.p; h_desc[si].gcta = d_gcta.p + gcta_off[si] * kGramCta; ++n_gram_sensors; }
      else if (h_desc[si].n_obs > 0) ++n_plain_sensors;
    }
    d_desc.upload(h_desc, h2d);
    {
      // accumulate_kernel's sensors and how many segments share one of its CTAs (one warp per (segment, sensor) pair)
      std::vector<int> plain;
      for (int si = 0; si < ns; ++si) if (!gram_sensor[si]) plain.push_back(si);
      n_plain_idx = int(plain.size());
      acc_spc = std::getenv("CB2_ACC_SPC1") ? 1 : (n_plain_idx <= 1 ? 4 : (n_plain_idx == 2 ? 2 : 1));
      if (plain.empty()) plain.push_back(0);
      d_plain_idx.upload(plain, h2d);
    }
    for (int b = 0; b < 2; ++b) d_state[b].upload(h_state, h2d);
    d_state0.upload(h_state, h2d);
    for (int b = 0; b < 2; ++b) d_ctrl[b].upload(ctrl, h2d);
    d_ctrl0.upload(ctrl, h2d);
    d_knots.upload(knots, h2d); d_basis.upload(basis, h2d);
    for (int b = 0; b < 2; ++b) { d_pw[b].upload(pw, h2d); d_body_q[b].upload(body_q, h2d); d_body_t[b].upload(body_t, h2d); d_pm[b].upload(pm, h2d); }
    d_body_q0.upload(body_q, h2d); d_body_t0.upload(body_t, h2d); d_pm0.upload(pm, h2d);
    d_pt_body.upload(pt_body, h2d); d_body_u.upload(h_body_u, h2d); d_pt_u.upload(h_pt_u, h2d);
    d_cp_ref.upload(cp_ref, h2d);
    n_cp_referenced = 0;
    for (auto v : cp_ref) n_cp_referenced += v;
    {
      std::vector<unsigned char> own(n_cp, kCpPeer);
      for (int c = chunk_lo; c < chunk_hi; ++c) for (int i = chunks[c].a; i < chunks[c].b; ++i) own[i] = kCpOwned;
      for (size_t c = 0; c + 1 < chunks.size(); ++c) for (int i = chunks[c].b; i < chunks[c].b + 5; ++i) own[i] = kCpShared;
      d_cp_own.upload(own, h2d);
    }
    d_scal.alloc(kScCount);
    d_grid_sync.alloc(4);
    d_lm_partial.alloc(16 * kLmMaxCtas); d_lm_ticket.alloc(4);   // grid_reduce_last: [gradient_norm | apply_step]
    cur = 0;
    // Normal-equation storage.
    d_c2off.upload(c2off, h2d);
    std::vector<CalibEntry> ce;
    const int nsl_ce = std::max(g_hi - g_lo, 0);
    for (int si = 0; si < ns; ++si) {
      const SensorDesc& d = h_desc[si];
      if (gram_sensor[si]) {
        // per-(CTA, warp) partials of the sweep (SensorDesc::gcta): [calib 0..7 x 0..7 | 8..15 x 0..7 | 8..15 x 8..15 | gradient 0..15]
        const long base = long(gcta_off[si]) * kGramCta;
        for (int li = 0; li < d.n_calib; ++li) for (int lj = 0; lj <= li; ++lj) {
          const int off = li < 8 ? li * 8 + lj : (lj < 8 ? 64 + (li - 8) * 8 + lj : 128 + (li - 8) * 8 + (lj - 8));
          ce.push_back(CalibEntry{base + off, kGramCta, gcta_cnt[si], 2, d.calib_off + li, d.calib_off + lj});
        }
        for (int li = 0; li < d.n_calib; ++li) ce.push_back(CalibEntry{base + 192 + li, kGramCta, gcta_cnt[si], 2, d.calib_off + li, -1});
      } else {
        for (int li = 0; li < d.n_calib; ++li) for (int lj = 0; lj <= li; ++lj) ce.push_back(CalibEntry{long(c2off[si] + li * d.n_calib + lj), csz, nsl_ce, 0, d.calib_off + li, d.calib_off + lj});
        for (int li = 0; li < d.n_calib; ++li) ce.push_back(CalibEntry{long(d.calib_off + li), N_c, nsl_ce, 1, d.calib_off + li, -1});
      }
    }
    d_centries.upload(ce, h2d);
    d_cpartial.alloc(std::max<size_t>(ce.size(), 1) * kCalibSlices);
    const size_t nsl = size_t(std::max(g_hi - g_lo, 1));
    d_segA.alloc(nsl * 36 * 36); d_segG.alloc(nsl * 36);
    const bool two_sources = n_gram_sensors > 0 && n_plain_sensors > 0;
    d_segA2.alloc(two_sources ? nsl * 36 * 36 : 0); d_segG2.alloc(two_sources ? nsl * 36 : 0);
    d_segB.alloc(nsl * 36 * std::max(N_c, 1)); d_segC.alloc(nsl * std::max(csz, 1)); d_segGc.alloc(nsl * std::max(N_c, 1));
    d_Aband.alloc(size_t(n_a) * 36); d_Bmat.alloc(size_t(n_a) * std::max(N_c, 1)); d_Cmat.alloc(size_t(std::max(N_c, 1)) * std::max(N_c, 1));
    d_grad.alloc(n_tot); d_diag.alloc(n_tot); d_scaling.alloc(n_tot); d_dtil2.alloc(n_tot); d_ytil.alloc(n_tot);
    d_Aband.zero(stream); d_Bmat.zero(stream); d_grad.zero(stream); d_ytil.zero(stream);
    if (world > 1) d_gradG.alloc(n_tot);
    rc = plan_schur();
    if (rc != CB2_OK) return rc;
    set_kernel_attributes();
    CB2_CUDA(cudaDeviceSynchronize());
    if (timing_log) std::fprintf(stderr, "[cb2 timing] upload: pack %.2f ms, layout + alloc + H2D + plan %.2f ms\n", 1e3 * (t_packed - t_upload0), 1e3 * (now_s() - t_packed));
    uploaded = true;
    scaling_set = false;
    return CB2_OK;
  }

  // ------------------------------------------------------------------------------------------------------------
  // Chunk plan for the substructured Schur elimination (cb2_schur.cuh).
  // ------------------------------------------------------------------------------------------------------------
  // Global chunk plan (identical on every rank) and this rank's share of it: chunks [chunk_lo, chunk_hi), segments [g_lo, g_hi).
  int plan_chunks() {
    int target = 110;   // interior control points per chunk of the band-factor path (tuned on C4, profiles/)
    use_cr = true;
    if (const char* e = std::getenv("CB2_SCHUR")) use_cr = std::string(e) != "band";
    if (use_cr) target = 1 << 28;   // cyclic reduction parallelises inside a chunk: one chunk per rank
    if (const char* e = std::getenv("CB2_CHUNK_CPS")) target = std::max(6, std::atoi(e));
    int P = std::max(1, (n_cp + 5) / (target + 5));
    P = (P + world - 1) / world * world;                       // same number of chunks on every rank
    while (P > world && (n_cp - 5 * (P - 1)) / P < 6) P -= world;
    if ((n_cp - 5 * (P - 1)) / P < 6 && P > 1) return fail(CB2_INVALID_ARGUMENT, "Trajectory too short to shard across this many GPUs.");
    const int interior = n_cp - 5 * (P - 1);
    chunks.clear();
    int a = 0;
    for (int p = 0; p < P; ++p) {
      const int len = interior / P + (p < interior % P ? 1 : 0);
      chunks.push_back(ChunkPlan{a, a + len});
      a += len + 5;
    }
    const int per = P / world;
    chunk_lo = rank * per; chunk_hi = chunk_lo + per;
    // Segment g touches control points g..g+5. The separator after chunk c starts at control point chunks[c].b: segments
    // below it belong to the left rank, segments from it on to the right rank.
    g_lo = chunk_lo == 0 ? 0 : chunks[chunk_lo - 1].b;
    g_hi = chunk_hi == P ? n_seg : chunks[chunk_hi - 1].b;
    return CB2_OK;
  }

  // Device storage of the substructured Schur elimination (cb2_schur.cuh) for the owned chunks + the replicated separator level.
  int plan_schur() {
    const int P = int(chunks.size()), PL = chunk_hi - chunk_lo;
    const int cal0 = P > 1 ? 2 * kSepDim : 0;   // a single chunk has no separator columns in its border
    const int nbw1 = cal0 + N_c + 1, nbw2 = N_c + 1;
    const int n2 = kSepDim * (P - 1);
    std::vector<int> rowidx, colidx;
    std::vector<size_t> row_off(PL + 1), col_off(PL + 1);
    size_t Lsz = 0, Wsz = 0, Tsz = 0;
    h_l1.assign(PL, BandSys{});
    const int nt1 = (nbw1 + 63) / 64;
    max_tilepairs1 = nt1 * (nt1 + 1) / 2;
    gram_dmma1 = nbw1 <= kGramMaxNbw && !std::getenv("CB2_GRAM_SIMT");
    max_ksplit1 = 1;
    std::vector<size_t> Loff(PL), Woff(PL), Toff(PL), Doff(PL);
    size_t Dsz = 0;
    for (int l = 0; l < PL; ++l) {
      const int p = chunk_lo + l;
      BandSys& sy = h_l1[l];
      sy.n = 6 * (chunks[p].b - chunks[p].a); sy.hb = 35; sy.nbw = nbw1; sy.cal0 = cal0;
      sy.nblk = (sy.n + kCrB - 1) / kCrB;
      sy.ksplit = std::max(1, std::min((sy.n + 63) / 64, (296 + PL * max_tilepairs1 - 1) / (PL * max_tilepairs1)));
      if (gram_dmma1) sy.ksplit = std::max(1, std::min((sy.n + 63) / 64, std::max(1, 148 / PL)));
      gram_parts.clear();
      if (gram_dmma1 && use_cr && PL == 1 && world == 1 && !std::getenv("CB2_NO_EARLY_GRAM")) {
        // The rows of the blocks the first La levels eliminate (15/16 of the chunk for La = 4) are multiplied beside the LATER levels, from
        // level Ls on: those levels are narrow (<= 16 blocks) and latency-bound, the product is bound by the FP64 tensor pipe and a CTA of
        // it fills an SM's register file, so it takes the SMs the levels leave idle (all but `reserve` of them). What is left for the
        // critical path after the last level is the 1/16 of the rows the late levels produce, one row tile per CTA.
        // (Measured on C4, profiles/r02_variants.md: Schur phase 0.442 -> 0.400 ms; starting earlier or reserving fewer SMs slows the levels.)
        static const int la_env = env_int("CB2_EARLY_GRAM_LEVELS", -1), ls_env = env_int("CB2_EARLY_GRAM_AFTER", -1), reserve = env_int("CB2_EARLY_GRAM_RESERVE", 64);
        const int nlev = cr_levels(sy.nblk);
        int ls_auto = 0;
        while (ls_auto < nlev - 2 && ((sy.nblk + (2 << ls_auto) - 1) >> (ls_auto + 1)) > 16) ++ls_auto;   // first level followed by <= 16 active blocks
        const int La = std::min(la_env >= 0 ? la_env : std::max(1, ls_auto), nlev - 1);
        int koff = 0;
        if (La >= 1) {
          const int Ls = std::max(La - 1, std::min(ls_env >= 0 ? ls_env : ls_auto, nlev - 2));
          const int M = 1 << La, nmult = (sy.nblk + M - 1) / M;
          const int rows_early = (sy.nblk - nmult) * kCrB, rows_last = nmult * kCrB;
          const int k_early = std::max(1, std::min(std::max(1, 148 - reserve), (rows_early + kGramRows - 1) / kGramRows));
          const int k_last = std::max(1, std::min(148, (rows_last + kGramRows - 1) / kGramRows));
          gram_parts.push_back(GramPart{-1, M, 0, k_early, Ls});
          gram_parts.push_back(GramPart{0, M, k_early, k_last, -1});
          koff = k_early + k_last;
        }
        if (!gram_parts.empty()) { sy.ksplit = koff; sy.kfirst = gram_prereduce ? gram_parts[0].k_cnt - 1 : 0; }
      }
      max_ksplit1 = std::max(max_ksplit1, sy.ksplit);
      row_off[l] = rowidx.size();
      for (int i = 0; i < sy.n; ++i) rowidx.push_back(6 * chunks[p].a + i);
      col_off[l] = colidx.size();
      if (cal0 > 0) {
        for (int j = 0; j < kSepDim; ++j) colidx.push_back(p > 0 ? 6 * (chunks[p].a - 5) + j : -1);
        for (int j = 0; j < kSepDim; ++j) colidx.push_back(p < P - 1 ? 6 * chunks[p].b + j : -1);
      }
      for (int c = 0; c < N_c; ++c) colidx.push_back(int(n_a) + c);
      Loff[l] = Lsz; Woff[l] = Wsz; Toff[l] = Tsz; Doff[l] = Dsz; Dsz += size_t(sy.n);
      Lsz += size_t(sy.n) * 36; Wsz += size_t(sy.n) * nbw1; Tsz += size_t(sy.ksplit) * nbw1 * nbw1;
      if (!use_cr && backsolve_smem_bytes(sy.n, nbw1, 36) > 220 * 1024) return fail(CB2_INTERNAL, "Schur chunk too large for shared memory; lower CB2_CHUNK_CPS.");
      if (nbw1 > 512) return fail(CB2_UNIMPLEMENTED, "Too many calibration unknowns for the band factor kernel.");   // PFW = 12 registers of border prefetch
    }
    const size_t row_off2 = rowidx.size();
    std::vector<int> shared_idx;
    for (int p = 0; p + 1 < P; ++p) for (int j = 0; j < kSepDim; ++j) { rowidx.push_back(6 * chunks[p].b + j); shared_idx.push_back(6 * chunks[p].b + j); }
    for (int c = 0; c < N_c; ++c) shared_idx.push_back(int(n_a) + c);
    n_shared = int(shared_idx.size());
    const size_t col_off2 = colidx.size();
    for (int c = 0; c < N_c; ++c) colidx.push_back(int(n_a) + c);
    d_rowidx.upload(rowidx); d_colidx.upload(colidx);
    d_shared_idx.upload(shared_idx);
    d_shared_buf.alloc(shared_buf_size(std::max(n_shared, 1), world));
    d_L1.alloc(use_cr ? 1 : Lsz); d_W1.alloc(Wsz); d_T1.alloc(Tsz); d_Dinv.alloc(Dsz + size_t(std::max(n2, 1)));
    for (int l = 0; l < PL; ++l) {
      BandSys& sy = h_l1[l];
      sy.row_gidx = d_rowidx.p + row_off[l]; sy.col_gidx = d_colidx.p + col_off[l];
      sy.L = use_cr ? nullptr : d_L1.p + Loff[l]; sy.W = d_W1.p + Woff[l]; sy.T = d_T1.p + Toff[l]; sy.Dinv = d_Dinv.p + Doff[l];
    }
    if (use_cr) {
      // Cyclic-reduction storage of the owned chunks.
      size_t blk_tot = 0, uslot_tot = 0;
      cr_max_nblk = 0;
      for (const auto& sy : h_l1) { blk_tot += size_t(sy.nblk); uslot_tot += size_t((sy.nblk + 1) / 2); cr_max_nblk = std::max(cr_max_nblk, sy.nblk); }
      cr_nlevels = cr_levels(std::max(cr_max_nblk, 1));
      const size_t usz = cr_u_size(nbw1);
      d_crD.alloc(blk_tot * kCrB * kCrB); d_crL.alloc(blk_tot * kCrB * kCrB); d_crBd.alloc(blk_tot * kCrB * nbw1);
      d_crWef.alloc(blk_tot * kCrB * 2 * kCrB); d_crU.alloc(2 * uslot_tot * usz + usz);
      CB2_CUDA(cudaMemsetAsync(d_crU.p + 2 * uslot_tot * usz, 0, usz * sizeof(double), stream));
      size_t b0 = 0, u0 = 0;
      for (auto& sy : h_l1) {
        sy.cr_uslots = (sy.nblk + 1) / 2;
        sy.crD = d_crD.p + b0 * kCrB * kCrB; sy.crL = d_crL.p + b0 * kCrB * kCrB; sy.crBd = d_crBd.p + b0 * kCrB * nbw1;
        sy.crWef = d_crWef.p + b0 * kCrB * 2 * kCrB; sy.crU = d_crU.p + 2 * u0 * usz; sy.crZero = d_crU.p + 2 * uslot_tot * usz;
        b0 += size_t(sy.nblk); u0 += size_t(sy.cr_uslots);
      }
      if (cr_smem_bytes(nbw1) > 227 * 1024) return fail(CB2_UNIMPLEMENTED, "Too many calibration unknowns for the shared-memory Schur kernels.");
    }
    d_l1.upload(h_l1);
    h_l2 = BandSys{};
    h_l2.n = n2; h_l2.hb = 59; h_l2.nbw = nbw2;
    const int nt2 = (nbw2 + 63) / 64;
    h_l2.ksplit = std::max(1, std::min((n2 + 63) / 64, (148 + nt2 * (nt2 + 1) / 2 - 1) / (nt2 * (nt2 + 1) / 2)));
    // The separator system and the calibration system live in ONE buffer: it is what the ranks sum (one allreduce per LM solve).
    const size_t szL2 = size_t(n2) * 60, szW2 = size_t(n2) * nbw2, szCw = size_t(N_c + 1) * (N_c + 1);
    d_red.alloc(szL2 + szW2 + szCw + shared_buf_size(std::max(n_shared, 1), world) + 1);   // tail: the deferred shared-row sums (see launch_step)
    d_T2.alloc(size_t(h_l2.ksplit) * nbw2 * nbw2);
    h_l2.row_gidx = d_rowidx.p + row_off2; h_l2.col_gidx = d_colidx.p + col_off2;
    h_l2.L = d_red.p; h_l2.W = d_red.p + szL2; h_l2.T = d_T2.p; h_l2.Dinv = d_Dinv.p + Dsz;
    red_Cw = d_red.p + szL2 + szW2;
    red_count = szL2 + szW2 + szCw;
    sep_cr = use_cr && n2 > 0 && std::getenv("CB2_SEP_BAND") == nullptr && cr_smem_bytes(nbw2) <= 227 * 1024;
    gram_dmma2 = sep_cr && nbw2 <= kGramMaxNbw && !std::getenv("CB2_GRAM_SIMT");
    if (sep_cr) {
      // separators = a block-tridiagonal chain of P - 1 blocks of 30 with the border [calibration | rhs]: cyclic reduction again
      h_l2.nblk = n2 / kCrB; h_l2.cr_uslots = (h_l2.nblk + 1) / 2; h_l2.cal0 = 0;
      cr2_nlevels = cr_levels(h_l2.nblk);
      const size_t usz2 = cr_u_size(nbw2);
      d_cr2D.alloc(size_t(h_l2.nblk) * kCrB * kCrB); d_cr2L.alloc(size_t(h_l2.nblk) * kCrB * kCrB); d_cr2Bd.alloc(size_t(h_l2.nblk) * kCrB * nbw2);
      d_cr2Wef.alloc(size_t(h_l2.nblk) * kCrB * 2 * kCrB); d_cr2U.alloc(2 * size_t(h_l2.cr_uslots) * usz2 + usz2);
      h_l2.crD = d_cr2D.p; h_l2.crL = d_cr2L.p; h_l2.crBd = d_cr2Bd.p; h_l2.crWef = d_cr2Wef.p; h_l2.crU = d_cr2U.p;
      h_l2.crZero = d_cr2U.p + 2 * size_t(h_l2.cr_uslots) * usz2;
      if (gram_dmma2) { h_l2.ksplit = std::max(1, std::min((n2 + kGramRows - 1) / kGramRows, 148)); d_T2.alloc(size_t(h_l2.ksplit) * nbw2 * nbw2); h_l2.T = d_T2.p; }
    }
    d_l2.upload(std::vector<BandSys>(1, h_l2));
    std::vector<int> chunk_sys(P, -1);
    for (int l = 0; l < PL; ++l) chunk_sys[chunk_lo + l] = l;
    d_chunk_sys.upload(chunk_sys);
    d_rawdiag.alloc(size_t(std::max(n2, 1)) + std::max(N_c, 1));
    const size_t smem_f1 = factor_smem_bytes(36, nbw1), smem_f2 = factor_smem_bytes(60, nbw2);
    if ((!use_cr && smem_f1 > 227 * 1024) || (h_l2.n > 0 && !sep_cr && smem_f2 > 227 * 1024)) return fail(CB2_UNIMPLEMENTED, "Too many calibration unknowns for the shared-memory Schur kernels.");
    return CB2_OK;
  }
  double* red_Cw = nullptr;
  size_t red_count = 0;

  void set_kernel_attributes() {
#ifndef CB2_EMUL
    // Opt in to > 48 KB of dynamic shared memory, per kernel, with the size actually used (static + dynamic <= 227 KB).
    auto set = [&](auto kernel, size_t bytes) {
      if (bytes > 48 * 1024) CB2_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
    };
    const size_t ev_max[3] = {size_t(rec_size(kCamera, 11)) * eval_rec_stride(kCamera) * 8, size_t(rec_size(kGyroscope, 12)) * eval_rec_stride(kGyroscope) * 8,
                              size_t(rec_size(kAccelerometer, 12)) * eval_rec_stride(kAccelerometer) * 8};
    set(eval_kernel<kCamera, kModeCost>, ev_max[0]); set(eval_kernel<kCamera, kModeResiduals>, ev_max[0]); set(eval_kernel<kCamera, kModeJacobian>, ev_max[0]);
    set(eval_kernel<kGyroscope, kModeCost>, ev_max[1]); set(eval_kernel<kGyroscope, kModeResiduals>, ev_max[1]); set(eval_kernel<kGyroscope, kModeJacobian>, ev_max[1]);
    set(eval_kernel<kAccelerometer, kModeCost>, ev_max[2]); set(eval_kernel<kAccelerometer, kModeResiduals>, ev_max[2]); set(eval_kernel<kAccelerometer, kModeJacobian>, ev_max[2]);
    set((accumulate_kernel<6, 37>), acc_smem_bytes()); set((accumulate_kernel<7, kAccCal0>), acc_smem_bytes()); set((accumulate_kernel<8, kAccCal0>), acc_smem_bytes());
    set(expand_gram_kernel, expand_smem_bytes());
    int max_n1 = 0;
    for (const auto& sy : h_l1) max_n1 = std::max(max_n1, sy.n);
    const int nbw1 = h_l1[0].nbw, nbw2 = h_l2.nbw;
    if (gram_dmma1) set(border_gram_dmma_kernel, gram_smem_bytes(nbw1));
    cr_zsplit = std::getenv("CB2_CR_ZSPLIT") ? std::max(1, std::atoi(std::getenv("CB2_CR_ZSPLIT"))) : std::max(1, std::min(4, (2 * kCrB + nbw1 + 7) / 8 / 6));
    cr_tma = use_cr && std::getenv("CB2_NO_TMA") == nullptr && cr_smem_bytes_tma(nbw1) <= 227 * 1024;   // staging area beside the working set
    if (use_cr) { set(cr_level_kernel<true>, cr_smem_bytes(nbw1)); set(cr_level_kernel<false>, cr_tma ? cr_smem_bytes_tma(nbw1) : cr_smem_bytes(nbw1)); }
    else set(band_factor_kernel<6>, factor_smem_bytes(36, nbw1));
    if (h_l2.n > 0 && !sep_cr) set(band_factor_kernel<10>, factor_smem_bytes(60, nbw2));
    if (gram_dmma2 && !gram_dmma1) set(border_gram_dmma_kernel, gram_smem_bytes(nbw2));
    set(reduced_solve_kernel, size_t(N_c + 1) * (kRedPanel + 1) * 8);
    if (reduced_smem_bytes(N_c) <= 227 * 1024 - 256) set(reduced_solve_smem_kernel, reduced_smem_bytes(N_c));
    set(band_backsolve_kernel, std::max(use_cr ? size_t(0) : backsolve_smem_bytes(max_n1, nbw1, 36), backsolve_smem_bytes(h_l2.n, nbw2, 60)));
#endif
  }

  // ------------------------------------------------------------------------------------------------------------
  // Kernel launches.
  // ------------------------------------------------------------------------------------------------------------
#ifdef CB2_EMUL
#define CB2_K(...) do { CB2_LAUNCH(__VA_ARGS__); ++stats.kernel_launches; } while (0)
#else
#define CB2_K(kernel, grid, block, smem, strm, ...) do { \
    if (kprof.on) { KernelProfiler::Rec pr_{#kernel, nullptr, nullptr}; cudaEventCreate(&pr_.a); cudaEventCreate(&pr_.b); cudaEventRecord(pr_.a, strm); \
                    CB2_LAUNCH(kernel, grid, block, smem, strm, __VA_ARGS__); cudaEventRecord(pr_.b, strm); kprof.open.push_back(pr_); } \
    else CB2_LAUNCH(kernel, grid, block, smem, strm, __VA_ARGS__); \
    ++stats.kernel_launches; } while (0)
#endif

  // K0: per-image records of every camera at parameter buffer `which`.
  void launch_frames(int which) {
    if (n_frames > 0)
      CB2_K(camera_frame_kernel, (n_frames + 31) / 32, 32, 0, stream, n_frames, d_frame_sensor.p, d_frame_seg.p, d_frame_stamp.p, d_state[which].p,
            d_ctrl[which].p, d_knots.p, d_basis.p, d_frames.p);
  }

  // Residual sweep over every sensor at parameter buffer `which`; scalars land in d_scal[slot], d_scal[slot + 1].
  template <int MODE>
  void launch_imu(cudaStream_t si, const SensorDesc* desc, const SensorState* st, const double* c, const int* nt, const double* pwp) {
    // Both IMU kernels are FP64-latency-bound single-warp CTAs at low occupancy: the gyroscope kernel runs beside the accelerometer kernel
    // on the second stream (they touch disjoint buffers) instead of after it. (Not inside graph capture / the emulation build.)
    cudaStream_t sg = si;
#ifndef CB2_EMUL
    const bool pair = imu_pair_streams && nt[1] && nt[2] && si == stream && !capturing;
    if (pair) { CB2_CUDA(cudaEventRecord(ev_fork, stream)); CB2_CUDA(cudaStreamWaitEvent(stream_imu, ev_fork, 0)); sg = stream_imu; }
#endif
    if (nt[2]) CB2_K((eval_kernel<kAccelerometer, MODE>), nt[2], eval_tile(kAccelerometer), smem_eval[2], si, desc, st, d_tiles.p + tile_off[2], c, d_knots.p, d_basis.p, pwp,
                     d_frames.p, gravity[0], gravity[1], gravity[2], d_cost_partial.p + tile_off[2], d_invalid_partial.p + tile_off[2], 1);
    if (nt[1]) CB2_K((eval_kernel<kGyroscope, MODE>), nt[1], eval_tile(kGyroscope), smem_eval[1], sg, desc, st, d_tiles.p + tile_off[1], c, d_knots.p, d_basis.p, pwp,
                     d_frames.p, gravity[0], gravity[1], gravity[2], d_cost_partial.p + tile_off[1], d_invalid_partial.p + tile_off[1], 1);
#ifndef CB2_EMUL
    if (pair) { CB2_CUDA(cudaEventRecord(ev_join, stream_imu)); CB2_CUDA(cudaStreamWaitEvent(stream, ev_join, 0)); }
#endif
  }
  // imu: kImuSame = IMU sensors in MODE too; kImuJacobian = IMU sensors in Jacobian mode whatever MODE is (speculative Jacobians of
  // the trial point, see launch_step); kImuSkip = their Jacobians, residuals and cost partials of this point already exist.
  enum { kImuSame = 0, kImuJacobian = 1, kImuSkip = 2 };
  template <int MODE>
  void launch_eval(int which, int slot, int imu = kImuSame) {
    const SensorDesc* desc = d_desc.p;
    const SensorState* st = d_state[which].p;
    const double* c = d_ctrl[which].p;
    const int nt[3] = {tile_off[1] - tile_off[0], tile_off[2] - tile_off[1], tile_off[3] - tile_off[2]};
    // One stream by default: the IMU kernels are FP64-latency-bound and, run beside the camera sweep, only fence SMs off from it
    // (measured: no net overlap), while a serial order lets the camera kernel — the dominant, HBM-bound one — be timed alone.
    // CB2_IMU_STREAM=1 moves them to a second (high-priority) stream.
    const bool fork = imu_side_stream && (nt[1] || nt[2]) && nt[0];
    cudaStream_t si = fork ? stream_imu : stream;
    if (fork) { CB2_CUDA(cudaEventRecord(ev_fork, stream)); CB2_CUDA(cudaStreamWaitEvent(stream_imu, ev_fork, 0)); }
    if (nt[0]) {
      launch_frames(which);
      if (MODE == kModeJacobian) timer.begin(kPhCamera, stream);
      CB2_K((eval_kernel<kCamera, MODE>), nt[0], eval_tile(kCamera), smem_eval[0], stream, desc, st, d_tiles.p + tile_off[0], c, d_knots.p, d_basis.p, d_pw[which].p,
            d_frames.p, gravity[0], gravity[1], gravity[2], d_cost_partial.p + tile_off[0], d_invalid_partial.p + tile_off[0], 1);
      if (MODE == kModeJacobian) timer.end(kPhCamera, stream);
    }
    if (imu != kImuSkip) {
      if (imu == kImuJacobian || MODE == kModeJacobian) launch_imu<kModeJacobian>(si, desc, st, c, nt, d_pw[which].p);
      else launch_imu<MODE>(si, desc, st, c, nt, d_pw[which].p);
    }
    if (fork) { CB2_CUDA(cudaEventRecord(ev_join, stream_imu)); CB2_CUDA(cudaStreamWaitEvent(stream, ev_join, 0)); }
    CB2_K(reduce_cost_kernel, 1, kRcThreads, 0, stream, d_cost_partial.p, d_invalid_partial.p, n_tiles, d_scal.p, slot);
  }

  // K1-K3 at x, then K4: normal equations, Hessian diagonal, gradient norms.
  void launch_jacobian_and_normal_equations(bool defer_shared_sums = false) {
    const bool defer_shared = defer_shared_sums && world > 1;
    ne_shared_pending = defer_shared;
    sweep_skipped = jac_point == cur;          // the trial pass that led to this point's acceptance was a full sweep (launch_trial)
    if (!sweep_skipped) {
      timer.begin(kPhJacobian, stream);
      // IMU Jacobians of this point may already have been produced by the trial-cost pass (speculative_imu).
      const bool imu_reuse = imu_jac_point == cur;
      launch_eval<kModeJacobian>(cur, kScCost, imu_reuse ? kImuSkip : kImuSame);
      imu_jac_point = cur;
      jac_point = cur;
      last_sweep_imu = !imu_reuse;
      timer.end(kPhJacobian, stream);
      ++stats.jacobian_sweeps;
      stats.jacobian_blocks += last_sweep_imu ? num_active_blocks() : num_active_blocks() - num_active_blocks(true);
      stats.jacobian_bytes += last_sweep_imu ? jacobian_bytes_per_sweep() : jacobian_bytes_per_sweep(kCamera);
      stats.camera_kernel_bytes += jacobian_bytes_per_sweep(kCamera);
      stats.camera_kernel_gram_bytes += gram_bytes_per_sweep();
    }
    timer.begin(kPhNormal, stream);
    graphed(g_normal[cur + (defer_shared ? 2 : 0)], [&] {
    const int ns = int(sensors.size());
    const int nsl = g_hi - g_lo;
    if (nsl > 0) {
      // IMU sensors (and cameras too wide for the compact slots) from their Jacobian rows; cameras from the Gram slots the sweep left.
      // Both kernels are latency-bound at low occupancy and independent (separate control-point partials, disjoint calibration columns):
      // they run side by side, accumulate_kernel on the second stream (a fork / join that a graph capture records as such).
      const bool plain = n_plain_sensors > 0 || n_gram_sensors == 0, gram = n_gram_sensors > 0;
      cudaStream_t s_acc = stream;
#ifndef CB2_EMUL
      if (plain && gram) { CB2_CUDA(cudaEventRecord(ev_fork, stream)); CB2_CUDA(cudaStreamWaitEvent(stream_imu, ev_fork, 0)); s_acc = stream_imu; }
#endif
      if (plain) {
        int max_nc = 0;
        for (const auto& d : h_desc) if (!d.gslots) max_nc = std::max(max_nc, d.n_calib);
        const int nacc = (nsl + acc_spc - 1) / acc_spc;
        if (37 + max_nc <= 48) CB2_K((accumulate_kernel<6, 37>), nacc, kAccThreads, acc_smem_bytes(), s_acc, d_desc.p, d_plain_idx.p, n_plain_idx, nsl, acc_spc, N_c, g_lo, d_c2off.p, csz, d_segA.p, d_segG.p, d_segB.p, d_segC.p, d_segGc.p);
        else if (kAccCal0 + max_nc <= 56) CB2_K((accumulate_kernel<7, kAccCal0>), nacc, kAccThreads, acc_smem_bytes(), s_acc, d_desc.p, d_plain_idx.p, n_plain_idx, nsl, acc_spc, N_c, g_lo, d_c2off.p, csz, d_segA.p, d_segG.p, d_segB.p, d_segC.p, d_segGc.p);
        else CB2_K((accumulate_kernel<8, kAccCal0>), nacc, kAccThreads, acc_smem_bytes(), s_acc, d_desc.p, d_plain_idx.p, n_plain_idx, nsl, acc_spc, N_c, g_lo, d_c2off.p, csz, d_segA.p, d_segG.p, d_segB.p, d_segC.p, d_segGc.p);
      }
      if (gram)
        CB2_K(expand_gram_kernel, nsl, kExpThreads, expand_smem_bytes(), stream, d_gslot_tab.p, d_gram_meta.p, n_gram_sensors, d_gslots.p, N_c, d_ext_tab.p, d_ext_dst.p,
              plain ? d_segA2.p : d_segA.p, plain ? d_segG2.p : d_segG.p, d_segB.p);
#ifndef CB2_EMUL
      if (plain && gram) { CB2_CUDA(cudaEventRecord(ev_join, stream_imu)); CB2_CUDA(cudaStreamWaitEvent(stream, ev_join, 0)); }
#endif
    }
    // The calibration blocks (from segC / segGc / the sweep's per-warp partials) and the control-point band + border (from segA / segB /
    // segG) are assembled from disjoint inputs into disjoint outputs (grad: calibration tail / control-point head): side by side.
    cudaStream_t s_cal = stream;
#ifndef CB2_EMUL
    if (calib_fork && N_c > 0) { CB2_CUDA(cudaEventRecord(ev_fork, stream)); CB2_CUDA(cudaStreamWaitEvent(stream_imu, ev_fork, 0)); s_cal = stream_imu; }
#endif
    const long total = n_a * 36 + n_a * N_c + n_a;
    CB2_K(assemble_band_kernel, int(std::min<long>((total + 255) / 256, 148 * 16)), 256, 0, stream, n_cp, g_lo, g_hi, N_c, d_segA.p, d_segG.p, d_segB.p,
          d_segA2.n ? d_segA2.p : nullptr, d_segG2.n ? d_segG2.p : nullptr, d_Aband.p, d_Bmat.p, d_grad.p);
    if (N_c > 0) {
      d_Cmat.zero(s_cal);
      const int ne = int(d_centries.n);
      CB2_K(assemble_calib_kernel, dim3((ne + 31) / 32, kCalibSlices), dim3(32, 8), 0, s_cal, ne, d_centries.p, d_segC.p, d_segGc.p, d_gcta.p, d_cpartial.p);
      CB2_K(assemble_calib_final_kernel, (ne + 255) / 256, 256, 0, s_cal, N_c, ne, kCalibSlices, d_centries.p, d_cpartial.p, d_Cmat.p, d_grad.p + n_a);
    }
#ifndef CB2_EMUL
    if (s_cal != stream) { CB2_CUDA(cudaEventRecord(ev_join, s_cal)); CB2_CUDA(cudaStreamWaitEvent(stream, ev_join, 0)); }
#endif
    if (world_free && tile_off[1] > tile_off[0]) {   // freed world-model blocks: their rows / columns of the normal equations, from the cameras' Jacobian rows
      if (N_c > n_sensor_unknowns) CB2_CUDA(cudaMemsetAsync(d_grad.p + n_a + n_sensor_unknowns, 0, sizeof(double) * (N_c - n_sensor_unknowns), stream));   // accumulated by atomics
      CB2_K(world_normal_kernel, tile_off[1] - tile_off[0], 128, 0, stream, d_desc.p, d_tiles.p + tile_off[0], n_a, N_c, d_pt_body.p, d_body_u.p, d_pt_u.p,
            d_body_q[cur].p, d_body_t[cur].p, d_pw[cur].p, d_Bmat.p, d_Cmat.p, d_grad.p);
    }
    // Single rank: the Hessian diagonal (for the damping of the coming solve) beside the gradient norms — two small independent kernels.
    cudaStream_t s_diag = stream;
#ifndef CB2_EMUL
    if (calib_fork && world == 1) { CB2_CUDA(cudaEventRecord(ev_fork, stream)); CB2_CUDA(cudaStreamWaitEvent(stream_imu, ev_fork, 0)); s_diag = stream_imu; }
#endif
    CB2_K(hess_diag_kernel, int(std::min<long>((n_tot + 255) / 256, 1024)), 256, 0, s_diag, n_a, N_c, d_Aband.p, d_Cmat.p, d_diag.p);
    if (world > 1) {
      // Separator rows and calibration receive contributions from several ranks (lmkernels: pack_shared_kernel). d_grad itself stays
      // rank-local; gradG carries the sums on the shared rows. When the host does not wait for this point's gradient norms (deferred
      // round trip, see minimize) the sums are NOT exchanged here: they ride in the tail of the next solve's collective (launch_step).
      CB2_CUDA(cudaMemcpyAsync(d_gradG.p, d_grad.p, sizeof(double) * n_tot, cudaMemcpyDeviceToDevice, stream));
      CB2_K(gradient_norm_kernel, lm_ctas(n_a), kLmThreads, 0, stream, n_a, d_grad.p, d_cp_own.p, 1, 0, 0, d_desc.p, d_state[cur].p, ns, world_refs(), d_body_q[cur].p, d_scal.p, d_lm_partial.p, d_lm_ticket.p);   // owned part
      if (!defer_shared) {
        CB2_K(pack_shared_kernel, (n_shared + 255) / 256, 256, 0, stream, n_shared, d_shared_idx.p, d_grad.p, d_diag.p, d_scal.p, world, rank, d_shared_buf.p);
        comm->allreduce_sum(d_shared_buf.p, shared_buf_size(n_shared, world), stream);
        CB2_K(unpack_shared_kernel, (n_shared + 255) / 256, 256, 0, stream, n_shared, d_shared_idx.p, d_shared_buf.p, world, d_gradG.p, d_diag.p, d_scal.p);
        CB2_K(gradient_norm_kernel, lm_ctas(n_a), kLmThreads, 0, stream, n_a, d_gradG.p, d_cp_own.p, 0, 1, 1, d_desc.p, d_state[cur].p, ns, world_refs(), d_body_q[cur].p, d_scal.p, d_lm_partial.p, d_lm_ticket.p);   // + shared part
      }
    } else {
      CB2_K(gradient_norm_kernel, lm_ctas(n_a), kLmThreads, 0, stream, n_a, d_grad.p, d_cp_own.p, 1, 1, 0, d_desc.p, d_state[cur].p, ns, world_refs(), d_body_q[cur].p, d_scal.p, d_lm_partial.p, d_lm_ticket.p);
    }
#ifndef CB2_EMUL
    if (s_diag != stream) { CB2_CUDA(cudaEventRecord(ev_join, s_diag)); CB2_CUDA(cudaStreamWaitEvent(stream, ev_join, 0)); }
#endif
    });
    timer.end(kPhNormal, stream);
  }

  const double* gradG() const { return world > 1 ? d_gradG.p : d_grad.p; }   // gradient with cross-rank sums on the shared rows

  // The LM loop ends before another solve would have carried the deferred shared-row sums: exchange them now (every rank takes this
  // branch together: it depends on the iteration count and the radius only).
  void finish_shared_sums() {
    if (!ne_shared_pending || world <= 1) return;
    ne_shared_pending = false;
    const int ns = int(sensors.size());
    CB2_K(pack_shared_kernel, (n_shared + 255) / 256, 256, 0, stream, n_shared, d_shared_idx.p, d_grad.p, d_diag.p, d_scal.p, world, rank, d_shared_buf.p);
    comm->allreduce_sum(d_shared_buf.p, shared_buf_size(n_shared, world), stream);
    CB2_K(unpack_shared_kernel, (n_shared + 255) / 256, 256, 0, stream, n_shared, d_shared_idx.p, d_shared_buf.p, world, d_gradG.p, d_diag.p, d_scal.p);
    CB2_K(gradient_norm_kernel, lm_ctas(n_a), kLmThreads, 0, stream, n_a, d_gradG.p, d_cp_own.p, 0, 1, 1, d_desc.p, d_state[cur].p, ns, world_refs(), d_body_q[cur].p, d_scal.p, d_lm_partial.p, d_lm_ticket.p);
  }

  // One LM linear solve + candidate point + candidate cost. Everything is enqueued; the caller syncs once.
  void launch_step(double radius, const cb2_options& opt) {
    const int P = int(chunks.size()), PL = chunk_hi - chunk_lo;
    const int ns = int(sensors.size());
    const int blocks = int(std::min<long>((n_tot + 255) / 256, 1024));
    timer.begin(kPhSchur, stream);
    h_param[0] = radius; h_param[1] = opt.min_lm_diagonal; h_param[2] = opt.max_lm_diagonal;
    CB2_CUDA(cudaMemcpyAsync(d_scal.p + kScRadius, h_param, 3 * sizeof(double), cudaMemcpyHostToDevice, stream));
    stats.h2d_bytes += 3 * sizeof(double);
    // The solve phase is ~25 small latency-bound kernels: replayed as ONE CUDA graph per parameter buffer.
    const bool tail = ne_shared_pending && world > 1;
    ne_shared_pending = false;
    graphed(g_solve[cur + (tail ? 2 : 0)], [&] {
    CB2_K(damping_kernel, blocks, 256, 0, stream, n_tot, d_diag.p, d_scaling.p, d_scal.p, d_dtil2.p);
    CB2_CUDA(cudaMemsetAsync(d_scal.p + kScSolveFail, 0, sizeof(double), stream));
    const int nbw1 = h_l1[0].nbw;
    int max_n1 = 0;
    for (const auto& sy : h_l1) max_n1 = std::max(max_n1, sy.n);
    if (use_cr) {
      // Level 1 by block cyclic reduction: one small kernel per level, ceil(log2(blocks)) + 1 levels.
      const size_t smem_cr = cr_smem_bytes(nbw1);
      for (int lv = 0; lv < cr_nlevels; ++lv) {
        const int nact = (cr_max_nblk + (1 << lv) - 1) >> lv;
        // Column split (CTAs per block): as many as keep the level's eliminated blocks within one wave of the SMs (survivor blocks cost
        // next to nothing); the wide first levels run one CTA per block.
        const int n_elim = std::max(1, nact / 2) * PL;
        const int nsm = 148;
        int z = 1;
        while (z < cr_zsplit && n_elim * (z * 2) <= nsm) z *= 2;
        if (lv == 0) CB2_K((cr_level_kernel<true>), dim3(nact, PL, 1), kCrThreads, smem_cr, stream, d_l1.p, lv, n_a, N_c, d_Aband.p, d_Bmat.p, d_Cmat.p, d_grad.p, d_dtil2.p, d_scal.p, 0);
        else CB2_K((cr_level_kernel<false>), dim3(nact, PL, z), kCrThreads, cr_tma ? cr_smem_bytes_tma(nbw1) : smem_cr, stream, d_l1.p, lv, n_a, N_c, d_Aband.p, d_Bmat.p, d_Cmat.p, d_grad.p, d_dtil2.p, d_scal.p, cr_tma ? 1 : 0);
        for (const GramPart& gp : gram_parts)
          if (gp.after_level == lv) {     // the W rows of the first levels' blocks are final: their share of the border Gram product, beside the next levels
            cudaStream_t sg = stream;
#ifndef CB2_EMUL
            if (stream_gram) { CB2_CUDA(cudaEventRecord(ev_gram[0], stream)); CB2_CUDA(cudaStreamWaitEvent(stream_gram, ev_gram[0], 0)); sg = stream_gram; }
#endif
            CB2_K(border_gram_dmma_kernel, dim3(gp.k_cnt, PL), 256, gram_smem_bytes(nbw1), sg, d_l1.p, gp.res, gp.mod, gp.k_off, gp.k_cnt);
            if (gram_prereduce && gp.k_cnt > 1)      // ... and its partials folded into one, still off the critical path
              CB2_K(gram_prereduce_kernel, dim3(std::min(148, (nbw1 * nbw1 + 255) / 256), PL), 256, 0, sg, d_l1.p, gp.k_off, gp.k_off + gp.k_cnt);
          }
      }
    } else {
      CB2_K(gather_level1_kernel, dim3(std::max(1, std::min(64, (max_n1 * (36 + nbw1) + 255) / 256)), PL), 256, 0, stream, d_l1.p, n_a, N_c, d_Aband.p,
            d_Bmat.p, d_Cmat.p, d_grad.p, d_dtil2.p);
      const size_t smem_f1 = factor_smem_bytes(36, nbw1);
      CB2_K((band_factor_kernel<6>), PL, kFacThreads, smem_f1, stream, d_l1.p, d_scal.p);
    }
    if (gram_dmma1 && !gram_parts.empty()) {
      const GramPart& gp = gram_parts.back();
      CB2_K(border_gram_dmma_kernel, dim3(gp.k_cnt, PL), 256, gram_smem_bytes(nbw1), stream, d_l1.p, gp.res, gp.mod, gp.k_off, gp.k_cnt);
#ifndef CB2_EMUL
      if (stream_gram) { CB2_CUDA(cudaEventRecord(ev_gram[3], stream_gram)); CB2_CUDA(cudaStreamWaitEvent(stream, ev_gram[3], 0)); }   // join
#endif
    } else if (gram_dmma1) CB2_K(border_gram_dmma_kernel, dim3(max_ksplit1, PL), 256, gram_smem_bytes(nbw1), stream, d_l1.p, 0, 1, 0, -1);
    else CB2_K(border_gram_kernel, dim3(max_tilepairs1, PL, max_ksplit1), dim3(16, 16), 0, stream, d_l1.p);
    // Separator + calibration systems: rank-local direct terms minus the Schur terms of the owned chunks, summed across ranks.
    if (h_l2.n > 0) {
      const long tot2 = long(h_l2.n) * (60 + h_l2.nbw);
      CB2_K(level2_build_kernel, int(std::min<long>((tot2 + 255) / 256, 1024)), 256, 0, stream, h_l2, d_l1.p, d_chunk_sys.p, P, n_a, N_c, d_Aband.p,
            d_Bmat.p, d_Cmat.p, d_grad.p, d_rawdiag.p);
    }
    if (N_c > 0) {
      const long tot3 = long(N_c + 1) * (N_c + 1);
      CB2_K(level3_build_kernel, int(std::min<long>((tot3 * 8 + 255) / 256, 4096)), 256, 0, stream, d_l1.p, PL, n_a, N_c, d_Cmat.p, d_grad.p, red_Cw,
            d_rawdiag.p + std::max(h_l2.n, 1));
    }
    if (world > 1) {
      // THE data-path collective of an LM iteration: separator + calibration system [+ the deferred shared-row sums of the normal equations].
      double* tailp = d_red.p + red_count;
      if (tail) CB2_K(pack_shared_kernel, (n_shared + 255) / 256, 256, 0, stream, n_shared, d_shared_idx.p, d_grad.p, d_diag.p, d_scal.p, world, rank, tailp);
      comm->allreduce_sum(d_red.p, red_count + (tail ? shared_buf_size(n_shared, world) : 0), stream);
      if (tail) {
        CB2_K(unpack_shared_kernel, (n_shared + 255) / 256, 256, 0, stream, n_shared, d_shared_idx.p, tailp, world, d_gradG.p, d_diag.p, d_scal.p);
        CB2_K(gradient_norm_kernel, lm_ctas(n_a), kLmThreads, 0, stream, n_a, d_gradG.p, d_cp_own.p, 0, 1, 1, d_desc.p, d_state[cur].p, ns, world_refs(), d_body_q[cur].p, d_scal.p, d_lm_partial.p, d_lm_ticket.p);
        CB2_K(damping_shared_kernel, (n_shared + 255) / 256, 256, 0, stream, n_shared, d_shared_idx.p, d_diag.p, d_scaling.p, d_scal.p, d_dtil2.p);
      }
    }
    if (h_l2.n > 0) {
      CB2_K(level2_damp_kernel, (h_l2.n + 255) / 256, 256, 0, stream, h_l2, d_dtil2.p);
      const int nt2 = (h_l2.nbw + 63) / 64;
      if (sep_cr) {
        // the separator chain by cyclic reduction: first level reads the summed separator system (band of width 60, border + rhs rows)
        const int nbw2 = h_l2.nbw;
        const bool tma2 = cr_tma && cr_smem_bytes_tma(nbw2) <= 227 * 1024;
        for (int lv = 0; lv < cr2_nlevels; ++lv) {
          const int nact = (h_l2.nblk + (1 << lv) - 1) >> lv;
          const int n_elim = std::max(1, nact / 2);
          int z = 1;
          while (z < cr_zsplit && n_elim * (z * 2) <= 148) z *= 2;
          if (lv == 0) CB2_K((cr_level_kernel<true>), dim3(nact, 1, z), kCrThreads, cr_smem_bytes(nbw2), stream, d_l2.p, lv, n_a, N_c, h_l2.L, h_l2.W, d_Cmat.p, h_l2.W + N_c,
                             static_cast<const double*>(nullptr), d_scal.p, 0, 60, nbw2, nbw2, 0L);
          else CB2_K((cr_level_kernel<false>), dim3(nact, 1, z), kCrThreads, tma2 ? cr_smem_bytes_tma(nbw2) : cr_smem_bytes(nbw2), stream, d_l2.p, lv, n_a, N_c, h_l2.L, h_l2.W,
                     d_Cmat.p, h_l2.W + N_c, static_cast<const double*>(nullptr), d_scal.p, tma2 ? 1 : 0, 60, nbw2, nbw2, 0L);
        }
        if (gram_dmma2) CB2_K(border_gram_dmma_kernel, dim3(h_l2.ksplit, 1), 256, gram_smem_bytes(nbw2), stream, d_l2.p);
        else CB2_K(border_gram_kernel, dim3(nt2 * (nt2 + 1) / 2, 1, h_l2.ksplit), dim3(16, 16), 0, stream, d_l2.p);
      } else {
        const size_t smem_f2 = factor_smem_bytes(60, h_l2.nbw);
        CB2_K((band_factor_kernel<10>), 1, kFacThreads, smem_f2, stream, d_l2.p, d_scal.p);
        CB2_K(border_gram_kernel, dim3(nt2 * (nt2 + 1) / 2, 1, h_l2.ksplit), dim3(16, 16), 0, stream, d_l2.p);
      }
    }
    if (N_c > 0) {
      const long tot3 = long(N_c + 1) * (N_c + 1);
      CB2_K(level3_finalize_kernel, int(std::min<long>((tot3 + 255) / 256, 1024)), 256, 0, stream, h_l2, N_c, n_a, red_Cw, d_dtil2.p);
      if (reduced_smem_bytes(N_c) <= 227 * 1024 - 256) CB2_K(reduced_solve_smem_kernel, 1, kRedThreads, reduced_smem_bytes(N_c), stream, N_c, n_a, red_Cw, d_ytil.p, d_scal.p);
      else CB2_K(reduced_solve_kernel, 1, kRedThreads, size_t(N_c + 1) * (kRedPanel + 1) * sizeof(double), stream, N_c, n_a, red_Cw, d_ytil.p, d_scal.p);
    }
    if (h_l2.n > 0) {
      CB2_K(border_matvec_kernel, dim3(std::max(1, std::min(148, (h_l2.n + 7) / 8)), 1), 256, size_t(h_l2.nbw) * sizeof(double), stream, d_l2.p, d_ytil.p);
      if (sep_cr) {
        int lv = cr2_nlevels - 1, lo = lv;   // levels with <= 8 eliminated blocks share one launch
        while (lo > 0 && std::max(1, ((h_l2.nblk + (1 << (lo - 1)) - 1) >> (lo - 1)) / 2) <= 8) --lo;
        CB2_K(cr_back_kernel, dim3(1, 1), 256, 0, stream, d_l2.p, lv, lo, d_ytil.p, static_cast<unsigned*>(nullptr));
        for (lv = lo - 1; lv >= 0; --lv) {
          const int nel = std::max(1, ((h_l2.nblk + (1 << lv) - 1) >> lv) / 2);
          CB2_K(cr_back_kernel, dim3((nel + 7) / 8, 1), 256, 0, stream, d_l2.p, lv, lv, d_ytil.p, static_cast<unsigned*>(nullptr));
        }
      } else CB2_K(band_backsolve_kernel, 1, kBackThreads, backsolve_smem_bytes(h_l2.n, h_l2.nbw, 60), stream, d_l2.p, d_ytil.p);
    }
    CB2_K(border_matvec_kernel, dim3(std::max(1, std::min(std::max(32, 592 / std::max(PL, 1)), (max_n1 + 7) / 8)), PL), 256, size_t(nbw1) * sizeof(double), stream, d_l1.p, d_ytil.p);
    if (use_cr) {
      const int nel0 = std::max(1, cr_max_nblk / 2);
      const int ctas0 = (nel0 + 7) / 8 * PL;
#ifndef CB2_EMUL
      // (opt-in: measured slower than one graph-captured launch per level on C4 — 57 us against ~45 us for the six launches)
      const bool fused_back = ctas0 <= 148 && std::getenv("CB2_FUSED_BACK") != nullptr;
#else
      const bool fused_back = false;
#endif
      if (fused_back) {
        // every level in ONE launch, the CTAs meeting at a global barrier between levels (all co-resident: at most one per SM)
        CB2_CUDA(cudaMemsetAsync(d_grid_sync.p, 0, sizeof(unsigned), stream));
        CB2_K(cr_back_kernel, dim3((nel0 + 7) / 8, PL), 256, 0, stream, d_l1.p, cr_nlevels - 1, 0, d_ytil.p, d_grid_sync.p);
      } else {
      // Top levels with <= 8 eliminated blocks each share one launch (block barrier between levels); then one launch per level.
      // With thread-block clusters: the levels with <= 64 eliminated blocks share one launch of ONE cluster of <= 8 CTAs per chunk.
      int lv = cr_nlevels - 1, lo = lv;
      const int fuse_cap = back_cluster ? 64 : 8;
      while (lo > 0 && std::max(1, ((cr_max_nblk + (1 << (lo - 1)) - 1) >> (lo - 1)) / 2) <= fuse_cap) --lo;
      const int cx = (std::max(1, ((cr_max_nblk + (1 << lo) - 1) >> lo) / 2) + 7) / 8;
#ifndef CB2_EMUL
      if (back_cluster && cx > 1) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(cx, PL); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cx; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        KernelProfiler::Rec pr{"cr_back_kernel(cluster)", nullptr, nullptr};
        if (kprof.on) { cudaEventCreate(&pr.a); cudaEventCreate(&pr.b); cudaEventRecord(pr.a, stream); }
        CB2_CUDA(cudaLaunchKernelEx(&cfg, cr_back_kernel, static_cast<const BandSys*>(d_l1.p), lv, lo, d_ytil.p, static_cast<unsigned*>(nullptr), 1));
        if (kprof.on) { cudaEventRecord(pr.b, stream); kprof.open.push_back(pr); }
        ++stats.kernel_launches;
      } else
#endif
      CB2_K(cr_back_kernel, dim3(1, PL), 256, 0, stream, d_l1.p, lv, lo, d_ytil.p, static_cast<unsigned*>(nullptr));
      for (lv = lo - 1; lv >= 0; --lv) {
        const int nact = (cr_max_nblk + (1 << lv) - 1) >> lv;
        const int nel = std::max(1, nact / 2);
        CB2_K(cr_back_kernel, dim3((nel + 7) / 8, PL), 256, 0, stream, d_l1.p, lv, lv, d_ytil.p, static_cast<unsigned*>(nullptr));
      }
      }
    } else {
      CB2_K(band_backsolve_kernel, PL, kBackThreads, backsolve_smem_bytes(max_n1, nbw1, 36), stream, d_l1.p, d_ytil.p);
    }
    CB2_K(apply_step_kernel, lm_ctas(n_a), kLmThreads, 0, stream, n_a, d_ytil.p, gradG(), d_dtil2.p, d_cp_ref.p, d_cp_own.p, rank == 0 ? 1 : 0, d_ctrl[cur].p,
          d_ctrl[cur ^ 1].p, d_desc.p, d_state[cur].p, d_state[cur ^ 1].p, ns, N_c, world_refs(), d_body_q[cur].p, d_body_t[cur].p, d_pm[cur].p,
          d_body_q[cur ^ 1].p, d_body_t[cur ^ 1].p, d_pm[cur ^ 1].p, d_scal.p, d_lm_partial.p + 8 * kLmMaxCtas, d_lm_ticket.p + 1);
    if (world_free && n_points > 0)
      CB2_K(world_points_kernel, (n_points + 255) / 256, 256, 0, stream, n_points, d_pt_body.p, d_body_q[cur ^ 1].p, d_body_t[cur ^ 1].p, d_pm[cur ^ 1].p, d_pw[cur ^ 1].p);
    });
    timer.end(kPhSchur, stream);
    launch_trial();
  }

  void launch_trial() {
    // The candidate buffer has just been overwritten by apply_step: whatever sweep it held belongs to an older (rejected) candidate.
    if (jac_point == (cur ^ 1)) jac_point = -1;
    if (imu_jac_point == (cur ^ 1)) imu_jac_point = -1;
    trial_was_sweep = speculative_sweep && speculate_next;
    if (trial_was_sweep) {
      timer.begin(kPhJacobian, stream);
      launch_eval<kModeJacobian>(cur ^ 1, kScCandCost, kImuSame);
      jac_point = cur ^ 1; imu_jac_point = cur ^ 1;
      if (world > 1) comm->allreduce_sum(d_scal.p + kScCandCost, 7, stream);
      timer.end(kPhJacobian, stream);
      ++stats.jacobian_sweeps;
      stats.jacobian_blocks += num_active_blocks();
      stats.jacobian_bytes += jacobian_bytes_per_sweep();
      stats.camera_kernel_bytes += jacobian_bytes_per_sweep(kCamera);
      stats.camera_kernel_gram_bytes += gram_bytes_per_sweep();
      return;
    }
    timer.begin(kPhCost, stream);
    // Trial point: cameras in cost-only mode; the (few, FP64-latency-bound) IMU blocks in Jacobian mode. Their Jacobians at x are not
    // needed any more (the normal equations of x are already assembled and survive a rejected step), and if the step is accepted the
    // Jacobians of the new x are then already there: the next sweep is the camera kernel alone.
    graphed(g_trial[cur], [&] { launch_eval<kModeCost>(cur ^ 1, kScCandCost, speculative_imu ? kImuJacobian : kImuSame); });
    if (speculative_imu) { imu_jac_point = cur ^ 1; stats.jacobian_blocks += num_active_blocks(true); }
    if (world > 1) comm->allreduce_sum(d_scal.p + kScCandCost, 7, stream);
    timer.end(kPhCost, stream);
  }

  void sync_scalars() {
    CB2_CUDA(cudaMemcpyAsync(h_scal, d_scal.p, sizeof(double) * kScCount, cudaMemcpyDeviceToHost, stream));
    sync_stream();
    CB2_CUDA(cudaGetLastError());
    stats.d2h_bytes += sizeof(double) * kScCount;
    timer.resolve(phase_ms);
    if (kprof.on) kprof.resolve();
    stats.jacobian_kernel_ms = phase_ms[kPhJacobian]; stats.normal_eq_ms = phase_ms[kPhNormal];
    stats.schur_ms = phase_ms[kPhSchur]; stats.cost_eval_ms = phase_ms[kPhCost]; stats.lm_loop_ms = phase_ms[kPhLoop]; stats.camera_kernel_ms = phase_ms[kPhCamera];
  }

  double jacobian_bytes_per_sweep(int only_kind = -1) const {   // obs_read + 8 m w + 8 m per residual block
    double total = 0.0;
    for (size_t si = 0; si < sensors.size(); ++si) {
      const SensorDesc& d = h_desc[si];
      if (only_kind >= 0 && d.kind != only_kind) continue;
      const double obs = d.kind == kCamera ? 32.0 : 40.0;
      total += double(d.n_obs) * (obs + 8.0 * d.m * d.jw + 8.0 * d.m);
    }
    return total;
  }
  double gram_bytes_per_sweep() const { return double(d_gslots.n + d_gcta.n) * sizeof(double); }
  long num_active_blocks(bool imu_only = false) const { long n = 0; for (const auto& d : h_desc) if (!imu_only || d.kind != kCamera) n += d.n_obs; return n; }

  // ------------------------------------------------------------------------------------------------------------
  // Trust-region Levenberg-Marquardt loop: TrustRegionMinimizer::Minimize (Ceres external), same control flow as the
  // reference's ceres::Solve call (batch_optimizer.cpp:73).
  // ------------------------------------------------------------------------------------------------------------
  void fill_summary_counts(cb2_summary& S) const {
    int nblocks = 0, nparams = 0, neff = 0;
    for (const auto& b : bodies) { nblocks += int(b.feature_ids.size()) + 2; nparams += 3 * int(b.feature_ids.size()) + 7; neff += 3 * int(b.feature_ids.size()) + 6; }
    nblocks += 1; nparams += 3; neff += 3;            // gravity
    nblocks += n_cp; nparams += 6 * n_cp; neff += 6 * n_cp;
    for (const auto& s : sensors) { nblocks += 4; nparams += int(s.intr.size()) + 8; neff += int(s.intr.size()) + 7; }
    S.num_parameter_blocks = nblocks; S.num_parameters = nparams; S.num_effective_parameters = neff;
    int rb = int(total_blocks), rr = int(total_residuals), pbr = 0, pr = 0, per = 0;
    for (const auto& d : h_desc) {
      if (d.u_intr >= 0) { ++pbr; pr += d.ni; per += d.ni; }
      if (d.u_rot >= 0) { pbr += 2; pr += 7; per += 6; }
      if (d.u_lat >= 0) { ++pbr; ++pr; ++per; }
    }
    const int ncp_ref = n_cp_referenced;
    pbr += ncp_ref; pr += 6 * ncp_ref; per += 6 * ncp_ref;
    pbr += 2 * n_world_pose_blocks + n_world_point_blocks;      // freed, referenced world-model blocks (pose = translation + quaternion block)
    pr += 7 * n_world_pose_blocks + 3 * n_world_point_blocks; per += 6 * n_world_pose_blocks + 3 * n_world_point_blocks;
    S.num_residual_blocks = rb; S.num_residuals = rr;
    S.num_residual_blocks_reduced = rb; S.num_residuals_reduced = rr;
    S.num_parameter_blocks_reduced = pbr; S.num_parameters_reduced = pr; S.num_effective_parameters_reduced = per;
  }
  int n_cp_referenced = 0, n_world_pose_blocks = 0, n_world_point_blocks = 0, n_sensor_unknowns = 0;

  int minimize(const cb2_options& opt, cb2_summary& S, std::vector<cb2_iteration>& L) {
    const double t_start = now_s();
    std::memset(&S, 0, sizeof(S));
    S.termination_type = CB2_FAILURE;
    fill_summary_counts(S);
    auto msg = [&](const char* fmt, double a = 0, double b = 0) { std::snprintf(S.message, sizeof S.message, fmt, a, b); };
    L.clear();
    if (total_blocks == 0 || n_cp_referenced == 0) {
      S.termination_type = CB2_CONVERGENCE;
      msg("Function tolerance reached. No non-constant parameter blocks found.");
      S.total_time = now_s() - t_start;
      return CB2_OK;
    }
    imu_jac_point = -1; jac_point = -1; speculate_next = true; ne_shared_pending = false;
    double radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
    double x_cost = 0, x_norm = 0, candidate_cost = 0, model_cost_change = 0, reference_cost = 0;
    int num_consecutive_invalid_steps = 0;
    cb2_iteration it{};
    timer.begin(kPhLoop, stream);
    const int blocks_tot = int(std::min<long>((n_tot + 255) / 256, 1024));

    auto evaluate_gradient_and_jacobian = [&](bool first) -> bool {
      const double t0 = now_s();
      launch_jacobian_and_normal_equations();
      if (first) CB2_K(jacobi_scaling_kernel, blocks_tot, 256, 0, stream, n_tot, d_diag.p, opt.jacobi_scaling, d_scaling.p);
      sync_scalars();
      S.jacobian_time += now_s() - t0;
      if (!sweep_skipped) {
        if (h_scal[kScInvalid] > 0) return false;
        x_cost = h_scal[kScCost];
      } else {
        x_cost = candidate_cost;   // the accepted trial point was evaluated by a full sweep: its cost is the candidate cost just read
      }
      it.cost = x_cost;
      it.gradient_max_norm = h_scal[kScGradMax];
      it.gradient_norm = std::sqrt(h_scal[kScGradSq]);
      return true;
    };

    bool pending_grad = false;
    auto resolve_pending_gradient = [&]() {
      if (!pending_grad) return;
      pending_grad = false;
      L.back().gradient_max_norm = h_scal[kScGradMax];
      L.back().gradient_norm = std::sqrt(h_scal[kScGradSq]);
    };
    double t_iter = now_s();
    it.iteration = 0; it.trust_region_radius = radius;
    if (!evaluate_gradient_and_jacobian(true)) {
      msg("Initial residual and Jacobian evaluation failed.");
      S.total_time = now_s() - t_start;
      return CB2_OK;
    }
    scaling_set = true;
    S.initial_cost = x_cost;
    it.step_is_valid = 1; it.step_is_successful = 1;
    reference_cost = x_cost;
    bool have_x_norm = false;
    if (opt.minimizer_progress_to_stdout)
      std::printf("iter      cost      cost_change  |gradient|   |step|    tr_ratio  tr_radius  ls_iter  iter_time  total_time\n");

    for (;;) {
      it.trust_region_radius = radius;
      it.iteration_time = now_s() - t_iter;
      L.push_back(it);
      if (opt.minimizer_progress_to_stdout)
        std::printf("% 4d % 8e   % 3.2e   % 3.2e  % 3.2e  % 3.2e % 3.2e     % 4d   % 3.2e   % 3.2e\n", it.iteration, it.cost, it.cost_change,
                    it.gradient_max_norm, it.step_norm, it.relative_decrease, it.trust_region_radius, 1, it.iteration_time, now_s() - t_start);
      if (pending_grad && (it.iteration >= opt.max_num_iterations || radius <= opt.min_trust_region_radius)) { finish_shared_sums(); sync_scalars(); resolve_pending_gradient(); }
      if (it.iteration >= opt.max_num_iterations) { S.termination_type = CB2_NO_CONVERGENCE; msg("Maximum number of iterations reached. Number of iterations: %.0f.", it.iteration); break; }
      if (!pending_grad && it.step_is_successful && it.gradient_max_norm <= opt.gradient_tolerance) {
        S.termination_type = CB2_CONVERGENCE; msg("Gradient tolerance reached. Gradient max norm: %e <= %e", it.gradient_max_norm, opt.gradient_tolerance); break;
      }
      if (radius <= opt.min_trust_region_radius) {
        S.termination_type = CB2_CONVERGENCE; msg("Minimum trust region radius reached. Trust region radius: %e <= %e", radius, opt.min_trust_region_radius); break;
      }
      t_iter = now_s();
      const double prev_gmax = it.gradient_max_norm, prev_gnorm = it.gradient_norm;
      const int next_iter = it.iteration + 1;
      it = cb2_iteration{};
      it.iteration = next_iter; it.gradient_max_norm = prev_gmax; it.gradient_norm = prev_gnorm;

      // LevenbergMarquardtStrategy::ComputeStep + candidate evaluation, one device round trip.
      const double t_ls = now_s();
      launch_step(radius, opt);
      sync_scalars();
      S.linear_solver_time += now_s() - t_ls;
      if (pending_grad) {
        // The normal equations of the point accepted in the previous iteration were enqueued without a host round trip; their gradient
        // norms arrive with this one. If they meet the gradient tolerance the solve just done is simply discarded: same termination
        // iteration, message and log as with an immediate check.
        resolve_pending_gradient();
        it.gradient_max_norm = L.back().gradient_max_norm; it.gradient_norm = L.back().gradient_norm;
        if (L.back().step_is_successful && L.back().gradient_max_norm <= opt.gradient_tolerance) {
          S.termination_type = CB2_CONVERGENCE;
          msg("Gradient tolerance reached. Gradient max norm: %e <= %e", L.back().gradient_max_norm, opt.gradient_tolerance);
          break;
        }
      }
      const bool solved = !(h_scal[kScSolveFail] > 0);
      if (debug_log)
        std::fprintf(stderr, "[cb2 debug] iter %d radius %.3e solve_fail %.0f model_change %.6e step2 %.3e cand_cost %.6e cand_invalid %.0f\n", next_iter, radius,
                     h_scal[kScSolveFail], h_scal[kScModelChange], h_scal[kScStepNorm2], h_scal[kScCandCost], h_scal[kScCandInvalid]);
      it.step_is_valid = 0;
      if (solved) {
        model_cost_change = h_scal[kScModelChange];
        it.step_is_valid = model_cost_change > 0.0;
        if (it.step_is_valid) num_consecutive_invalid_steps = 0;
      }
      if (!have_x_norm) { x_norm = std::sqrt(h_scal[kScXNorm2]); have_x_norm = true; }
      if (!it.step_is_valid) {
        if (++num_consecutive_invalid_steps >= opt.max_num_consecutive_invalid_steps) {
          S.termination_type = CB2_FAILURE;
          msg("Number of consecutive invalid steps more than Solver::Options::max_num_consecutive_invalid_steps: %.0f", opt.max_num_consecutive_invalid_steps);
          break;
        }
        radius *= 0.5;
        speculate_next = false;
        it.cost = x_cost; it.cost_change = 0.0; it.step_norm = 0.0; it.relative_decrease = 0.0;
        continue;
      }
      candidate_cost = h_scal[kScCandInvalid] > 0 ? std::numeric_limits<double>::max() : h_scal[kScCandCost];   // "Step failed to evaluate."
      it.step_norm = std::sqrt(h_scal[kScStepNorm2]);
      if (it.step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) {
        S.termination_type = CB2_CONVERGENCE;
        msg("Parameter tolerance reached. Relative step_norm: %e <= %e.", it.step_norm / (x_norm + opt.parameter_tolerance), opt.parameter_tolerance);
        break;
      }
      it.cost_change = x_cost - candidate_cost;
      if (std::fabs(it.cost_change) <= opt.function_tolerance * x_cost) {
        S.termination_type = CB2_CONVERGENCE;
        msg("Function tolerance reached. |cost_change|/cost: %e <= %e", std::fabs(it.cost_change) / x_cost, opt.function_tolerance);
        break;
      }
      if (candidate_cost >= std::numeric_limits<double>::max()) it.relative_decrease = std::numeric_limits<double>::lowest();
      else it.relative_decrease = std::max((x_cost - candidate_cost) / model_cost_change, (reference_cost - candidate_cost) / model_cost_change);
      if (it.relative_decrease > opt.min_relative_decrease) {
        x_norm = std::sqrt(h_scal[kScCandXNorm2]);
        cur ^= 1;   // x = candidate_x
        if (defer_normal_sync && jac_point == cur && !opt.minimizer_progress_to_stdout) {
          // The trial pass was a full sweep of this point: only the normal equations remain, and nothing the host must decide before the
          // next solve depends on them except the gradient-tolerance test, which is applied one round trip later (see above).
          launch_jacobian_and_normal_equations(/*defer_shared_sums=*/true);
          x_cost = candidate_cost;
          it.cost = x_cost;
          pending_grad = true;
        } else if (!evaluate_gradient_and_jacobian(false)) { S.termination_type = CB2_FAILURE; msg("Residual and Jacobian evaluation failed."); break; }
        it.step_is_successful = 1;
        speculate_next = true;
        radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * it.relative_decrease - 1.0, 3));
        radius = std::min(opt.max_trust_region_radius, radius);
        decrease_factor = 2.0;
        reference_cost = x_cost;
        ++S.num_successful_steps;
      } else {
        it.step_is_successful = 0;
        speculate_next = false;
        it.cost = candidate_cost;
        radius = radius / decrease_factor; decrease_factor *= 2.0;
        ++S.num_unsuccessful_steps;
      }
    }
    // Ceres (solver.cc, SetSummaryFinalCost): the minimum over ALL recorded iterations, rejected ones included (their cost column is the
    // candidate cost, trust_region_minimizer.cc HandleUnsuccessfulStep) — reproduced as is so that Summary::final_cost reads the same.
    S.final_cost = S.initial_cost;
    for (const auto& e : L) S.final_cost = std::min(S.final_cost, e.cost);
    S.num_iterations = int(L.size());
    stats.lm_iterations += int(L.size()) - 1;
    timer.end(kPhLoop, stream);
    sync_stream();
    timer.resolve(phase_ms);
    stats.lm_loop_ms = phase_ms[kPhLoop];
    S.total_time = now_s() - t_start;
    return CB2_OK;
  }

  // Device -> host write-back of the optimised parameters (the reference mutates the user's objects in place).
  void download_parameters() {
    if (world > 1) {
      // Every rank holds its own interiors + the separators; one sum of the masked vectors gives everybody the full trajectory.
      double* tmp = d_ctrl[cur ^ 1].p;
      CB2_K(mask_ctrl_kernel, int(std::min<long>((n_a + 255) / 256, 1024)), 256, 0, stream, n_a, d_cp_own.p, rank == 0 ? 1 : 0, d_ctrl[cur].p, tmp);
      comm->allreduce_sum(tmp, size_t(n_a), stream);
      CB2_CUDA(cudaMemcpyAsync(d_ctrl[cur].p, tmp, sizeof(double) * n_a, cudaMemcpyDeviceToDevice, stream));
      sync_stream();
    }
    d_ctrl[cur].download(ctrl, &stats.d2h_bytes);
    d_state[cur].download(h_state, &stats.d2h_bytes);
    if (world_free) {   // the freed rigid-body poses / model points (world_model.cpp:52-70: Ceres mutates them in place)
      std::vector<double> bq, bt, bp;
      d_body_q[cur].download(bq, &stats.d2h_bytes); d_body_t[cur].download(bt, &stats.d2h_bytes); d_pm[cur].download(bp, &stats.d2h_bytes);
      for (size_t bi = 0; bi < bodies.size(); ++bi) {
        HostBody& b = bodies[bi];
        for (int j = 0; j < 4; ++j) b.q[j] = bq[4 * bi + j];
        for (int j = 0; j < 3; ++j) b.t[j] = bt[3 * bi + j];
        for (size_t f = 0; f < b.feature_ids.size(); ++f) for (int j = 0; j < 3; ++j) b.pts[3 * f + j] = bp[3 * (b.pw0 + f) + j];
      }
    }
    for (size_t si = 0; si < sensors.size(); ++si) {
      HostSensor& s = sensors[si];
      const SensorState& st = h_state[si];
      for (size_t j = 0; j < s.intr.size(); ++j) s.intr[j] = st.intr[j];
      s.q[0] = st.q.x; s.q[1] = st.q.y; s.q[2] = st.q.z; s.q[3] = st.q.w;
      s.t[0] = st.t.x; s.t[1] = st.t.y; s.t[2] = st.t.z;
      s.latency = st.latency;
    }
  }

  // Sensor::UpdateResiduals (camera.cpp:70-80, gyroscope.cpp:171-182, accelerometer.cpp:58-69): EVERY measurement's un-robustified residual.
  // The sweep leaves residuals in the device order (sorted by spline segment, this rank's shard only); a scatter kernel puts them back into
  // the caller's observation order on the device — all sensors in one buffer [r of sensor 0 | r of sensor 1 | ...] + a flag byte per observation —
  // so that (i) with several ranks ONE sum over the ranks gives every rank every residual (an observation is evaluated by exactly one
  // rank), and (ii) the host receives final-layout arrays: two device -> host copies, no host-side un-permutation.
  DevBuf<double> d_rfull;
  DevBuf<unsigned char> d_vfull;
  PinnedBuf h_res;
  int refresh_residuals() {
    launch_eval<kModeResiduals>(cur, kScCost);
    if (world > 1) comm->allreduce_sum(d_scal.p + kScCost, 2, stream);
    const int ns = int(sensors.size());
    std::vector<size_t> roff(ns + 1, 0), voff(ns + 1, 0);
    for (int si = 0; si < ns; ++si) { roff[si + 1] = roff[si] + size_t(sensors[si].n_obs()) * sensors[si].m(); voff[si + 1] = voff[si] + size_t(sensors[si].n_obs()); }
    if (d_rfull.n != roff[ns] || d_vfull.n != voff[ns]) { d_rfull.alloc(roff[ns], false); d_vfull.alloc(voff[ns], false); CB2_CUDA(cudaDeviceSynchronize()); }
    d_rfull.zero(stream); d_vfull.zero(stream);     // outliers (and, before the sum, the other ranks' observations) stay 0 / invalid
    for (int si = 0; si < ns; ++si) {
      HostSensor& s = sensors[si];
      if (s.n_active > 0)
        CB2_K(scatter_residuals_kernel, std::min((s.n_active + 255) / 256, 1184), 256, 0, stream, s.n_active, s.m(), s.d_perm, s.d_r.p, s.d_valid.p, d_rfull.p + roff[si],
              d_vfull.p + voff[si]);
    }
    if (world > 1 && roff[ns] > 0) { comm->allreduce_sum(d_rfull.p, roff[ns], stream); comm->allreduce_sum_u8(d_vfull.p, voff[ns], stream); }
    sync_scalars();
    // Device -> host: two asynchronous copies into the problem's pinned result block; cb2_get_residuals hands out slices of it.
    const size_t rbytes = roff[ns] * sizeof(double);
    h_res.acquire(std::max<size_t>(rbytes + voff[ns], 256));
    unsigned char* hb = static_cast<unsigned char*>(h_res.p);
    if (roff[ns]) {
      CB2_CUDA(cudaMemcpyAsync(hb, d_rfull.p, rbytes, cudaMemcpyDeviceToHost, stream));
      CB2_CUDA(cudaMemcpyAsync(hb + rbytes, d_vfull.p, voff[ns], cudaMemcpyDeviceToHost, stream));
      stats.d2h_bytes += int64_t(rbytes + voff[ns]);
    }
    for (int si = 0; si < ns; ++si) {
      sensors[si].residuals = reinterpret_cast<const double*>(hb) + roff[si];
      sensors[si].residual_valid = hb + rbytes + voff[si];
    }
    sync_stream();
    int rc = CB2_OK;
    if (h_scal[kScInvalid] > 0) {
      // Some residual block could not be evaluated at the final point: the reference fails with kInternal on the first such sensor (camera.cpp:73-76).
      for (int si = 0; si < ns && rc == CB2_OK; ++si) {
        const HostSensor& s = sensors[si];
        bool bad = false;
        for (int o = 0; o < s.n_obs() && !bad; ++o) bad = !s.residual_valid[o] && !(s.kind == kCamera && !s.outlier.empty() && s.outlier[o]);
        if (bad) {
          const char* kind = s.kind == kCamera ? "camera " : (s.kind == kGyroscope ? "gyroscope " : "accelerometer ");
          rc = fail(CB2_INTERNAL, std::string("Failed to update residual for ") + kind + s.name);
        }
      }
    }
    return rc;
  }
};

// ================================================================================================================
// C ABI
// ================================================================================================================
// No C++ exception may cross the C ABI: every entry point runs its body through one of these.
template <class F>
static int guarded(cb2_problem* p, F&& body) {
  try {
    return body();
  } catch (const CudaFail& f) {
    return p ? p->fail(CB2_INTERNAL, f.msg) : CB2_INTERNAL;
  } catch (const std::bad_alloc&) {
    try { return p ? p->fail(CB2_INTERNAL, "Out of host memory.") : CB2_INTERNAL; } catch (...) { return CB2_INTERNAL; }
  } catch (const std::exception& e) {
    try { return p ? p->fail(CB2_INTERNAL, std::string("Unexpected error: ") + e.what()) : CB2_INTERNAL; } catch (...) { return CB2_INTERNAL; }
  } catch (...) {
    return CB2_INTERNAL;
  }
}

extern "C" {

const char* cb2_version(void) {
#ifdef CB2_EMUL
  return "calico_b200 0.1 (SIMT-emulation test build)";
#else
  return "calico_b200 0.1 (sm_100a)";
#endif
}

void cb2_default_options(cb2_options* o) {
  o->max_num_iterations = 50;
  o->function_tolerance = 1e-8;
  o->gradient_tolerance = 1e-10;
  o->parameter_tolerance = 1e-10;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->max_num_consecutive_invalid_steps = 5;
  o->jacobi_scaling = 1;
  o->num_threads = 1;
  o->minimizer_progress_to_stdout = 1;
  o->linear_solver = 0;
  o->use_cuda_graph = 0;
}

int cb2_problem_create(cb2_problem** out) {
  if (!out) return CB2_INVALID_ARGUMENT;
  *out = nullptr;
  return guarded(nullptr, [&]() -> int { *out = new cb2_problem(); return CB2_OK; });
}
void cb2_problem_destroy(cb2_problem* p) { try { delete p; } catch (...) {} }
const char* cb2_last_error(cb2_problem* p) { return p ? p->error.c_str() : ""; }

// The number of intrinsics of a model: CameraModel::NumberOfParameters (camera_models.h:79,231,395,596,716,848,961) and the IMU
// models' (accelerometer_models.h / gyroscope_models.h: 1, 4, 12); -1 for an unknown kind / model. One definition for every host layer.
int cb2_num_intrinsics(int kind, int model) {
  if (kind == kCamera) return camera_num_params(model);
  if (kind == kGyroscope || kind == kAccelerometer) return imu_num_params(model);
  return -1;
}

int cb2_set_device(int device) { return cudaSetDevice(device) == cudaSuccess ? CB2_OK : CB2_INTERNAL; }

int cb2_set_trajectory(cb2_problem* p, int spline_order, int n_knots, const double* knots, int n_cp, const double* ctrl) {
  return guarded(p, [&]() -> int {
  if (spline_order < 2) return p->fail(CB2_INVALID_ARGUMENT, "Spline order must be greater than 2.");   // bspline.hpp:27-30
  if (n_knots != n_cp + spline_order) return p->fail(CB2_INVALID_ARGUMENT, "Knot vector size must equal control points + spline order.");
  p->k = spline_order;
  p->knots.assign(knots, knots + n_knots);
  p->ctrl.assign(ctrl, ctrl + size_t(n_cp) * 6);
  p->uploaded = false;
  return CB2_OK;
  });
}
int cb2_set_gravity(cb2_problem* p, const double* g) { std::memcpy(p->gravity, g, 24); p->uploaded = false; return CB2_OK; }

int cb2_add_rigid_body(cb2_problem* p, int id, const double* q, const double* t, int n_pts, const int* feature_ids, const double* pts,
                       int pose_const, int model_const) {
  return guarded(p, [&]() -> int {
  if (p->body_slot.count(id)) return p->fail(CB2_INVALID_ARGUMENT, "Rigid body with id " + std::to_string(id) + " already exists in world model.");   // world_model.cpp:32-35
  HostBody b;
  b.id = id; std::memcpy(b.q, q, 32); std::memcpy(b.t, t, 24);
  b.pose_const = pose_const != 0; b.model_const = model_const != 0;
  b.feature_ids.assign(feature_ids, feature_ids + n_pts);
  b.pts.assign(pts, pts + size_t(n_pts) * 3);
  int max_id = -1, min_id = 0;
  for (int i = 0; i < n_pts; ++i) { b.slot[feature_ids[i]] = i; max_id = std::max(max_id, feature_ids[i]); min_id = std::min(min_id, feature_ids[i]); }
  if (min_id >= 0 && max_id < (1 << 22)) {
    b.dense_slot.assign(size_t(max_id) + 1, -1);
    for (int i = 0; i < n_pts; ++i) b.dense_slot[feature_ids[i]] = i;
  }
  p->body_slot[id] = int(p->bodies.size());
  p->bodies.push_back(std::move(b));
  p->uploaded = false;
  return CB2_OK;
  });
}

int cb2_add_sensor(cb2_problem* p, int kind, int model, const char* name, int n_intr, const double* intr, const double* q, const double* t,
                   double latency, double sigma, int loss_type, double loss_scale, int en_intr, int en_extr, int en_lat, int* sensor_id) {
  return guarded(p, [&]() -> int {
  if (kind < 0 || kind > 2) return p->fail(CB2_INVALID_ARGUMENT, "Unknown sensor kind.");
  if (sigma <= 0.0) return p->fail(CB2_INVALID_ARGUMENT, "Sigma must be greater than 0.");   // camera.cpp:62-65
  p->sensors.emplace_back();
  HostSensor& s = p->sensors.back();
  s.kind = kind; s.model = model; s.name = name ? name : "";
  s.intr.assign(intr, intr + n_intr);
  std::memcpy(s.q, q, 32); std::memcpy(s.t, t, 24);
  s.latency = latency; s.sigma = sigma; s.loss_type = loss_type; s.loss_scale = loss_scale;
  s.en_intr = en_intr != 0; s.en_extr = en_extr != 0; s.en_lat = en_lat != 0;
  *sensor_id = int(p->sensors.size()) - 1;
  p->uploaded = false;
  return CB2_OK;
  });
}

int cb2_add_camera_observations(cb2_problem* p, int sid, int n, const double* stamp, const int* image_id, const int* model_id,
                                const int* feature_id, const double* pixel, const uint8_t* outlier) {
  (void)image_id;
  return guarded(p, [&]() -> int {
  if (sid < 0 || sid >= int(p->sensors.size()) || p->sensors[sid].kind != kCamera) return p->fail(CB2_INVALID_ARGUMENT, "Not a camera sensor id.");
  if (n < 0) return p->fail(CB2_INVALID_ARGUMENT, "Negative observation count.");
  HostSensor& s = p->sensors[sid];
  // All ids are resolved into local vectors first: a failed call leaves the sensor exactly as it was.
  std::vector<int> bslot(n), fslot(n);
  int last_model = 0, last_bs = -2;   // consecutive observations almost always name the same rigid body
  for (int i = 0; i < n; ++i) {
    if (last_bs == -2 || model_id[i] != last_model) {
      auto it = p->body_slot.find(model_id[i]);
      last_bs = it != p->body_slot.end() ? it->second : -1;
      last_model = model_id[i];
    }
    const int bs = last_bs;
    int fs = -1;
    if (bs >= 0) {
      const HostBody& b = p->bodies[bs];
      const int fid = feature_id[i];
      if (!b.dense_slot.empty()) fs = (fid >= 0 && size_t(fid) < b.dense_slot.size()) ? b.dense_slot[fid] : -1;
      else { auto f = b.slot.find(fid); fs = f != b.slot.end() ? f->second : -1; }
      if (fs < 0) return p->fail(CB2_INVALID_ARGUMENT, "Feature id not in rigid body model definition.");
    }
    bslot[i] = bs; fslot[i] = fs;
  }
  const size_t base = s.stamp.size();
  s.stamp.reserve(base + n); s.meas.reserve((base + n) * 2); s.body_slot.reserve(base + n); s.feat_slot.reserve(base + n); s.outlier.reserve(base + n);
  s.stamp.append(stamp, n);
  s.meas.append(pixel, size_t(n) * 2);
  s.body_slot.append(bslot.data(), n);
  s.feat_slot.append(fslot.data(), n);
  if (outlier) s.outlier.append(outlier, n); else s.outlier.append_fill(size_t(n), uint8_t(0));
  s.residuals = nullptr; s.residual_valid = nullptr;   // the residuals of an earlier Optimize no longer match the measurement set
  p->uploaded = false;
  return CB2_OK;
  });
}

int cb2_add_imu_observations(cb2_problem* p, int sid, int n, const double* stamp, const int* seq, const double* xyz) {
  (void)seq;
  return guarded(p, [&]() -> int {
  if (sid < 0 || sid >= int(p->sensors.size()) || p->sensors[sid].kind == kCamera) return p->fail(CB2_INVALID_ARGUMENT, "Not an IMU sensor id.");
  if (n < 0) return p->fail(CB2_INVALID_ARGUMENT, "Negative observation count.");
  HostSensor& s = p->sensors[sid];
  s.stamp.reserve(s.stamp.size() + n); s.meas.reserve(s.meas.size() + size_t(n) * 3); s.outlier.reserve(s.outlier.size() + n);
  s.stamp.append(stamp, n);
  s.meas.append(xyz, size_t(n) * 3);
  s.outlier.append_fill(size_t(n), uint8_t(0));
  s.residuals = nullptr; s.residual_valid = nullptr;
  p->uploaded = false;
  return CB2_OK;
  });
}

int cb2_upload(cb2_problem* p) { return guarded(p, [&]() -> int { return p->uploaded ? CB2_OK : p->upload(); }); }

int cb2_optimize(cb2_problem* p, const cb2_options* opts, cb2_summary* summary, cb2_iteration* log, int log_cap, int* n_log) {
  cb2_options o;
  if (opts) o = *opts; else cb2_default_options(&o);
  return guarded(p, [&]() -> int {
    if (!p->uploaded) { const int rc = p->upload(); if (rc != CB2_OK) return rc; }
    std::vector<cb2_iteration> L;
    cb2_summary S;
    const double t0 = now_s();
    int rc = p->minimize(o, S, L);
    const double t1 = now_s();
    if (summary) *summary = S;
    if (n_log) *n_log = int(L.size());
    if (log) for (int i = 0; i < int(L.size()) && i < log_cap; ++i) log[i] = L[i];
    if (rc != CB2_OK) return rc;
    p->download_parameters();
    const double t2 = now_s();
    rc = p->refresh_residuals();
    if (p->timing_log) std::fprintf(stderr, "[cb2 timing] optimize: LM loop %.2f ms, parameter write-back %.2f ms, residual refresh %.2f ms\n", 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (now_s() - t2));
    return rc;
  });
}

int cb2_evaluate_sensor(cb2_problem* p, int sid, double* residuals, double* jacobians, uint8_t* valid) {
  return guarded(p, [&]() -> int {
    if (!p->uploaded) { const int rc = p->upload(); if (rc != CB2_OK) return rc; }
    if (sid < 0 || sid >= int(p->sensors.size())) return p->fail(CB2_INVALID_ARGUMENT, "Unknown sensor id.");
    HostSensor& s = p->sensors[sid];
    const SensorDesc& d0 = p->h_desc[sid];
    const int m = s.m(), ni = d0.ni, W = kCpCols + ni + 7, na = s.n_active;
    if (na == 0) return CB2_OK;
    // Temporary descriptor storing every canonical column, un-robustified.
    SensorDesc d = d0;
    d.gslots = nullptr; d.gcta = nullptr;
    d.n_jcal = ni + 7; d.jw = W;
    for (int j = 0; j < ni + 7; ++j) { d.jcanon[j] = j; d.junk[j] = j; }
    DevBuf<double> J, r;
    DevBuf<unsigned char> v;
    DevBuf<SensorDesc> dd;
    DevBuf<EvalTile> dt;
    DevBuf<double> cpart;
    DevBuf<int> ipart;
    if (jacobians) J.alloc(size_t(na) * m * W);
    r.alloc(size_t(na) * m); v.alloc(na);
    d.J = J.p; d.r = r.p; d.valid = v.p;
    dd.upload(std::vector<SensorDesc>(1, d));
    std::vector<EvalTile> tiles;
    for (int o0 = 0, T = eval_tile(s.kind); o0 < na; o0 += T) tiles.push_back(EvalTile{0, o0, std::min(T, na - o0)});
    dt.upload(tiles);
    cpart.alloc(tiles.size()); ipart.alloc(tiles.size());
    const SensorState* st = p->d_state[p->cur].p + sid;
    const double* c = p->d_ctrl[p->cur].p;
    const size_t smem = size_t(rec_size(s.kind, ni)) * eval_rec_stride(s.kind) * sizeof(double);
    const int nt = int(tiles.size());
    cudaStream_t stream = p->stream;
    cb2_stats& stats = p->stats;
    KernelProfiler& kprof = p->kprof;
#define CB2_EVAL(KIND, MODE)                                                                                                             \
  CB2_K((eval_kernel<KIND, MODE>), nt, eval_tile(KIND), smem, stream, dd.p, st, dt.p, c, p->d_knots.p, p->d_basis.p, p->d_pw[p->cur].p, p->d_frames.p, \
        p->gravity[0], p->gravity[1], p->gravity[2], cpart.p, ipart.p, 0)
    // residuals + validity (un-robustified), then the Jacobian pass if requested
    if (s.kind == kCamera) p->launch_frames(p->cur);
    if (s.kind == kCamera) CB2_EVAL(kCamera, kModeResiduals); else if (s.kind == kGyroscope) CB2_EVAL(kGyroscope, kModeResiduals); else CB2_EVAL(kAccelerometer, kModeResiduals);
    CB2_CUDA(cudaStreamSynchronize(stream));
    std::vector<double> hr; std::vector<unsigned char> hv;
    r.download(hr); v.download(hv);
    std::vector<double> hJ;
    if (jacobians) {
      if (s.kind == kCamera) CB2_EVAL(kCamera, kModeJacobian); else if (s.kind == kGyroscope) CB2_EVAL(kGyroscope, kModeJacobian); else CB2_EVAL(kAccelerometer, kModeJacobian);
      CB2_CUDA(cudaStreamSynchronize(stream));
      J.download(hJ);
    }
#undef CB2_EVAL
    CB2_CUDA(cudaGetLastError());
    for (int i = 0; i < na; ++i) {
      const int o = s.perm[i];
      if (valid) valid[o] = hv[i];
      if (!hv[i]) continue;
      if (residuals) for (int q = 0; q < m; ++q) residuals[size_t(o) * m + q] = hr[size_t(i) * m + q];
      if (jacobians) std::memcpy(jacobians + size_t(o) * m * W, hJ.data() + size_t(i) * m * W, sizeof(double) * m * W);
    }
    return CB2_OK;
  });
}

int cb2_cost(cb2_problem* p, double* cost, int* ok) {
  return guarded(p, [&]() -> int {
    if (!p->uploaded) { const int rc = p->upload(); if (rc != CB2_OK) return rc; }
    p->launch_eval<kModeCost>(p->cur, kScCost);
    if (p->world > 1) p->comm->allreduce_sum(p->d_scal.p + kScCost, 2, p->stream);
    p->sync_scalars();
    *cost = p->h_scal[kScCost];
    *ok = p->h_scal[kScInvalid] > 0 ? 0 : 1;
    return CB2_OK;
  });
}

int cb2_get_sensor(cb2_problem* p, int sid, double* intr, double* q, double* t, double* latency) {
  if (sid < 0 || sid >= int(p->sensors.size())) return p->fail(CB2_INVALID_ARGUMENT, "Unknown sensor id.");
  const HostSensor& s = p->sensors[sid];
  if (intr) std::memcpy(intr, s.intr.data(), s.intr.size() * 8);
  if (q) std::memcpy(q, s.q, 32);
  if (t) std::memcpy(t, s.t, 24);
  if (latency) *latency = s.latency;
  return CB2_OK;
}
int cb2_set_sensor(cb2_problem* p, int sid, const double* intr, const double* q, const double* t, double latency) {
  if (sid < 0 || sid >= int(p->sensors.size())) return p->fail(CB2_INVALID_ARGUMENT, "Unknown sensor id.");
  HostSensor& s = p->sensors[sid];
  if (intr) std::memcpy(s.intr.data(), intr, s.intr.size() * 8);
  if (q) std::memcpy(s.q, q, 32);
  if (t) std::memcpy(s.t, t, 24);
  s.latency = latency;
  p->uploaded = false;
  return CB2_OK;
}
// RigidBody write-back (world_model.h:41-69): pose and model definition after cb2_optimize; pts_xyz in the order the points were added.
int cb2_get_rigid_body(cb2_problem* p, int id, double* q, double* t, double* pts) {
  auto it = p->body_slot.find(id);
  if (it == p->body_slot.end()) return p->fail(CB2_INVALID_ARGUMENT, "Unknown rigid body id.");
  const HostBody& b = p->bodies[it->second];
  if (q) std::memcpy(q, b.q, 32);
  if (t) std::memcpy(t, b.t, 24);
  if (pts) std::memcpy(pts, b.pts.data(), b.pts.size() * 8);
  return CB2_OK;
}
int cb2_get_trajectory(cb2_problem* p, double* ctrl) { std::memcpy(ctrl, p->ctrl.data(), p->ctrl.size() * 8); return CB2_OK; }
int cb2_get_residuals(cb2_problem* p, int sid, double* out, uint8_t* valid) {
  if (sid < 0 || sid >= int(p->sensors.size())) return p->fail(CB2_INVALID_ARGUMENT, "Unknown sensor id.");
  const HostSensor& s = p->sensors[sid];
  if (!s.residuals && s.n_obs() > 0) return p->fail(CB2_FAILED_PRECONDITION, "Residuals have not been computed.");
  if (out && s.n_obs() > 0) std::memcpy(out, s.residuals, size_t(s.n_obs()) * s.m() * 8);
  if (valid && s.n_obs() > 0) std::memcpy(valid, s.residual_valid, size_t(s.n_obs()));
  return CB2_OK;
}

int cb2_reset_parameters(cb2_problem* p) {
  if (!p->uploaded) return p->fail(CB2_FAILED_PRECONDITION, "Problem has not been uploaded.");
  return guarded(p, [&]() -> int {
    p->cur = 0;
    CB2_CUDA(cudaMemcpyAsync(p->d_ctrl[0].p, p->d_ctrl0.p, p->d_ctrl0.n * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
    CB2_CUDA(cudaMemcpyAsync(p->d_state[0].p, p->d_state0.p, p->d_state0.n * sizeof(SensorState), cudaMemcpyDeviceToDevice, p->stream));
    if (p->world_free && p->n_points > 0) {
      CB2_CUDA(cudaMemcpyAsync(p->d_body_q[0].p, p->d_body_q0.p, p->d_body_q0.n * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
      CB2_CUDA(cudaMemcpyAsync(p->d_body_t[0].p, p->d_body_t0.p, p->d_body_t0.n * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
      CB2_CUDA(cudaMemcpyAsync(p->d_pm[0].p, p->d_pm0.p, p->d_pm0.n * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
      CB2_LAUNCH(world_points_kernel, (p->n_points + 255) / 256, 256, 0, p->stream, p->n_points, p->d_pt_body.p, p->d_body_q[0].p, p->d_body_t[0].p, p->d_pm[0].p, p->d_pw[0].p);
    }
    CB2_CUDA(cudaStreamSynchronize(p->stream));
    return CB2_OK;
  });
}

int cb2_stats_reset(cb2_problem* p) { p->stats = cb2_stats{}; for (auto& v : p->phase_ms) v = 0.0; return CB2_OK; }
int cb2_stats_get(cb2_problem* p, cb2_stats* out) { *out = p->stats; return CB2_OK; }

// ---- multi-GPU: one process per GPU; call cb2_set_device first, then cb2_comm_init on every rank before cb2_upload/optimize ----
// ---- trajectory spline fit: BSpline<6>::FitToData / FitSpline (bspline.hpp:20-38,247-297) and
//      Trajectory::FitSpline (trajectory.cpp:14-49) ----
static thread_local std::string g_fit_error;
static int fit_fail(int code, const std::string& msg) { g_fit_error = msg; return code; }
const char* cb2_fit_last_error(void) { return g_fit_error.c_str(); }

int cb2_fit_spline_size(int n, const double* times, int spline_order, double knot_frequency, int* n_knots, int* n_cp) {
  // CheckDataForSplineFit (bspline.hpp:299-327) + ComputeKnotVector (bspline.hpp:164-180)
  if (n <= 0 || !times) return fit_fail(CB2_INVALID_ARGUMENT, "Attempted to fit data on empty time vector.");
  for (int i = 1; i < n; ++i) if (times[i - 1] > times[i]) return fit_fail(CB2_INVALID_ARGUMENT, "Time vector is not monotonically increasing.");
  if (spline_order < 2) return fit_fail(CB2_INVALID_ARGUMENT, "Spline order must be greater than 2. Got " + std::to_string(spline_order));
  if (!(knot_frequency > 0)) return fit_fail(CB2_INVALID_ARGUMENT, "Knot frequency must be greater than 0.");
  if (spline_order != kK) return fit_fail(CB2_UNIMPLEMENTED, "Only spline order 6 (calico::Trajectory::kSplineOrder, trajectory.h:28) is supported.");
  const double duration = times[n - 1] - times[0];
  const int num_valid = 1 + int(std::ceil(duration * knot_frequency));
  if (num_valid < 2) return fit_fail(CB2_INVALID_ARGUMENT, "Time span too short for a spline segment.");
  if (n_knots) *n_knots = num_valid + 2 * (spline_order - 1);
  if (n_cp) *n_cp = num_valid + spline_order - 2;
  return CB2_OK;
}

int cb2_fit_spline(int n, const double* times, const double* data6, int spline_order, double knot_frequency, int n_knots, double* knots_out,
                   int n_cp, double* ctrl_out) {
  int nk = 0, ncp = 0;
  int rc = cb2_fit_spline_size(n, times, spline_order, knot_frequency, &nk, &ncp);
  if (rc != CB2_OK) return rc;
  if (!data6) return fit_fail(CB2_INVALID_ARGUMENT, "Attempted to fit on empty data.");
  if (nk != n_knots || ncp != n_cp || !knots_out || !ctrl_out) return fit_fail(CB2_INVALID_ARGUMENT, "Output sizes do not match cb2_fit_spline_size.");
  try {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return fit_fail(CB2_INTERNAL, "No CUDA device available: calico_b200 has no CPU fallback.");
    const int deg = spline_order - 1, n_seg = nk - 2 * deg - 1;
    std::vector<double> knots(nk);
    const double dt = 1.0 / knot_frequency;
    for (int i = -deg; i < nk - deg; ++i) knots[i + deg] = times[0] + dt * i;
    std::vector<double> basis(size_t(n_seg) * 36);
    for (int sg = 0; sg < n_seg; ++sg) cb2_problem::basis_matrix(knots, sg + deg, &basis[size_t(sg) * 36]);
    // Segment of every sample (FitSpline's lookup, bspline.hpp:255-267; times are sorted): CSR over segments.
    const double* vk = knots.data() + deg;
    const int nv = n_seg + 1;
    std::vector<int> seg_start(n_seg + 1, 0);
    for (int j = 0; j < n; ++j) {
      const double t = times[j];
      int sg;
      if (t == vk[nv - 1]) sg = n_seg - 1;
      else if (t < vk[nv - 1]) sg = int(std::upper_bound(vk, vk + nv, t) - vk) - 1;
      else return fit_fail(CB2_INVALID_ARGUMENT, "Sample time is past the last valid knot.");
      ++seg_start[sg + 1];
    }
    for (int g = 0; g < n_seg; ++g) seg_start[g + 1] += seg_start[g];
    cudaStream_t st;
    CB2_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    {
      DevBuf<double> d_t, d_d, d_kn, d_bs, d_sg, d_sr, d_A, d_B, d_L, d_c;
      DevBuf<int> d_ss, d_fail;
      d_t.upload(std::vector<double>(times, times + n)); d_d.upload(std::vector<double>(data6, data6 + size_t(n) * 6));
      d_kn.upload(knots); d_bs.upload(basis); d_ss.upload(seg_start);
      d_sg.alloc(size_t(n_seg) * kFitGram); d_sr.alloc(size_t(n_seg) * 36);
      d_A.alloc(size_t(ncp) * 6); d_B.alloc(size_t(ncp) * 6); d_L.alloc(size_t(ncp) * 6); d_c.alloc(size_t(ncp) * 6); d_fail.alloc(1);
      CB2_CUDA(cudaDeviceSynchronize());
      CB2_LAUNCH(fit_accumulate_kernel, (n_seg + 3) / 4, 128, 0, st, n_seg, d_ss.p, d_t.p, d_d.p, d_kn.p, d_bs.p, d_sg.p, d_sr.p);
      CB2_LAUNCH(fit_assemble_kernel, std::min((ncp * 12 + 255) / 256, 1024), 256, 0, st, ncp, n_seg, d_sg.p, d_sr.p, d_A.p, d_B.p);
      CB2_LAUNCH(fit_solve_kernel, 1, 32, 0, st, ncp, d_A.p, d_B.p, 1e-14, d_L.p, d_c.p, d_fail.p);
      CB2_CUDA(cudaStreamSynchronize(st));
      CB2_CUDA(cudaGetLastError());
      std::vector<double> c; std::vector<int> f;
      d_c.download(c); d_fail.download(f);
      if (f[0]) { cudaStreamDestroy(st); return fit_fail(CB2_INTERNAL, "Spline fit: normal equations are not positive definite."); }
      std::memcpy(ctrl_out, c.data(), sizeof(double) * c.size());
      std::memcpy(knots_out, knots.data(), sizeof(double) * knots.size());
    }
    cudaStreamDestroy(st);
    return CB2_OK;
  } catch (const CudaFail& f) {
    return fit_fail(CB2_INTERNAL, f.msg);
  } catch (const std::exception& e) {
    return fit_fail(CB2_INTERNAL, std::string("Unexpected error: ") + e.what());
  }
}

int cb2_fit_trajectory(int n, const double* stamps, const double* q_xyzw, const double* t3, int spline_order, double knot_frequency, int n_knots,
                       double* knots_out, int n_cp, double* ctrl_out) {
  if (n <= 0 || !stamps || !q_xyzw || !t3) return fit_fail(CB2_INVALID_ARGUMENT, "Attempted to fit data on empty time vector.");
  // trajectory.cpp:19-24: timestamps sorted; :29-37: Eigen::AngleAxisd(rotation) -> axis * angle; :38 UnwrapPhaseLogMap (:81-93).
  try {
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return stamps[a] < stamps[b]; });
  std::vector<double> ts(n), data(size_t(n) * 6);
  for (int i = 0; i < n; ++i) {
    const int o = order[i];
    ts[i] = stamps[o];
    const double x = q_xyzw[4 * o], y = q_xyzw[4 * o + 1], z = q_xyzw[4 * o + 2], w = q_xyzw[4 * o + 3];
    // Eigen::AngleAxis = Quaternion: angle = 2 atan2(|vec|, |w|), axis = vec / |vec| (negated when w < 0); identity -> axis (1,0,0), angle 0.
    const double nrm = std::sqrt(x * x + y * y + z * z);
    double ax = 1.0, ay = 0.0, az = 0.0, angle = 0.0;
    if (nrm > 0.0) {
      angle = 2.0 * std::atan2(nrm, std::fabs(w));
      const double sgn = w < 0.0 ? -1.0 : 1.0;
      ax = sgn * x / nrm; ay = sgn * y / nrm; az = sgn * z / nrm;
    }
    double* d = &data[size_t(i) * 6];
    d[0] = ax * angle; d[1] = ay * angle; d[2] = az * angle;
    d[3] = t3[3 * o]; d[4] = t3[3 * o + 1]; d[5] = t3[3 * o + 2];
  }
  for (int i = 1; i < n; ++i) {   // UnwrapPhaseLogMap, trajectory.cpp:81-93
    double* p = &data[size_t(i) * 6];
    const double* q = &data[size_t(i - 1) * 6];
    const double theta = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    if (theta == 0.0) continue;
    const double k = std::round((p[0] * q[0] + p[1] * q[1] + p[2] * q[2] - theta * theta) / (2.0 * M_PI * theta));
    const double sc = 1.0 + 2.0 * M_PI * k / theta;
    p[0] *= sc; p[1] *= sc; p[2] *= sc;
  }
  return cb2_fit_spline(n, ts.data(), data.data(), spline_order, knot_frequency, n_knots, knots_out, n_cp, ctrl_out);
  } catch (const std::exception& e) {
    return fit_fail(CB2_INTERNAL, std::string("Unexpected error: ") + e.what());
  }
}

int cb2_comm_unique_id(uint8_t* id128) {
#if defined(CB2_EMUL)
  (void)id128;
  return CB2_UNIMPLEMENTED;
#else
  if (!nccl().load()) return CB2_INTERNAL;
  NcclApi::UniqueId id;
  if (nccl().GetUniqueId(&id) != 0) return CB2_INTERNAL;
  std::memcpy(id128, &id, sizeof id);
  return CB2_OK;
#endif
}
int cb2_comm_init(cb2_problem* p, int world_size, int rank, const uint8_t* id128) {
  if (world_size < 1 || rank < 0 || rank >= world_size) return p->fail(CB2_INVALID_ARGUMENT, "Invalid world size / rank.");
  p->uploaded = false;
  if (world_size == 1) { p->comm.reset(); p->world = 1; p->rank = 0; return CB2_OK; }
#if defined(CB2_EMUL)
  (void)id128;
  return p->fail(CB2_UNIMPLEMENTED, "The emulation build has no NCCL; use cb2_comm_init_local.");
#else
  try {
    if (!nccl().load()) return p->fail(CB2_INTERNAL, nccl().error);
    const int rc0 = p->ensure_device();
    if (rc0 != CB2_OK) return rc0;
    std::unique_ptr<NcclComm> c(new NcclComm());
    c->world = world_size; c->rank = rank;
    NcclApi::UniqueId id;
    std::memcpy(&id, id128, sizeof id);
    c->check(nccl().CommInitRank(&c->comm, world_size, id, rank), "ncclCommInitRank");
    c->start_watchdog();
    p->comm = std::shared_ptr<Comm>(c.release());
    p->world = world_size; p->rank = rank;
    return CB2_OK;
  } catch (const CudaFail& f) {
    return p->fail(CB2_INTERNAL, f.msg);
  } catch (const std::exception& e) {
    return p->fail(CB2_INTERNAL, std::string("Unexpected error: ") + e.what());
  }
#endif
}
#if defined(CB2_EMUL)
// Test hook of the emulation build: ranks are host threads of one process that meet in the group named `group`.
int cb2_comm_init_local(cb2_problem* p, int world_size, int rank, const char* group) {
  p->uploaded = false;
  std::unique_ptr<LocalComm> c(new LocalComm());
  c->world = world_size; c->rank = rank;
  {
    std::lock_guard<std::mutex> lk(g_groups_mu);
    LocalGroup*& g = local_groups()[group];
    if (!g) { g = new LocalGroup(); g->world = world_size; }
    c->grp = g;
  }
  p->comm = std::shared_ptr<Comm>(c.release());
  p->world = world_size; p->rank = rank;
  return CB2_OK;
}
#endif
// Re-uses the communicator of another handle of this process (creating an NCCL communicator is a one-time, seconds-long setup).
int cb2_comm_clone(cb2_problem* dst, cb2_problem* src) {
  dst->comm = src->comm; dst->world = src->world; dst->rank = src->rank; dst->uploaded = false;
  return CB2_OK;
}
// The shard this rank takes of a trajectory with n_cp control points (host-side plan only; no device needed):
// chunks [chunk_lo, chunk_hi) of n_chunks, spline segments [seg_lo, seg_hi).
int cb2_shard_plan(cb2_problem* p, int world_size, int rank, int* n_chunks, int* chunk_lo, int* chunk_hi, int* seg_lo, int* seg_hi) {
  if (p->ctrl.empty()) return p->fail(CB2_FAILED_PRECONDITION, "Trajectory has not been set.");
  if (world_size < 1 || rank < 0 || rank >= world_size) return p->fail(CB2_INVALID_ARGUMENT, "Invalid world size / rank.");
  const int w0 = p->world, r0 = p->rank;
  p->world = world_size; p->rank = rank;
  p->n_cp = int(p->ctrl.size() / 6); p->n_seg = p->n_cp - (kK - 1);
  const int rc = p->plan_chunks();
  if (rc == CB2_OK) {
    *n_chunks = int(p->chunks.size()); *chunk_lo = p->chunk_lo; *chunk_hi = p->chunk_hi; *seg_lo = p->g_lo; *seg_hi = p->g_hi;
  }
  p->world = w0; p->rank = r0;
  p->uploaded = false;
  return rc;
}

}  // extern "C"
