// TEST HARNESS ONLY — a minimal SIMT emulator so the calico_b200 kernel sources can be compiled by g++ and their
// indexing / synchronisation logic debugged without a GPU (tests/emul/build.py → libcalico_b200_emul.so).
// Never part of the product: calico_b200/ loads only the nvcc-built libcalico_b200.so.
//
// Model: a launch runs its blocks one after another; inside a block every CUDA thread is an OS thread;
// __syncthreads() is a std::barrier over the block, warp collectives exchange through per-warp slots.
// `__shared__` becomes `static` (safe because only one block is live at a time).
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

struct alignas(16) double2 { double x, y; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };

namespace cb2emul {
struct Tls { dim3 tIdx, bIdx, bDim, gDim; int linear_tid = 0; };
inline Tls& tls() { static thread_local Tls t; return t; }
struct Block {
  std::barrier<>* block_bar = nullptr;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<uint64_t> slots;   // 32 per warp
  unsigned char* dyn = nullptr;
};
inline Block& blk() { static Block b; return b; }
inline unsigned char* dyn_smem_base() { return blk().dyn; }

inline std::mutex& launch_mutex() { static std::mutex m; return m; }

template <class F>
void launch(dim3 grid, dim3 block, size_t smem, F&& f) {
  std::lock_guard<std::mutex> launch_lock(launch_mutex());   // one emulated kernel at a time (several host threads = ranks)
  const int nt = int(block.x * block.y * block.z);
  const long nb = long(grid.x) * grid.y * grid.z;
  if (nt <= 0 || nb <= 0) return;
  std::barrier<> bar(nt);
  Block& B = blk();
  B.block_bar = &bar;
  const int nwarps = (nt + 31) / 32;
  B.warp_bar.clear();
  for (int w = 0; w < nwarps; ++w) B.warp_bar.emplace_back(new std::barrier<>(std::min(32, nt - 32 * w)));
  B.slots.assign(size_t(nwarps) * 32, 0);
  std::vector<unsigned char> dyn(smem + 64);
  B.dyn = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn.data()) + 63) & ~uintptr_t(63));
  std::vector<std::thread> th;
  th.reserve(nt);
  for (int t = 0; t < nt; ++t) {
    th.emplace_back([&, t]() {
      Tls& T = tls();
      T.bDim = block; T.gDim = grid; T.linear_tid = t;
      T.tIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
      for (long b = 0; b < nb; ++b) {
        T.bIdx = dim3(unsigned(b % grid.x), unsigned((b / grid.x) % grid.y), unsigned(b / (long(grid.x) * grid.y)));
        f();
        bar.arrive_and_wait();
      }
    });
  }
  for (auto& t : th) t.join();
  B.block_bar = nullptr;
}
}  // namespace cb2emul

#define threadIdx (::cb2emul::tls().tIdx)
#define blockIdx (::cb2emul::tls().bIdx)
#define blockDim (::cb2emul::tls().bDim)
#define gridDim (::cb2emul::tls().gDim)
#define warpSize 32
using std::min;
using std::max;

inline void __syncthreads() { ::cb2emul::blk().block_bar->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { ::cb2emul::blk().warp_bar[::cb2emul::tls().linear_tid / 32]->arrive_and_wait(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }

namespace cb2emul {
template <typename T>
inline T warp_exchange(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "");
  Block& B = blk();
  const int tid = tls().linear_tid, w = tid / 32, lane = tid % 32;
  uint64_t raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  B.slots[size_t(w) * 32 + lane] = raw;
  B.warp_bar[w]->arrive_and_wait();
  const int nlanes = std::min(32, int(tls().bDim.x * tls().bDim.y * tls().bDim.z) - 32 * w);
  T out = v;
  if (src_lane >= 0 && src_lane < nlanes) { raw = B.slots[size_t(w) * 32 + src_lane]; std::memcpy(&out, &raw, sizeof(T)); }
  B.warp_bar[w]->arrive_and_wait();
  return out;
}
}  // namespace cb2emul

namespace cb2emul {
// mma.sync.aligned.m8n8k4.row.col.f64: lane l holds a = A[l / 4][l % 4], b = B[l % 4][l / 4], c0, c1 = D[l / 4][2 (l % 4) + {0, 1}].
inline void dmma_8x8x4(double& c0, double& c1, double a, double b) {
  Block& B = blk();
  const int tid = tls().linear_tid, w = tid / 32, lane = tid % 32;
  double A[32], Bv[32];
  uint64_t raw;
  std::memcpy(&raw, &a, 8);
  B.slots[size_t(w) * 32 + lane] = raw;
  B.warp_bar[w]->arrive_and_wait();
  for (int l = 0; l < 32; ++l) { raw = B.slots[size_t(w) * 32 + l]; std::memcpy(&A[l], &raw, 8); }
  B.warp_bar[w]->arrive_and_wait();
  std::memcpy(&raw, &b, 8);
  B.slots[size_t(w) * 32 + lane] = raw;
  B.warp_bar[w]->arrive_and_wait();
  for (int l = 0; l < 32; ++l) { raw = B.slots[size_t(w) * 32 + l]; std::memcpy(&Bv[l], &raw, 8); }
  B.warp_bar[w]->arrive_and_wait();
  const int mrow = lane / 4;
  for (int i = 0; i < 2; ++i) {
    const int n = 2 * (lane % 4) + i;
    double acc = i == 0 ? c0 : c1;
    for (int k = 0; k < 4; ++k) acc = std::fma(A[mrow * 4 + k], Bv[n * 4 + k], acc);
    (i == 0 ? c0 : c1) = acc;
  }
}
}  // namespace cb2emul

template <typename T> inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  const int lane = ::cb2emul::tls().linear_tid % 32;
  return ::cb2emul::warp_exchange(v, (lane / width) * width + (src % width));
}
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
  const int lane = ::cb2emul::tls().linear_tid % 32;
  (void)width;
  return ::cb2emul::warp_exchange(v, lane ^ m);
}
template <typename T> inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
  const int lane = ::cb2emul::tls().linear_tid % 32;
  const int src = lane + int(d);
  return ::cb2emul::warp_exchange(v, (src / width == lane / width) ? src : lane);
}
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
  const int lane = ::cb2emul::tls().linear_tid % 32;
  const int src = lane - int(d);
  return ::cb2emul::warp_exchange(v, (src >= 0 && src / width == lane / width) ? src : lane);
}
inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned out = 0;
  for (int l = 0; l < 32; ++l) out |= (::cb2emul::warp_exchange<int>(pred ? 1 : 0, l) ? 1u : 0u) << l;
  // lanes beyond the block return their own value through warp_exchange; mask them out
  const int tid = ::cb2emul::tls().linear_tid, w = tid / 32;
  const int nt = int(::cb2emul::tls().bDim.x * ::cb2emul::tls().bDim.y * ::cb2emul::tls().bDim.z);
  const int nlanes = std::min(32, nt - 32 * w);
  if (nlanes < 32) out &= (1u << nlanes) - 1u;
  return out;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(unsigned v) { return __builtin_ffs(int(v)); }
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) {
  const int tid = ::cb2emul::tls().linear_tid, w = tid / 32;
  const int nt = int(::cb2emul::tls().bDim.x * ::cb2emul::tls().bDim.y * ::cb2emul::tls().bDim.z);
  const int nlanes = std::min(32, nt - 32 * w);
  const unsigned full = nlanes < 32 ? (1u << nlanes) - 1u : 0xffffffffu;
  return __ballot_sync(m, pred) == full;
}

inline double atomicAdd(double* addr, double v) {
  uint64_t* a = reinterpret_cast<uint64_t*>(addr);
  uint64_t old = __atomic_load_n(a, __ATOMIC_RELAXED), nw;
  double od;
  do { std::memcpy(&od, &old, 8); const double nd = od + v; std::memcpy(&nw, &nd, 8); }
  while (!__atomic_compare_exchange_n(a, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED));
  return od;
}
inline int atomicAdd(int* a, int v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned* a, unsigned v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* a, unsigned long long v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
inline int atomicOr(int* a, int v) { return __atomic_fetch_or(a, v, __ATOMIC_SEQ_CST); }
inline int atomicMax(int* a, int v) {
  int old = __atomic_load_n(a, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(a, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {}
  return old;
}
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline double __ldg(const double* p) { return *p; }
inline int __ldg(const int* p) { return *p; }

// ---- runtime API subset ----
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorEmul = 1, cudaErrorNotReady = 600 };
typedef struct cb2emul_stream* cudaStream_t;
struct cb2emul_event { std::chrono::steady_clock::time_point t; };
typedef cb2emul_event* cudaEvent_t;
typedef void* cudaGraphExec_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDefault = 0 };
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = 0; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = reinterpret_cast<cudaStream_t>(1); return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? cudaSuccess : cudaErrorEmul; }
template <typename T> inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n); }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
template <typename T> inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n); }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new cb2emul_event(); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
constexpr unsigned cudaEventDisableTiming = 2;
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
