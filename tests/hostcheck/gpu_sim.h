// Test-only: the execution environment shared by the CPU models of the kernels (tc_sim.h: tcgen05 kernels, te_sim.h:
// tile-engine kernels): one OS thread per GPU thread, one CTA at a time; CUDA built-ins the kernels use (threadIdx /
// blockIdx, __syncthreads, __syncthreads_or, __shfl_xor_sync, bit casts, float2 / float4 / uint4); a per-CTA state
// object with the dynamic shared memory, mbarrier table, warp scratch and the list of model violations.
#pragma once
#include <math.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <barrier>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct SimDim3 { unsigned x = 0, y = 0, z = 0; };
inline thread_local SimDim3 threadIdx, blockIdx, blockDim, gridDim;
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
template <class T> static inline T min(T a, T b) { return a < b ? a : b; }
template <class T> static inline void __stcg(T* p, T v) { *p = v; }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }

namespace sim {

constexpr size_t DYN_SMEM = 232448;
struct Mbar { int count = 0, pending = 0; long long tx = 0; unsigned phase = 0; bool init = false; };

struct State {
  alignas(1024) unsigned char smem[DYN_SMEM];
  float tmem[128 * 512];
  bool tmem_allocated = false;
  std::mutex m;
  std::condition_variable cv;
  std::unordered_map<uint32_t, Mbar> bars;
  std::unordered_map<const void*, Mbar> pbars;          // mbarriers addressed by pointer (tile engine)
  std::unique_ptr<std::barrier<>> named_barrier;        // bar.sync 1, 256
  std::unique_ptr<std::barrier<>> named_barrier64;      // bar.sync 2, 64
  std::atomic<int> or_accum{0};
  std::function<void()> on_thread_exit;                 // e.g. complete the thread's outstanding bulk stores
  std::vector<const void*> static_ptrs;                 // handles of shared objects outside the dynamic buffer
  std::unique_ptr<std::barrier<>> cta_barrier;
  std::vector<std::unique_ptr<std::barrier<>>> warp_barrier;
  std::vector<std::vector<float>> warp_buf;
  std::vector<std::string> errors;
  long long mma_count = 0;
};
inline State& S() { static State s; return s; }
inline void fail(const std::string& msg) {
  std::lock_guard<std::mutex> lk(S().m);
  if (S().errors.size() < 32) S().errors.push_back(msg);
}
inline std::vector<std::string>& errors() { return S().errors; }

// run `body` once per thread of every CTA of the grid (CTAs one after the other)
inline void launch(int grid, int block, const std::function<void()>& body) {
  State& st = S();
  for (int b = 0; b < grid; ++b) {
    memset(st.tmem, 0xff, sizeof st.tmem);              // NaN pattern: reading an unwritten accumulator shows
    if (getenv("APG_SIM_POISON_SMEM")) {                  // stress: shared memory starts as NaNs, not as zeros
      const uint32_t nan_bits = 0x7fc00000u;
      for (size_t i = 0; i + 4 <= DYN_SMEM; i += 4) memcpy(st.smem + i, &nan_bits, 4);
    }
    st.tmem_allocated = false;
    st.bars.clear();
    st.pbars.clear();
    st.static_ptrs.clear();
    st.or_accum = 0;
    if (block >= 256) st.named_barrier.reset(new std::barrier<>(256));
    st.named_barrier64.reset(new std::barrier<>(64));
    st.cta_barrier.reset(new std::barrier<>(block));
    const int nwarp = (block + 31) / 32;
    st.warp_barrier.clear();
    st.warp_buf.assign(nwarp, std::vector<float>(32, 0.f));
    for (int w = 0; w < nwarp; ++w) st.warp_barrier.emplace_back(new std::barrier<>(std::min(32, block - 32 * w)));
    std::vector<std::thread> th;
    th.reserve(block);
    for (int t = 0; t < block; ++t)
      th.emplace_back([=, &body]() {
        threadIdx.x = (unsigned)t; blockIdx.x = (unsigned)b; blockDim.x = (unsigned)block; gridDim.x = (unsigned)grid;
        body();
        if (sim::S().on_thread_exit) sim::S().on_thread_exit();
      });
    for (auto& t : th) t.join();
    if (st.tmem_allocated) fail("CTA " + std::to_string(b) + " exited without tcgen05.dealloc");
  }
}
}  // namespace sim

// ---- the few CUDA runtime names the launchers and the C-ABI layer use -------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
typedef void* cudaStream_t;
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum { cudaStreamNonBlocking = 1 };
namespace sim {
inline int& max_dynamic_smem_set() { static int v = 0; return v; }     // last cudaFuncSetAttribute value
}
template <class K>
static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int bytes) {
  if (bytes < 0 || (size_t)bytes > sim::DYN_SMEM) { sim::fail("dynamic shared memory request of " + std::to_string(bytes) + " B exceeds the 227 KB limit"); return cudaErrorInvalidValue; }
  sim::max_dynamic_smem_set() = bytes;
  return cudaSuccess;
}
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "simulated CUDA error"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = reinterpret_cast<void*>(1); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
// the "device" of the model: a handful of SMs (APG_SIM_SMS, default 3) keeps the grids small
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) {
  const char* e = getenv("APG_SIM_SMS");
  *v = e ? atoi(e) : 3;
  return cudaSuccess;
}

namespace sim {
// kernel<<<grid, block, smem>>>(args...) on the model.  Shared-memory guard: everything beyond the requested dynamic
// shared memory is filled with a canary before every CTA and checked afterwards (a kernel that writes past the size
// its launcher computed is caught here; on hardware it would fault or corrupt silently).
struct LaunchCfg { dim3 grid; int block; size_t smem; };
template <class F>
struct BoundLaunch {
  LaunchCfg l;
  F f;
  template <class... A>
  void operator()(A&&... args) {
    State& st = S();
    if (l.smem > DYN_SMEM) { fail("launch with more dynamic shared memory than an SM has"); return; }
    if ((int)l.smem > max_dynamic_smem_set() && l.smem > 48 * 1024)
      fail("launch with " + std::to_string(l.smem) + " B of dynamic shared memory without cudaFuncSetAttribute");
    // canary = quiet NaNs: a READ past the requested size poisons the result, a WRITE is found afterwards
    const size_t first = (l.smem + 3) / 4 * 4, words = std::min<size_t>((DYN_SMEM - first) / 4, 4096);
    const uint32_t nan_bits = 0x7fc00000u;
    for (unsigned by = 0; by < l.grid.y; ++by) {
      for (size_t i = 0; i < words; ++i) memcpy(st.smem + first + 4 * i, &nan_bits, 4);
      launch((int)l.grid.x, l.block, [&]() { blockIdx.y = by; gridDim.y = l.grid.y; f(args...); });
      for (size_t i = 0; i < words; ++i) {
        uint32_t v;
        memcpy(&v, st.smem + first + 4 * i, 4);
        if (v != nan_bits) {
          fail("a kernel wrote beyond its dynamic shared memory (" + std::to_string(l.smem) + " B requested, word +" +
               std::to_string(i) + ")");
          break;
        }
      }
    }
  }
};
struct Launcher {
  LaunchCfg cfg;
  Launcher(dim3 g, int b, size_t s) : cfg{g, b, s} {}
  Launcher(int g, int b, size_t s) : cfg{dim3((unsigned)g), b, s} {}
  Launcher(unsigned g, int b, size_t s) : cfg{dim3(g), b, s} {}
  template <class F>
  BoundLaunch<F> bind(F f) { return BoundLaunch<F>{cfg, f}; }
};
}  // namespace sim

static inline void __syncthreads() { sim::S().cta_barrier->arrive_and_wait(); }
// barrier + OR of the predicate over the CTA
static inline int __syncthreads_or(int pred) {
  sim::State& st = sim::S();
  if (pred) st.or_accum.store(1);
  st.cta_barrier->arrive_and_wait();
  const int r = st.or_accum.load();
  st.cta_barrier->arrive_and_wait();
  if (threadIdx.x == 0) st.or_accum.store(0);
  st.cta_barrier->arrive_and_wait();
  return r;
}
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_add(v); }
// warp vote: true iff the predicate holds on every lane of the warp
static inline int __all_sync(unsigned, int pred) {
  sim::State& st = sim::S();
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  st.warp_buf[w][l] = pred ? 1.f : 0.f;
  st.warp_barrier[w]->arrive_and_wait();
  int all = 1;
  for (size_t i = 0; i < 32 && 32 * (size_t)w + i < blockDim.x; ++i) all &= st.warp_buf[w][i] != 0.f;
  st.warp_barrier[w]->arrive_and_wait();
  return all;
}
static inline void __syncwarp(unsigned = 0xffffffffu) { sim::S().warp_barrier[threadIdx.x >> 5]->arrive_and_wait(); }
static inline void __trap() { sim::fail("__trap()"); throw std::runtime_error("__trap"); }
static inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
  sim::State& st = sim::S();
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  st.warp_buf[w][l] = v;
  st.warp_barrier[w]->arrive_and_wait();
  const float r = st.warp_buf[w][l ^ lane_mask];
  st.warp_barrier[w]->arrive_and_wait();
  return r;
}

