// Test-only software model of the execution environment of the tcgen05 kernels (csrc/hutter_tc_kernels.cu,
// csrc/adj_dw_tc_kernels.cu compiled with -DAPG_TC_SIM): one OS thread per GPU thread, one CTA at a time.
//   * CUDA built-ins the kernels use (threadIdx / blockIdx, __syncthreads, __shfl_xor_sync, bit casts, float2 / uint4)
//   * csrc/tc_prims.cuh implemented in software:
//       TMEM      128 lanes x 512 columns of 32 bits; tcgen05.ld / st check the lane-quarter rule (warp w may only
//                 touch lanes 32*(w%4) .. +31) and the column range
//       mbarrier  expected-arrival count + phase parity, try_wait / test_wait with the PTX parity semantics
//       tcgen05.mma  queued by the issuing thread and EXECUTED AT COMMIT TIME - the latest moment the hardware may
//                 read its operands - by the descriptor-decoding model of tc_emu.h (canonical unswizzled K-major /
//                 MN-major forms, TF32 truncation); an epilogue that overwrites an operand before it has waited for
//                 the commit therefore corrupts the result here too
//   * every violation (bad descriptor, lane rule, column overflow, barrier misuse) is recorded in sim::errors().
// What this cannot show: that the hardware agrees with the model (descriptor encodings, A-from-TMEM, MN-major B) -
// that is what tools/micro/tcgen05_gemm.cu and the GPU parity tests are for.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <barrier>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
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
static thread_local SimDim3 threadIdx, blockIdx, blockDim, gridDim;
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct uint4 { unsigned x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
template <class T> static inline T min(T a, T b) { return a < b ? a : b; }

namespace sim {

constexpr size_t DYN_SMEM = 232448;
struct Mbar { int count = 0, pending = 0; unsigned phase = 0; bool init = false; };

struct State {
  alignas(1024) unsigned char smem[DYN_SMEM];
  float tmem[128 * 512];
  bool tmem_allocated = false;
  std::mutex m;
  std::condition_variable cv;
  std::unordered_map<uint32_t, Mbar> bars;
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
    st.tmem_allocated = false;
    st.bars.clear();
    st.static_ptrs.clear();
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
      });
    for (auto& t : th) t.join();
    if (st.tmem_allocated) fail("CTA " + std::to_string(b) + " exited without tcgen05.dealloc");
  }
}
}  // namespace sim

static inline void __syncthreads() { sim::S().cta_barrier->arrive_and_wait(); }
static inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
  sim::State& st = sim::S();
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  st.warp_buf[w][l] = v;
  st.warp_barrier[w]->arrive_and_wait();
  const float r = st.warp_buf[w][l ^ lane_mask];
  st.warp_barrier[w]->arrive_and_wait();
  return r;
}

namespace apg {
// what the kernels take from tile_engine.cuh
enum Act { ACT_NONE = 0, ACT_TANH = 1, ACT_RELU = 2, ACT_SIGMOID = 3 };
static inline float act_apply(float v, int act) {
  if (act == ACT_TANH) return tanhf(v);
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  return v;
}
// "shared-window address": byte offset inside the dynamic buffer, or a handle (bit 30) for the kernels' static
// __shared__ objects
static inline uint32_t smem_u32(const void* p) {
  sim::State& st = sim::S();
  const unsigned char* c = static_cast<const unsigned char*>(p);
  if (c >= st.smem && c < st.smem + sim::DYN_SMEM) return (uint32_t)(c - st.smem);
  std::lock_guard<std::mutex> lk(st.m);
  for (size_t i = 0; i < st.static_ptrs.size(); ++i)
    if (st.static_ptrs[i] == p) return 0x40000000u + (uint32_t)i;
  st.static_ptrs.push_back(p);
  return 0x40000000u + (uint32_t)(st.static_ptrs.size() - 1);
}
}  // namespace apg

#include "tc_emu.h"

namespace apg {
namespace tcp {

struct QueuedMma { uint32_t d, a_tmem; uint64_t a_desc, b_desc; uint32_t idesc, acc; bool ts; };
static thread_local std::vector<QueuedMma> t_queue;

inline unsigned char* dynamic_smem() { return sim::S().smem; }
inline void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  t_queue.push_back({d, a, 0, b, idesc, acc, true});
}
inline void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  t_queue.push_back({d, 0, a, b, idesc, acc, false});
}
inline void mbar_init(uint32_t bar, int count) {
  std::lock_guard<std::mutex> lk(sim::S().m);
  sim::Mbar& b = sim::S().bars[bar];
  b.count = b.pending = count; b.phase = 0; b.init = true;
}
inline void mbar_arrive(uint32_t bar) {
  sim::State& st = sim::S();
  bool bad = false;
  {
    std::lock_guard<std::mutex> lk(st.m);
    sim::Mbar& b = st.bars[bar];
    if (!b.init) bad = true;
    else if (--b.pending == 0) { b.phase ^= 1u; b.pending = b.count; }
  }
  if (bad) sim::fail("arrive on an uninitialised mbarrier");
  st.cv.notify_all();
}
// true once the phase with the given parity has completed (PTX mbarrier.try_wait.parity)
inline bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  sim::State& st = sim::S();
  std::unique_lock<std::mutex> lk(st.m);
  sim::Mbar& b = st.bars[bar];
  return st.cv.wait_for(lk, std::chrono::microseconds(500), [&] { return b.init && b.phase != (parity & 1u); });
}
inline bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  bool ok;
  {
    std::lock_guard<std::mutex> lk(sim::S().m);
    sim::Mbar& b = sim::S().bars[bar];
    ok = b.init && b.phase != (parity & 1u);
  }
  if (!ok) std::this_thread::yield();
  return ok;
}
inline void commit(uint32_t bar) {
  sim::State& st = sim::S();
  for (const QueuedMma& q : t_queue) {
    if ((q.d >> 16) != 0 || (q.ts && (q.a_tmem >> 16) != 0)) sim::fail("tcgen05.mma: TMEM operand with a lane offset");
    const int d_col = (int)(q.d & 0xffff), a_col = q.ts ? (int)(q.a_tmem & 0xffff) : -1;
    if (!emu::mma(st.tmem, st.smem, sim::DYN_SMEM, d_col, a_col, q.a_desc, q.b_desc, q.idesc, q.acc != 0))
      sim::fail("tcgen05.mma: descriptor / shape rejected by the model");
    ++st.mma_count;
  }
  t_queue.clear();
  mbar_arrive(bar);
}
inline bool lane_rule(uint32_t addr, int* lane, int* col) {
  const int lane0 = (int)(addr >> 16), c = (int)(addr & 0xffff);
  const int warp = (int)(threadIdx.x >> 5);
  *lane = lane0 + (int)(threadIdx.x & 31);
  *col = c;
  if (lane0 != 32 * (warp & 3)) { sim::fail("tcgen05.ld/st: warp " + std::to_string(warp) + " addressed lane base " +
                                            std::to_string(lane0)); return false; }
  if (c < 0 || c + 8 > 512) { sim::fail("tcgen05.ld/st: column range " + std::to_string(c)); return false; }
  return true;
}
inline void tmem_ld8(uint32_t addr, uint32_t* r) {
  int lane, col;
  if (!lane_rule(addr, &lane, &col)) { for (int j = 0; j < 8; ++j) r[j] = 0x7fc00000u; return; }
  memcpy(r, &sim::S().tmem[lane * 512 + col], 32);
}
inline void tmem_st8(uint32_t addr, const uint32_t* r) {
  int lane, col;
  if (!lane_rule(addr, &lane, &col)) return;
  memcpy(&sim::S().tmem[lane * 512 + col], r, 32);
}
inline void wait_st() {}
inline void fence_before_thread_sync() {}
inline void fence_after_thread_sync() {}
inline void fence_mbar_init() {}
inline void fence_proxy_async_smem() {}
inline void tmem_alloc512(uint32_t* slot) {
  std::lock_guard<std::mutex> lk(sim::S().m);
  sim::S().tmem_allocated = true;
  *slot = 0u;
}
inline void tmem_dealloc512(uint32_t addr) {
  if (addr != 0u) sim::fail("tcgen05.dealloc with a foreign address");
  std::lock_guard<std::mutex> lk(sim::S().m);
  sim::S().tmem_allocated = false;
}
// slow clock: the kernels' 2e9-tick timeouts become ~200 s of wall time here
inline long long clock_now() {
  return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch())
             .count() / 100;
}

}  // namespace tcp
}  // namespace apg
