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
#include "gpu_sim.h"

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
inline thread_local std::vector<QueuedMma> t_queue;

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
inline void settle(sim::Mbar& b) {
  if (b.pending == 0 && b.tx == 0) { b.phase ^= 1u; b.pending = b.count; }
}
inline void mbar_arrive(uint32_t bar) {
  sim::State& st = sim::S();
  bool bad = false;
  {
    std::lock_guard<std::mutex> lk(st.m);
    sim::Mbar& b = st.bars[bar];
    if (!b.init) bad = true;
    else { --b.pending; settle(b); }
  }
  if (bad) sim::fail("arrive on an uninitialised mbarrier");
  st.cv.notify_all();
}
// arrive + expected transaction bytes; the phase completes when the arrivals AND the bytes are in
inline void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  sim::State& st = sim::S();
  bool bad = false;
  {
    std::lock_guard<std::mutex> lk(st.m);
    sim::Mbar& b = st.bars[bar];
    if (!b.init) bad = true;
    else { b.tx += bytes; --b.pending; settle(b); }
  }
  if (bad) sim::fail("expect_tx on an uninitialised mbarrier");
  st.cv.notify_all();
}
// cp.async.bulk global -> shared.  Default: copied at issue, then complete_tx on the barrier.
// APG_SIM_BULK_DELAY_US=<max>: the copy LANDS LATER - after a pseudo-random delay of up to <max> microseconds, different
// for every copy, so that copies complete out of order like on hardware - and until then its destination holds NaN
// patterns.  A thread that passes a wait it should not have passed (a parity wait that took the phase before for the
// one it needs, a stage handed on too early) then reads NaNs, and the result shows it.  Pending copies are completed
// by whichever thread next polls or signals a barrier.
struct PendingBulk { uint32_t dst; const void* src; uint32_t bytes, bar; long long due_ns; };
inline std::vector<PendingBulk>& pending_bulk() { static std::vector<PendingBulk> v; return v; }
inline long long sim_now_ns() {
  return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline long long bulk_delay_max_ns() {
  static const long long v = getenv("APG_SIM_BULK_DELAY_US") ? atoll(getenv("APG_SIM_BULK_DELAY_US")) * 1000LL : 0LL;
  return v;
}
// (st.m held)
inline void land_due_copies_locked(sim::State& st, bool all) {
  std::vector<PendingBulk>& v = pending_bulk();
  if (v.empty()) return;
  const long long now = sim_now_ns();
  bool any = false;
  for (size_t i = 0; i < v.size();) {
    if (all || v[i].due_ns <= now) {
      memcpy(st.smem + v[i].dst, v[i].src, v[i].bytes);
      sim::Mbar& b = st.bars[v[i].bar];
      b.tx -= v[i].bytes;
      settle(b);
      v[i] = v.back();
      v.pop_back();
      any = true;
    } else {
      ++i;
    }
  }
  if (any) st.cv.notify_all();
}
inline void land_due_copies() {
  if (!bulk_delay_max_ns()) return;
  sim::State& st = sim::S();
  std::lock_guard<std::mutex> lk(st.m);
  land_due_copies_locked(st, false);
}
inline void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  sim::State& st = sim::S();
  if ((bytes & 15u) || (dst_smem & 15u) || ((uintptr_t)src & 15u)) sim::fail("cp.async.bulk: size / address not 16-byte aligned");
  if ((dst_smem >> 30) != 0 || (size_t)dst_smem + bytes > sim::DYN_SMEM) { sim::fail("cp.async.bulk: destination outside the dynamic shared memory"); return; }
  if (bulk_delay_max_ns()) {
    static std::atomic<uint64_t> lcg{0x9e3779b97f4a7c15ull};
    const uint64_t x = lcg.fetch_add(0x9e3779b97f4a7c15ull) * 0xbf58476d1ce4e5b9ull;
    std::lock_guard<std::mutex> lk(st.m);
    memset(st.smem + dst_smem, 0xff, bytes);            // in flight: the destination is garbage
    pending_bulk().push_back({dst_smem, src, bytes, bar, sim_now_ns() + (long long)((x >> 20) % (uint64_t)bulk_delay_max_ns())});
    return;
  }
  memcpy(st.smem + dst_smem, src, bytes);
  {
    std::lock_guard<std::mutex> lk(st.m);
    sim::Mbar& b = st.bars[bar];
    b.tx -= bytes;
    settle(b);
  }
  st.cv.notify_all();
}
inline void prefetch_l2(const void*) {}
inline void griddep_wait() {}
inline void griddep_launch() {}
inline bool elect_one() { return (threadIdx.x & 31) == 0; }
// true once the phase with the given parity has completed (PTX mbarrier.try_wait.parity)
inline bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  sim::State& st = sim::S();
  std::unique_lock<std::mutex> lk(st.m);
  if (bulk_delay_max_ns()) land_due_copies_locked(st, false);
  sim::Mbar& b = st.bars[bar];
  return st.cv.wait_for(lk, std::chrono::microseconds(500), [&] { return b.init && b.phase != (parity & 1u); });
}
inline bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  bool ok;
  {
    std::lock_guard<std::mutex> lk(sim::S().m);
    if (bulk_delay_max_ns()) land_due_copies_locked(sim::S(), false);
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
inline bool lane_rule(uint32_t addr, int* lane, int* col, int ncol = 8) {
  const int lane0 = (int)(addr >> 16), c = (int)(addr & 0xffff);
  const int warp = (int)(threadIdx.x >> 5);
  *lane = lane0 + (int)(threadIdx.x & 31);
  *col = c;
  if (lane0 != 32 * (warp & 3)) { sim::fail("tcgen05.ld/st: warp " + std::to_string(warp) + " addressed lane base " +
                                            std::to_string(lane0)); return false; }
  if (c < 0 || c + ncol > 512) { sim::fail("tcgen05.ld/st: column range " + std::to_string(c)); return false; }
  return true;
}
inline void tmem_ldn(uint32_t addr, uint32_t* r, int ncol) {
  int lane, col;
  if (!lane_rule(addr, &lane, &col, ncol)) { for (int j = 0; j < ncol; ++j) r[j] = 0x7fc00000u; return; }
  memcpy(r, &sim::S().tmem[lane * 512 + col], 4 * ncol);
}
inline void tmem_stn(uint32_t addr, const uint32_t* r, int ncol) {
  int lane, col;
  if (!lane_rule(addr, &lane, &col, ncol)) return;
  memcpy(&sim::S().tmem[lane * 512 + col], r, 4 * ncol);
}
inline void tmem_ld8(uint32_t addr, uint32_t* r) { tmem_ldn(addr, r, 8); }
inline void tmem_st8(uint32_t addr, const uint32_t* r) { tmem_stn(addr, r, 8); }
inline void tmem_ld16(uint32_t addr, uint32_t* r) { tmem_ldn(addr, r, 16); }
inline void tmem_st16(uint32_t addr, const uint32_t* r) { tmem_stn(addr, r, 16); }
inline void tmem_ld32(uint32_t addr, uint32_t* r) { tmem_ldn(addr, r, 32); }
inline void tmem_st32(uint32_t addr, const uint32_t* r) { tmem_stn(addr, r, 32); }
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
// (APG_SIM_FAST_TIMEOUT=1: ~5 s, for debugging a protocol dead-lock)
inline long long clock_now() {
  static const long long div = getenv("APG_SIM_FAST_TIMEOUT") ? 2 : 100;
  return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch())
             .count() / div;
}

}  // namespace tcp
}  // namespace apg
