// Test-only software model of the PTX helpers of csrc/tile_engine.cuh (-DAPG_SIM): the unchanged sources of the
// tile-engine kernels (evaluation rollouts, split-adjoint dX chain, ...) run on the CPU, one OS thread per GPU thread
// (gpu_sim.h).
//   mbarrier          arrival count + transaction bytes + phase parity (expect_tx / complete_tx semantics)
//   cp.async.bulk     global -> shared: copied at issue, then complete_tx on the barrier;
//                     shared -> global (bulk groups): the copy is performed when the issuing thread WAITS for the
//                     group (wait_group.read / wait_group) or exits - the latest moment the hardware may read the
//                     source, so a tile that is overwritten before its store was waited for shows up as wrong data
//   mma.sync.m16n8k8  warp-collective: fragments exchanged through per-warp scratch, inputs truncated to TF32,
//                     fp32 result (fragment coordinates as documented in tile_engine.cuh)
//   red.global.add    atomic float add;  bar.sync 1, 256: a 256-thread barrier
#pragma once
#include "gpu_sim.h"

namespace simte {

inline unsigned char* dynamic_smem() { return sim::S().smem; }

inline void settle(sim::Mbar& b) {
  if (b.pending == 0 && b.tx == 0) { b.phase ^= 1u; b.pending = b.count; }
}
inline void mbar_init(uint64_t* bar, uint32_t count) {
  std::lock_guard<std::mutex> lk(sim::S().m);
  sim::Mbar& b = sim::S().pbars[bar];
  b.count = b.pending = (int)count; b.tx = 0; b.phase = 0; b.init = true;
}
inline void mbar_arrive(uint64_t* bar) {
  {
    std::lock_guard<std::mutex> lk(sim::S().m);
    sim::Mbar& b = sim::S().pbars[bar];
    if (!b.init) { sim::S().errors.push_back("arrive on an uninitialised mbarrier"); return; }
    --b.pending;
    settle(b);
  }
  sim::S().cv.notify_all();
}
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  {
    std::lock_guard<std::mutex> lk(sim::S().m);
    sim::Mbar& b = sim::S().pbars[bar];
    if (!b.init) { sim::S().errors.push_back("expect_tx on an uninitialised mbarrier"); return; }
    b.tx += bytes;
    --b.pending;
    settle(b);
  }
  sim::S().cv.notify_all();
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  sim::State& st = sim::S();
  std::unique_lock<std::mutex> lk(st.m);
  sim::Mbar& b = st.pbars[bar];
  if (!st.cv.wait_for(lk, std::chrono::seconds(120), [&] { return b.init && b.phase != (parity & 1u); })) {
    st.errors.push_back("mbar_wait timed out (the kernel would __trap)");
    lk.unlock();
    throw std::runtime_error("mbar_wait");
  }
}
inline void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  if ((bytes & 15u) || ((uintptr_t)dst_smem & 15u) || ((uintptr_t)src_gmem & 15u))
    sim::fail("cp.async.bulk g2s: size / address not 16-byte aligned");
  memcpy(dst_smem, src_gmem, bytes);
  {
    std::lock_guard<std::mutex> lk(sim::S().m);
    sim::Mbar& b = sim::S().pbars[bar];
    b.tx -= bytes;
    settle(b);
  }
  sim::S().cv.notify_all();
}
struct PendingStore { void* dst; const void* src; uint32_t bytes; };
inline thread_local std::vector<std::vector<PendingStore>> t_groups;      // committed groups, oldest first
inline thread_local std::vector<PendingStore> t_open;
inline void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  if ((bytes & 15u) || ((uintptr_t)dst_gmem & 15u) || ((uintptr_t)src_smem & 15u))
    sim::fail("cp.async.bulk s2g: size / address not 16-byte aligned");
  t_open.push_back({dst_gmem, src_smem, bytes});
}
inline void bulk_commit() { t_groups.push_back(t_open); t_open.clear(); }
// all but the newest `keep` groups have been read (and, here, written)
inline void bulk_wait(int keep) {
  while ((int)t_groups.size() > keep) {
    for (const PendingStore& p : t_groups.front()) memcpy(p.dst, p.src, p.bytes);
    t_groups.erase(t_groups.begin());
  }
}
inline void thread_exit() {
  if (!t_open.empty()) bulk_commit();
  bulk_wait(0);
}
inline void named_barrier_256() { sim::S().named_barrier->arrive_and_wait(); }
inline void named_barrier_64() { sim::S().named_barrier64->arrive_and_wait(); }
inline void red_add(float* addr, float v) {
  std::atomic_ref<float> a(*addr);
  a.fetch_add(v, std::memory_order_relaxed);
}
// D(16x8) = A(16x8) * B(8x8) + C;  lane = 4*g + t:
//   a0 (g,t) a1 (g+8,t) a2 (g,t+4) a3 (g+8,t+4);  b0 (k=t,n=g) b1 (k=t+4,n=g);  c0 (g,2t) c1 (g,2t+1) c2 (g+8,2t) c3 (g+8,2t+1)
struct WarpFrag { uint32_t a[32][4]; uint32_t b[32][2]; };
inline std::vector<WarpFrag>& frags() { static std::vector<WarpFrag> f(64); return f; }
inline float tf32(uint32_t u) { u &= 0xffffe000u; float f; memcpy(&f, &u, 4); return f; }
inline void mma_m16n8k8_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  sim::State& st = sim::S();
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, g = l >> 2, t = l & 3;
  WarpFrag& F = frags()[w];
  for (int i = 0; i < 4; ++i) F.a[l][i] = a[i];
  F.b[l][0] = b0; F.b[l][1] = b1;
  st.warp_barrier[w]->arrive_and_wait();
  auto A = [&](int r, int k) { return tf32(F.a[(r & 7) * 4 + (k & 3)][(r >> 3) + 2 * (k >> 2)]); };
  auto B = [&](int k, int n) { return tf32(F.b[n * 4 + (k & 3)][k >> 2]); };
  float out[4];
  for (int e = 0; e < 4; ++e) {
    const int r = g + 8 * (e >> 1), n = 2 * t + (e & 1);
    double s = 0.0;
    for (int k = 0; k < 8; ++k) s += (double)A(r, k) * (double)B(k, n);
    out[e] = (float)((double)c[e] + s);
  }
  st.warp_barrier[w]->arrive_and_wait();
  for (int e = 0; e < 4; ++e) c[e] = out[e];
}

inline void st_release(unsigned* p, unsigned v) { std::atomic_ref<unsigned>(*p).store(v); }
inline unsigned ld_acquire(const unsigned* p) { return std::atomic_ref<unsigned>(*const_cast<unsigned*>(p)).load(); }
inline unsigned atomic_inc(unsigned* p) { return std::atomic_ref<unsigned>(*p).fetch_add(1u); }
inline long long slow_clock() {
  return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch())
             .count() / 1000;
}

struct Install { Install() { sim::S().on_thread_exit = thread_exit; } };
inline Install install_hooks;

}  // namespace simte
