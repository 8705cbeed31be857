// tcgen05 / TMEM forward kernel of the quadrotor CONCURRENT rollout (Net(15,10,9,40,conv), h = 10):
// policy forward on the 5th-generation tensor cores with the accumulators and the activations in TMEM, then the h
// dynamics steps + tracking loss by the thread that owns the drone, straight from its tcgen05.ld registers.
// Writes the SAME activation / action / state stash as hutter_fwd_kernel, so hutter_adj_kernel consumes it unchanged.
//
// OPTIONAL PATH: selected by APG_TC_FWD=1 (capi.cu); the default forward stays hutter_fwd_kernel until this kernel
// has passed the GPU parity tests (it was written after the round-1 GPU budget was spent; the index arithmetic is
// host-checked, tests/test_tc_layout_host.py).  Prototype with timing: tools/micro/tcgen05_policy.cu.
//
// Roles (288 threads): warps 0-3 / 4-7 = epilogue groups of TMEM slot 0 / 1 (two tiles in flight, thread r of a
// group owns TMEM lane r = drone r of the tile); warp 8 lane 0 issues every tcgen05.mma.  Hand-off by mbarriers:
// a_ready[s] (128 arrivals: the A operand of the next op is in TMEM), d_ready[s] (tcgen05.commit: the op is done).
#include "tc_layout.cuh"
#include "tc_prims.cuh"
#include "rollout_args.h"
#ifndef APG_TC_SIM
#include "tile_engine.cuh"
#endif
#include "kernels.h"

namespace apg {

using namespace tc;

namespace {

constexpr int TC_THREADS = 288;
constexpr int TC_SMEM_BYTES = 1024 + BLOB_BYTES;
static_assert(TC_SMEM_BYTES <= 232448 - 512, "weight images do not fit in shared memory");

using tcp::mma_ts;
__device__ __forceinline__ void mma_commit(uint32_t bar) { tcp::commit(bar); }
__device__ __forceinline__ void tc_mbar_init(uint32_t bar, int count) { tcp::mbar_init(bar, count); }
__device__ __forceinline__ void tc_mbar_arrive(uint32_t bar) { tcp::mbar_arrive(bar); }
// bounded wait: a protocol error must end the launch (wrong results are caught by the parity tests), never hang
// the GPU.  A wait that lasts longer than ~1 s of SM clocks sets the CTA's abort flag; from then on every wait of
// the CTA returns at once and the CTA reports a NaN loss.
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag) {
  const long long t0 = tcp::clock_now();
  for (int spin = 0;; ++spin) {
    if (tcp::mbar_try_wait(bar, parity)) return;
    if ((spin & 63) == 63) {
      if (*abort_flag) return;
      if (tcp::clock_now() - t0 > 2000000000LL) { *abort_flag = 1; return; }
    }
  }
}
// non-blocking phase test
__device__ __forceinline__ bool tc_mbar_test(uint32_t bar, uint32_t parity) { return tcp::mbar_test_wait(bar, parity); }
__device__ __forceinline__ void tmem_ld8(uint32_t addr, float* v) {
  uint32_t r[8];
  tcp::tmem_ld8(addr, r);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}
// split 8 values into (hi, lo) and store them into the A-operand columns [col, col + 8)
__device__ __forceinline__ void tmem_st8_split(uint32_t a_hi, uint32_t a_lo, const float* x) {
  uint32_t h[8], l[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    h[j] = __float_as_uint(x[j]) & 0xffffe000u;
    l[j] = __float_as_uint(x[j] - __uint_as_float(h[j]));
  }
  tcp::tmem_st8(a_hi, h);
  tcp::tmem_st8(a_lo, l);
}
__device__ __forceinline__ void a_operand_ready(uint32_t bar) {
  tcp::wait_st();
  tcp::fence_before_thread_sync();
  tc_mbar_arrive(bar);
}

struct TcBars {
  unsigned long long a_ready[2];
  unsigned long long d_ready[2];
};

}  // namespace

// weights (torch-flat) -> (hi, lo) K-major images + bias block, once per forward call
__global__ void apg_pack_tc_kernel(const float* __restrict__ params, const HutterLayout y,
                                   unsigned char* __restrict__ blob) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < PAIRS_TOTAL + B_TOTAL) pack_body(e, params, y, blob);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
    hutter_fwd_tc_kernel(const unsigned char* __restrict__ blob, const HutterLayout y, const RolloutArgs g) {
  APG_TC_DYNAMIC_SMEM(smem_raw);
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const float* s_bias = (const float*)(base + IMG_TOTAL);
  __shared__ __align__(8) TcBars s_bars;
  __shared__ uint32_t s_tmem;
  __shared__ int s_abort;
  __shared__ float s_red[8];
  using Sys = Quad<float>;
  constexpr int S = Sys::S, A = Sys::A, R = Sys::REFW;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < BLOB_BYTES / 16; i += blockDim.x) ((uint4*)base)[i] = ((const uint4*)blob)[i];
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      tc_mbar_init(smem_u32(&s_bars.a_ready[s]), 128);
      tc_mbar_init(smem_u32(&s_bars.d_ready[s]), 1);
    }
    s_abort = 0;
    tcp::fence_mbar_init();
  }
  tcp::fence_proxy_async_smem();     // generic-proxy image writes -> tensor core reads
  if (warp == 8) {
    tcp::tmem_alloc512(&s_tmem);
  }
  tcp::fence_before_thread_sync();
  __syncthreads();
  tcp::fence_after_thread_sync();
  const uint32_t tmem = s_tmem;
  const int n = g.N;
  const int ntiles = (n + TMT - 1) / TMT;                     // tcgen05 tiles of 128 drones
  const int ntiles64 = (n + TM - 1) / TM;                     // stash tiles of 64 drones (what the adjoint walks)
  // this CTA's tiles: blockIdx.x + j * gridDim.x, j = 0, 1, ...; tile j runs in slot j & 1
  const int my_tiles = (ntiles > (int)blockIdx.x) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  volatile int* abort_flag = &s_abort;
  float my_loss = 0.f;

  if (warp == 8) {
    if (lane == 0) {
      // The two slots advance independently: whichever slot has its next A operand ready gets its next op issued
      // (non-blocking mbarrier test), so one group's dynamics phase does not hold back the other group's GEMMs.
      uint32_t par[2] = {0, 0};
      int op_i[2] = {0, 0}, tile_j[2] = {0, 1};
      int remaining = my_tiles * NOPS;
      long long t_idle = tcp::clock_now();
      while (remaining > 0) {
        bool progressed = false;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (tile_j[s] >= my_tiles) continue;
          if (!tc_mbar_test(smem_u32(&s_bars.a_ready[s]), par[s])) continue;
          par[s] ^= 1;
          tcp::fence_after_thread_sync();
          const Op op = op_of(op_i[s]);
          const uint32_t idesc = idesc_tf32(TMT, op.N);
          const uint32_t whi = smem_u32(base + op.img_off), wlo = whi + img_bytes(op.rows, op.K);
          const uint32_t slot = tmem + s * SLOT_COLS;
          const uint32_t d = slot + op.d_col, ahi = slot + C_AHI, alo = slot + C_ALO;
          for (int ks = 0; ks < op.K / 8; ++ks) {
            const uint64_t bh = kmajor_desc(whi, ks, op.K), bl = kmajor_desc(wlo, ks, op.K);
            mma_ts(d, alo + ks * 8, bh, idesc, (ks > 0 || !op.clear) ? 1u : 0u);
            mma_ts(d, ahi + ks * 8, bl, idesc, 1u);
            mma_ts(d, ahi + ks * 8, bh, idesc, 1u);
          }
          mma_commit(smem_u32(&s_bars.d_ready[s]));
          if (++op_i[s] == NOPS) { op_i[s] = 0; tile_j[s] += 2; }
          --remaining;
          progressed = true;
        }
        if (progressed) {
          t_idle = tcp::clock_now();
        } else if (*abort_flag || tcp::clock_now() - t_idle > 2000000000LL) {
          *abort_flag = 1;                                    // protocol error: give up, the CTA reports NaN
          break;
        }
      }
    }
  } else {
    const int s = warp >> 2;
    const int row = (warp & 3) * 32 + lane;                  // TMEM lane = drone of the tile
    const uint32_t slot = tmem + s * SLOT_COLS + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t d_main = slot + C_DMAIN, d_conv = slot + C_DCONV, ahi = slot + C_AHI, alo = slot + C_ALO;
    const uint32_t bar_a = smem_u32(&s_bars.a_ready[s]), bar_d = smem_u32(&s_bars.d_ready[s]);
    uint32_t par = 0;
    auto wait_d = [&]() {
      tc_mbar_wait(bar_d, par, abort_flag);
      par ^= 1;
      tcp::fence_after_thread_sync();
    };
    for (int j = s; j < my_tiles; j += 2) {
      const int tile = (int)blockIdx.x + j * (int)gridDim.x;
      const size_t drone = (size_t)tile * TMT + row;
      const bool live = drone < (size_t)n;
      const bool stash = tile * 2 + (row >> 6) < ntiles64;    // this drone's 64-tile exists in the stash
      // D_main (64 columns) -> tanh(x + b) -> A operand; the activation also goes to the stash (rows row0 + c)
      auto dense_epilogue = [&](const float* b, float* st, int rows, int row0) {
#pragma unroll
        for (int c0 = 0; c0 < HID; c0 += 8) {
          float v[8];
          tmem_ld8(d_main + c0, v);
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = act_apply(v[q] + b[c0 + q], ACT_TANH);
          tmem_st8_split(ahi + c0, alo + c0, v);
          if (stash) {
#pragma unroll
            for (int q = 0; q < 8; ++q) st[stash_index(tile, row, rows, row0 + c0 + q)] = v[q];
          }
        }
        a_operand_ready(bar_a);
      };
      // op 0 operand: in_state, K = 16 (column 15 zero)
      {
        float x[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = (live && k < F0) ? g.in_state[drone * F0 + k] : 0.f;
        tmem_st8_split(ahi, alo, x);
        tmem_st8_split(ahi + 8, alo + 8, x + 8);
        a_operand_ready(bar_a);
      }
      wait_d();                                               // op 0: states_in
      dense_epilogue(s_bias + B_S, g.st_x1, K1, 0);           // s -> X1 rows [0, 64), operand of op 1
      const float* rr = g.in_ref + drone * REFW;
      for (int gq = 0; gq < 4; ++gq) {
        wait_d();        // op 1 (gq = 0) or the fc1 piece of the previous pair: the A columns are free again
        {
          float x[40];
#pragma unroll
          for (int k = 0; k < 36; k += 2) {
            const float2 t = live ? *(const float2*)(rr + 18 * gq + k) : make_float2(0.f, 0.f);
            x[k] = t.x;
            x[k + 1] = t.y;
          }
          x[36] = x[37] = x[38] = x[39] = 0.f;
#pragma unroll
          for (int c0 = 0; c0 < 40; c0 += 8) tmem_st8_split(ahi + c0, alo + c0, x + c0);
          a_operand_ready(bar_a);
        }
        wait_d();        // conv of this position pair
        {
          const float* b = s_bias + B_C;
#pragma unroll
          for (int c0 = 0; c0 < 40; c0 += 8) {
            float v[8];
            tmem_ld8(d_conv + c0, v);
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = act_apply(v[q] + b[c0 + q], ACT_RELU);
            tmem_st8_split(ahi + c0, alo + c0, v);
            if (stash) {
#pragma unroll
              for (int q = 0; q < 8; ++q) g.st_x1[stash_index(tile, row, K1, x1_row_of_conv(gq, c0 + q))] = v[q];
            }
          }
          a_operand_ready(bar_a);
        }
      }
      wait_d();                                               // last fc1 piece
      dense_epilogue(s_bias + B_1, g.st_h1, HID, 0);
      wait_d();                                               // fc2
      dense_epilogue(s_bias + B_2, g.st_h2, HID, 0);
      wait_d();                                               // fc3
      dense_epilogue(s_bias + B_3, g.st_h3, HID, 0);
      wait_d();                                               // fc_out
      float act[MO];
      {
        const float* b = s_bias + B_O;
#pragma unroll
        for (int c0 = 0; c0 < MO; c0 += 8) {
          float v[8];
          tmem_ld8(d_main + c0, v);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            act[c0 + q] = act_apply(v[q] + b[c0 + q], ACT_SIGMOID);       // train_base.py:203
            if (stash) g.st_act[stash_index(tile, row, MO, c0 + q)] = act[c0 + q];
          }
        }
      }
      // every tcgen05.ld of this tile has completed (wait::ld inside tmem_ld8): the slot's D columns may be
      // overwritten by the next tile's op 0 as soon as this group signals its next A operand.
      // ---- h dynamics steps + tracking loss of this drone (train_drone.py:175-203), states to the stash
      if (live) {
        float sc[S], s0[S], sn[S], rf[R];
        const float* cur_g = g.cur + drone * S;
        const float* ref_g = g.ref + drone * g.ref_rows * R;
#pragma unroll
        for (int q = 0; q < S; ++q) s0[q] = sc[q] = cur_g[q];
#pragma unroll
        for (int k = 0; k < H; ++k) {
          const float* a = act + k * A;
#pragma unroll
          for (int c = 0; c < R; ++c) rf[c] = ref_g[k * R + c];
          Sys::step(sc, a, g.dt, g.pc.v, sn);
          my_loss += Sys::loss(sn, rf, a, s0, k, H);
#pragma unroll
          for (int q = 0; q < S; ++q) {
            sc[q] = sn[q];
            g.st_states[stash_index(tile, row, H * S, k * S + q)] = sn[q];
          }
          if (g.states_out) {
#pragma unroll
            for (int q = 0; q < S; ++q) g.states_out[(drone * H + k) * S + q] = sn[q];
          }
          if (g.actions_out) {
#pragma unroll
            for (int c = 0; c < A; ++c) g.actions_out[(drone * H + k) * A + c] = a[c];
          }
        }
      }
    }
  }
  // ---- loss of this CTA: fixed-order sum over the 256 epilogue threads
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) my_loss += __shfl_xor_sync(0xffffffffu, my_loss, o);
  if (lane == 0 && warp < 8) s_red[warp] = my_loss;
  tcp::fence_before_thread_sync();
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_red[w];
    // a protocol timeout poisons the loss on purpose: the caller must never take such a launch for a result
    g.loss_partials[blockIdx.x] = s_abort ? __int_as_float(0x7fc00000) : t;
  }
  if (warp == 8) {
    tcp::tmem_dealloc512(tmem);
  }
}

// =========================================================================================================
// tcgen05 dX chain of the split adjoint (OPTIONAL PATH, APG_TC_DX=1 on top of APG_TC_DW=1): the tcgen05 counterpart
// of hutter_adj_dx_kernel.  Same roles and hand-off protocol as hutter_fwd_tc_kernel; the A operands are the dZ_l
// written to TMEM by the owning threads, the B operands are the SAME forward weight images read MN-major: the
// unswizzled K-major image of W[out][in] is byte for byte the unswizzled MN-major image of W^T (CUTLASS canonical
// forms, cute/atom/mma_traits_sm100.hpp; microbenchmark tools/micro/tcgen05_gemm.cu variants 7 / 8), so
//     dX[drone][in] = sum_out dZ[drone][out] W[out][in]
// needs b_major = MN, LBO = (K_f / 4) * 128 (8-row groups of the image = k groups), SBO = 128 (adjacent 16-byte
// chunks = adjacent groups of four `in` features) and start address + ks * LBO.  Reads the activation stash, writes
// the dZ stash adj_dw_tc_kernel consumes.
// =========================================================================================================
__global__ void __launch_bounds__(TC_THREADS, 1)
    hutter_adj_dx_tc_kernel(const unsigned char* __restrict__ blob, const HutterLayout y, const RolloutArgs g,
                            const DzStash z) {
  APG_TC_DYNAMIC_SMEM(smem_raw);
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) TcBars s_bars;
  __shared__ uint32_t s_tmem;
  __shared__ int s_abort;
  using Sys = Quad<float>;
  constexpr int S = Sys::S, A = Sys::A, R = Sys::REFW;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < BLOB_BYTES / 16; i += blockDim.x) ((uint4*)base)[i] = ((const uint4*)blob)[i];
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      tc_mbar_init(smem_u32(&s_bars.a_ready[s]), 128);
      tc_mbar_init(smem_u32(&s_bars.d_ready[s]), 1);
    }
    s_abort = 0;
    tcp::fence_mbar_init();
  }
  tcp::fence_proxy_async_smem();
  if (warp == 8) {
    tcp::tmem_alloc512(&s_tmem);
  }
  tcp::fence_before_thread_sync();
  __syncthreads();
  tcp::fence_after_thread_sync();
  const uint32_t tmem = s_tmem;
  const int n = g.N;
  const int ntiles = (n + TMT - 1) / TMT;
  const int ntiles64 = (n + TM - 1) / TM;
  const int my_tiles = (ntiles > (int)blockIdx.x) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  volatile int* abort_flag = &s_abort;

  if (warp == 8) {
    if (lane == 0) {
      uint32_t par[2] = {0, 0};
      int op_i[2] = {0, 0}, tile_j[2] = {0, 1};
      int remaining = my_tiles * NROPS;
      long long t_idle = tcp::clock_now();
      while (remaining > 0) {
        bool progressed = false;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (tile_j[s] >= my_tiles) continue;
          if (!tc_mbar_test(smem_u32(&s_bars.a_ready[s]), par[s])) continue;
          par[s] ^= 1;
          tcp::fence_after_thread_sync();
          const ROp op = rop_of(op_i[s]);
          const uint32_t idesc = idesc_tf32(TMT, op.N, 1);
          const uint32_t whi = smem_u32(base + op.img_off), wlo = whi + img_bytes(op.rows, op.Kf);
          const uint32_t slot = tmem + s * SLOT_COLS;
          const uint32_t d = slot + op.d_col, ahi = slot + C_AHI, alo = slot + C_ALO;
          for (int ks = 0; ks < op.K / 8; ++ks) {
            const uint64_t bh = mnmajor_desc(whi, ks, op.Kf), bl = mnmajor_desc(wlo, ks, op.Kf);
            mma_ts(d, alo + ks * 8, bh, idesc, ks > 0 ? 1u : 0u);
            mma_ts(d, ahi + ks * 8, bl, idesc, 1u);
            mma_ts(d, ahi + ks * 8, bh, idesc, 1u);
          }
          mma_commit(smem_u32(&s_bars.d_ready[s]));
          if (++op_i[s] == NROPS) { op_i[s] = 0; tile_j[s] += 2; }
          --remaining;
          progressed = true;
        }
        if (progressed) {
          t_idle = tcp::clock_now();
        } else if (*abort_flag || tcp::clock_now() - t_idle > 2000000000LL) {
          *abort_flag = 1;
          break;
        }
      }
    }
  } else {
    const int s = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t slot = tmem + s * SLOT_COLS + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t d_main = slot + C_DMAIN, d_conv = slot + C_DCONV, ahi = slot + C_AHI, alo = slot + C_ALO;
    const uint32_t bar_a = smem_u32(&s_bars.a_ready[s]), bar_d = smem_u32(&s_bars.d_ready[s]);
    const float poison = __int_as_float(0x7fc00000);
    uint32_t par = 0;
    auto wait_d = [&]() {
      tc_mbar_wait(bar_d, par, abort_flag);
      par ^= 1;
      tcp::fence_after_thread_sync();
    };
    for (int j = s; j < my_tiles; j += 2) {
      const int tile = (int)blockIdx.x + j * (int)gridDim.x;
      const size_t drone = (size_t)tile * TMT + row;
      const bool live = drone < (size_t)n;
      const bool stash = tile * 2 + (row >> 6) < ntiles64;
      // ---- reverse dynamics sweep of this drone (dyn_phase.cuh dyn_adjoint_conc, on the stash layout): dlog[40]
      float dlog[MO];
#pragma unroll
      for (int q = 0; q < MO; ++q) dlog[q] = 0.f;
      if (live) {
        float s0[S], sk[S], sn[S], a[A], rf[R], gq[S], gs[S], ga[A], ga2[A];
        const float* cur_g = g.cur + drone * S;
        const float* ref_g = g.ref + drone * g.ref_rows * R;
#pragma unroll
        for (int q = 0; q < S; ++q) { s0[q] = cur_g[q]; gq[q] = 0.f; }
#pragma unroll
        for (int q = 0; q < S; ++q) sn[q] = g.st_states[stash_index(tile, row, H * S, (H - 1) * S + q)];
#pragma unroll
        for (int k = H - 1; k >= 0; --k) {
#pragma unroll
          for (int c = 0; c < A; ++c) { a[c] = g.st_act[stash_index(tile, row, MO, k * A + c)]; ga[c] = 0.f; }
#pragma unroll
          for (int c = 0; c < R; ++c) rf[c] = ref_g[k * R + c];
          if (k > 0) {
#pragma unroll
            for (int q = 0; q < S; ++q) sk[q] = g.st_states[stash_index(tile, row, H * S, (k - 1) * S + q)];
          } else {
#pragma unroll
            for (int q = 0; q < S; ++q) sk[q] = s0[q];
          }
          Sys::loss_grad(sn, rf, a, s0, k, H, gq, ga);
          Sys::step_adj(sk, a, g.dt, g.pc.v, gq, gs, ga2);
#pragma unroll
          for (int c = 0; c < A; ++c) dlog[k * A + c] = (ga[c] + ga2[c]) * a[c] * (1.f - a[c]);      // sigmoid'
#pragma unroll
          for (int q = 0; q < S; ++q) { gq[q] = gs[q]; sn[q] = sk[q]; }
        }
      }
      if (stash) {
#pragma unroll
        for (int q = 0; q < MO; ++q) z.o[stash_index(tile, row, MO, q)] = dlog[q];
      }
#pragma unroll
      for (int c0 = 0; c0 < MO; c0 += 8) tmem_st8_split(ahi + c0, alo + c0, dlog + c0);
      a_operand_ready(bar_a);
      // ---- dZ_l = (dZ_{l+1} W_{l+1}) (.) (1 - X_l^2) for h3, h2, h1: D_main -> A operand + dZ stash
      const float* xs[3] = {g.st_h3, g.st_h2, g.st_h1};
      float* zs[3] = {z.z3, z.z2, z.z1};
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        wait_d();
#pragma unroll
        for (int c0 = 0; c0 < HID; c0 += 8) {
          float v[8];
          tmem_ld8(d_main + c0, v);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float yv = stash ? xs[l][stash_index(tile, row, HID, c0 + q)] : 0.f;
            v[q] *= 1.f - yv * yv;
            if (stash) zs[l][stash_index(tile, row, HID, c0 + q)] = *abort_flag ? poison : v[q];
          }
          tmem_st8_split(ahi + c0, alo + c0, v);
        }
        a_operand_ready(bar_a);
      }
      // ---- first layer: ds = (dZ1 W1[:, :64]) (.) (1 - s^2); the A operand (dZ1) stays for the conv pieces
      wait_d();
#pragma unroll
      for (int c0 = 0; c0 < HID; c0 += 8) {
        float v[8];
        tmem_ld8(d_main + c0, v);
        if (stash) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float yv = g.st_x1[stash_index(tile, row, K1, c0 + q)];
            z.x[stash_index(tile, row, K1, c0 + q)] = *abort_flag ? poison : v[q] * (1.f - yv * yv);
          }
        }
      }
      tcp::fence_before_thread_sync();
      tc_mbar_arrive(bar_a);                                   // D_main has been read: go on with the conv pieces
#pragma unroll
      for (int gp = 0; gp < 4; ++gp) {
        wait_d();
#pragma unroll
        for (int c0 = 0; c0 < 40; c0 += 8) {
          float v[8];
          tmem_ld8(d_conv + c0, v);
          if (stash) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int xr = x1_row_of_conv(gp, c0 + q);
              const float yv = g.st_x1[stash_index(tile, row, K1, xr)];
              z.x[stash_index(tile, row, K1, xr)] = *abort_flag ? poison : (yv > 0.f ? v[q] : 0.f);      // relu'
            }
          }
        }
        if (gp < 3) {
          tcp::fence_before_thread_sync();
          tc_mbar_arrive(bar_a);                               // D_conv is free for the next position pair
        }
      }
    }
  }
  tcp::fence_before_thread_sync();
  __syncthreads();
  if (warp == 8) {
    tcp::tmem_dealloc512(tmem);
  }
}

cudaError_t launch_hutter_adj_dx_tc(const HutterLayout& y, const float* params, unsigned char* blob,
                                    const RolloutArgs& a, const DzStash& z, int grid, cudaStream_t st) {
  cudaError_t e = cudaSuccess;
  if (params) {                                  // nullptr: the images of the tcgen05 forward call are still valid
    APG_LAUNCH((PAIRS_TOTAL + B_TOTAL + 255) / 256, 256, 0, st, apg_pack_tc_kernel)(params, y, blob);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  e = cudaFuncSetAttribute(hutter_adj_dx_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  APG_LAUNCH(grid, TC_THREADS, TC_SMEM_BYTES, st, hutter_adj_dx_tc_kernel)(blob, y, a, z);
  return cudaGetLastError();
}

size_t tc_blob_bytes() { return (size_t)BLOB_BYTES; }

bool tc_fwd_supported(const HutterLayout& y, int h) {
  return y.conv && y.F0 == F0 && y.L == H && y.RD == RD && y.Mo == MO && h == H;
}

cudaError_t launch_hutter_fwd_tc(const HutterLayout& y, const float* params, unsigned char* blob,
                                 const RolloutArgs& a, int grid, cudaStream_t st) {
  APG_LAUNCH((PAIRS_TOTAL + B_TOTAL + 255) / 256, 256, 0, st, apg_pack_tc_kernel)(params, y, blob);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(hutter_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  APG_LAUNCH(grid, TC_THREADS, TC_SMEM_BYTES, st, hutter_fwd_tc_kernel)(blob, y, a);
  return cudaGetLastError();
}

}  // namespace apg
