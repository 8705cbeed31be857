// Fused rollout kernels for the "hutter" policy MLP (models/hutter_model.py) in CONCURRENT mode:
//   quadrotor  : Net(15, 10, 9, 4h, conv=True )  (scripts/train_drone.py:82-88,  175-203)
//   fixed wing : Net( 9,  1, 3, 4h, conv=False)  (scripts/train_fixed_wing.py:67-73, 90-116)
//
// hutter_fwd_kernel  one persistent CTA per SM (8 GEMM warps + 2 dynamics warps); per tile of 64 drones: TMA-staged
//                    input tiles -> policy forward on the tensor path (all weights resident in shared memory,
//                    [in][out] packing) -> sigmoid -> h dynamics steps -> tracking loss.  Activations / actions /
//                    states are stashed tile-major for the adjoint.
// hutter_adj_kernel  replays the horizon in reverse (hand-written adjoint of the dynamics, no autograd tape),
//                    back-propagates through the MLP ([out][in] weights resident) and accumulates the weight
//                    gradient of its tiles into a per-CTA partial (one owner thread per entry -> deterministic;
//                    reduced over CTAs in a fixed order by apg_reduce_kernel).
#include "dyn_phase.cuh"
#include "layouts.h"
#include "rollout_args.h"
#include "tile_engine.cuh"
#include "hutter_policy.cuh"

#ifdef APG_PROFILE
__device__ long long g_apg_prof[2][148][APG_NPROF];
extern "C" __attribute__((visibility("default"))) int apg_debug_profile(long long* out_host) {
  return (int)cudaMemcpyFromSymbol(out_host, g_apg_prof, sizeof(long long) * 2 * 148 * APG_NPROF);
}
#endif

namespace apg {

// Warp specialisation: threads 0..255 (8 warps) are the GEMM group, threads 256..319 (2 warps) run the
// thread-per-drone horizon loops.  The two groups work on neighbouring tiles and hand tiles over through mbarriers:
//   forward : GEMM group fills s_act(tile i) -> act_full; the dynamics warps integrate tile i while the GEMM group
//             already computes tile i+1; act_empty releases s_act again.
//   adjoint : the dynamics warps sweep tile i+1 backwards into s_dlog while the GEMM group back-propagates tile i;
//             dlog_full / dlog_empty.
constexpr int NTH = NT + TM;     // 320 threads

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
template <template <typename> class SysT, bool CONV>
__global__ void __launch_bounds__(NTH, 1) hutter_fwd_kernel(const HutterLayout y, const RolloutArgs g) {
  APG_DYNAMIC_SMEM_F32(smem);
  using Sys = SysT<float>;
  constexpr int S = Sys::S, A = Sys::A, R = Sys::REFW;
  float* s_w = smem;
  float* s_ins = s_w + y.f_total;
  float* s_inr = s_ins + pad4(TM * y.F0);
  float* s_x1 = s_inr + pad4(TM * y.LR);
  float* s_h = s_x1 + y.K1 * TMP;
  float* s_act = s_h + HID * TMP;
  float* s_red = s_act + y.Mo4 * TMP;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(s_red + 8);
  uint64_t* bar_in = bar_w + 1;
  uint64_t* act_full = bar_w + 2;
  uint64_t* act_empty = bar_w + 3;

  const int tid = threadIdx.x;
  const int ntiles = (g.N + TM - 1) / TM;
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_in, 1);
    mbar_init(act_full, 1);
    mbar_init(act_empty, TM);
    fence_mbar_init();
  }
  __syncthreads();

  if (tid < NT) {
    // ===================================================== GEMM group
    const Lane L;
    const uint32_t ins_bytes = TM * y.F0 * 4, inr_bytes = TM * y.LR * 4;
    auto issue_inputs = [&](int tile) {   // thread 0, full tiles only
      mbar_expect_tx(bar_in, ins_bytes + inr_bytes);
      bulk_g2s(s_ins, g.in_state + (size_t)tile * TM * y.F0, ins_bytes, bar_in);
      bulk_g2s(s_inr, g.in_ref + (size_t)tile * TM * y.LR, inr_bytes, bar_in);
    };
    auto tile_full = [&](int tile) { return (tile + 1) * TM <= g.N; };
    if (tid == 0) {
      mbar_expect_tx(bar_w, y.f_total * 4);
      bulk_g2s_chunked(s_w, g.wf, y.f_total * 4, bar_w);
      if ((int)blockIdx.x < ntiles && tile_full(blockIdx.x)) issue_inputs(blockIdx.x);
    }
    mbar_wait(bar_w, 0);
    uint32_t in_phase = 0, it = 0;
    PROF_DECL
    PROF(0);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int valid = min(TM, g.N - tile * TM);
      if (valid == TM) {
        mbar_wait(bar_in, in_phase);
        in_phase ^= 1;
      } else {
        load_tile_manual(s_ins, g.in_state + (size_t)tile * TM * y.F0, y.F0, valid);
        load_tile_manual(s_inr, g.in_ref + (size_t)tile * TM * y.LR, y.LR, valid);
        gsync<true>();
      }
      PROF(1);
      // ---- first layer: state branch and reference branch -> X1 = [s | r]
      hutter_first_layer<CONV>(L, y, s_w, s_ins, s_inr, s_x1);
      fence_proxy_async();
      gsync<true>();
      PROF(2);
      if (tid == 0) {
        const int next = tile + gridDim.x;
        if (next < ntiles && tile_full(next)) issue_inputs(next);        // input buffers are free again
        bulk_s2g(g.st_x1 + (size_t)tile * y.K1 * TMP, s_x1, y.K1 * TMP * 4);
        bulk_commit();
      }
      // ---- fc1, fc2, fc3, fc_out + sigmoid (train_base.py:202-203); every activation is stashed for the adjoint
      hutter_trunk<true>(L, y, s_w, s_x1, s_h, s_act, g.st_h1 + (size_t)tile * HID * TMP,
                         g.st_h2 + (size_t)tile * HID * TMP, g.st_h3 + (size_t)tile * HID * TMP,
                         g.st_act + (size_t)tile * y.Mo4 * TMP, it > 0 ? act_empty : nullptr, (it - 1) & 1);
      PROF(3);
      if (tid == 0) {
        mbar_arrive(act_full);        // actions of this tile are in s_act: hand the tile to the dynamics warps
        bulk_wait_read<0>();          // every stash store has finished reading shared memory
      }
      gsync<true>();
      PROF(4);
    }
    PROF_FLUSH(0);
  } else {
    // ===================================================== dynamics warps: one thread per drone of the tile
    const int d = tid - NT;
    float my_loss = 0.f;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int valid = min(TM, g.N - tile * TM);
      mbar_wait(act_full, it & 1);
      if (d < valid) {
        const size_t drone = (size_t)tile * TM + d;
        my_loss += dyn_forward_conc<SysT>(s_act, d, g.cur + drone * S, g.ref + drone * g.ref_rows * R, g.h, g.dt,
                                          g.pc.v, g.st_states + (size_t)tile * g.h * S * TMP,
                                          g.states_out ? g.states_out + drone * g.h * S : nullptr,
                                          g.actions_out ? g.actions_out + drone * g.h * A : nullptr);
      }
      mbar_arrive(act_empty);
    }
    // fixed-order sum of the 64 per-thread losses
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_loss += __shfl_xor_sync(0xffffffffu, my_loss, o);
    if ((d & 31) == 0) s_red[d >> 5] = my_loss;
    dyn_group_sync();
    if (d == 0) g.loss_partials[blockIdx.x] = s_red[0] + s_red[1];
  }
  if (tid == 0) bulk_wait_all();
}

// ------------------------------------------------------------------------------------------------------------
// adjoint
// ------------------------------------------------------------------------------------------------------------
template <template <typename> class SysT, bool CONV>
__global__ void __launch_bounds__(NTH, 1) hutter_adj_kernel(const HutterLayout y, const RolloutArgs g) {
  APG_DYNAMIC_SMEM_F32(smem);
  using Sys = SysT<float>;
  constexpr int S = Sys::S, R = Sys::REFW;
  const int wb_floats = y.b_ws;              // concurrent mode needs no first-layer dX weights
  float* s_w = smem;
  float* bufA = s_w + wb_floats;
  float* bufB = bufA + y.K1 * TMP;
  float* bufD = bufB + HID * TMP;
  float* bufC = bufD + HID * TMP;
  float* s_dlog = bufC + HID * TMP;          // [Mo4][TMP]
  float* s_red = s_dlog + y.Mo4 * TMP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_red + 8);
  uint64_t *bar_w = bars, *bar_A = bars + 1, *bar_B = bars + 2, *bar_D = bars + 3, *bar_C = bars + 4,
           *bar_in = bars + 5, *dlog_full = bars + 6, *dlog_empty = bars + 7;
  float* s_ins = bufB;                       // input tiles reuse bufB|bufD once those are dead
  float* s_inr = bufB + pad4(TM * y.F0);

  const int tid = threadIdx.x;
  const int ntiles = (g.N + TM - 1) / TM;
  if (tid == 0) {
    for (int b = 0; b < 6; ++b) mbar_init(bars + b, 1);
    mbar_init(dlog_full, TM);
    mbar_init(dlog_empty, 1);
    fence_mbar_init();
  }
  __syncthreads();
  // tiles are visited in the reverse order of the forward kernel: the most recently written stash is still in L2
  const int first = ntiles - 1 - (int)blockIdx.x;

  if (tid < NT) {
    // ===================================================== GEMM group
    const Lane L;
    float* P = g.grad_partials + (size_t)blockIdx.x * y.n_params;
    for (int i = tid; i < y.n_params; i += NT) __stcg(P + i, 0.f);
    const uint32_t hbytes = HID * TMP * 4;
    auto issue_stage_loads = [&](int tile) {    // thread 0
      mbar_expect_tx(bar_A, y.K1 * TMP * 4);
      bulk_g2s_chunked(bufA, g.st_x1 + (size_t)tile * y.K1 * TMP, y.K1 * TMP * 4, bar_A);
      mbar_expect_tx(bar_B, hbytes);
      bulk_g2s(bufB, g.st_h3 + (size_t)tile * HID * TMP, hbytes, bar_B);
      mbar_expect_tx(bar_D, hbytes);
      bulk_g2s(bufD, g.st_h2 + (size_t)tile * HID * TMP, hbytes, bar_D);
      mbar_expect_tx(bar_C, hbytes);
      bulk_g2s(bufC, g.st_h1 + (size_t)tile * HID * TMP, hbytes, bar_C);
    };
    if (tid == 0) {
      mbar_expect_tx(bar_w, wb_floats * 4);
      bulk_g2s_chunked(s_w, g.wb, wb_floats * 4, bar_w);
      if (first >= 0) issue_stage_loads(first);
    }
    mbar_wait(bar_w, 0);
    gsync<true>();                             // partial zeroed before the first reduction
    uint32_t ph = 0, ph_in = 0;
    PROF_DECL
    PROF(0);
    for (int tile = first; tile >= 0; tile -= gridDim.x) {
      const int valid = min(TM, g.N - tile * TM);
      // ---- d loss / d logits of this tile comes from the dynamics warps
      mbar_wait(dlog_full, ph);
      PROF(1);
      // ---- fc_out
      mbar_wait(bar_B, ph);
      PROF(2);
      dw_auto(L, s_dlog, y.Mo, bufB, HID, P + y.t_wo, HID, P + y.t_bo);
      gsync<true>();
      PROF(3);
      dense_auto<EPI_DTANH>(L, s_dlog, y.Mo, s_w + y.b_wo, HID, mma_sw(HID), nullptr, HID, bufB, 0, 0);  // dz3 over h3
      gsync<true>();
      if (tid == 0) mbar_arrive(dlog_empty);     // s_dlog may be overwritten with the next tile's sweep
      PROF(4);
      // ---- fc3
      mbar_wait(bar_D, ph);
      PROF(5);
      dw_auto(L, bufB, HID, bufD, HID, P + y.t_w3, HID, P + y.t_b3);
      gsync<true>();
      PROF(6);
      dense_auto<EPI_DTANH>(L, bufB, HID, s_w + y.b_w3, HID, mma_sw(HID), nullptr, HID, bufD, 0, 0);    // dz2 over h2
      gsync<true>();
      PROF(7);
      // ---- fc2
      mbar_wait(bar_C, ph);
      PROF(8);
      dw_auto(L, bufD, HID, bufC, HID, P + y.t_w2, HID, P + y.t_b2);
      gsync<true>();
      PROF(9);
      dense_auto<EPI_DTANH>(L, bufD, HID, s_w + y.b_w2, HID, mma_sw(HID), nullptr, HID, bufC, 0, 0);    // dz1 over h1
      fence_proxy_async();
      gsync<true>();
      PROF(10);
      // bufB | bufD are dead: fetch the input tiles into them while fc1 is processed
      if (valid == TM) {
        if (tid == 0) {
          mbar_expect_tx(bar_in, TM * (y.F0 + y.LR) * 4);
          bulk_g2s(s_ins, g.in_state + (size_t)tile * TM * y.F0, TM * y.F0 * 4, bar_in);
          bulk_g2s(s_inr, g.in_ref + (size_t)tile * TM * y.LR, TM * y.LR * 4, bar_in);
        }
      } else {
        load_tile_manual(s_ins, g.in_state + (size_t)tile * TM * y.F0, y.F0, valid);
        load_tile_manual(s_inr, g.in_ref + (size_t)tile * TM * y.LR, y.LR, valid);
      }
      // ---- fc1
      mbar_wait(bar_A, ph);
      PROF(11);
      dw_auto(L, bufC, HID, bufA, y.K1, P + y.t_w1, y.K1, P + y.t_b1, y.perm_npos);
      gsync<true>();
      PROF(12);
      dense_auto<EPI_DTANH>(L, bufC, HID, s_w + y.b_w1, y.ld_bw1, mma_sw(y.ld_bw1), nullptr, HID, bufA, 0, 0);  // ds
      if (CONV)
        dense_auto<EPI_DRELU>(L, bufC, HID, s_w + y.b_w1, y.ld_bw1, mma_sw(y.ld_bw1), nullptr, y.NRtot, bufA, HID, 0, HID);
      else
        dense_auto<EPI_DTANH>(L, bufC, HID, s_w + y.b_w1, y.ld_bw1, mma_sw(y.ld_bw1), nullptr, y.NRtot, bufA, HID, 0, HID);
      gsync<true>();
      PROF(13);
      // ---- first layer weight gradients (no dX: the inputs need no gradient in concurrent mode)
      if (valid == TM) {
        mbar_wait(bar_in, ph_in);
        ph_in ^= 1;
      }
      PROF(14);
      aos_linear64_dw(L, bufA, s_ins, y.F0, y.F0, P + y.t_ws, P + y.t_bs);
      if (CONV)
        conv_dw(L, y, bufA + HID * TMP, s_inr, nullptr, P);
      else
        aos_linear64_dw(L, bufA + HID * TMP, s_inr, y.LR, y.LR, P + y.t_wr, P + y.t_br);
      fence_proxy_async();
      gsync<true>();
      PROF(15);
      ph ^= 1;
      const int next = tile - gridDim.x;
      if (tid == 0 && next >= 0) issue_stage_loads(next);
    }
    PROF_FLUSH(1);
  } else {
    // ===================================================== dynamics warps: reverse sweep, one thread per drone
    const int d = tid - NT;
    uint32_t it = 0;
    for (int tile = first; tile >= 0; tile -= gridDim.x, ++it) {
      const int valid = min(TM, g.N - tile * TM);
      if (it > 0) mbar_wait(dlog_empty, (it - 1) & 1);
      if (d < valid) {
        const size_t drone = (size_t)tile * TM + d;
        dyn_adjoint_conc<SysT>(g.st_act + (size_t)tile * y.Mo4 * TMP, g.st_states + (size_t)tile * g.h * S * TMP, d,
                               g.cur + drone * S, g.ref + drone * g.ref_rows * R, g.h, g.dt, g.pc.v, s_dlog);
      } else {
        for (int r = 0; r < y.Mo4; ++r) s_dlog[r * TMP + d] = 0.f;
      }
      mbar_arrive(dlog_full);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------------------
size_t hutter_fwd_smem_bytes(const HutterLayout& y) {
  return sizeof(float) * (size_t)(y.f_total + pad4(TM * y.F0) + pad4(TM * y.LR) + y.K1 * TMP + HID * TMP +
                                  y.Mo4 * TMP + 8) + 48;
}
size_t hutter_adj_smem_bytes(const HutterLayout& y) {
  return sizeof(float) * (size_t)(y.b_ws + y.K1 * TMP + 3 * HID * TMP + y.Mo4 * TMP + 8) + 80;
}

template <typename K>
static cudaError_t set_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t launch_hutter_fwd(int system, const HutterLayout& y, const RolloutArgs& a, int grid, cudaStream_t st) {
  const size_t smem = hutter_fwd_smem_bytes(y);
  cudaError_t e;
  if (system == SYS_QUAD && y.conv) {
    if ((e = set_smem(hutter_fwd_kernel<Quad, true>, smem)) != cudaSuccess) return e;
    APG_LAUNCH(grid, NTH, smem, st, hutter_fwd_kernel<Quad, true>)(y, a);
  } else if (system == SYS_WING && !y.conv) {
    if ((e = set_smem(hutter_fwd_kernel<Wing, false>, smem)) != cudaSuccess) return e;
    APG_LAUNCH(grid, NTH, smem, st, hutter_fwd_kernel<Wing, false>)(y, a);
  } else {
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_hutter_adj(int system, const HutterLayout& y, const RolloutArgs& a, int grid, cudaStream_t st) {
  const size_t smem = hutter_adj_smem_bytes(y);
  cudaError_t e;
  if (system == SYS_QUAD && y.conv) {
    if ((e = set_smem(hutter_adj_kernel<Quad, true>, smem)) != cudaSuccess) return e;
    APG_LAUNCH(grid, NTH, smem, st, hutter_adj_kernel<Quad, true>)(y, a);
  } else if (system == SYS_WING && !y.conv) {
    if ((e = set_smem(hutter_adj_kernel<Wing, false>, smem)) != cudaSuccess) return e;
    APG_LAUNCH(grid, NTH, smem, st, hutter_adj_kernel<Wing, false>)(y, a);
  } else {
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}


}  // namespace apg
