// Fused rollout kernels for the RECURRENT train modes of the quadrotor (scripts/train_drone.py:113-173):
// the policy is evaluated inside the horizon loop on (state_preprocessing(current_state), next-h reference rows
// relative to the drone) and the loss gradient is back-propagated through time (policy <- dynamics <- policy ...).
//
//   autoregressive : hutter MLP Net(15, h, 9, 4)        (rec_fwd_kernel / rec_adj_kernel)
//
// Window semantics: the reference subtracts the current position IN PLACE from a view of
// the shared (N,2h,9) buffer, so the rows seen at step k are
//     cumulative: in_ref0[k+r,:3] - (P_k - P_{k+r-h}),  P_m = sum_{i<=m} pos_i  (P_{<0} = 0)      [the reference]
//     relative  : in_ref0[k+r,:3] - pos_k                                                       [documented intent]
// Both are implemented, forward and adjoint (the reference's own backward() raises for these modes, so the
// gradient is defined by the functional restatement pinned in oracle/apg_oracle.py::rollout_recurrent).
#include "dyn_phase.cuh"
#include "hutter_policy.cuh"
#include "layouts.h"
#include "rollout_args.h"
#include "tile_engine.cuh"
#include "rec_common.cuh"

namespace apg {

// ------------------------------------------------------------------------------------------------------------
// autoregressive forward
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1) rec_fwd_kernel(const HutterLayout y, const RolloutArgs g) {
  APG_DYNAMIC_SMEM_F32(smem);
  using Sys = Quad<float>;
  constexpr int S = Sys::S, A = Sys::A, R = Sys::REFW;
  const int h = g.h;
  float* s_w = smem;
  float* s_ins = s_w + y.f_total;
  float* s_win = s_ins + pad4(TM * y.F0);
  float* s_x1 = s_win + pad4(TM * y.LR);
  float* s_h = s_x1 + y.XR * TMP;
  float* s_P = s_h + HID * TMP;
  float* s_pos = s_P + h * 3 * TMP;
  float* s_red = s_pos + 4 * TMP;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(s_red + 8);
  const Lane L;
  const int tid = threadIdx.x;
  const int ntiles = (g.N + TM - 1) / TM;
  if (tid == 0) {
    mbar_init(bar_w, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar_w, y.f_total * 4);
    bulk_g2s_chunked(s_w, g.wf, y.f_total * 4, bar_w);
  }
  mbar_wait(bar_w, 0);
  float cta_loss = 0.f;
  float* s_act = s_x1 + HID * TMP;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int valid = min(TM, g.N - tile * TM);
    const size_t drone = (size_t)tile * TM + tid;
    float s[S], P[3] = {0.f, 0.f, 0.f};
    float my_loss = 0.f;
#pragma unroll
    for (int i = 0; i < S; ++i) s[i] = (tid < valid) ? g.cur[drone * S + i] : 0.f;
    for (int k = 0; k < h; ++k) {
      if (tid < TM) {
        float f[15];
        Sys::features(s, f);
#pragma unroll
        for (int i = 0; i < 15; ++i) s_ins[tid * y.F0 + i] = f[i];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          P[c] += s[c];
          s_P[(k * 3 + c) * TMP + tid] = P[c];
          s_pos[c * TMP + tid] = s[c];
        }
      }
      __syncthreads();
      build_window(s_win, g.in_ref + (size_t)tile * TM * 2 * y.LR, s_P, s_pos, k, h, y.RD, valid, g.window);
      __syncthreads();
      hutter_first_layer<true>(L, y, s_w, s_ins, s_win, s_x1);
      fence_proxy_async();
      __syncthreads();
      const size_t sk = (size_t)tile * h + k;
      if (tid == 0) {
        bulk_s2g(g.st_x1 + sk * y.K1 * TMP, s_x1, y.K1 * TMP * 4);
        bulk_commit();
      }
      hutter_trunk<false>(L, y, s_w, s_x1, s_h, s_act, g.st_h1 + sk * HID * TMP, g.st_h2 + sk * HID * TMP,
                          g.st_h3 + sk * HID * TMP, g.st_act + sk * y.Mo4 * TMP);
      if (tid < valid) {
        float a[A], rf[R], sn[S];
#pragma unroll
        for (int c = 0; c < A; ++c) a[c] = s_act[c * TMP + tid];
#pragma unroll
        for (int c = 0; c < R; ++c) rf[c] = g.ref[(drone * g.ref_rows + k) * R + c];
        Sys::step(s, a, g.dt, g.pc.v, sn);
        my_loss += Sys::loss(sn, rf, a, nullptr, k, h);
        float* st = g.st_states + (size_t)tile * h * S * TMP;
#pragma unroll
        for (int i = 0; i < S; ++i) {
          s[i] = sn[i];
          st[(k * S + i) * TMP + tid] = sn[i];
        }
        if (g.states_out) {
#pragma unroll
          for (int i = 0; i < S; ++i) g.states_out[(drone * h + k) * S + i] = sn[i];
        }
        if (g.actions_out) {
#pragma unroll
          for (int c = 0; c < A; ++c) g.actions_out[(drone * h + k) * A + c] = a[c];
        }
      }
      if (tid == 0) bulk_wait_read<0>();
      __syncthreads();
    }
    const float tl = block_sum(my_loss, s_red);
    if (tid == 0) cta_loss += tl;
  }
  if (tid == 0) {
    g.loss_partials[blockIdx.x] = cta_loss;
    bulk_wait_all();
  }
}

// ------------------------------------------------------------------------------------------------------------
// autoregressive adjoint (back-propagation through time)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1) rec_adj_kernel(const HutterLayout y, const RolloutArgs g) {
  APG_DYNAMIC_SMEM_F32(smem);
  using Sys = Quad<float>;
  constexpr int S = Sys::S, A = Sys::A, R = Sys::REFW;
  const int h = g.h;
  float* s_w = smem;
  float* bufA = s_w + y.b_total;
  float* bufB = bufA + y.XR * TMP;
  float* bufD = bufB + HID * TMP;
  float* bufC = bufD + HID * TMP;
  float* s_dlog = bufC + HID * TMP;
  float* s_P = s_dlog + 4 * TMP;
  float* s_dP = s_P + h * 3 * TMP;
  float* s_pos = s_dP + h * 3 * TMP;
  float* s_red = s_pos + 4 * TMP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_red + 8);
  uint64_t *bar_w = bars, *bar_A = bars + 1, *bar_B = bars + 2, *bar_D = bars + 3, *bar_C = bars + 4;
  float* s_ins = bufB;
  float* s_win = bufB + pad4(TM * y.F0);
  float* scratch = s_win + pad4(TM * y.LR);
  float* s_din = bufC;                 // [16][TMP]  d loss / d features
  float* s_dwin = bufC + 16 * TMP;     // [3h][TMP]  d loss / d window position columns

  const Lane L;
  const int tid = threadIdx.x;
  const int ntiles = (g.N + TM - 1) / TM;
  float* P = g.grad_partials + (size_t)blockIdx.x * y.n_params;
  for (int i = tid; i < y.n_params; i += NT) __stcg(P + i, 0.f);
  if (tid == 0) {
    for (int b = 0; b < 5; ++b) mbar_init(bars + b, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const uint32_t hbytes = HID * TMP * 4;
  auto issue_step_loads = [&](size_t sk) {     // thread 0
    mbar_expect_tx(bar_A, y.K1 * TMP * 4);
    bulk_g2s_chunked(bufA, g.st_x1 + sk * y.K1 * TMP, y.K1 * TMP * 4, bar_A);
    mbar_expect_tx(bar_B, hbytes);
    bulk_g2s(bufB, g.st_h3 + sk * HID * TMP, hbytes, bar_B);
    mbar_expect_tx(bar_D, hbytes);
    bulk_g2s(bufD, g.st_h2 + sk * HID * TMP, hbytes, bar_D);
    mbar_expect_tx(bar_C, hbytes);
    bulk_g2s(bufC, g.st_h1 + sk * HID * TMP, hbytes, bar_C);
  };
  const int first = ntiles - 1 - (int)blockIdx.x;
  if (tid == 0) {
    mbar_expect_tx(bar_w, y.b_total * 4);
    bulk_g2s_chunked(s_w, g.wb, y.b_total * 4, bar_w);
    if (first >= 0) issue_step_loads((size_t)first * h + (h - 1));
  }
  mbar_wait(bar_w, 0);

  uint32_t ph = 0;
  for (int tile = first; tile >= 0; tile -= gridDim.x) {
    const int valid = min(TM, g.N - tile * TM);
    const size_t drone = (size_t)tile * TM + tid;
    const float* st = g.st_states + (size_t)tile * h * S * TMP;
    float gS[S], srun[3] = {0.f, 0.f, 0.f}, sk_[S];
#pragma unroll
    for (int i = 0; i < S; ++i) gS[i] = 0.f;
    if (tid < TM) {
      float Pc[3] = {0.f, 0.f, 0.f};
      for (int m = 0; m < h; ++m) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float pos = 0.f;
          if (tid < valid) pos = (m == 0) ? g.cur[drone * S + c] : st[((m - 1) * S + c) * TMP + tid];
          Pc[c] += pos;
          s_P[(m * 3 + c) * TMP + tid] = Pc[c];
          s_dP[(m * 3 + c) * TMP + tid] = 0.f;
        }
      }
    }
    __syncthreads();
    for (int k = h - 1; k >= 0; --k) {
      const size_t sk = (size_t)tile * h + k;
      // ---- dynamics + loss adjoint of step k -> d loss / d logits, state cotangent through the step
      if (tid < TM) {
        if (tid < valid) {
          float sn[S], a[A], rf[R], ga[A], ga2[A], gs[S];
#pragma unroll
          for (int i = 0; i < S; ++i) {
            sn[i] = st[(k * S + i) * TMP + tid];
            sk_[i] = (k > 0) ? st[((k - 1) * S + i) * TMP + tid] : g.cur[drone * S + i];
          }
#pragma unroll
          for (int c = 0; c < A; ++c) { a[c] = g.st_act[sk * y.Mo4 * TMP + c * TMP + tid]; ga[c] = 0.f; }
#pragma unroll
          for (int c = 0; c < R; ++c) rf[c] = g.ref[(drone * g.ref_rows + k) * R + c];
          Sys::loss_grad(sn, rf, a, nullptr, k, h, gS, ga);
          Sys::step_adj(sk_, a, g.dt, g.pc.v, gS, gs, ga2);
#pragma unroll
          for (int c = 0; c < A; ++c) s_dlog[c * TMP + tid] = (ga[c] + ga2[c]) * a[c] * (1.f - a[c]);
#pragma unroll
          for (int i = 0; i < S; ++i) gS[i] = gs[i];
#pragma unroll
          for (int c = 0; c < 3; ++c) s_pos[c * TMP + tid] = sk_[c];
        } else {
#pragma unroll
          for (int i = 0; i < S; ++i) sk_[i] = 0.f;
#pragma unroll
          for (int c = 0; c < A; ++c) s_dlog[c * TMP + tid] = 0.f;
#pragma unroll
          for (int c = 0; c < 3; ++c) s_pos[c * TMP + tid] = 0.f;
        }
      }
      __syncthreads();
      // ---- fc_out
      mbar_wait(bar_B, ph);
      dw_auto(L, s_dlog, y.Mo, bufB, HID, P + y.t_wo, HID, P + y.t_bo);
      __syncthreads();
      dense_auto<EPI_DTANH>(L, s_dlog, y.Mo, s_w + y.b_wo, HID, mma_sw(HID), nullptr, HID, bufB, 0, 0);
      __syncthreads();
      // ---- fc3
      mbar_wait(bar_D, ph);
      dw_auto(L, bufB, HID, bufD, HID, P + y.t_w3, HID, P + y.t_b3);
      __syncthreads();
      dense_auto<EPI_DTANH>(L, bufB, HID, s_w + y.b_w3, HID, mma_sw(HID), nullptr, HID, bufD, 0, 0);
      __syncthreads();
      // ---- fc2
      mbar_wait(bar_C, ph);
      dw_auto(L, bufD, HID, bufC, HID, P + y.t_w2, HID, P + y.t_b2);
      __syncthreads();
      dense_auto<EPI_DTANH>(L, bufD, HID, s_w + y.b_w2, HID, mma_sw(HID), nullptr, HID, bufC, 0, 0);
      __syncthreads();
      // ---- rebuild the policy inputs of step k in bufB|bufD (dead now): features(S_k) and the window
      if (tid < TM) {
        float f[15];
        Sys::features(sk_, f);
#pragma unroll
        for (int i = 0; i < 15; ++i) s_ins[tid * y.F0 + i] = f[i];
      }
      build_window(s_win, g.in_ref + (size_t)tile * TM * 2 * y.LR, s_P, s_pos, k, h, y.RD, valid, g.window);
      __syncthreads();
      // ---- fc1
      mbar_wait(bar_A, ph);
      dw_auto(L, bufC, HID, bufA, y.K1, P + y.t_w1, y.K1, P + y.t_b1, y.perm_npos);
      __syncthreads();
      dense_auto<EPI_DTANH>(L, bufC, HID, s_w + y.b_w1, y.ld_bw1, mma_sw(y.ld_bw1), nullptr, HID, bufA, 0, 0);
      dense_auto<EPI_DRELU>(L, bufC, HID, s_w + y.b_w1, y.ld_bw1, mma_sw(y.ld_bw1), nullptr, y.NRtot, bufA, HID, 0, HID);
      __syncthreads();
      // ---- first layer: weight gradients, then the input gradients (they feed the state cotangent)
      aos_linear64_dw(L, bufA, s_ins, y.F0, y.F0, P + y.t_ws, P + y.t_bs);
      conv_dw(L, y, bufA + HID * TMP, s_win, scratch, P);
      __syncthreads();
      dense<SrcT, EPI_DNONE>(L, SrcT{bufA}, HID, s_w + y.b_ws, y.ld_bws, nullptr, y.ld_bws / 4, s_din, 0, 1, 0);
      conv_dx_pos(y, bufA + HID * TMP, s_w + y.b_wr, s_dwin);
      __syncthreads();
      if (tid < valid) state_input_adjoint(sk_, s_din, s_dwin, s_dP, srun, gS, k, h, g.window, tid);
      fence_proxy_async();
      __syncthreads();
      ph ^= 1;
      if (tid == 0) {
        if (k > 0) issue_step_loads(sk - 1);
        else if (tile - (int)gridDim.x >= 0) issue_step_loads((size_t)(tile - gridDim.x) * h + (h - 1));
      }
    }
  }
}

size_t rec_fwd_smem_bytes(const HutterLayout& y, int h) {
  return sizeof(float) * (size_t)(y.f_total + pad4(TM * y.F0) + pad4(TM * y.LR) + y.XR * TMP + HID * TMP +
                                  (3 * h + 4) * TMP + 8) + 16;
}
size_t rec_adj_smem_bytes(const HutterLayout& y, int h) {
  return sizeof(float) * (size_t)(y.b_total + y.XR * TMP + 3 * HID * TMP + (4 + 6 * h + 4) * TMP + 8) + 48;
}

cudaError_t launch_rec_fwd(const HutterLayout& y, const RolloutArgs& a, int grid, cudaStream_t st) {
  const size_t smem = rec_fwd_smem_bytes(y, a.h);
  cudaError_t e = cudaFuncSetAttribute(rec_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  APG_LAUNCH(grid, NT, smem, st, rec_fwd_kernel)(y, a);
  return cudaGetLastError();
}

cudaError_t launch_rec_adj(const HutterLayout& y, const RolloutArgs& a, int grid, cudaStream_t st) {
  const size_t smem = rec_adj_smem_bytes(y, a.h);
  cudaError_t e = cudaFuncSetAttribute(rec_adj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  APG_LAUNCH(grid, NT, smem, st, rec_adj_kernel)(y, a);
  return cudaGetLastError();
}

}  // namespace apg
