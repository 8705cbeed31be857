// Fused rollout kernels for the LSTM train mode of the quadrotor (scripts/train_drone.py:113-173 with
// train_mode == "LSTM"; policy models/rnn.py LSTM_NEW: conv encoder -> LSTMCell(175, 8), gate order i,f,g,o ->
// Linear(8, 4)).  The initial hidden / cell state (reference: torch.randn in reset_hidden_state, rnn.py:29-32) is
// an INPUT (h0c0 [2][N][8]) so that parity does not depend on RNG order.
//
// Every step's activation arena (features | conv | h_prev | c_prev | gates | c' | h' | actions, LstmLayout) is
// stashed with one bulk copy; the adjoint restores it and back-propagates through the cell, the conv encoder, the
// featurizer, the reference window and the dynamics (BPTT).
#include "dyn_phase.cuh"
#include "hutter_policy.cuh"
#include "layouts.h"
#include "rec_common.cuh"
#include "rollout_args.h"
#include "tile_engine.cuh"

namespace apg {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(NT, 1) lstm_fwd_kernel(const LstmLayout y, const RolloutArgs g) {
  APG_DYNAMIC_SMEM_F32(smem);
  using Sys = Quad<float>;
  constexpr int S = Sys::S, A = Sys::A, R = Sys::REFW, HS = LSTM_HS;
  const int h = g.h;
  float* s_w = smem;
  float* s_win = s_w + y.f_total;
  float* arena0 = s_win + pad4(TM * y.LR);
  float* arena1 = arena0 + y.ROWS * TMP;
  float* s_P = arena1 + y.ROWS * TMP;
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
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int valid = min(TM, g.N - tile * TM);
    const size_t drone = (size_t)tile * TM + tid;
    float s[S], P[3] = {0.f, 0.f, 0.f};
    float my_loss = 0.f;
#pragma unroll
    for (int i = 0; i < S; ++i) s[i] = (tid < valid) ? g.cur[drone * S + i] : 0.f;
    for (int k = 0; k < h; ++k) {
      float* ar = (k & 1) ? arena1 : arena0;
      const float* prev = (k & 1) ? arena0 : arena1;
      if (tid < TM) {
        float f[15];
        Sys::features(s, f);
#pragma unroll
        for (int i = 0; i < 15; ++i) ar[i * TMP + tid] = f[i];
        for (int i = y.F0; i < pad4(y.F0); ++i) ar[i * TMP + tid] = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          P[c] += s[c];
          s_P[(k * 3 + c) * TMP + tid] = P[c];
          s_pos[c * TMP + tid] = s[c];
        }
#pragma unroll
        for (int u = 0; u < HS; ++u) {
          float hp = 0.f, cp = 0.f;
          if (k == 0) {
            if (tid < valid) {
              hp = g.h0c0[drone * HS + u];
              cp = g.h0c0[((size_t)g.N + drone) * HS + u];
            }
          } else {
            hp = prev[(y.R_H + u) * TMP + tid];
            cp = prev[(y.R_C + u) * TMP + tid];
          }
          ar[(y.R_HP + u) * TMP + tid] = hp;
          ar[(y.R_CP + u) * TMP + tid] = cp;
        }
      }
      __syncthreads();
      build_window(s_win, g.in_ref + (size_t)tile * TM * 2 * y.LR, s_P, s_pos, k, h, y.RD, valid, g.window);
      __syncthreads();
      conv_layer_fwd(L, y.npos, y.RD, y.LR, y.KC, s_w + y.f_wc, s_w + y.f_bc, s_win, ar, pad4(y.F0), y.npos, 1);
      __syncthreads();
      // gate pre-activations: [x | h_prev] (KG rows) x Wg -> rows R_G .. R_G+32 (biases added in the cell pass)
      dense<SrcT, EPI_ACT>(L, SrcT{ar}, y.KG, s_w + y.f_wg, 4 * HS, nullptr, HS, ar, y.R_G, 1, ACT_NONE);
      __syncthreads();
      for (int idx = tid; idx < TM * HS; idx += NT) {
        const int u = idx / TM, d = idx - u * TM;
        const float* bi = s_w + y.f_bih;
        const float* bh = s_w + y.f_bhh;
        float* gp = ar + y.R_G * TMP + d;
        const float ig = sigmoidf_(gp[(u) * TMP] + bi[u] + bh[u]);
        const float fg = sigmoidf_(gp[(HS + u) * TMP] + bi[HS + u] + bh[HS + u]);
        const float gg = tanhf(gp[(2 * HS + u) * TMP] + bi[2 * HS + u] + bh[2 * HS + u]);
        const float og = sigmoidf_(gp[(3 * HS + u) * TMP] + bi[3 * HS + u] + bh[3 * HS + u]);
        const float cn = fg * ar[(y.R_CP + u) * TMP + d] + ig * gg;
        gp[(u) * TMP] = ig; gp[(HS + u) * TMP] = fg; gp[(2 * HS + u) * TMP] = gg; gp[(3 * HS + u) * TMP] = og;
        ar[(y.R_C + u) * TMP + d] = cn;
        ar[(y.R_H + u) * TMP + d] = og * tanhf(cn);
      }
      __syncthreads();
      {   // fc_out + sigmoid: one (drone, action) per thread
        const int c = tid / TM, d = tid - c * TM;
        float acc = s_w[y.f_bo + c];
#pragma unroll
        for (int u = 0; u < HS; ++u) acc = fmaf(ar[(y.R_H + u) * TMP + d], s_w[y.f_wo + u * pad4(y.Mo) + c], acc);
        ar[(y.R_A + c) * TMP + d] = sigmoidf_(acc);
      }
      fence_proxy_async();
      __syncthreads();
      const size_t sk = (size_t)tile * h + k;
      if (tid == 0) {
        bulk_s2g(g.st_x1 + sk * y.ROWS * TMP, ar, y.ROWS * TMP * 4);
        bulk_commit();
        bulk_wait_read<1>();        // the store of step k-1 (other arena) has finished reading
      }
      if (tid < valid) {
        float a[A], rf[R], sn[S];
#pragma unroll
        for (int c = 0; c < A; ++c) a[c] = ar[(y.R_A + c) * TMP + tid];
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
      __syncthreads();
    }
    if (tid == 0) bulk_wait_read<0>();
    const float tl = block_sum(my_loss, s_red);
    if (tid == 0) cta_loss += tl;
  }
  if (tid == 0) {
    g.loss_partials[blockIdx.x] = cta_loss;
    bulk_wait_all();
  }
}

// dW of the gate GEMM with the arena-row -> torch-column mapping:
//   arena row k < F0        -> weight_ih[j][k]
//   pad rows [F0, pad4(F0)) -> nothing
//   rows [pad4(F0), KX)     -> weight_ih[j][k - (pad4(F0) - F0)]
//   rows [KX, KG)           -> weight_hh[j][k - KX]
template <int NKI>
__device__ __forceinline__ void lstm_dw_gates(const Lane& L, const LstmLayout& y, const float* __restrict__ dz,
                                              const float* __restrict__ x, float* __restrict__ P) {
  const int M = 4 * LSTM_HS, K = y.KG, f0p = pad4(y.F0);
  for (int j0 = 8 * L.warp; j0 < M; j0 += 8 * NWARP) {
    float acc[8][NKI], accb[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      accb[jj] = 0.f;
#pragma unroll
      for (int i = 0; i < NKI; ++i) acc[jj][i] = 0.f;
    }
    for (int d4 = 0; d4 < TM / 4; ++d4) {
      float4 xv[NKI];
#pragma unroll
      for (int i = 0; i < NKI; ++i) {
        const int k = L.lane + 32 * i;
        xv[i] = k < K ? *reinterpret_cast<const float4*>(x + k * TMP + 4 * d4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float4 z = *reinterpret_cast<const float4*>(dz + (j0 + jj) * TMP + 4 * d4);
        accb[jj] += (z.x + z.y) + (z.z + z.w);
#pragma unroll
        for (int i = 0; i < NKI; ++i) acc[jj][i] = dot4(z, xv[i], acc[jj][i]);
      }
    }
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int j = j0 + jj;
#pragma unroll
      for (int i = 0; i < NKI; ++i) {
        const int k = L.lane + 32 * i;
        if (k < y.F0) red_add(P + y.t_wih + j * y.IH + k, acc[jj][i]);
        else if (k >= f0p && k < y.KX) red_add(P + y.t_wih + j * y.IH + k - (f0p - y.F0), acc[jj][i]);
        else if (k >= y.KX && k < K) red_add(P + y.t_whh + j * LSTM_HS + (k - y.KX), acc[jj][i]);
      }
      if (L.lane == 0) {
        red_add(P + y.t_bih + j, accb[jj]);
        red_add(P + y.t_bhh + j, accb[jj]);
      }
    }
  }
}

__global__ void __launch_bounds__(NT, 1) lstm_adj_kernel(const LstmLayout y, const RolloutArgs g) {
  APG_DYNAMIC_SMEM_F32(smem);
  using Sys = Quad<float>;
  constexpr int S = Sys::S, A = Sys::A, R = Sys::REFW, HS = LSTM_HS;
  const int h = g.h;
  float* s_w = smem;
  float* ar = s_w + y.b_total;                 // arena of the current step [ROWS][TMP]
  float* s_win = ar + y.ROWS * TMP;            // [TM][LR] drone-major window
  float* scratch = s_win + pad4(TM * y.LR);    // conv_dw scratch: NWARP * (20*KC + 20)
  float* s_din = scratch + NWARP * (CONV_CH * y.KC + CONV_CH);   // [16][TMP]
  float* s_dwin = s_din + 16 * TMP;            // [3h][TMP]
  float* s_dh = s_dwin + 3 * h * TMP;          // [8][TMP] carried d loss / d h_prev
  float* s_dc = s_dh + HS * TMP;               // [8][TMP] carried d loss / d c_prev
  float* s_dlog = s_dc + HS * TMP;             // [4][TMP]
  float* s_P = s_dlog + 4 * TMP;
  float* s_dP = s_P + h * 3 * TMP;
  float* s_pos = s_dP + h * 3 * TMP;
  float* s_red = s_pos + 4 * TMP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_red + 8);
  uint64_t *bar_w = bars, *bar_a = bars + 1;

  const Lane L;
  const int tid = threadIdx.x;
  const int ntiles = (g.N + TM - 1) / TM;
  float* P = g.grad_partials + (size_t)blockIdx.x * y.n_params;
  for (int i = tid; i < y.n_params; i += NT) __stcg(P + i, 0.f);
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_a, 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue_arena = [&](size_t sk) {      // thread 0
    mbar_expect_tx(bar_a, y.ROWS * TMP * 4);
    bulk_g2s_chunked(ar, g.st_x1 + sk * y.ROWS * TMP, y.ROWS * TMP * 4, bar_a);
  };
  const int first = ntiles - 1 - (int)blockIdx.x;
  if (tid == 0) {
    mbar_expect_tx(bar_w, y.b_total * 4);
    bulk_g2s_chunked(s_w, g.wb, y.b_total * 4, bar_w);
    if (first >= 0) issue_arena((size_t)first * h + (h - 1));
  }
  mbar_wait(bar_w, 0);
  uint32_t ph = 0;
  const int f0p = pad4(y.F0);
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
#pragma unroll
      for (int u = 0; u < HS; ++u) { s_dh[u * TMP + tid] = 0.f; s_dc[u * TMP + tid] = 0.f; }
    }
    __syncthreads();
    for (int k = h - 1; k >= 0; --k) {
      const size_t sk = (size_t)tile * h + k;
      mbar_wait(bar_a, ph);
      ph ^= 1;
      // ---- dynamics + loss adjoint of step k
      if (tid < TM) {
        if (tid < valid) {
          float sn[S], a[A], rf[R], ga[A], ga2[A], gs[S];
#pragma unroll
          for (int i = 0; i < S; ++i) {
            sn[i] = st[(k * S + i) * TMP + tid];
            sk_[i] = (k > 0) ? st[((k - 1) * S + i) * TMP + tid] : g.cur[drone * S + i];
          }
#pragma unroll
          for (int c = 0; c < A; ++c) { a[c] = ar[(y.R_A + c) * TMP + tid]; ga[c] = 0.f; }
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
      // ---- fc_out weight gradient + cell backward (elementwise; gates rows become pre-activation gradients)
      dw_T<1>(L, s_dlog, y.Mo, ar + y.R_H * TMP, HS, P + y.t_wo, HS, P + y.t_bo);
      for (int idx = tid; idx < TM * HS; idx += NT) {
        const int u = idx / TM, d = idx - u * TM;
        float dh = s_dh[u * TMP + d];
#pragma unroll
        for (int c = 0; c < A; ++c) dh = fmaf(s_dlog[c * TMP + d], s_w[y.b_wo + c * HS + u], dh);
        float* gp = ar + y.R_G * TMP + d;
        const float ig = gp[u * TMP], fg = gp[(HS + u) * TMP], gg = gp[(2 * HS + u) * TMP], og = gp[(3 * HS + u) * TMP];
        const float tc = tanhf(ar[(y.R_C + u) * TMP + d]);
        const float dc = dh * og * (1.f - tc * tc) + s_dc[u * TMP + d];
        const float cp = ar[(y.R_CP + u) * TMP + d];
        gp[u * TMP] = dc * gg * ig * (1.f - ig);
        gp[(HS + u) * TMP] = dc * cp * fg * (1.f - fg);
        gp[(2 * HS + u) * TMP] = dc * ig * (1.f - gg * gg);
        gp[(3 * HS + u) * TMP] = dh * tc * og * (1.f - og);
        s_dc[u * TMP + d] = dc * fg;
      }
      __syncthreads();
      // ---- gate GEMM: weight gradients, then input gradients
      lstm_dw_gates<6>(L, y, ar + y.R_G * TMP, ar, P);
      build_window(s_win, g.in_ref + (size_t)tile * TM * 2 * y.LR, s_P, s_pos, k, h, y.RD, valid, g.window);
      __syncthreads();
      const float* dz = ar + y.R_G * TMP;
      dense<SrcT, EPI_DNONE>(L, SrcT{dz}, 4 * HS, s_w + y.b_wg, y.KG, nullptr, f0p / 4, s_din, 0, 1, 0);
      dense<SrcT, EPI_DRELU>(L, SrcT{dz}, 4 * HS, s_w + y.b_wg + f0p, y.KG, nullptr, (y.KX - f0p) / 4, ar, f0p, 1, 0);
      dense<SrcT, EPI_DNONE>(L, SrcT{dz}, 4 * HS, s_w + y.b_wg + y.KX, y.KG, nullptr, HS / 4, s_dh, 0, 1, 0);
      __syncthreads();
      // ---- conv encoder: weight gradient and d loss / d window positions
      conv_dw(L, y.cv, ar + f0p * TMP, s_win, scratch, P);
      conv_dx_pos(y.cv, ar + f0p * TMP, s_w + y.b_wc, s_dwin);
      __syncthreads();
      if (tid < valid) state_input_adjoint(sk_, s_din, s_dwin, s_dP, srun, gS, k, h, g.window, tid);
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        if (k > 0) issue_arena(sk - 1);
        else if (tile - (int)gridDim.x >= 0) issue_arena((size_t)(tile - gridDim.x) * h + (h - 1));
      }
    }
  }
}

size_t lstm_fwd_smem_bytes(const LstmLayout& y, int h) {
  return sizeof(float) * (size_t)(y.f_total + pad4(TM * y.LR) + 2 * y.ROWS * TMP + (3 * h + 4) * TMP + 8) + 16;
}
size_t lstm_adj_smem_bytes(const LstmLayout& y, int h) {
  return sizeof(float) * (size_t)(y.b_total + y.ROWS * TMP + pad4(TM * y.LR) + NWARP * (CONV_CH * y.KC + CONV_CH) +
                                  (16 + 3 * h + 2 * LSTM_HS + 4 + 6 * h + 4) * TMP + 8) + 32;
}

cudaError_t launch_lstm_fwd(const LstmLayout& y, const RolloutArgs& a, int grid, cudaStream_t st) {
  const size_t smem = lstm_fwd_smem_bytes(y, a.h);
  cudaError_t e = cudaFuncSetAttribute(lstm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  APG_LAUNCH(grid, NT, smem, st, lstm_fwd_kernel)(y, a);
  return cudaGetLastError();
}

cudaError_t launch_lstm_adj(const LstmLayout& y, const RolloutArgs& a, int grid, cudaStream_t st) {
  const size_t smem = lstm_adj_smem_bytes(y, a.h);
  cudaError_t e = cudaFuncSetAttribute(lstm_adj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  APG_LAUNCH(grid, NT, smem, st, lstm_adj_kernel)(y, a);
  return cudaGetLastError();
}

}  // namespace apg
