// Closed-loop evaluation rollout on table references: the batched, no-grad counterpart of
// QuadEvaluator.follow_trajectory("rand") (scripts/evaluate_drone.py:81-194).  Every drone walks its own reference
// table: window -> QuadDataset.prepare_data -> hutter policy (first predicted action) -> dynamics step ->
// divergence / stability -> stop or reset, for up to `steps` steps in ONE launch.  Same tile engine as the training
// kernels (64-drone tiles, weights resident in shared memory, 3xTF32 tensor path), no stash, no loss.
#include "eval_math.cuh"
#include "hutter_policy.cuh"
#include "layouts.h"
#include "tile_engine.cuh"
#include "kernels.h"

namespace apg {

struct EvalArgs {
  const float* wf;            // packed forward weights (apg_pack_kernel)
  const float* tables;        // [n_tables][RL][9]
  const int* table_index;     // [N] table of each drone (nullptr: drone i uses table i)
  const float* init_states;   // [N][12]
  int N, h;
  float dt;
  PhysConsts pc;
  EvalParams ev;
  float* states_out;          // optional [N][steps+1][12]
  float* div_out;             // optional [N][steps]
  float* actions_out;         // optional [N][steps][4]
  int* n_steps_out;           // optional [N]
  const float* h0c0;          // LSTM policy: [2][N][8] hidden / cell state before the first policy call
  float* hc_out;              // LSTM policy, optional: [2][N][8] after the last one
};

__global__ void __launch_bounds__(NT, 1) eval_rollout_kernel(const HutterLayout y, const EvalArgs g) {
  APG_DYNAMIC_SMEM_F32(smem);
  using Sys = Quad<float>;
  constexpr int S = Sys::S, A = Sys::A;
  const int h = g.h, RL = g.ev.table_rows, steps = g.ev.steps;
  float* s_w = smem;
  float* s_ins = s_w + y.f_total;
  float* s_win = s_ins + pad4(TM * y.F0);
  float* s_x1 = s_win + pad4(TM * y.LR);
  float* s_h = s_x1 + y.XR * TMP;
  float* s_mf = s_h + HID * TMP;                          // [6][TM] drone position / velocity
  int* s_mi = reinterpret_cast<int*>(s_mf + 6 * TM);      // [2][TM] window start / real rows
  const float** s_tab = reinterpret_cast<const float**>(s_mi + 2 * TM);   // [TM] table of each drone
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(s_tab + TM);
  float* s_act = s_x1 + HID * TMP;
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
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int valid = min(TM, g.N - tile * TM);
    const size_t drone = (size_t)tile * TM + tid;
    const bool mine = tid < valid;                        // this thread owns a drone
    float s[S];
    int ci = 0, alive = mine ? 1 : 0, nsteps = 0;
    const float* tab = nullptr;
    if (mine) {
      tab = g.tables + (size_t)(g.table_index ? g.table_index[drone] : (int)drone) * RL * 9;
#pragma unroll
      for (int i = 0; i < S; ++i) s[i] = g.init_states[drone * S + i];
      if (g.states_out) {
#pragma unroll
        for (int i = 0; i < S; ++i) g.states_out[drone * (steps + 1) * S + i] = s[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < S; ++i) s[i] = 0.f;
    }
    if (tid < TM) s_tab[tid] = tab;
    for (int i = 0; i < steps; ++i) {
      if (tid < TM) {
        int start = 0, nreal = 0, ci_next = ci;
        if (alive) eval_window_plan(ci, RL, h, &start, &nreal, &ci_next);
        ci = ci_next;
        float c0[S], f[15];
        c0[0] = c0[1] = c0[2] = 0.f;
#pragma unroll
        for (int j = 3; j < S; ++j) c0[j] = s[j];
        Sys::features(c0, f);
#pragma unroll
        for (int j = 0; j < 15; ++j) s_ins[tid * y.F0 + j] = f[j];
#pragma unroll
        for (int c = 0; c < 3; ++c) { s_mf[c * TM + tid] = s[c]; s_mf[(3 + c) * TM + tid] = s[6 + c]; }
        s_mi[tid] = start;
        s_mi[TM + tid] = alive ? nreal : -1;              // -1: no live drone in this slot -> zero window
      }
      __syncthreads();
      for (int idx = tid; idx < TM * y.LR; idx += NT) {
        const int d = idx / y.LR, e = idx - d * y.LR;
        const int r = e / 9, c = e - r * 9;
        float v = 0.f;
        const int nreal = s_mi[TM + d];
        if (nreal >= 0) {
          const float pos_c = c < 3 ? s_mf[c * TM + d] : 0.f;
          const float vel_c = c >= 6 ? s_mf[(c - 3) * TM + d] : 0.f;      // rows 3..5 of s_mf hold the velocity
          v = eval_in_ref_elem(s_tab[d], RL, s_mi[d], nreal, r, c, pos_c, vel_c);
        }
        s_win[idx] = v;
      }
      __syncthreads();
      hutter_first_layer<true>(L, y, s_w, s_ins, s_win, s_x1);
      __syncthreads();
      dense_auto<EPI_ACT>(L, s_x1, y.K1, s_w + y.f_w1, HID, mma_sw(HID), s_w + y.f_b1, HID, s_h, 0, ACT_TANH);
      __syncthreads();
      dense_auto<EPI_ACT>(L, s_h, HID, s_w + y.f_w2, HID, mma_sw(HID), s_w + y.f_b2, HID, s_x1, 0, ACT_TANH);
      __syncthreads();
      dense_auto<EPI_ACT>(L, s_x1, HID, s_w + y.f_w3, HID, mma_sw(HID), s_w + y.f_b3, HID, s_h, 0, ACT_TANH);
      __syncthreads();
      dense_auto<EPI_ACT>(L, s_h, HID, s_w + y.f_wo, y.ld_fwo, mma_sw(y.ld_fwo), s_w + y.f_bo, y.Mo4, s_act, 0,
                          ACT_SIGMOID);
      __syncthreads();
      if (alive) {
        float a[A], sn[S];
#pragma unroll
        for (int c = 0; c < A; ++c) a[c] = fminf(fmaxf(s_act[c * TMP + tid], 0.f), 1.f);   // first predicted action
        Sys::step(s, a, g.dt, g.pc.v, sn);
        if (g.states_out) {
#pragma unroll
          for (int j = 0; j < S; ++j) g.states_out[(drone * (steps + 1) + i + 1) * S + j] = sn[j];
        }
        if (g.actions_out) {
#pragma unroll
          for (int c = 0; c < A; ++c) g.actions_out[(drone * steps + i) * A + c] = a[c];
        }
        const float div = eval_post_step(sn, tab, ci, g.ev, &alive);
        if (g.div_out) g.div_out[drone * steps + i] = div;
#pragma unroll
        for (int j = 0; j < S; ++j) s[j] = sn[j];
        ++nsteps;
        if (i >= RL) alive = 0;                           // evaluate_drone.py:187-188
      }
      // barrier before the next step overwrites s_ins / s_act; also the tile-wide early exit
      if (!__syncthreads_or(alive)) break;
    }
    if (mine && g.n_steps_out) g.n_steps_out[drone] = nsteps;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------
// The same closed loop with the LSTM policy (train_mode "LSTM": models/rnn.py LSTM_NEW, conv encoder ->
// LSTMCell(175, 8) -> Linear(8, 4); evaluate_drone.py:156-157 applies the net's 4 outputs).  Every policy call
// advances the drone's hidden / cell state (rnn.py:45-48); a reset of the drone does not touch it (the reference
// re-draws it only when an evaluator is constructed, evaluate_drone.py:55-57), a stopped drone keeps its last one.
// One arena of LstmLayout rows per tile (no stash): h', c' of step i are copied to the h_prev / c_prev rows of
// step i + 1 by the thread that owns the drone.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float eval_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(NT, 1) eval_rollout_lstm_kernel(const LstmLayout y, const EvalArgs g) {
  APG_DYNAMIC_SMEM_F32(smem);
  using Sys = Quad<float>;
  constexpr int S = Sys::S, A = Sys::A, HS = LSTM_HS;
  const int h = g.h, RL = g.ev.table_rows, steps = g.ev.steps;
  float* s_w = smem;
  float* s_win = s_w + y.f_total;
  float* ar = s_win + pad4(TM * y.LR);                    // [ROWS][TMP] activation arena of the current step
  float* s_mf = ar + y.ROWS * TMP;                        // [6][TM] drone position / velocity
  int* s_mi = reinterpret_cast<int*>(s_mf + 6 * TM);      // [2][TM] window start / real rows
  const float** s_tab = reinterpret_cast<const float**>(s_mi + 2 * TM);
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(s_tab + TM);
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
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int valid = min(TM, g.N - tile * TM);
    const size_t drone = (size_t)tile * TM + tid;
    const bool mine = tid < valid;
    float s[S], hp[HS], cp[HS];
    int ci = 0, alive = mine ? 1 : 0, nsteps = 0;
    const float* tab = nullptr;
#pragma unroll
    for (int u = 0; u < HS; ++u) hp[u] = cp[u] = 0.f;
    if (mine) {
      tab = g.tables + (size_t)(g.table_index ? g.table_index[drone] : (int)drone) * RL * 9;
#pragma unroll
      for (int i = 0; i < S; ++i) s[i] = g.init_states[drone * S + i];
#pragma unroll
      for (int u = 0; u < HS; ++u) {
        hp[u] = g.h0c0[drone * HS + u];
        cp[u] = g.h0c0[((size_t)g.N + drone) * HS + u];
      }
      if (g.states_out) {
#pragma unroll
        for (int i = 0; i < S; ++i) g.states_out[drone * (steps + 1) * S + i] = s[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < S; ++i) s[i] = 0.f;
    }
    if (tid < TM) s_tab[tid] = tab;
    for (int i = 0; i < steps; ++i) {
      if (tid < TM) {
        int start = 0, nreal = 0, ci_next = ci;
        if (alive) eval_window_plan(ci, RL, h, &start, &nreal, &ci_next);
        ci = ci_next;
        float c0[S], f[15];
        c0[0] = c0[1] = c0[2] = 0.f;
#pragma unroll
        for (int j = 3; j < S; ++j) c0[j] = s[j];
        Sys::features(c0, f);
#pragma unroll
        for (int j = 0; j < 15; ++j) ar[j * TMP + tid] = f[j];
        for (int j = y.F0; j < pad4(y.F0); ++j) ar[j * TMP + tid] = 0.f;
#pragma unroll
        for (int u = 0; u < HS; ++u) {
          ar[(y.R_HP + u) * TMP + tid] = hp[u];
          ar[(y.R_CP + u) * TMP + tid] = cp[u];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) { s_mf[c * TM + tid] = s[c]; s_mf[(3 + c) * TM + tid] = s[6 + c]; }
        s_mi[tid] = start;
        s_mi[TM + tid] = alive ? nreal : -1;
      }
      __syncthreads();
      for (int idx = tid; idx < TM * y.LR; idx += NT) {
        const int d = idx / y.LR, e = idx - d * y.LR;
        const int r = e / 9, c = e - r * 9;
        float v = 0.f;
        const int nreal = s_mi[TM + d];
        if (nreal >= 0) {
          const float pos_c = c < 3 ? s_mf[c * TM + d] : 0.f;
          const float vel_c = c >= 6 ? s_mf[(c - 3) * TM + d] : 0.f;
          v = eval_in_ref_elem(s_tab[d], RL, s_mi[d], nreal, r, c, pos_c, vel_c);
        }
        s_win[idx] = v;
      }
      __syncthreads();
      conv_layer_fwd(L, y.npos, y.RD, y.LR, y.KC, s_w + y.f_wc, s_w + y.f_bc, s_win, ar, pad4(y.F0), y.npos, 1);
      __syncthreads();
      dense<SrcT, EPI_ACT>(L, SrcT{ar}, y.KG, s_w + y.f_wg, 4 * HS, nullptr, HS, ar, y.R_G, 1, ACT_NONE);
      __syncthreads();
      for (int idx = tid; idx < TM * HS; idx += NT) {
        const int u = idx / TM, d = idx - u * TM;
        const float* bi = s_w + y.f_bih;
        const float* bh = s_w + y.f_bhh;
        const float* gp = ar + y.R_G * TMP + d;
        const float ig = eval_sigmoid(gp[(u) * TMP] + bi[u] + bh[u]);
        const float fg = eval_sigmoid(gp[(HS + u) * TMP] + bi[HS + u] + bh[HS + u]);
        const float gg = tanhf(gp[(2 * HS + u) * TMP] + bi[2 * HS + u] + bh[2 * HS + u]);
        const float og = eval_sigmoid(gp[(3 * HS + u) * TMP] + bi[3 * HS + u] + bh[3 * HS + u]);
        const float cn = fg * ar[(y.R_CP + u) * TMP + d] + ig * gg;
        ar[(y.R_C + u) * TMP + d] = cn;
        ar[(y.R_H + u) * TMP + d] = og * tanhf(cn);
      }
      __syncthreads();
      {   // fc_out + sigmoid: one (drone, action) per thread
        const int c = tid / TM, d = tid - c * TM;
        float acc = s_w[y.f_bo + c];
#pragma unroll
        for (int u = 0; u < HS; ++u) acc = fmaf(ar[(y.R_H + u) * TMP + d], s_w[y.f_wo + u * pad4(y.Mo) + c], acc);
        ar[(y.R_A + c) * TMP + d] = eval_sigmoid(acc);
      }
      __syncthreads();
      if (alive) {
        float a[A], sn[S];
#pragma unroll
        for (int c = 0; c < A; ++c) a[c] = fminf(fmaxf(ar[(y.R_A + c) * TMP + tid], 0.f), 1.f);
#pragma unroll
        for (int u = 0; u < HS; ++u) { hp[u] = ar[(y.R_H + u) * TMP + tid]; cp[u] = ar[(y.R_C + u) * TMP + tid]; }
        Sys::step(s, a, g.dt, g.pc.v, sn);
        if (g.states_out) {
#pragma unroll
          for (int j = 0; j < S; ++j) g.states_out[(drone * (steps + 1) + i + 1) * S + j] = sn[j];
        }
        if (g.actions_out) {
#pragma unroll
          for (int c = 0; c < A; ++c) g.actions_out[(drone * steps + i) * A + c] = a[c];
        }
        const float div = eval_post_step(sn, tab, ci, g.ev, &alive);
        if (g.div_out) g.div_out[drone * steps + i] = div;
#pragma unroll
        for (int j = 0; j < S; ++j) s[j] = sn[j];
        ++nsteps;
        if (i >= RL) alive = 0;
      }
      if (!__syncthreads_or(alive)) break;
    }
    if (mine) {
      if (g.n_steps_out) g.n_steps_out[drone] = nsteps;
      if (g.hc_out) {
#pragma unroll
        for (int u = 0; u < HS; ++u) {
          g.hc_out[drone * HS + u] = hp[u];
          g.hc_out[((size_t)g.N + drone) * HS + u] = cp[u];
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------
// fixed wing: FixedWingEvaluator.fly_to_point (scripts/evaluate_fixed_wing.py:46-130).  Per step: WingDataset.
// prepare_data on (observed state, current target) -> hutter "linear ref" policy -> wing dynamics step -> divergence
// to the line towards the target, target switching, stop or reset (per-drone logic in eval_math.cuh).
// ------------------------------------------------------------------------------------------------------------
struct WingEvalArgs {
  const float* wf;
  const float* targets;       // [N][K][3]
  const float* init_states;   // [N][12]
  int N;
  float dt;                   // environment step
  PhysConsts pc;
  NormConsts nc;              // dataset mean / std
  WingEvalParams ev;
  float* states_out;          // optional [N][steps+1][12]
  float* div_out;             // optional [N][steps]
  float* actions_out;         // optional [N][steps][4]
  int* n_steps_out;           // optional [N]
  float* dt_sum_out;          // optional [N]  sum of the div_target list
  float* dt_cnt_out;          // optional [N]  its length
};

__global__ void __launch_bounds__(NT, 1) eval_wing_kernel(const HutterLayout y, const WingEvalArgs g) {
  APG_DYNAMIC_SMEM_F32(smem);
  using Sys = Wing<float>;
  constexpr int S = Sys::S, A = Sys::A;
  const int steps = g.ev.steps, K = g.ev.n_targets;
  float* s_w = smem;
  float* s_ins = s_w + y.f_total;
  float* s_inr = s_ins + pad4(TM * y.F0);
  float* s_x1 = s_inr + pad4(TM * y.LR);
  float* s_h = s_x1 + y.XR * TMP;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(s_h + HID * TMP);
  float* s_act = s_x1 + HID * TMP;
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
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int valid = min(TM, g.N - tile * TM);
    const size_t drone = (size_t)tile * TM + tid;
    const bool mine = tid < valid;
    WingEvalDrone D;
    const float* tg = nullptr;
    {
      float init[S];
#pragma unroll
      for (int j = 0; j < S; ++j) init[j] = mine ? g.init_states[drone * S + j] : 0.f;
      wing_eval_init(D, init, mine ? 1 : 0);
    }
    if (mine) {
      tg = g.targets + drone * K * 3;
      if (g.states_out) {
#pragma unroll
        for (int j = 0; j < S; ++j) g.states_out[drone * (steps + 1) * S + j] = D.env[j];
      }
    }
    for (int i = 0; i < steps; ++i) {
      if (tid < TM) {
        float f[9], r3[3];
        if (D.alive) {
          float t3[3];
#pragma unroll
          for (int j = 0; j < 3; ++j) t3[j] = tg[D.ti * 3 + j];
          WingPrep<float>::drone(D.obs, t3, g.nc.mean, g.nc.std_, g.ev.vlen, g.ev.h, f, r3);
        } else {
#pragma unroll
          for (int j = 0; j < 9; ++j) f[j] = 0.f;
          r3[0] = r3[1] = r3[2] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 9; ++j) s_ins[tid * y.F0 + j] = f[j];
#pragma unroll
        for (int j = 0; j < 3; ++j) s_inr[tid * y.LR + j] = r3[j];
      }
      __syncthreads();
      hutter_first_layer<false>(L, y, s_w, s_ins, s_inr, s_x1);
      __syncthreads();
      dense_auto<EPI_ACT>(L, s_x1, y.K1, s_w + y.f_w1, HID, mma_sw(HID), s_w + y.f_b1, HID, s_h, 0, ACT_TANH);
      __syncthreads();
      dense_auto<EPI_ACT>(L, s_h, HID, s_w + y.f_w2, HID, mma_sw(HID), s_w + y.f_b2, HID, s_x1, 0, ACT_TANH);
      __syncthreads();
      dense_auto<EPI_ACT>(L, s_x1, HID, s_w + y.f_w3, HID, mma_sw(HID), s_w + y.f_b3, HID, s_h, 0, ACT_TANH);
      __syncthreads();
      dense_auto<EPI_ACT>(L, s_h, HID, s_w + y.f_wo, y.ld_fwo, mma_sw(y.ld_fwo), s_w + y.f_bo, y.Mo4, s_act, 0,
                          ACT_SIGMOID);
      __syncthreads();
      if (D.alive) {
        float a[A], nxt[S];
#pragma unroll
        for (int c = 0; c < A; ++c) a[c] = s_act[c * TMP + tid];            // first of the h predicted actions
        Sys::step(D.env, a, g.dt, g.pc.v, nxt);
        if (g.states_out) {
#pragma unroll
          for (int j = 0; j < S; ++j) g.states_out[(drone * (steps + 1) + i + 1) * S + j] = nxt[j];
        }
        if (g.actions_out) {
#pragma unroll
          for (int c = 0; c < A; ++c) g.actions_out[(drone * steps + i) * A + c] = a[c];
        }
        const float div = wing_eval_post_step(D, nxt, tg, g.ev);
        if (g.div_out) g.div_out[drone * steps + i] = div;
      }
      if (!__syncthreads_or(D.alive)) break;
    }
    if (mine) {
      wing_eval_finish(D, g.ev);
      if (g.n_steps_out) g.n_steps_out[drone] = D.nsteps;
      if (g.dt_sum_out) g.dt_sum_out[drone] = D.dt_sum;
      if (g.dt_cnt_out) g.dt_cnt_out[drone] = D.dt_cnt;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------
// cartpole: Evaluator.evaluate_in_environment (scripts/evaluate_cartpole.py:78-262) with CartpoleWrapper -- how long
// the simple Net (4->32->64->64->32->h, tanh everywhere) keeps the pole upright.  Per step: policy on the state with
// column 0 zeroed (first of the h predicted actions) -> CartPoleEnv._step -> upright test (per-drone logic in
// eval_math.cuh, including the in-place zeroing of the environment's cart position).
// ------------------------------------------------------------------------------------------------------------
struct CartpoleEvalArgs {
  const float* wf;            // packed forward weights (apg_pack_kernel, SimpleLayout)
  const float* init_states;   // [N][4]
  int N;
  float dt;
  PhysConsts pc;
  CartpoleEvalParams ev;
  float* states_out;          // optional [N][steps][4]  (states returned by _step)
  float* actions_out;         // optional [N][steps]
  int* n_steps_out;           // optional [N]
  float* angle_sum_out;       // optional [N]  sum of |theta| for step index > burn_in
  float* angle_cnt_out;       // optional [N]  number of terms in it
  float* vel_sum_out;         // optional [N]  sum of |x_dot| over the steps taken
};

__global__ void __launch_bounds__(NT, 1) eval_cartpole_kernel(const SimpleLayout y, const CartpoleEvalArgs g) {
  APG_DYNAMIC_SMEM_F32(smem);
  using Sys = Cartpole<float>;
  constexpr int S = Sys::S;
  const int steps = g.ev.steps;
  float* s_w = smem;
  float* s_in = s_w + y.f_total;
  float* s_a = s_in + pad4(TM * y.F0);               // activation arena [rows_total][TMP]
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(s_a + y.rows_total * TMP);
  const Lane L;
  const int tid = threadIdx.x;
  const int ntiles = (g.N + TM - 1) / TM;
  const int LAST = SIMPLE_NL - 1;
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
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int valid = min(TM, g.N - tile * TM);
    const size_t drone = (size_t)tile * TM + tid;
    const bool mine = tid < valid;
    CartpoleEvalDrone D;
    {
      float init[S];
#pragma unroll
      for (int j = 0; j < S; ++j) init[j] = mine ? g.init_states[drone * S + j] : 0.f;
      cartpole_eval_init(D, init, mine ? 1 : 0);
    }
    for (int i = 0; i < steps; ++i) {
      if (tid < TM) {
        if (D.alive) cartpole_eval_before_policy(D, i);
        s_in[tid * y.F0 + 0] = 0.f;                                      // simple_model.py:21
#pragma unroll
        for (int j = 1; j < S; ++j) s_in[tid * y.F0 + j] = D.alive ? D.s[j] : 0.f;
      }
      __syncthreads();
      dense<SrcAoS, EPI_ACT>(L, SrcAoS{s_in, y.F0, 0}, y.din[0], s_w + y.f_w[0], y.ldf[0], s_w + y.f_b[0],
                             y.ldf[0] / 4, s_a, y.row[0], 1, ACT_TANH);
      __syncthreads();
      for (int l = 1; l < SIMPLE_NL; ++l) {
        dense<SrcT, EPI_ACT>(L, SrcT{s_a + y.row[l - 1] * TMP}, y.din[l], s_w + y.f_w[l], y.ldf[l], s_w + y.f_b[l],
                             y.ldf[l] / 4, s_a, y.row[l], 1, ACT_TANH);
        __syncthreads();
      }
      if (D.alive) {
        float a[1], nxt[S];
        a[0] = s_a[y.row[LAST] * TMP + tid];                              // first of the h predicted actions
        Sys::step(D.s, a, g.dt, g.pc.v, nxt);
        cartpole_eval_post_step(D, nxt, i, g.ev);
        if (g.states_out) {
#pragma unroll
          for (int j = 0; j < S; ++j) g.states_out[(drone * steps + i) * S + j] = nxt[j];
        }
        if (g.actions_out) g.actions_out[drone * steps + i] = a[0];
      }
      // barrier before the next step overwrites s_in / the arena; also the tile-wide early exit
      if (!__syncthreads_or(D.alive)) break;
    }
    if (mine) {
      if (g.n_steps_out) g.n_steps_out[drone] = D.nsteps;
      if (g.angle_sum_out) g.angle_sum_out[drone] = D.ang_sum;
      if (g.angle_cnt_out) g.angle_cnt_out[drone] = D.ang_cnt;
      if (g.vel_sum_out) g.vel_sum_out[drone] = D.vel_sum;
    }
    __syncthreads();
  }
}

size_t eval_cartpole_smem_bytes(const SimpleLayout& y) {
  return sizeof(float) * (size_t)(y.f_total + pad4(TM * y.F0) + y.rows_total * TMP) + 16;
}

cudaError_t launch_eval_cartpole(const SimpleLayout& y, const float* wf, const float* init_states, int n, float dt,
                                 const PhysConsts& pc, const CartpoleEvalParams& ev, float* states_out,
                                 float* actions_out, int* n_steps_out, float* angle_sum_out, float* angle_cnt_out,
                                 float* vel_sum_out, int grid, cudaStream_t st) {
  CartpoleEvalArgs a;
  a.wf = wf; a.init_states = init_states; a.N = n; a.dt = dt; a.pc = pc; a.ev = ev;
  a.states_out = states_out; a.actions_out = actions_out; a.n_steps_out = n_steps_out;
  a.angle_sum_out = angle_sum_out; a.angle_cnt_out = angle_cnt_out; a.vel_sum_out = vel_sum_out;
  const size_t smem = eval_cartpole_smem_bytes(y);
  cudaError_t e = cudaFuncSetAttribute(eval_cartpole_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  APG_LAUNCH(grid, NT, smem, st, eval_cartpole_kernel)(y, a);
  return cudaGetLastError();
}

size_t eval_wing_smem_bytes(const HutterLayout& y) {
  return sizeof(float) * (size_t)(y.f_total + pad4(TM * y.F0) + pad4(TM * y.LR) + y.XR * TMP + HID * TMP) + 16;
}

cudaError_t launch_eval_wing(const HutterLayout& y, const float* wf, const float* targets, const float* init_states,
                             int n, float dt_env, const PhysConsts& pc, const float* mean_host, const float* std_host,
                             const WingEvalParams& ev, float* states_out, float* div_out, float* actions_out,
                             int* n_steps_out, float* dt_sum_out, float* dt_cnt_out, int grid, cudaStream_t st) {
  WingEvalArgs a;
  a.wf = wf; a.targets = targets; a.init_states = init_states; a.N = n; a.dt = dt_env; a.pc = pc; a.ev = ev;
  for (int j = 0; j < 12; ++j) { a.nc.mean[j] = mean_host[j]; a.nc.std_[j] = std_host[j]; }
  a.states_out = states_out; a.div_out = div_out; a.actions_out = actions_out; a.n_steps_out = n_steps_out;
  a.dt_sum_out = dt_sum_out; a.dt_cnt_out = dt_cnt_out;
  const size_t smem = eval_wing_smem_bytes(y);
  cudaError_t e = cudaFuncSetAttribute(eval_wing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  APG_LAUNCH(grid, NT, smem, st, eval_wing_kernel)(y, a);
  return cudaGetLastError();
}

size_t eval_smem_bytes(const HutterLayout& y) {
  return sizeof(float) * (size_t)(y.f_total + pad4(TM * y.F0) + pad4(TM * y.LR) + y.XR * TMP + HID * TMP + 6 * TM) +
         sizeof(int) * 2 * TM + sizeof(void*) * TM + 16;
}

cudaError_t launch_eval_rollout(const HutterLayout& y, const float* wf, const float* tables, const int* table_index,
                                const float* init_states, int n, float dt, const PhysConsts& pc, const EvalParams& ev,
                                float* states_out, float* div_out, float* actions_out, int* n_steps_out, int grid,
                                cudaStream_t st) {
  EvalArgs a;
  a.wf = wf; a.tables = tables; a.table_index = table_index; a.init_states = init_states;
  a.N = n; a.h = y.L; a.dt = dt; a.pc = pc; a.ev = ev;
  a.states_out = states_out; a.div_out = div_out; a.actions_out = actions_out; a.n_steps_out = n_steps_out;
  a.h0c0 = nullptr; a.hc_out = nullptr;
  const size_t smem = eval_smem_bytes(y);
  cudaError_t e = cudaFuncSetAttribute(eval_rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  APG_LAUNCH(grid, NT, smem, st, eval_rollout_kernel)(y, a);
  return cudaGetLastError();
}

size_t eval_lstm_smem_bytes(const LstmLayout& y) {
  return sizeof(float) * (size_t)(y.f_total + pad4(TM * y.LR) + y.ROWS * TMP + 6 * TM) + sizeof(int) * 2 * TM +
         sizeof(void*) * TM + 16;
}

cudaError_t launch_eval_rollout_lstm(const LstmLayout& y, const float* wf, const float* h0c0, const float* tables,
                                     const int* table_index, const float* init_states, int n, float dt,
                                     const PhysConsts& pc, const EvalParams& ev, float* states_out, float* div_out,
                                     float* actions_out, int* n_steps_out, float* hc_out, int grid, cudaStream_t st) {
  EvalArgs a;
  a.wf = wf; a.tables = tables; a.table_index = table_index; a.init_states = init_states;
  a.N = n; a.h = y.L; a.dt = dt; a.pc = pc; a.ev = ev;
  a.states_out = states_out; a.div_out = div_out; a.actions_out = actions_out; a.n_steps_out = n_steps_out;
  a.h0c0 = h0c0; a.hc_out = hc_out;
  const size_t smem = eval_lstm_smem_bytes(y);
  cudaError_t e = cudaFuncSetAttribute(eval_rollout_lstm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  APG_LAUNCH(grid, NT, smem, st, eval_rollout_lstm_kernel)(y, a);
  return cudaGetLastError();
}


}  // namespace apg
