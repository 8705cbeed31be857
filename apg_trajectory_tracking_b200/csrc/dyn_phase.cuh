// Thread-per-drone phases of the rollout kernels: the horizon loop over the dynamics (forward: states + loss;
// adjoint: reverse sweep emitting d loss / d logits).  Actions / logit-gradients are exchanged with the policy
// GEMMs through feature-major shared-memory tiles (row = k*A + c, column = drone).
//
// Reference loops restated: scripts/train_drone.py:175-203 (quad), scripts/train_fixed_wing.py:90-116 (wing),
// scripts/train_cartpole.py:127-155 (cartpole).
#pragma once
#include "apg_math.cuh"
#include "tile_engine.cuh"

namespace apg {

// Forward horizon for drone column d of the tile (concurrent mode: all actions known up front).
//   act      : smem, feature-major rows [h*A][TMP]  (already squashed: sigmoid / tanh)
//   cur_g    : this drone's start state (S floats, global);  ref_g: this drone's reference [h][REFW] (global)
//   st_states: global stash of the tile, feature-major [h*S][TMP]  (states AFTER each step)
// returns this drone's loss.
template <template <typename> class SysT>
__device__ __forceinline__ float dyn_forward_conc(const float* __restrict__ act, int d, const float* __restrict__ cur_g,
                                                  const float* __restrict__ ref_g, int h, float dt, const float* pc,
                                                  float* __restrict__ st_states, float* __restrict__ states_out,
                                                  float* __restrict__ actions_out) {
  using Sys = SysT<float>;
  constexpr int S = Sys::S, A = Sys::A, R = Sys::REFW;
  float s[S], s0[S], sn[S], a[A], rf[R > 0 ? R : 1], rf_next[R > 0 ? R : 1];
#pragma unroll
  for (int i = 0; i < S; ++i) s0[i] = s[i] = cur_g[i];
#pragma unroll
  for (int c = 0; c < R; ++c) rf_next[c] = ref_g[c];
  float loss = 0.f;
  for (int k = 0; k < h; ++k) {
#pragma unroll
    for (int c = 0; c < A; ++c) a[c] = act[(k * A + c) * TMP + d];
    // the reference row of step k + 1 is requested now: the h steps of a drone are one dependent chain on two warps per
    // CTA, and a global load issued where it is needed adds its whole latency to every link of that chain
#pragma unroll
    for (int c = 0; c < R; ++c) rf[c] = rf_next[c];
    if (k + 1 < h) {
#pragma unroll
      for (int c = 0; c < R; ++c) rf_next[c] = ref_g[(k + 1) * R + c];
    }
    Sys::step(s, a, dt, pc, sn);
    loss += Sys::loss(sn, rf, a, s0, k, h);
#pragma unroll
    for (int i = 0; i < S; ++i) {
      s[i] = sn[i];
      st_states[(k * S + i) * TMP + d] = sn[i];
    }
    if (states_out) {
#pragma unroll
      for (int i = 0; i < S; ++i) states_out[k * S + i] = sn[i];
    }
    if (actions_out) {
#pragma unroll
      for (int c = 0; c < A; ++c) actions_out[k * A + c] = a[c];
    }
  }
  return loss;
}

// Reverse sweep for drone column d.  Reads the stashed actions / states of the tile (global, feature-major),
// writes d loss / d logit into the smem tile dlog (rows [h*A][TMP]).
template <template <typename> class SysT>
__device__ __forceinline__ void dyn_adjoint_conc(const float* __restrict__ st_act, const float* __restrict__ st_states,
                                                 int d, const float* __restrict__ cur_g, const float* __restrict__ ref_g,
                                                 int h, float dt, const float* pc, float* __restrict__ dlog) {
  using Sys = SysT<float>;
  constexpr int S = Sys::S, A = Sys::A, R = Sys::REFW;
  float s0[S], sk[S], sn[S], a[A], rf[R > 0 ? R : 1], g[S], gs[S], ga[A], ga2[A];
  float a_nx[A], rf_nx[R > 0 ? R : 1], sk_nx[S];         // operands of the NEXT (earlier) step, requested one step ahead
#pragma unroll
  for (int i = 0; i < S; ++i) { s0[i] = cur_g[i]; g[i] = 0.f; }
#pragma unroll
  for (int i = 0; i < S; ++i) sn[i] = st_states[((h - 1) * S + i) * TMP + d];
  auto request = [&](int k) {                             // stashed action, reference row and state BEFORE step k
#pragma unroll
    for (int c = 0; c < A; ++c) a_nx[c] = st_act[(k * A + c) * TMP + d];
#pragma unroll
    for (int c = 0; c < R; ++c) rf_nx[c] = ref_g[k * R + c];
    if (k > 0) {
#pragma unroll
      for (int i = 0; i < S; ++i) sk_nx[i] = st_states[((k - 1) * S + i) * TMP + d];
    } else {
#pragma unroll
      for (int i = 0; i < S; ++i) sk_nx[i] = s0[i];
    }
  };
  request(h - 1);
  for (int k = h - 1; k >= 0; --k) {
#pragma unroll
    for (int c = 0; c < A; ++c) { a[c] = a_nx[c]; ga[c] = 0.f; }
#pragma unroll
    for (int c = 0; c < R; ++c) rf[c] = rf_nx[c];
#pragma unroll
    for (int i = 0; i < S; ++i) sk[i] = sk_nx[i];
    // (the reverse sweep of a drone is one dependent chain on two warps per CTA: loads issued where they are needed
    // add their whole latency to every link of it)
    if (k > 0) request(k - 1);
    Sys::loss_grad(sn, rf, a, s0, k, h, g, ga);          // g += dl_k/ds_{k+1}, ga += dl_k/da_k
    Sys::step_adj(sk, a, dt, pc, g, gs, ga2);             // through the step
#pragma unroll
    for (int c = 0; c < A; ++c) {
      const float da = ga[c] + ga2[c];
      const float dsq = Sys::SIGMOID_ACTIONS ? a[c] * (1.f - a[c]) : (1.f - a[c] * a[c]);
      dlog[(k * A + c) * TMP + d] = da * dsq;
    }
#pragma unroll
    for (int i = 0; i < S; ++i) { g[i] = gs[i]; sn[i] = sk[i]; }
  }
}

}  // namespace apg
