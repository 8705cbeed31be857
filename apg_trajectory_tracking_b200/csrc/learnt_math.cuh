// Per-drone math of the learnt residual quadrotor dynamics:
//   next = simulate_quadrotor(linear_at @ action, state, dt) + linear_state_2(relu(linear_state_1([state, at])))
// forward and hand-written adjoint w.r.t. state, action and every parameter.  `__host__ __device__`, templated on the
// scalar type like apg_math.cuh (kernels: csrc/learnt_kernels.cu; CPU check: tests/hostcheck/hostcheck_learnt.cpp).
//
// Reference: neural_control/dynamics/quad_dynamics_trained.py:10-69 (LearntDynamics).  Its quirks are kept:
//   * `torch_kinv_ang_vel_tau` / `torch_inertia_J` are built ONCE from the parameters in __init__ (:47-48), so the
//     simulator keeps using the construction-time values (here: the `phys` constants) while the parameters still
//     RECEIVE gradients: d/d kinv_i = dt (br_i - w_i) g_{9+i};  d/d J_i = -dt rot_drag_i / J_i^2 g_{9+i} (J cancels
//     otherwise);  d/d mass = 0 (mass * thrust / mass).
// Flat parameter vector (named_parameters() order): linear_at [4][4] | mass | inertia (3) | kinv (3) |
//   linear_state_1.weight [64][16] | .bias (64) | linear_state_2.weight [12][64] | .bias (12)
#pragma once
#include "apg_math.cuh"

namespace apg {

struct LearntLayout {
  static constexpr int XD = 16, HD = 64, SD = 12, AD = 4;
  static constexpr int O_LAT = 0, O_MASS = 16, O_J = 17, O_K = 20, O_W1 = 23, O_B1 = O_W1 + HD * XD,
                       O_W2 = O_B1 + HD, O_B2 = O_W2 + SD * HD, NP = O_B2 + SD;      // 1891
};

// Factor rows the adjoint kernel leaves per drone (one shared-memory row per factor component) and the mapping from
// a flat parameter-gradient entry to the two rows whose dot product over the drones it is.  Generic in the number of
// "physical" parameters NPH that precede the residual MLP in the flat vector: each of them is a per-drone scalar
// cotangent (row R_DP + e) summed over the drones (dot with the row of ones).
template <int NPH>
struct LearntRows {
  static constexpr int XD = 16, HD = 64, SD = 12;
  static constexpr int R_G = 0, R_H = R_G + SD, R_DH = R_H + HD, R_X = R_DH + HD, R_DP = R_X + XD, R_ONE = R_DP + NPH,
                       R_TOTAL = R_ONE + 1;
  static constexpr int O_W1 = NPH, O_B1 = O_W1 + HD * XD, O_W2 = O_B1 + HD, O_B2 = O_W2 + SD * HD, NP = O_B2 + SD;
  // entry e of the flat parameter gradient = dot(row ra, row rb) over the drones
  APG_HD static void entry_rows(int e, int* ra, int* rb) {
    if (e < O_W1) { *ra = R_DP + e; *rb = R_ONE; }
    else if (e < O_B1) { const int q = e - O_W1; *ra = R_DH + q / XD; *rb = R_X + q % XD; }
    else if (e < O_W2) { *ra = R_DH + (e - O_B1); *rb = R_ONE; }
    else if (e < O_B2) { const int q = e - O_W2; *ra = R_G + q / HD; *rb = R_H + q % HD; }
    else { *ra = R_G + (e - O_B2); *rb = R_ONE; }
  }
};

template <typename T>
struct LearntQuad {
  using Y = LearntLayout;

  // at = L a
  APG_HD static void transform_action(const T* P, const T* a, T* at) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      T v = 0;
#pragma unroll
      for (int c = 0; c < 4; ++c) v += P[Y::O_LAT + r * 4 + c] * a[c];
      at[r] = v;
    }
  }
  // hidden pre-activation j of the residual MLP on x = [s, at]
  APG_HD static T hidden(const T* P, const T* s, const T* at, int j) {
    T v = P[Y::O_B1 + j];
    const T* w = P + Y::O_W1 + j * Y::XD;
#pragma unroll
    for (int k = 0; k < 12; ++k) v += w[k] * s[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) v += w[12 + k] * at[k];
    return v > T(0) ? v : T(0);
  }
  // out = step(s, at) + W2 h + b2.  h: caller-provided storage of HD values with stride hs (registers or a smem column)
  APG_HD static void forward(const T* P, const float* pc, const T* s, const T* a, T dt, T* out, T* at, T* h, int hs) {
    transform_action(P, a, at);
    Quad<T>::step(s, at, dt, pc, out);
    for (int j = 0; j < Y::HD; ++j) h[j * hs] = hidden(P, s, at, j);
#pragma unroll
    for (int i = 0; i < Y::SD; ++i) {
      T v = P[Y::O_B2 + i];
      const T* w = P + Y::O_W2 + i * Y::HD;
      for (int j = 0; j < Y::HD; ++j) v += w[j] * h[j * hs];
      out[i] += v;
    }
  }
  // Adjoint for one drone given the cotangent g (12) of `out`, at and h of the forward.
  //   gs (12), ga (4): cotangents of state / action
  //   dh (HD, stride hs): cotangent of the hidden pre-activations   -> dW1 = sum dh (x) [s, at], db1 = sum dh
  //   gat (4): cotangent of at                                       -> dL = sum gat (x) a
  //   dk (3), dj (3): cotangents of the kinv / inertia vectors;  dW2 = sum g (x) h, db2 = sum g
  APG_HD static void adjoint(const T* P, const float* pc, const T* s, const T* a, const T* at, const T* h, int hs,
                             T dt, const T* g, T* gs, T* ga, T* dh, T* gat, T* dk, T* dj) {
    Quad<T>::step_adj(s, at, dt, pc, g, gs, gat);
    for (int j = 0; j < Y::HD; ++j) {
      T v = 0;
      if (h[j * hs] > T(0)) {
#pragma unroll
        for (int i = 0; i < Y::SD; ++i) v += P[Y::O_W2 + i * Y::HD + j] * g[i];
      }
      dh[j * hs] = v;
    }
#pragma unroll
    for (int k = 0; k < Y::XD; ++k) {
      T v = 0;
      for (int j = 0; j < Y::HD; ++j) v += P[Y::O_W1 + j * Y::XD + k] * dh[j * hs];
      if (k < 12) gs[k] += v; else gat[k - 12] += v;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      T v = 0;
#pragma unroll
      for (int r = 0; r < 4; ++r) v += P[Y::O_LAT + r * 4 + c] * gat[r];
      ga[c] = v;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      dk[i] = dt * ((at[1 + i] - T(0.5)) - s[9 + i]) * g[9 + i];
      const T J = T(pc[Q_JX + i]);
      dj[i] = -dt * T(pc[Q_RDX + i]) / (J * J) * g[9 + i];
    }
  }

  // ---- streaming forms for the fused horizon kernels (controller trained THROUGH the learnt dynamics,
  //      scripts/train_drone.py:175-199 with train_dynamics = LearntDynamics): state / action cotangents only, hidden
  //      units produced and consumed one at a time (nothing but registers), same summation order as forward / adjoint
  //      above.  `pc` holds the analytic constants, P the flat parameter vector (shared memory in the kernels).
  APG_HD static void step_sa(const T* P, const float* pc, const T* s, const T* a, T dt, T* out) {
    T at[4], v[Y::SD];
    transform_action(P, a, at);
    Quad<T>::step(s, at, dt, pc, out);
#pragma unroll
    for (int i = 0; i < Y::SD; ++i) v[i] = P[Y::O_B2 + i];
#pragma unroll 4
    for (int j = 0; j < Y::HD; ++j) {
      const T hj = hidden(P, s, at, j);
#pragma unroll
      for (int i = 0; i < Y::SD; ++i) v[i] += P[Y::O_W2 + i * Y::HD + j] * hj;
    }
#pragma unroll
    for (int i = 0; i < Y::SD; ++i) out[i] += v[i];
  }
  APG_HD static void step_adj_sa(const T* P, const float* pc, const T* s, const T* a, T dt, const T* g, T* gs, T* ga) {
    T at[4], gat[4], acc[Y::XD];
    transform_action(P, a, at);
    Quad<T>::step_adj(s, at, dt, pc, g, gs, gat);
#pragma unroll
    for (int k = 0; k < Y::XD; ++k) acc[k] = T(0);
#pragma unroll 4
    for (int j = 0; j < Y::HD; ++j) {
      T dh = T(0);
      if (hidden(P, s, at, j) > T(0)) {
#pragma unroll
        for (int i = 0; i < Y::SD; ++i) dh += P[Y::O_W2 + i * Y::HD + j] * g[i];
      }
#pragma unroll
      for (int k = 0; k < Y::XD; ++k) acc[k] += P[Y::O_W1 + j * Y::XD + k] * dh;
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) gs[k] += acc[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) gat[k] += acc[12 + k];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      T v = 0;
#pragma unroll
      for (int r = 0; r < 4; ++r) v += P[Y::O_LAT + r * 4 + c] * gat[r];
      ga[c] = v;
    }
  }

  // ---- the interface the kernels use (same for the fixed wing): x = the 16 inputs of the residual MLP,
  //      dph = per-drone cotangents of the NPH = 23 parameters that precede the MLP in the flat vector
  static constexpr int NPH = Y::O_W1;
  APG_HD static void fwd(const T* P, const float* pc, const T* s, const T* a, T dt, T* out, T* x, T* h, int hs) {
    T at[4];
    forward(P, pc, s, a, dt, out, at, h, hs);
    for (int k = 0; k < 12; ++k) x[k] = s[k];
    for (int k = 0; k < 4; ++k) x[12 + k] = at[k];
  }
  APG_HD static void adj(const T* P, const float* pc, const T* s, const T* a, const T* x, const T* h, int hs, T dt,
                         const T* g, T* gs, T* ga, T* dh, T* dph) {
    T gat[4], dk[3], dj[3];
    adjoint(P, pc, s, a, x + 12, h, hs, dt, g, gs, ga, dh, gat, dk, dj);
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) dph[Y::O_LAT + r * 4 + c] = gat[r] * a[c];
    dph[Y::O_MASS] = T(0);
    for (int i = 0; i < 3; ++i) { dph[Y::O_J + i] = dj[i]; dph[Y::O_K + i] = dk[i]; }
  }
};

}  // namespace apg
