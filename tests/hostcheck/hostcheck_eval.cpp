// Test-only harness for the closed-loop evaluation rollout: the per-drone logic of csrc/eval_math.cuh driven exactly
// like the kernel's per-thread code (csrc/eval_kernels.cu), with a plain scalar hutter-conv policy standing in for the
// tile-engine GEMMs (which are covered by the GPU parity tests).  Not part of the product.
#include <math.h>
#include <vector>
#include "eval_math.cuh"

using namespace apg;

namespace {
// models/hutter_model.py:32-49 for one drone; params = torch-flat Net(15, h, 9, Mo); returns the first 4 sigmoids
void policy_first_action(const float* p, int h, int Mo, const float* f /*15*/, const float* in_ref /*[h][9]*/,
                         float* a4) {
  const int npos = h - 2, K1 = 64 + 20 * npos;
  const float* ws = p;            p += 64 * 15;
  const float* bs = p;            p += 64;
  const float* wc = p;            p += 20 * 9 * 3;
  const float* bc = p;            p += 20;
  p += 64 * 9 * h + 64;           // ref_in (unused by the conv net)
  const float* w1 = p;            p += 64 * K1;
  const float* b1 = p;            p += 64;
  const float* w2 = p;            p += 64 * 64;
  const float* b2 = p;            p += 64;
  const float* w3 = p;            p += 64 * 64;
  const float* b3 = p;            p += 64;
  const float* wo = p;            p += Mo * 64;
  const float* bo = p;
  std::vector<float> x(K1), h1(64), h2(64), h3(64);
  for (int j = 0; j < 64; ++j) {
    float s = bs[j];
    for (int k = 0; k < 15; ++k) s += ws[j * 15 + k] * f[k];
    x[j] = tanhf(s);
  }
  for (int c = 0; c < 20; ++c)
    for (int t = 0; t < npos; ++t) {
      float s = bc[c];
      for (int d = 0; d < 9; ++d)
        for (int j = 0; j < 3; ++j) s += wc[(c * 9 + d) * 3 + j] * in_ref[(t + j) * 9 + d];
      x[64 + c * npos + t] = s > 0.f ? s : 0.f;
    }
  for (int j = 0; j < 64; ++j) { float s = b1[j]; for (int k = 0; k < K1; ++k) s += w1[j * K1 + k] * x[k]; h1[j] = tanhf(s); }
  for (int j = 0; j < 64; ++j) { float s = b2[j]; for (int k = 0; k < 64; ++k) s += w2[j * 64 + k] * h1[k]; h2[j] = tanhf(s); }
  for (int j = 0; j < 64; ++j) { float s = b3[j]; for (int k = 0; k < 64; ++k) s += w3[j * 64 + k] * h2[k]; h3[j] = tanhf(s); }
  for (int j = 0; j < 4; ++j) {
    float s = bo[j];
    for (int k = 0; k < 64; ++k) s += wo[j * 64 + k] * h3[k];
    a4[j] = 1.f / (1.f + expf(-s));
  }
}
}  // namespace

// mirrors the per-thread code of eval_rollout_kernel for drones 0..n-1
extern "C" void hc_eval_rollout(const float* params, int h, int Mo, const float* tables, const int* table_index,
                                int RL, const float* init_states, int n, int steps, float dt, const float* pc,
                                float thresh_div, float thresh_stable, int test_time, float* states_out,
                                float* div_out, float* actions_out, int* n_steps_out) {
  EvalParams ev;
  ev.steps = steps; ev.table_rows = RL; ev.test_time = test_time; ev.thresh_div = thresh_div;
  ev.thresh_stable = thresh_stable;
  std::vector<float> win(h * 9);
  for (int d = 0; d < n; ++d) {
    const float* tab = tables + (size_t)(table_index ? table_index[d] : d) * RL * 9;
    float s[12];
    int ci = 0, alive = 1, nsteps = 0;
    for (int j = 0; j < 12; ++j) { s[j] = init_states[d * 12 + j]; states_out[(size_t)d * (steps + 1) * 12 + j] = s[j]; }
    for (int i = 0; i < steps && alive; ++i) {
      int start, nreal, ci_next;
      eval_window_plan(ci, RL, h, &start, &nreal, &ci_next);
      ci = ci_next;
      float c0[12], f[15];
      c0[0] = c0[1] = c0[2] = 0.f;
      for (int j = 3; j < 12; ++j) c0[j] = s[j];
      Quad<float>::features(c0, f);
      for (int r = 0; r < h; ++r)
        for (int c = 0; c < 9; ++c)
          win[r * 9 + c] = eval_in_ref_elem(tab, RL, start, nreal, r, c, c < 3 ? s[c] : 0.f, c >= 6 ? s[c] : 0.f);
      float a[4], sn[12];
      policy_first_action(params, h, Mo, f, win.data(), a);
      for (int c = 0; c < 4; ++c) a[c] = fminf(fmaxf(a[c], 0.f), 1.f);
      Quad<float>::step(s, a, dt, pc, sn);
      for (int j = 0; j < 12; ++j) states_out[((size_t)d * (steps + 1) + i + 1) * 12 + j] = sn[j];
      for (int c = 0; c < 4; ++c) actions_out[((size_t)d * steps + i) * 4 + c] = a[c];
      div_out[(size_t)d * steps + i] = eval_post_step(sn, tab, ci, ev, &alive);
      for (int j = 0; j < 12; ++j) s[j] = sn[j];
      ++nsteps;
      if (i >= RL) alive = 0;
    }
    n_steps_out[d] = nsteps;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fixed wing: mirrors the per-thread code of eval_wing_kernel (policy: scalar hutter "linear ref" net)
// ---------------------------------------------------------------------------------------------------------------
namespace {
void wing_policy_first_action(const float* p, int h, const float* f /*9*/, const float* r3, float* a4) {
  const int Mo = 4 * h;
  const float* ws = p;  p += 64 * 9;
  const float* bs = p;  p += 64;
  p += 20 * 3 * 3 + 20;                 // conv_ref (unused by the linear-ref net)
  const float* wr = p;  p += 64 * 3;
  const float* br = p;  p += 64;
  const float* w1 = p;  p += 64 * 128;
  const float* b1 = p;  p += 64;
  const float* w2 = p;  p += 64 * 64;
  const float* b2 = p;  p += 64;
  const float* w3 = p;  p += 64 * 64;
  const float* b3 = p;  p += 64;
  const float* wo = p;  p += Mo * 64;
  const float* bo = p;
  float x[128], h1[64], h2[64], h3[64];
  for (int j = 0; j < 64; ++j) {
    float s = bs[j];
    for (int k = 0; k < 9; ++k) s += ws[j * 9 + k] * f[k];
    x[j] = tanhf(s);
    float q = br[j];
    for (int k = 0; k < 3; ++k) q += wr[j * 3 + k] * r3[k];
    x[64 + j] = tanhf(q);
  }
  for (int j = 0; j < 64; ++j) { float s = b1[j]; for (int k = 0; k < 128; ++k) s += w1[j * 128 + k] * x[k]; h1[j] = tanhf(s); }
  for (int j = 0; j < 64; ++j) { float s = b2[j]; for (int k = 0; k < 64; ++k) s += w2[j * 64 + k] * h1[k]; h2[j] = tanhf(s); }
  for (int j = 0; j < 64; ++j) { float s = b3[j]; for (int k = 0; k < 64; ++k) s += w3[j * 64 + k] * h2[k]; h3[j] = tanhf(s); }
  for (int j = 0; j < 4; ++j) {
    float s = bo[j];
    for (int k = 0; k < 64; ++k) s += wo[j * 64 + k] * h3[k];
    a4[j] = 1.f / (1.f + expf(-s));
  }
}
}  // namespace

extern "C" void hc_eval_wing(const float* params, int h, const float* targets, int K, const float* init_states, int n,
                             const float* mean, const float* std_, float dt_data, float dt_env, const float* pc,
                             int steps, float thresh_div, float thresh_stable, int test_time, float* states_out,
                             float* div_out, float* actions_out, int* n_steps_out, float* dts_out, float* dtc_out) {
  WingEvalParams e;
  e.steps = steps; e.n_targets = K; e.test_time = test_time; e.h = h; e.thresh_div = thresh_div;
  e.thresh_stable = thresh_stable; e.vlen = (float)(12.0 * (double)dt_data); e.des_speed = 11.5f;
  for (int d = 0; d < n; ++d) {
    WingEvalDrone D;
    wing_eval_init(D, init_states + d * 12, 1);
    const float* tg = targets + (size_t)d * K * 3;
    for (int j = 0; j < 12; ++j) states_out[(size_t)d * (steps + 1) * 12 + j] = D.env[j];
    for (int i = 0; i < steps && D.alive; ++i) {
      float f[9], r3[3], a[4], nxt[12];
      WingPrep<float>::drone(D.obs, tg + D.ti * 3, mean, std_, e.vlen, h, f, r3);
      wing_policy_first_action(params, h, f, r3, a);
      Wing<float>::step(D.env, a, dt_env, pc, nxt);
      for (int j = 0; j < 12; ++j) states_out[((size_t)d * (steps + 1) + i + 1) * 12 + j] = nxt[j];
      for (int c = 0; c < 4; ++c) actions_out[((size_t)d * steps + i) * 4 + c] = a[c];
      div_out[(size_t)d * steps + i] = wing_eval_post_step(D, nxt, tg, e);
    }
    wing_eval_finish(D, e);
    n_steps_out[d] = D.nsteps; dts_out[d] = D.dt_sum; dtc_out[d] = D.dt_cnt;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// cartpole: mirrors the per-thread code of eval_cartpole_kernel (policy: scalar models/simple_model.py Net)
// ---------------------------------------------------------------------------------------------------------------
namespace {
float simple_first_action(const float* p, int h, const float* s4) {
  const int dims[6] = {4, 32, 64, 64, 32, h};
  float x[64], yv[64];
  x[0] = 0.f;                                         // simple_model.py:21
  for (int j = 1; j < 4; ++j) x[j] = s4[j];
  for (int l = 0; l < 5; ++l) {
    const float* w = p;  p += dims[l + 1] * dims[l];
    const float* b = p;  p += dims[l + 1];
    for (int j = 0; j < dims[l + 1]; ++j) {
      float s = b[j];
      for (int k = 0; k < dims[l]; ++k) s += w[j * dims[l] + k] * x[k];
      yv[j] = tanhf(s);
    }
    for (int j = 0; j < dims[l + 1]; ++j) x[j] = yv[j];
  }
  return x[0];
}
}  // namespace

extern "C" void hc_eval_cartpole(const float* params, int h, const float* init_states, int n, float dt,
                                 const float* pc, int steps, float thresh_div, int burn_in, float* states_out,
                                 float* actions_out, int* n_steps_out, float* ang_sum_out, float* ang_cnt_out,
                                 float* vel_sum_out) {
  CartpoleEvalParams e;
  e.steps = steps; e.burn_in = burn_in; e.thresh_div = thresh_div;
  for (int d = 0; d < n; ++d) {
    CartpoleEvalDrone D;
    cartpole_eval_init(D, init_states + d * 4, 1);
    for (int i = 0; i < steps && D.alive; ++i) {
      cartpole_eval_before_policy(D, i);
      float a[1], nxt[4];
      a[0] = simple_first_action(params, h, D.s);
      Cartpole<float>::step(D.s, a, dt, pc, nxt);
      cartpole_eval_post_step(D, nxt, i, e);
      for (int j = 0; j < 4; ++j) states_out[((size_t)d * steps + i) * 4 + j] = nxt[j];
      actions_out[(size_t)d * steps + i] = a[0];
    }
    n_steps_out[d] = D.nsteps; ang_sum_out[d] = D.ang_sum; ang_cnt_out[d] = D.ang_cnt; vel_sum_out[d] = D.vel_sum;
  }
}
