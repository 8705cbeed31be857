// Test-only: the evaluation kernels THEMSELVES (csrc/eval_kernels.cu, unchanged source, -DAPG_SIM) running on the CPU
// on top of the software model of te_sim.h - tile engine GEMMs (mma.sync fragments), TMA bulk copies, mbarriers,
// __syncthreads_or early exit, per-drone logic, output addressing, end to end.
#define APG_SIM 1
#include "te_sim.h"

#include "../../apg_trajectory_tracking_b200/csrc/eval_kernels.cu"
#include "../../apg_trajectory_tracking_b200/csrc/pack_tables.h"

using namespace apg;

namespace {
// apg_pack_kernel (misc_kernels.cu, GPU-verified) restated for the host: only used to produce the kernels' inputs
int pack_perm_host(int k, int npos) {
  if (npos <= 0 || k < 64) return k;
  const int tt = (k - 64) / 20, c = (k - 64) - tt * 20;
  return 64 + c * npos + tt;
}
void pack_host(const PackTable& t, const float* params, float* wf, float* wb) {
  for (int s = 0; s < t.n; ++s) {
    const PackSeg g = t.seg[s];
    const float* src = params + g.src;
    float* dst = (g.which ? wb : wf) + g.dst;
    int total;
    if (g.mode == PK_COPY_PAD) total = g.rows * g.wcols;
    else if (g.mode == PK_CONV_BWD) total = g.rows * g.ldd;
    else if (g.mode == PK_TRANSPOSE) total = g.cols * g.wcols;
    else total = g.cols * g.ldd;
    for (int i = 0; i < total; ++i) {
      float v = 0.f;
      int di = i;
      if (g.mode == PK_COPY_PAD) {
        const int r = i / g.wcols, c = i - r * g.wcols;
        if (c < g.cols) v = src[r * g.sld + pack_perm_host(c, g.perm)];
        di = r * g.ldd + (g.sw ? (c ^ ((r & 3) << 3)) : c);
      } else if (g.mode == PK_TRANSPOSE) {
        const int c = i / g.wcols, r = i - c * g.wcols;
        if (r < g.rows) v = src[r * g.sld + pack_perm_host(c, g.perm)];
        di = c * g.ldd + (g.sw ? (r ^ ((c & 3) << 3)) : r);
      } else if (g.mode == PK_CONV_FWD) {
        const int kk = i / g.ldd, c = i - kk * g.ldd;
        const int rd = g.cols / 3, j = kk / rd, d = kk - j * rd;
        if (c < g.rows) v = src[c * g.cols + d * 3 + j];
      } else {
        const int c = i / g.ldd, kk = i - c * g.ldd;
        const int rd = g.cols / 3;
        if (kk < g.cols) { const int j = kk / rd, d = kk - j * rd; v = src[c * g.cols + d * 3 + j]; }
      }
      dst[di] = v;
    }
  }
}
int report(char* err, int err_len) {
  std::vector<std::string>& e = sim::errors();
  std::string all;
  for (const std::string& s : e) all += s + "; ";
  if (err && err_len > 0) { strncpy(err, all.c_str(), (size_t)err_len - 1); err[err_len - 1] = 0; }
  const int n = (int)e.size();
  e.clear();
  return n;
}
template <class F>
int run(int grid, int block, F&& f, char* err, int err_len) {
  try {
    sim::launch(grid, block, f);
  } catch (const std::exception& ex) {
    sim::fail(std::string("exception: ") + ex.what());
  }
  return report(err, err_len);
}
}  // namespace

// eval_rollout_kernel<<<grid, 256>>>  (quadrotor, hutter conv net with out_dim outputs: 4h concurrent / 4 autoregressive)
extern "C" int hc_tesim_eval_rollout(const float* params, int h, int out_dim, const float* tables, const int* index,
                                     int RL, const float* init, int n, int steps, float dt, const float* pc,
                                     float thresh_div, float thresh_stable, int test_time, int grid, float* states_out,
                                     float* div_out, float* actions_out, int* n_steps_out, char* err, int err_len) {
  const HutterLayout y = make_hutter_layout(15, h, 9, out_dim, 1);
  std::vector<float> wf(y.f_total + 64, 0.f), wb(y.b_total + 64, 0.f);
  pack_host(hutter_pack_table(y), params, wf.data(), wb.data());
  EvalArgs a;
  a.wf = wf.data(); a.tables = tables; a.table_index = index; a.init_states = init; a.N = n; a.h = h; a.dt = dt;
  memcpy(a.pc.v, pc, sizeof(float) * MAX_PHYS);
  a.ev.steps = steps; a.ev.table_rows = RL; a.ev.test_time = test_time; a.ev.thresh_div = thresh_div;
  a.ev.thresh_stable = thresh_stable;
  a.states_out = states_out; a.div_out = div_out; a.actions_out = actions_out; a.n_steps_out = n_steps_out;
  return run(grid, NT, [&]() { eval_rollout_kernel(y, a); }, err, err_len);
}

// eval_wing_kernel<<<grid, 256>>>
extern "C" int hc_tesim_eval_wing(const float* params, int h, const float* targets, int K, const float* init, int n,
                                  const float* mean, const float* std_, float dt_data, float dt_env, const float* pc,
                                  int steps, float thresh_div, float thresh_stable, int test_time, int grid,
                                  float* states_out, float* div_out, float* actions_out, int* n_steps_out,
                                  float* dts_out, float* dtc_out, char* err, int err_len) {
  const HutterLayout y = make_hutter_layout(9, 1, 3, 4 * h, 0);
  std::vector<float> wf(y.f_total + 64, 0.f), wb(y.b_total + 64, 0.f);
  pack_host(hutter_pack_table(y), params, wf.data(), wb.data());
  WingEvalArgs a;
  a.wf = wf.data(); a.targets = targets; a.init_states = init; a.N = n; a.dt = dt_env;
  memcpy(a.pc.v, pc, sizeof(float) * MAX_PHYS);
  for (int j = 0; j < 12; ++j) { a.nc.mean[j] = mean[j]; a.nc.std_[j] = std_[j]; }
  a.ev.steps = steps; a.ev.n_targets = K; a.ev.test_time = test_time; a.ev.h = h; a.ev.thresh_div = thresh_div;
  a.ev.thresh_stable = thresh_stable; a.ev.vlen = (float)(12.0 * (double)dt_data); a.ev.des_speed = 11.5f;
  a.states_out = states_out; a.div_out = div_out; a.actions_out = actions_out; a.n_steps_out = n_steps_out;
  a.dt_sum_out = dts_out; a.dt_cnt_out = dtc_out;
  return run(grid, NT, [&]() { eval_wing_kernel(y, a); }, err, err_len);
}

// eval_cartpole_kernel<<<grid, 256>>>
extern "C" int hc_tesim_eval_cartpole(const float* params, int h, const float* init, int n, float dt, const float* pc,
                                      int steps, float thresh_div, int burn_in, int grid, float* states_out,
                                      float* actions_out, int* n_steps_out, float* ang_sum, float* ang_cnt,
                                      float* vel_sum, char* err, int err_len) {
  const SimpleLayout y = make_simple_layout(4, h);
  std::vector<float> wf(y.f_total + 64, 0.f), wb(y.b_total + 64, 0.f);
  pack_host(simple_pack_table(y), params, wf.data(), wb.data());
  CartpoleEvalArgs a;
  a.wf = wf.data(); a.init_states = init; a.N = n; a.dt = dt;
  memcpy(a.pc.v, pc, sizeof(float) * MAX_PHYS);
  a.ev.steps = steps; a.ev.burn_in = burn_in; a.ev.thresh_div = thresh_div;
  a.states_out = states_out; a.actions_out = actions_out; a.n_steps_out = n_steps_out;
  a.angle_sum_out = ang_sum; a.angle_cnt_out = ang_cnt; a.vel_sum_out = vel_sum;
  return run(grid, NT, [&]() { eval_cartpole_kernel(y, a); }, err, err_len);
}
