// Test-only: the concurrent hutter rollout kernels THEMSELVES on the CPU model of te_sim.h (-DAPG_SIM, unchanged
// sources): hutter_fwd_kernel / hutter_adj_kernel (csrc/hutter_kernels.cu - GPU-verified, so this run also validates
// the MODEL: warp specialisation, named barrier, mbarrier hand-offs, TMA bulk loads / stores, mma fragments).
#define APG_SIM 1
#include "te_sim.h"

#include "../../apg_trajectory_tracking_b200/csrc/hutter_kernels.cu"
#include "../../apg_trajectory_tracking_b200/csrc/pack_tables.h"

using namespace apg;

namespace {
int pack_perm_host(int k, int npos) {
  if (npos <= 0 || k < 64) return k;
  const int tt = (k - 64) / 20, c = (k - 64) - tt * 20;
  return 64 + c * npos + tt;
}
// apg_pack_kernel (misc_kernels.cu, GPU-verified) restated for the host: only used to produce the kernels' inputs
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
struct Ctx {
  HutterLayout y;
  std::vector<float> wf, wb;
  RolloutArgs a;
};
Ctx make_ctx(const float* params, const float* in_state, const float* cur, const float* in_ref, const float* ref, int n,
             int h, float dt, const float* pc, float* st_x1, float* st_h1, float* st_h2, float* st_h3, float* st_act,
             float* st_states) {
  Ctx c;
  c.y = make_hutter_layout(15, h, 9, 4 * h, 1);
  c.wf.assign(c.y.f_total + 64, 0.f);
  c.wb.assign(c.y.b_total + 64, 0.f);
  pack_host(hutter_pack_table(c.y), params, c.wf.data(), c.wb.data());
  memset(&c.a, 0, sizeof c.a);
  c.a.in_state = in_state; c.a.cur = cur; c.a.in_ref = in_ref; c.a.ref = ref;
  c.a.N = n; c.a.h = h; c.a.ref_rows = h; c.a.dt = dt;
  memcpy(c.a.pc.v, pc, sizeof(float) * MAX_PHYS);
  c.a.st_x1 = st_x1; c.a.st_h1 = st_h1; c.a.st_h2 = st_h2; c.a.st_h3 = st_h3; c.a.st_act = st_act;
  c.a.st_states = st_states;
  return c;
}
}  // namespace

extern "C" int hc_hksim_num_params(int h) { return make_hutter_layout(15, h, 9, 4 * h, 1).n_params; }

// hutter_fwd_kernel<Quad, true><<<grid, 320>>>
extern "C" int hc_hksim_forward(const float* params, const float* in_state, const float* cur, const float* in_ref,
                                const float* ref, int n, int h, float dt, const float* pc, int grid, float* st_x1,
                                float* st_h1, float* st_h2, float* st_h3, float* st_act, float* st_states,
                                float* loss_partials, char* err, int err_len) {
  Ctx c = make_ctx(params, in_state, cur, in_ref, ref, n, h, dt, pc, st_x1, st_h1, st_h2, st_h3, st_act, st_states);
  c.a.wf = c.wf.data(); c.a.wb = c.wb.data(); c.a.loss_partials = loss_partials;
  return run(grid, NTH, [&]() { hutter_fwd_kernel<Quad, true>(c.y, c.a); }, err, err_len);
}

// hutter_adj_kernel<Quad, true><<<grid, 320>>>: per-CTA gradient partials [grid][n_params] (kernel column order)
extern "C" int hc_hksim_adjoint(const float* params, const float* in_state, const float* cur, const float* in_ref,
                                const float* ref, int n, int h, float dt, const float* pc, int grid, float* st_x1,
                                float* st_h1, float* st_h2, float* st_h3, float* st_act, float* st_states,
                                float* grad_partials, char* err, int err_len) {
  Ctx c = make_ctx(params, in_state, cur, in_ref, ref, n, h, dt, pc, st_x1, st_h1, st_h2, st_h3, st_act, st_states);
  c.a.wf = c.wf.data(); c.a.wb = c.wb.data(); c.a.grad_partials = grad_partials;
  return run(grid, NTH, [&]() { hutter_adj_kernel<Quad, true>(c.y, c.a); }, err, err_len);
}

// torch entry p of the fc1 block <- kernel (position-major) column, as apg_reduce_kernel maps it
extern "C" void hc_hksim_reduce(const float* partials, int grid, int h, float* grad) {
  const HutterLayout y = make_hutter_layout(15, h, 9, 4 * h, 1);
  for (int p = 0; p < y.n_params; ++p) {
    int q = p;
    if (y.perm_npos > 0 && p >= y.t_w1 && p < y.t_w1 + 64 * y.K1) {
      const int j = (p - y.t_w1) / y.K1, k = (p - y.t_w1) - j * y.K1;
      if (k >= 64) { const int c = (k - 64) / y.perm_npos, tt = (k - 64) - c * y.perm_npos; q = y.t_w1 + j * y.K1 + 64 + tt * 20 + c; }
    }
    float s = 0.f;
    for (int c = 0; c < grid; ++c) s += partials[(size_t)c * y.n_params + q];
    grad[p] = s;
  }
}
