// Test-only: adj_dw_tc_kernel ITSELF (csrc/adj_dw_tc_kernels.cu, unchanged source) running on the CPU on top of the
// software model of tc_sim.h (see hostcheck_tcsim.cpp).
#define APG_TC_SIM 1
#define APG_SIM 1
#include "tc_sim.h"

#include "../../apg_trajectory_tracking_b200/csrc/adj_dw_tc_kernels.cu"

using namespace apg;

namespace {
HutterLayout layout() { return make_hutter_layout(tc::F0, tc::H, tc::RD, tc::MO, 1); }
int report(char* err, int err_len) {
  std::vector<std::string>& e = sim::errors();
  std::string all;
  for (const std::string& s : e) all += s + "; ";
  if (err && err_len > 0) { strncpy(err, all.c_str(), (size_t)err_len - 1); err[err_len - 1] = 0; }
  const int n = (int)e.size();
  e.clear();
  return n;
}
}  // namespace

extern "C" int hc_simdw_num_params() { return layout().n_params; }

// adj_dw_tc_kernel<<<grid, 288>>>: per-CTA gradient partials [grid][n_params]
extern "C" int hc_simdw_adj_dw(const float* in_state, const float* in_ref, int n, int grid, float* st_x1, float* st_h1,
                             float* st_h2, float* st_h3, float* dzo, float* dz3, float* dz2, float* dz1, float* dzx,
                             float* grad_partials, char* err, int err_len) {
  const HutterLayout y = layout();
  RolloutArgs a;
  memset(&a, 0, sizeof a);
  a.in_state = in_state; a.in_ref = in_ref; a.N = n; a.h = tc::H;
  a.st_x1 = st_x1; a.st_h1 = st_h1; a.st_h2 = st_h2; a.st_h3 = st_h3; a.grad_partials = grad_partials;
  DzStash z{dzo, dz3, dz2, dz1, dzx};
  sim::launch(grid, DW_THREADS, [&]() { adj_dw_tc_kernel(y, a, z); });
  return report(err, err_len);
}

