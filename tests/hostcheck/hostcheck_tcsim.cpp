// Test-only: the tcgen05 kernels THEMSELVES (csrc/hutter_tc_kernels.cu here, csrc/adj_dw_tc_kernels.cu in
// hostcheck_tcsim_dw.cpp - the two files open different layout namespaces; unchanged source)
// running on the CPU on top of the software model of tc_sim.h - roles, mbarrier hand-off, TMEM operand placement,
// issue order / accumulate flags, epilogues, stash addressing, gradient mapping, end to end.
#define APG_TC_SIM 1
#define APG_SIM 1
#include "tc_sim.h"

#include "../../apg_trajectory_tracking_b200/csrc/hutter_tc_kernels.cu"

using namespace apg;

namespace {
HutterLayout layout() { return make_hutter_layout(tc::F0, tc::H, tc::RD, tc::MO, 1); }

RolloutArgs make_args(const float* in_state, const float* cur, const float* in_ref, const float* ref, int n, float dt,
                      const float* pc, float* st_x1, float* st_h1, float* st_h2, float* st_h3, float* st_act,
                      float* st_states) {
  RolloutArgs a;
  memset(&a, 0, sizeof a);
  a.in_state = in_state; a.cur = cur; a.in_ref = in_ref; a.ref = ref;
  a.N = n; a.h = tc::H; a.ref_rows = tc::H; a.dt = dt;
  memcpy(a.pc.v, pc, sizeof(float) * MAX_PHYS);
  a.st_x1 = st_x1; a.st_h1 = st_h1; a.st_h2 = st_h2; a.st_h3 = st_h3; a.st_act = st_act; a.st_states = st_states;
  return a;
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
}  // namespace

extern "C" int hc_sim_blob_bytes() { return tc::BLOB_BYTES; }
extern "C" int hc_sim_num_params() { return layout().n_params; }

// weights -> images (the kernel's body over its flat index; the launch geometry of apg_pack_tc_kernel is trivial)
extern "C" void hc_sim_pack(const float* params, unsigned char* blob) {
  const HutterLayout y = layout();
  for (int e = 0; e < tc::PAIRS_TOTAL + tc::B_TOTAL; ++e) tc::pack_body(e, params, y, blob);
}

// hutter_fwd_tc_kernel<<<grid, 288>>>: returns the number of model violations (messages in err)
extern "C" int hc_sim_forward(const unsigned char* blob, const float* in_state, const float* cur, const float* in_ref,
                              const float* ref, int n, float dt, const float* pc, int grid, float* st_x1, float* st_h1,
                              float* st_h2, float* st_h3, float* st_act, float* st_states, float* loss_partials,
                              float* states_out, float* actions_out, char* err, int err_len) {
  const HutterLayout y = layout();
  RolloutArgs a = make_args(in_state, cur, in_ref, ref, n, dt, pc, st_x1, st_h1, st_h2, st_h3, st_act, st_states);
  a.loss_partials = loss_partials; a.states_out = states_out; a.actions_out = actions_out;
  sim::launch(grid, TC_THREADS, [&]() { hutter_fwd_tc_kernel(blob, y, a); });
  return report(err, err_len);
}

// hutter_adj_dx_tc_kernel<<<grid, 288>>> on the stash the forward left
extern "C" int hc_sim_adj_dx(const unsigned char* blob, const float* in_state, const float* cur, const float* in_ref,
                             const float* ref, int n, float dt, const float* pc, int grid, float* st_x1, float* st_h1,
                             float* st_h2, float* st_h3, float* st_act, float* st_states, float* dzo, float* dz3,
                             float* dz2, float* dz1, float* dzx, char* err, int err_len) {
  const HutterLayout y = layout();
  RolloutArgs a = make_args(in_state, cur, in_ref, ref, n, dt, pc, st_x1, st_h1, st_h2, st_h3, st_act, st_states);
  DzStash z{dzo, dz3, dz2, dz1, dzx};
  sim::launch(grid, TC_THREADS, [&]() { hutter_adj_dx_tc_kernel(blob, y, a, z); });
  return report(err, err_len);
}

extern "C" long long hc_sim_mma_count() { return sim::S().mma_count; }
