// Test-only: the second-generation tcgen05 kernels THEMSELVES (csrc/tq_kernels.cu: forward + dX chain,
// csrc/tq_dw_kernels.cu: streaming weight-gradient GEMM; unchanged source) on the CPU model of tc_sim.h: four epilogue
// groups over two TMEM slots (slot hand-over), operand-image stashes, bulk-copy pipeline, 128B-swizzled descriptors.
#define APG_TC_SIM 1
#define APG_SIM 1
#include "tc_sim.h"

#include "../../apg_trajectory_tracking_b200/csrc/tq_kernels.cu"
#include "../../apg_trajectory_tracking_b200/csrc/tq_dw_kernels.cu"

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

// sizes: [blob, tblob, fstash, zstash, n_params, f_tile_bytes, z_tile_bytes]
extern "C" void hc_tq_sizes(int n, long long* out) {
  out[0] = (long long)tq_blob_bytes(); out[1] = (long long)tq_tblob_bytes();
  out[2] = (long long)tq_fstash_bytes(n); out[3] = (long long)tq_zstash_bytes(n);
  out[4] = layout().n_params; out[5] = (long long)tq::F_TILE_BYTES; out[6] = (long long)tq::Z_TILE_BYTES;
}

// pack -> forward -> dX chain -> dW GEMM through the REAL launchers; all buffers from the caller (fstash / zstash
// 1024-byte aligned).  Returns the number of model violations.
extern "C" int hc_tq_step(const float* params, const float* in_state, const float* cur, const float* in_ref,
                          const float* ref, int n, float dt, const float* pc, int grid, unsigned char* blob,
                          unsigned char* tblob, unsigned char* fstash, unsigned char* zstash, float* loss_partials,
                          float* grad_partials, float* states_out, float* actions_out, int stages, int dyn_grid, float* loss_total,
                          char* err,
                          int err_len) {
  const HutterLayout y = layout();
  RolloutArgs a;
  memset(&a, 0, sizeof a);
  a.in_state = in_state; a.cur = cur; a.in_ref = in_ref; a.ref = ref;
  a.raw_inputs = (in_state == nullptr && in_ref == nullptr) ? 1 : 0;   // capi.cu make_args: RAW samples, prepare in the prologue
  a.N = n; a.h = tc::H; a.ref_rows = tc::H; a.dt = dt;
  memcpy(a.pc.v, pc, sizeof(float) * MAX_PHYS);
  a.loss_partials = loss_partials; a.grad_partials = grad_partials; a.states_out = states_out; a.actions_out = actions_out;
  unsigned char stamp[16];
  memset(stamp, 3, sizeof stamp);
  if (stages >= 1) {
    launch_tq_pack(y, params, blob, tblob, nullptr);
    launch_tq_fwd(blob, a, fstash, grid, nullptr);
    unsigned ticket = 77u;
    float loss_sum = 0.f;
    launch_tq_dyn(a, fstash, zstash, &loss_sum, &ticket, 77u, dyn_grid, nullptr);
    if (loss_total) *loss_total = loss_sum;
  }
  if (stages >= 2) launch_tq_dx(tblob, a, fstash, zstash, stamp, 3, grid, nullptr);
  if (stages >= 3) launch_tq_dw(y, a, fstash, zstash, grid, nullptr);
  return report(err, err_len);
}

extern "C" long long hc_tq_mma_count() { return sim::S().mma_count; }
