// Test-only: the learnt-dynamics kernels THEMSELVES (csrc/learnt_kernels.cu, unchanged source, -DAPG_SIM) on the CPU
// thread model of gpu_sim.h: shared-memory parameter / factor rows, tile loop, per-thread entry accumulation,
// per-block partial vectors (reduced here in block order like apg_reduce_kernel).
#define APG_SIM 1
#include "te_sim.h"

#include "../../apg_trajectory_tracking_b200/csrc/learnt_kernels.cu"
#include "../../apg_trajectory_tracking_b200/csrc/misc_kernels.cu"      // launch_reduce_grad of the real launcher

using namespace apg;

namespace {
int report(char* err, int err_len) {
  std::vector<std::string>& e = sim::errors();
  std::string all;
  for (const std::string& s : e) all += s + "; ";
  if (err && err_len > 0) { strncpy(err, all.c_str(), (size_t)err_len - 1); err[err_len - 1] = 0; }
  const int n = (int)e.size();
  e.clear();
  return n;
}
template <class M>
int run(const float* params, const float* pcv, const float* s, const float* a, float dt, int n, int grid,
        const float* g, float* out, float* gs, float* ga, float* gp, char* err, int err_len) {
  using RW = LearntRows<M::NPH>;
  PhysConsts pc;
  memcpy(pc.v, pcv, sizeof(float) * MAX_PHYS);
  try {
    sim::launch(grid, LT, [&]() { learnt_fwd_kernel<M>(params, pc, s, a, dt, n, out); });
    std::vector<float> partials((size_t)grid * RW::NP, NAN);
    sim::launch(grid, LT, [&]() { learnt_adj_kernel<M>(params, pc, s, a, dt, n, g, gs, ga, partials.data()); });
    for (int p = 0; p < RW::NP; ++p) {
      float acc = 0.f;
      for (int c = 0; c < grid; ++c) acc += partials[(size_t)c * RW::NP + p];
      gp[p] = acc;
    }
  } catch (const std::exception& ex) {
    sim::fail(std::string("exception: ") + ex.what());
  }
  return report(err, err_len);
}
}  // namespace

extern "C" int hc_lnsim_num_params(int system) {
  return system == 1 ? LearntRows<LearntWing<float>::NPH>::NP : LearntRows<LearntQuad<float>::NPH>::NP;
}
// learnt_fwd_kernel + learnt_adj_kernel <<<grid, 128>>>, system 0 = quadrotor, 1 = fixed wing
extern "C" int hc_lnsim_step_and_adjoint(int system, const float* params, const float* pc, const float* s,
                                         const float* a, float dt, int n, int grid, const float* g, float* out,
                                         float* gs, float* ga, float* gp, char* err, int err_len) {
  if (system == 1) return run<LearntWing<float>>(params, pc, s, a, dt, n, grid, g, out, gs, ga, gp, err, err_len);
  return run<LearntQuad<float>>(params, pc, s, a, dt, n, grid, g, out, gs, ga, gp, err, err_len);
}
