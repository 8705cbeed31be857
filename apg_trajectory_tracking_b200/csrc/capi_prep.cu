// extern "C" entry points of the data-format kernels (include/apg_b200.h, "input side" section).
#include <stdint.h>

#include "../../include/apg_b200.h"
#include "kernels.h"

using namespace apg;

namespace {
inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
}  // namespace

#define APG_API extern "C" __attribute__((visibility("default")))

APG_API int apg_prepare_quad(const float* states, const float* ref_states, int n, int ref_rows, float* in_state,
                             float* cur_out, float* in_ref, float* ref_out, void* stream) {
  if (!states || n < 0 || ref_rows < 0) return APG_ERR_BAD_CONFIG;
  if ((in_ref || ref_out) && !ref_states) return APG_ERR_BAD_CONFIG;
  if (!al16(states) || !al16(ref_states) || !al16(in_state) || !al16(cur_out) || !al16(in_ref) || !al16(ref_out))
    return APG_ERR_ALIGNMENT;
  return (int)launch_prepare_quad(states, ref_states, n, ref_rows, in_state, cur_out, in_ref, ref_out,
                                  static_cast<cudaStream_t>(stream));
}

APG_API int apg_prepare_wing(const float* states, const float* targets, const float* mean_host, const float* std_host,
                             float dt, int horizon, int n, float* in_state, float* cur_out, float* in_ref,
                             float* ref_out, void* stream) {
  if (!states || !targets || !mean_host || !std_host || n < 0 || horizon <= 0) return APG_ERR_BAD_CONFIG;
  if (!al16(states) || !al16(in_state) || !al16(cur_out) || !al16(ref_out)) return APG_ERR_ALIGNMENT;
  return (int)launch_prepare_wing(states, targets, mean_host, std_host, dt, horizon, n, in_state, cur_out, in_ref,
                                  ref_out, static_cast<cudaStream_t>(stream));
}

APG_API int apg_poly_reference(const float* coef, int n, int rows, float t_first, float dt, float* ref_out,
                               void* stream) {
  if (!coef || !ref_out || n < 0 || rows < 0) return APG_ERR_BAD_CONFIG;
  return (int)launch_poly_reference(coef, n, rows, t_first, dt, ref_out, static_cast<cudaStream_t>(stream));
}

APG_API int apg_sample_windows(const float* traj, int traj_rows, int traj_cols, int ref_rows, int stride, int n,
                               float* states, float* ref_states, void* stream) {
  if (!traj || !states || !ref_states || n < 0 || ref_rows < 0 || stride <= 0 || traj_cols < 9)
    return APG_ERR_BAD_CONFIG;
  // the last sample reads rows (n-1)*stride .. (n-1)*stride + ref_rows
  if (n > 0 && (long long)(n - 1) * stride + ref_rows > (long long)traj_rows - 1) return APG_ERR_BAD_CONFIG;
  return (int)launch_sample_windows(traj, traj_cols, ref_rows, stride, n, states, ref_states,
                                    static_cast<cudaStream_t>(stream));
}

APG_API int apg_reference_table(const float* traj, int traj_rows, int traj_cols, int take_every_nth,
                                float speed_factor, float z_offset, int table_rows, float* table_out, void* stream) {
  if (!traj || !table_out || traj_cols < 10 || take_every_nth < 1 || table_rows < 0 || traj_rows < 0)
    return APG_ERR_BAD_CONFIG;
  // table row k reads raw row k * take_every_nth (numpy's traj[::nth] has ceil(T / nth) rows)
  if (table_rows > 0 && (long long)(table_rows - 1) * take_every_nth > (long long)traj_rows - 1)
    return APG_ERR_BAD_CONFIG;
  return (int)launch_reference_table(traj, traj_cols, take_every_nth, speed_factor, z_offset, table_rows, table_out,
                                     static_cast<cudaStream_t>(stream));
}

APG_API int apg_polynomial_points(const double* coef, int degree, const double* rot, const double* start, int n,
                                  double x_start, double x_range, double dist_points, int hover_steps, int max_rows,
                                  float* points_out, int* ref_len_out, void* stream) {
  if (!coef || !rot || !points_out || n < 0 || degree < 1 || degree > 11 || max_rows < 1 || hover_steps < 0)
    return APG_ERR_BAD_CONFIG;
  if (!(dist_points > 0.0) || !(x_range >= 0.0)) return APG_ERR_BAD_CONFIG;
  return (int)launch_polynomial_points(coef, degree, rot, start, n, x_start, x_range, dist_points, hover_steps,
                                       max_rows, points_out, ref_len_out, static_cast<cudaStream_t>(stream));
}

// ---- learnt residual dynamics (learnt_kernels.cu): system = APG_SYS_QUAD / APG_SYS_WING
#include <string.h>

APG_API int apg_learnt_num_params(int system) {
  if (system != APG_SYS_QUAD && system != APG_SYS_WING) return APG_ERR_UNSUPPORTED;
  return learnt_num_params(system);
}

APG_API size_t apg_learnt_workspace_bytes(int system, int n) {
  if (system != APG_SYS_QUAD && system != APG_SYS_WING) return 0;
  if (n <= 0) return 256;
  return sizeof(float) * learnt_partials_floats(system, n, apg_sm_count()) + 256;
}

APG_API int apg_learnt_step(int system, const float* params, const float* phys, const float* state,
                            const float* action, float dt, int n, float* out, void* stream) {
  if (system != APG_SYS_QUAD && system != APG_SYS_WING) return APG_ERR_UNSUPPORTED;
  if (!params || !phys || !state || !action || !out || n < 0) return APG_ERR_BAD_CONFIG;
  PhysConsts pc;
  memcpy(pc.v, phys, sizeof(float) * MAX_PHYS);
  return (int)launch_learnt_fwd(system, params, pc, state, action, dt, n, out, apg_sm_count(),
                                static_cast<cudaStream_t>(stream));
}

APG_API int apg_learnt_step_adjoint(int system, const float* params, const float* phys, const float* state,
                                    const float* action, float dt, int n, const float* grad_out, float* grad_state,
                                    float* grad_action, float* grad_params, void* workspace, void* stream) {
  if (system != APG_SYS_QUAD && system != APG_SYS_WING) return APG_ERR_UNSUPPORTED;
  if (!params || !phys || !state || !action || !grad_out || n < 0) return APG_ERR_BAD_CONFIG;
  if (grad_params && !workspace) return APG_ERR_BAD_CONFIG;
  PhysConsts pc;
  memcpy(pc.v, phys, sizeof(float) * MAX_PHYS);
  return (int)launch_learnt_adj(system, params, pc, state, action, dt, n, grad_out, grad_state, grad_action,
                                grad_params, static_cast<float*>(workspace), apg_sm_count(),
                                static_cast<cudaStream_t>(stream));
}
