// Test-only harness: the bodies of the data-format kernels (csrc/prep_math.cuh) run in a loop over the flat thread
// index on the CPU, in the order the launchers of csrc/prep_kernels.cu issue them.  Not part of the product.
#include "prep_math.cuh"

using namespace apg;

extern "C" void hc_prepare_quad(const float* states, const float* ref, int n, int L, float* in_state, float* cur_out,
                                float* in_ref, float* ref_out) {
  const size_t total = (size_t)n * L * 9;
  if (in_ref || ref_out)
    for (size_t idx = 0; idx < total; ++idx) prep_quad_rows_body(idx, states, ref, L, in_ref, ref_out);
  if (in_state || cur_out)
    for (size_t i = 0; i < (size_t)n; ++i) prep_quad_state_body(i, states, cur_out, in_state);
}

extern "C" void hc_prepare_wing(const float* states, const float* targets, const float* mean, const float* std_,
                                float dt, int h, int n, float* in_state, float* cur_out, float* in_ref,
                                float* ref_out) {
  NormConsts nc;
  for (int j = 0; j < 12; ++j) { nc.mean[j] = mean[j]; nc.std_[j] = std_[j]; }
  const float vlen = (float)(12.0 * (double)dt);
  const size_t total = (size_t)n * h * 3;
  if (ref_out)
    for (size_t idx = 0; idx < total; ++idx) prep_wing_line_body(idx, states, targets, vlen, h, ref_out);
  if (in_state || in_ref || cur_out)
    for (size_t i = 0; i < (size_t)n; ++i)
      prep_wing_state_body(i, states, targets, nc, vlen, h, in_state, in_ref, cur_out);
}

extern "C" void hc_poly_reference(const float* coef, int n, int L, float t_first, float dt, float* out) {
  for (size_t row = 0; row < (size_t)n * L; ++row) poly_rows_body(row, coef, L, t_first, dt, out);
}

extern "C" void hc_sample_windows(const float* traj, int W, int L, int stride, int n, float* states, float* refs) {
  const size_t total_ref = (size_t)n * L * 9, total = total_ref + (size_t)n * 12;
  for (size_t idx = 0; idx < total; ++idx) sample_windows_body(idx, traj, W, L, stride, total_ref, states, refs);
}

extern "C" void hc_reference_table(const float* traj, int W, int nth, float speed, float z_offset, int rows,
                                   float* out) {
  for (size_t k = 0; k < (size_t)rows; ++k) ref_table_body(k, traj, W, nth, speed, z_offset, out);
}

extern "C" void hc_polynomial_points(const double* coef, int degree, const double* rot, const double* start, int n,
                                     double x_start, double x_range, double dist_points, int hover, int max_rows,
                                     float* out, int* ref_len) {
  for (int i = 0; i < n; ++i)
    ref_len[i] = poly_march_body(coef + (size_t)i * (degree + 1), degree, rot + (size_t)i * 9,
                                 start ? start + (size_t)i * 3 : nullptr, x_start, x_range, dist_points, hover,
                                 max_rows, out + (size_t)i * max_rows * 3);
}
