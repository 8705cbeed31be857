// Data-format kernels on the input side of the rollout: the reference's `prepare_data`
// layouts, window sampling and polynomial reference rows, produced on the device so that a train step only needs
// the RAW samples (quad: 408 B per drone instead of 828 B of prepared tensors).  All of them are element-wise,
// HBM-bound and fully coalesced: consecutive threads own consecutive output floats (or consecutive drones for the
// 12 / 15-float per-drone records).  The bodies live in prep_math.cuh (`__host__ __device__`, host-checked); the
// kernels here only map threads to indices.
#include "prep_math.cuh"
#include "kernels.h"

namespace apg {

namespace {
constexpr int PREP_THREADS = 256;
inline unsigned blocks_for(size_t total) { return (unsigned)((total + PREP_THREADS - 1) / PREP_THREADS); }
__device__ __forceinline__ size_t flat_tid() { return (size_t)blockIdx.x * blockDim.x + threadIdx.x; }
}  // namespace

__global__ void apg_prep_quad_rows_kernel(const float* s, const float* ref, size_t total, int L, float* in_ref,
                                          float* ref_out) {
  const size_t idx = flat_tid();
  if (idx < total) prep_quad_rows_body(idx, s, ref, L, in_ref, ref_out);
}

__global__ void apg_prep_quad_state_kernel(const float* s, size_t n, float* cur_out, float* in_state) {
  const size_t i = flat_tid();
  if (i < n) prep_quad_state_body(i, s, cur_out, in_state);
}

cudaError_t launch_prepare_quad(const float* states, const float* ref, int n, int L, float* in_state, float* cur_out,
                                float* in_ref, float* ref_out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const size_t total = (size_t)n * L * 9;
  // rows first: they read the raw drone position that the state kernel may zero in place
  if ((in_ref || ref_out) && total)
    APG_LAUNCH(blocks_for(total), PREP_THREADS, 0, st, apg_prep_quad_rows_kernel)(states, ref, total, L, in_ref, ref_out);
  if (in_state || cur_out)
    APG_LAUNCH(blocks_for((size_t)n), PREP_THREADS, 0, st, apg_prep_quad_state_kernel)(states, (size_t)n, cur_out, in_state);
  return cudaGetLastError();
}

__global__ void apg_prep_wing_line_kernel(const float* __restrict__ s, const float* __restrict__ target, float vlen,
                                          int h, size_t total, float* __restrict__ ref_out) {
  const size_t idx = flat_tid();
  if (idx < total) prep_wing_line_body(idx, s, target, vlen, h, ref_out);
}

__global__ void apg_prep_wing_state_kernel(const float* s, const float* target, const NormConsts nc, float vlen,
                                           int h, size_t n, float* in_state, float* in_ref, float* cur_out) {
  const size_t i = flat_tid();
  if (i < n) prep_wing_state_body(i, s, target, nc, vlen, h, in_state, in_ref, cur_out);
}

cudaError_t launch_prepare_wing(const float* states, const float* targets, const float* mean_host,
                                const float* std_host, float dt, int h, int n, float* in_state, float* cur_out,
                                float* in_ref, float* ref_out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  NormConsts nc;
  for (int j = 0; j < 12; ++j) { nc.mean[j] = mean_host[j]; nc.std_[j] = std_host[j]; }
  const float vlen = (float)(12.0 * (double)dt);          // `12 * self.dt` is a Python double, cast once
  const size_t total = (size_t)n * h * 3;
  if (ref_out && total)
    APG_LAUNCH(blocks_for(total), PREP_THREADS, 0, st, apg_prep_wing_line_kernel)(states, targets, vlen, h, total, ref_out);
  if (in_state || in_ref || cur_out)
    APG_LAUNCH(blocks_for((size_t)n), PREP_THREADS, 0, st, apg_prep_wing_state_kernel)(states, targets, nc, vlen, h, (size_t)n,
                                                                               in_state, in_ref, cur_out);
  return cudaGetLastError();
}

__global__ void apg_poly_rows_kernel(const float* __restrict__ coef, size_t rows, int L, float t_first, float dt,
                                     float* __restrict__ out) {
  const size_t row = flat_tid();
  if (row < rows) poly_rows_body(row, coef, L, t_first, dt, out);
}

cudaError_t launch_poly_reference(const float* coef, int n, int L, float t_first, float dt, float* out,
                                  cudaStream_t st) {
  const size_t rows = (size_t)n * L;
  if (rows == 0) return cudaSuccess;
  APG_LAUNCH(blocks_for(rows), PREP_THREADS, 0, st, apg_poly_rows_kernel)(coef, rows, L, t_first, dt, out);
  return cudaGetLastError();
}

__global__ void apg_sample_windows_kernel(const float* __restrict__ traj, int W, int L, int stride, size_t total_ref,
                                          size_t total, float* __restrict__ states, float* __restrict__ refs) {
  const size_t idx = flat_tid();
  if (idx < total) sample_windows_body(idx, traj, W, L, stride, total_ref, states, refs);
}

cudaError_t launch_sample_windows(const float* traj, int W, int L, int stride, int n, float* states, float* refs,
                                  cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const size_t total_ref = (size_t)n * L * 9, total = total_ref + (size_t)n * 12;
  APG_LAUNCH(blocks_for(total), PREP_THREADS, 0, st, apg_sample_windows_kernel)(traj, W, L, stride, total_ref, total, states,
                                                                        refs);
  return cudaGetLastError();
}

__global__ void apg_ref_table_kernel(const float* __restrict__ traj, int W, int nth, float speed, float z_offset,
                                     size_t rows, float* __restrict__ out) {
  const size_t k = flat_tid();
  if (k < rows) ref_table_body(k, traj, W, nth, speed, z_offset, out);
}

cudaError_t launch_reference_table(const float* traj, int W, int nth, float speed, float z_offset, int rows,
                                   float* out, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  APG_LAUNCH(blocks_for((size_t)rows), PREP_THREADS, 0, st, apg_ref_table_kernel)(traj, W, nth, speed, z_offset, (size_t)rows,
                                                                         out);
  return cudaGetLastError();
}

__global__ void apg_poly_march_kernel(const double* __restrict__ coef, int degree, const double* __restrict__ rot,
                                      const double* __restrict__ start, int n, double x_start, double x_range,
                                      double dist_points, int hover, int max_rows, float* __restrict__ out,
                                      int* __restrict__ ref_len) {
  const size_t i = flat_tid();
  if (i >= (size_t)n) return;
  const int len = poly_march_body(coef + i * (degree + 1), degree, rot + i * 9, start ? start + i * 3 : nullptr,
                                  x_start, x_range, dist_points, hover, max_rows, out + i * (size_t)max_rows * 3);
  if (ref_len) ref_len[i] = len;
}

cudaError_t launch_polynomial_points(const double* coef, int degree, const double* rot, const double* start, int n,
                                     double x_start, double x_range, double dist_points, int hover, int max_rows,
                                     float* out, int* ref_len, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  // one thread per trajectory (sequential march): small blocks spread few trajectories over many SMs
  APG_LAUNCH((n + 31) / 32, 32, 0, st, apg_poly_march_kernel)(coef, degree, rot, start, n, x_start, x_range, dist_points,
                                                      hover, max_rows, out, ref_len);
  return cudaGetLastError();
}

}  // namespace apg
