// Per-drone / per-element math of the data formats on the INPUT side of the rollout:
// turning raw (state, reference) samples into the four tensors of a train batch, and producing the raw samples
// themselves from polynomial trajectories.  Same conventions as apg_math.cuh: `__host__ __device__`, templated on
// the scalar type, compiled into the sm_100a kernels (csrc/prep_kernels.cu) AND with g++ into the test-only
// harness (tests/hostcheck) that pins it on the reference's own `prepare_data` outputs (tests/golden/prep_data.npz).
//
// Reference behaviour restated here (paths relative to the reference checkout):
//   quad prepare   neural_control/dataset.py:155-204   (QuadDataset.prepare_data)
//   wing prepare   neural_control/dataset.py:309-350   (WingDataset._compute_target_pos / prepare_data)
//   windows        neural_control/environments/drone_env.py:232-269 (full_state_training_data)
//   polynomials    SURVEY.md 8d synthetic quad inputs (degree-5 per axis, positions + analytic velocities)
#pragma once
#include <stddef.h>
#include "apg_math.cuh"

namespace apg {

// ---------------------------------------------------------------------------------------------------------
// Quadrotor.  Reference row layout: [pos(3), euler(3), vel(3)].
// ---------------------------------------------------------------------------------------------------------
template <typename T>
struct QuadPrep {
  static constexpr int S = 12, F = 15, RW = 9;

  // Element c (0..8) of one reference row r (9 floats) of a drone with raw state s (12 floats):
  //   ref_out[c] = r[c] - pos[c]            (c < 3; dataset.py:170-173)      else r[c]
  //   in_ref[c]  = [rel. position, reference velocity, reference velocity - drone velocity][c]   (:194-201)
  APG_HD static T ref_elem(const T* s, const T* r, int c) { return c < 3 ? r[c] - s[c] : r[c]; }
  APG_HD static T in_ref_elem(const T* s, const T* r, int c) {
    if (c < 3) return r[c] - s[c];
    if (c < 6) return r[c + 3];
    return r[c] - s[c];                          // c in 6..8: ref velocity minus drone velocity (state cols 6..8)
  }
  // Per drone: cur = s with the position zeroed (:174), in_state = state_preprocessing(cur) (:177-191)
  APG_HD static void drone(const T* s, T* cur, T* in_state) {
    T c[S];
    c[0] = c[1] = c[2] = T(0);
    for (int j = 3; j < S; ++j) c[j] = s[j];
    if (in_state) Quad<T>::features(c, in_state);
    if (cur) for (int j = 0; j < S; ++j) cur[j] = c[j];
  }
};

// ---------------------------------------------------------------------------------------------------------
// Fixed wing.  Raw sample = (state (12), target position (3)); the loss reference is a straight line of
// 12 m/s towards the target, the policy sees the normalised state without position and the line's last point.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
struct WingPrep {
  static constexpr int S = 12, F = 9;

  // unit vector from the drone to the target (dataset.py:339-343)
  APG_HD static void unit(const T* s, const T* target, T* u) {
    const T r0 = target[0] - s[0], r1 = target[1] - s[1], r2 = target[2] - s[2];
    const T nrm = sqrt_(r0 * r0 + r1 * r1 + r2 * r2);
    u[0] = r0 / nrm; u[1] = r1 / nrm; u[2] = r2 / nrm;
  }
  // reference point k (0-based) of the straight line: pos + (unit * (12 dt)) * (k + 1)   (:311-321)
  APG_HD static T line_elem(T sj, T uj, T vlen, int k) { return sj + (uj * vlen) * T(k + 1); }
  // per drone: in_state = ((s - mean) / std)[3:] (:336), in_ref = line[h-1] - pos (:346)
  APG_HD static void drone(const T* s, const T* target, const float* mean, const float* std_, T vlen, int h,
                           T* in_state, T* in_ref) {
    for (int j = 0; j < F; ++j) in_state[j] = (s[3 + j] - T(mean[3 + j])) / T(std_[3 + j]);
    T u[3];
    unit(s, target, u);
    for (int j = 0; j < 3; ++j) in_ref[j] = line_elem(s[j], u[j], vlen, h - 1) - s[j];
  }
};

// ---------------------------------------------------------------------------------------------------------
// Polynomial reference trajectories: per axis p(t) = sum_i c[i] t^i (degree DEG), row = [p(t), 0 0 0, p'(t)].
// ---------------------------------------------------------------------------------------------------------
template <typename T>
struct PolyTraj {
  static constexpr int DEG = 5, NC = DEG + 1;
  // coef: [3][NC] of one drone; out: 9 floats
  APG_HD static void row(const T* coef, T t, T* out) {
    for (int a = 0; a < 3; ++a) {
      const T* c = coef + a * NC;
      T p = c[DEG], d = T(0);
      for (int i = DEG - 1; i >= 0; --i) { d = d * t + p; p = p * t + c[i]; }     // Horner, value and derivative
      out[a] = p; out[3 + a] = T(0); out[6 + a] = d;
    }
  }
};

// =========================================================================================================
// Kernel bodies, indexed by the flat thread id.  The __global__ wrappers in prep_kernels.cu only compute the
// thread id and the bounds check; tests/hostcheck runs the SAME bodies in a loop over idx on the CPU.
// =========================================================================================================
struct NormConsts { float mean[12], std_[12]; };

// idx over the floats of in_ref / ref_out ([n][L][9]).  ref_out may alias ref: columns >= 3 are copied unchanged
// and a position column is read and written by its own thread only.
APG_HD void prep_quad_rows_body(size_t idx, const float* s, const float* ref, int L, float* in_ref, float* ref_out) {
  const size_t row = idx / 9;
  const int c = (int)(idx - row * 9);
  const size_t i = row / (size_t)L;
  const float* r = ref + row * 9;
  const float* si = s + i * 12;
  const float a = QuadPrep<float>::in_ref_elem(si, r, c);
  const float b = QuadPrep<float>::ref_elem(si, r, c);
  if (in_ref) in_ref[idx] = a;
  if (ref_out) ref_out[idx] = b;
}

// i over drones: position-zeroed state and the 15 policy features.  cur_out may alias s.
APG_HD void prep_quad_state_body(size_t i, const float* s, float* cur_out, float* in_state) {
  float si[12], ci[12], fi[15];
#pragma unroll
  for (int j = 0; j < 12; ++j) si[j] = s[i * 12 + j];
  QuadPrep<float>::drone(si, ci, fi);
  if (cur_out) {
#pragma unroll
    for (int j = 0; j < 12; ++j) cur_out[i * 12 + j] = ci[j];
  }
  if (in_state) {
#pragma unroll
    for (int j = 0; j < 15; ++j) in_state[i * 15 + j] = fi[j];
  }
}

// idx over the floats of the straight-line reference [n][h][3]
APG_HD void prep_wing_line_body(size_t idx, const float* s, const float* target, float vlen, int h, float* ref_out) {
  const size_t row = idx / 3;
  const int j = (int)(idx - row * 3);
  const size_t i = row / (size_t)h;
  const int k = (int)(row - i * h);
  float u[3];
  WingPrep<float>::unit(s + i * 12, target + i * 3, u);
  const float uj = j == 0 ? u[0] : (j == 1 ? u[1] : u[2]);
  ref_out[idx] = WingPrep<float>::line_elem(s[i * 12 + j], uj, vlen, k);
}

// i over drones: normalised state without position, relative last reference point, copy of the state
APG_HD void prep_wing_state_body(size_t i, const float* s, const float* target, const NormConsts& nc, float vlen,
                                 int h, float* in_state, float* in_ref, float* cur_out) {
  float si[12], ti[3], fi[9], ri[3];
#pragma unroll
  for (int j = 0; j < 12; ++j) si[j] = s[i * 12 + j];
#pragma unroll
  for (int j = 0; j < 3; ++j) ti[j] = target[i * 3 + j];
  WingPrep<float>::drone(si, ti, nc.mean, nc.std_, vlen, h, fi, ri);
  if (in_state) {
#pragma unroll
    for (int j = 0; j < 9; ++j) in_state[i * 9 + j] = fi[j];
  }
  if (in_ref) {
#pragma unroll
    for (int j = 0; j < 3; ++j) in_ref[i * 3 + j] = ri[j];
  }
  if (cur_out && cur_out != s) {
#pragma unroll
    for (int j = 0; j < 12; ++j) cur_out[i * 12 + j] = si[j];
  }
}

// row over the reference rows [n][L]: [p(t), 0 0 0, p'(t)] for t = t_first + k dt
APG_HD void poly_rows_body(size_t row, const float* coef, int L, float t_first, float dt, float* out) {
  constexpr int NC3 = 3 * PolyTraj<float>::NC;
  const size_t i = row / (size_t)L;
  const int k = (int)(row - i * L);
  float c[NC3], o[9];
#pragma unroll
  for (int j = 0; j < NC3; ++j) c[j] = coef[i * NC3 + j];
  PolyTraj<float>::row(c, t_first + (float)k * dt, o);
#pragma unroll
  for (int j = 0; j < 9; ++j) out[row * 9 + j] = o[j];
}

// Window sampling of one trajectory table [T][W] (W >= 9): sample i starts at row i*stride,
//   states[i] = [traj[i*stride][0:9], 0 0 0],  refs[i][k] = traj[i*stride + k + 1][0:9]
// idx < total_ref: a float of refs; otherwise a float of states.
APG_HD void sample_windows_body(size_t idx, const float* traj, int W, int L, int stride, size_t total_ref,
                                float* states, float* refs) {
  if (idx < total_ref) {
    const size_t row = idx / 9;
    const int c = (int)(idx - row * 9);
    const size_t i = row / (size_t)L;
    const int k = (int)(row - i * L);
    refs[idx] = traj[(i * stride + k + 1) * W + c];
  } else {
    const size_t q = idx - total_ref;
    const size_t i = q / 12;
    const int c = (int)(q - i * 12);
    states[q] = c < 9 ? traj[(i * stride) * W + c] : 0.f;
  }
}

// Reference table of the evaluation / training data from one raw trajectory file: load_prepare_trajectory
// (neural_control/trajectory/generate_trajectory.py:566-603) + the z offset of Random.__init__
// (trajectory/random_traj.py:35).  Raw rows [T][W >= 10] = [pos(3), quaternion w x y z (4), vel(3), ...] sampled at
// 0.01 s; table row k = raw row k * nth -> [pos (z + z_offset), euler(q) * speed, vel * speed * 2].
// euler(q) = q_funcs.quaternion_to_euler (:38-41) = pyquaternion's Quaternion.yaw_pitch_roll on the normalised
// quaternion, returned as [roll, pitch, yaw].  pyquaternion is NOT in this image (nor pinned by the reference's
// setup.py); its published formula is restated here:
//   yaw = atan2(2(wz - xy), 1 - 2(y^2 + z^2)),  pitch = asin(2(wy + zx)),  roll = atan2(2(wx - yz), 1 - 2(x^2 + y^2))
template <typename T>
struct RefTable {
  APG_HD static void euler(const T* q4, T* rpy) {
    const T nrm = sqrt_(q4[0] * q4[0] + q4[1] * q4[1] + q4[2] * q4[2] + q4[3] * q4[3]);
    const T w = q4[0] / nrm, x = q4[1] / nrm, y = q4[2] / nrm, z = q4[3] / nrm;
    rpy[2] = atan2_(T(2) * (w * z - x * y), T(1) - T(2) * (y * y + z * z));
    rpy[1] = asin_(T(2) * (w * y + z * x));
    rpy[0] = atan2_(T(2) * (w * x - y * z), T(1) - T(2) * (x * x + y * y));
  }
  APG_HD static void row(const T* raw, T speed, T z_offset, T* o) {
    T rpy[3];
    euler(raw + 3, rpy);
    o[0] = raw[0]; o[1] = raw[1]; o[2] = raw[2] + z_offset;
    o[3] = rpy[0] * speed; o[4] = rpy[1] * speed; o[5] = rpy[2] * speed;
    o[6] = raw[7] * speed * T(2); o[7] = raw[8] * speed * T(2); o[8] = raw[9] * speed * T(2);
  }
};

// k over the table rows
APG_HD void ref_table_body(size_t k, const float* traj, int W, int nth, float speed, float z_offset, float* out) {
  float raw[10], o[9];
#pragma unroll
  for (int j = 0; j < 10; ++j) raw[j] = traj[(k * (size_t)nth) * W + j];
  RefTable<float>::row(raw, speed, z_offset, o);
#pragma unroll
  for (int j = 0; j < 9; ++j) out[k * 9 + j] = o[j];
}

// Polynomial reference of the evaluation (neural_control/trajectory/polynomial.py): random_polynomial's march along
// the fitted polynomial in steps of dist_points of arc length (:99-113), the lift to 3D and rotation (:115-125), the
// shift to the drone's position and the hover padding at both ends (:41-47).  One trajectory per call (the march is
// sequential); the march itself runs in double like the reference's numpy code, so the number of points agrees.
//   coef [degree + 1] highest power first (np.poly1d order), rot [9] row-major 3x3, start [3] (may be null: no shift),
//   all double like the numpy arrays they come from (x^5 at x ~ 20 amplifies a float32 rounding of the fit to ~1e-3)
//   out [max_rows][3]: hover copies of the first point | the marched points | hover copies of the last point
// Returns the number of rows the full reference has (rows beyond max_rows are not written).
APG_HD int poly_march_body(const double* coef, int degree, const double* rot, const double* start, double x_start,
                           double x_range, double dist_points, int hover, int max_rows, float* out) {
  double c[12];
  for (int i = 0; i <= degree; ++i) c[i] = coef[i];
  auto poly = [&](double x) { double y = 0.0; for (int i = 0; i <= degree; ++i) y = y * x + c[i]; return y; };
  auto grad = [&](double x) {
    double gsum = 0.0;
    for (int i = 0; i < degree; ++i) {
      double pw = 1.0;
      for (int e = 0; e < degree - i - 1; ++e) pw *= x;
      gsum += (double)(degree - i) * c[i] * pw;
    }
    return gsum;
  };
  auto lift = [&](double x, double y, double* p3) {
    for (int j = 0; j < 3; ++j) p3[j] = x * rot[j] + y * rot[6 + j];      // [x, 0, y] @ rot
  };
  double x = x_start;
  const double x_final = x + x_range, step = dist_points;
  double p0[3], p[3], sh[3];
  lift(x, poly(x), p0);
  for (int j = 0; j < 3; ++j) sh[j] = start ? start[j] - p0[j] : 0.0;
  int row = 0;
  auto emit = [&](const double* q) {
    if (row < max_rows) { for (int j = 0; j < 3; ++j) out[(size_t)row * 3 + j] = (float)(q[j] + sh[j]); }
    ++row;
  };
  for (int k = 0; k < hover; ++k) emit(p0);
  emit(p0);
  for (int j = 0; j < 3; ++j) p[j] = p0[j];
  while (x < x_final) {
    const double gr = grad(x);
    x += (1.0 / sqrt(1.0 + gr * gr)) * step;                     // (vec / norm(vec) * dist_points)[0]
    lift(x, poly(x), p);
    emit(p);
  }
  for (int k = 0; k < hover; ++k) emit(p);
  return row;
}

}  // namespace apg
