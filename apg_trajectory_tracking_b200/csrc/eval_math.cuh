// Per-drone logic of the closed-loop evaluation rollout on table references: which reference rows
// the policy sees, the divergence / stability test and the reset rule.  `__host__ __device__` like apg_math.cuh: the
// sm_100a kernel (csrc/eval_kernels.cu) and the g++ test harness (tests/hostcheck) compile the same source.
//
// Reference behaviour restated here (paths relative to the reference checkout):
//   window      neural_control/trajectory/random_traj.py:62-81   (Random.get_ref_traj)
//   policy in   neural_control/dataset.py:155-204                (QuadDataset.prepare_data on (state, window))
//   divergence  neural_control/trajectory/random_traj.py:83-92, scripts/evaluate_drone.py:171-185
//   stability   neural_control/environments/drone_env.py:66-74   (get_is_stable: |roll|, |pitch| < thresh)
#pragma once
#include "apg_math.cuh"

namespace apg {

struct EvalParams {
  int steps;            // max_nr_steps
  int table_rows;       // RL: rows per reference table
  int test_time;        // 1: a diverged / unstable drone stops; 0: it is reset onto the reference and goes on
  float thresh_div, thresh_stable;
};

// Random.get_ref_traj: rows the policy sees at current index ci, and the index afterwards.
//   regular (ci < RL - h): rows ci+1 .. ci+h, index advances;  end of table: rows ci .. RL-1 then padding, index stays
APG_HD void eval_window_plan(int ci, int RL, int h, int* start, int* nreal, int* ci_next) {
  if (ci >= RL - h) { *start = ci; *nreal = RL - ci; *ci_next = ci; }
  else { *start = ci + 1; *nreal = h; *ci_next = ci + 1; }
}

// Element (r, c) of the policy's reference input: prepare_data applied to window row r of table `tab` [RL][9].
// pos_c: the drone's position component c (used for c < 3); vel_c: its velocity component c - 6 (used for c >= 6).
// Rows beyond nreal are [last table position, 0 0 0, 0 0 0].
APG_HD float eval_in_ref_elem(const float* tab, int RL, int start, int nreal, int r, int c, float pos_c, float vel_c) {
  const bool real = r < nreal;
  const float* row = tab + (size_t)(real ? start + r : RL - 1) * 9;
  if (c < 3) return row[c] - pos_c;
  const float v = real ? row[c < 6 ? c + 3 : c] : 0.f;      // reference velocity column (6..8)
  return c < 6 ? v : v - vel_c;
}

// After the dynamics step: divergence to the table point at the walking index, stability, stop / reset.
// s: the new state (in), possibly replaced by the reference state (out).  Returns the divergence.
APG_HD float eval_post_step(float* s, const float* tab, int ci, const EvalParams& e, int* alive) {
  const float* row = tab + (size_t)ci * 9;
  const float dx = row[0] - s[0], dy = row[1] - s[1], dz = row[2] - s[2];
  const float div = sqrt_(dx * dx + dy * dy + dz * dz);
  float ar = s[3] < 0.f ? -s[3] : s[3], ap = s[4] < 0.f ? -s[4] : s[4];
  const bool stable = ar < e.thresh_stable && ap < e.thresh_stable;
  if (div > e.thresh_div || !stable) {
    if (e.test_time) {
      *alive = 0;
    } else {                                   // Random.get_current_full_state: [table row (9), 0 0 0]
#pragma unroll
      for (int j = 0; j < 9; ++j) s[j] = row[j];
      s[9] = s[10] = s[11] = 0.f;
    }
  }
  return div;
}

}  // namespace apg
