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
#include "prep_math.cuh"

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

// ---------------------------------------------------------------------------------------------------------
// Fixed wing: FixedWingEvaluator.fly_to_point (scripts/evaluate_fixed_wing.py:46-130), project_to_line
// (neural_control/trajectory/q_funcs.py:6-18), SimpleWingEnv.step (environments/wing_env.py:44-58).
// ---------------------------------------------------------------------------------------------------------
struct WingEvalParams {
  int steps;            // max_steps
  int n_targets;        // K target points per drone
  int test_time;        // 1: stop at divergence / instability; 0: reset onto the line and go on
  int h;                // horizon of the dataset's straight-line reference
  float thresh_div, thresh_stable;
  float vlen;           // 12 * dataset dt (length of one reference step)
  float des_speed;      // 11.5: speed a reset drone is given towards the target
};

// projection of p onto the line through a and b; a itself when a == b
APG_HD void project_to_line3(const float* a, const float* b, const float* p, float* out) {
  const float abx = b[0] - a[0], aby = b[1] - a[1], abz = b[2] - a[2];
  if (abx == 0.f && aby == 0.f && abz == 0.f) { out[0] = a[0]; out[1] = a[1]; out[2] = a[2]; return; }
  const float nrm = abx * abx + aby * aby + abz * abz;
  const float t = (p[0] - a[0]) * abx + (p[1] - a[1]) * aby + (p[2] - a[2]) * abz;
  out[0] = a[0] + abx * t / nrm; out[1] = a[1] + aby * t / nrm; out[2] = a[2] + abz * t / nrm;
}

// Per-drone bookkeeping of one flight.  env: the environment's state; obs: the state the policy is shown (the
// evaluator's local `state`, which is NOT refreshed by a reset, :118-129); prev_pos: position after the previous step.
struct WingEvalDrone {
  float env[12], obs[12], prev_pos[3], line_start[3];
  int ti, alive, nsteps;
  float dt_sum, dt_cnt;      // sum / count of the evaluator's div_target list
};

APG_HD void wing_eval_init(WingEvalDrone& D, const float* init_state, int alive) {
  for (int j = 0; j < 12; ++j) { D.env[j] = init_state[j]; D.obs[j] = init_state[j]; }
  for (int j = 0; j < 3; ++j) { D.prev_pos[j] = init_state[j]; D.line_start[j] = init_state[j]; }
  D.ti = 0; D.alive = alive; D.nsteps = 0; D.dt_sum = 0.f; D.dt_cnt = 0.f;
}

// After env.step returned nxt: divergence to the current line, target switching, stop / reset (:82-129).
// targets: [K][3] of this drone.  Returns the divergence to the line (div_to_linear entry of this step).
APG_HD float wing_eval_post_step(WingEvalDrone& D, const float* nxt, const float* targets, const WingEvalParams& e) {
  const float* tgt = targets + D.ti * 3;
  const bool stable = (nxt[6] < 0.f ? -nxt[6] : nxt[6]) < e.thresh_stable &&
                      (nxt[7] < 0.f ? -nxt[7] : nxt[7]) < e.thresh_stable;
  float on_line[3];
  project_to_line3(D.line_start, tgt, nxt, on_line);
  const float dx = on_line[0] - nxt[0], dy = on_line[1] - nxt[1], dz = on_line[2] - nxt[2];
  const float div = sqrt_(dx * dx + dy * dy + dz * dz);
  ++D.nsteps;
  bool finished = false;
  if (nxt[0] > tgt[0]) {                                        // passed the target (:93-110)
    float t_on[3];
    project_to_line3(D.prev_pos, nxt, tgt, t_on);
    const float ex = t_on[0] - tgt[0], ey = t_on[1] - tgt[1], ez = t_on[2] - tgt[2];
    D.dt_sum += sqrt_(ex * ex + ey * ey + ez * ez);
    D.dt_cnt += 1.f;
    if (D.ti < e.n_targets - 1) {
      ++D.ti;
      for (int j = 0; j < 3; ++j) D.line_start[j] = nxt[j];
    } else {
      finished = true;
    }
  }
  for (int j = 0; j < 12; ++j) { D.obs[j] = nxt[j]; D.env[j] = nxt[j]; }
  for (int j = 0; j < 3; ++j) D.prev_pos[j] = nxt[j];
  if (finished) { D.alive = 0; return div; }
  if (!stable || div > e.thresh_div) {                          // judged against the target of THIS step (:112-129)
    D.dt_cnt += 1.f;
    if (e.test_time) {
      const float ex = nxt[0] - tgt[0], ey = nxt[1] - tgt[1], ez = nxt[2] - tgt[2];
      D.dt_sum += sqrt_(ex * ex + ey * ey + ez * ez);
      D.alive = 0;
    } else {
      D.dt_sum += e.thresh_div;
      const float vx = tgt[0] - on_line[0], vy = tgt[1] - on_line[1], vz = tgt[2] - on_line[2];
      const float vn = sqrt_(vx * vx + vy * vy + vz * vz);
      for (int j = 0; j < 12; ++j) D.env[j] = 0.f;
      D.env[0] = on_line[0]; D.env[1] = on_line[1]; D.env[2] = on_line[2];
      D.env[3] = vx / vn * e.des_speed; D.env[4] = vy / vn * e.des_speed; D.env[5] = vz / vn * e.des_speed;
    }
  }
  return div;
}

// end of the flight: a drone that used all its steps gets one more div_target entry (:130-132)
APG_HD void wing_eval_finish(WingEvalDrone& D, const WingEvalParams& e) {
  if (D.alive && D.nsteps == e.steps) { D.dt_sum += e.thresh_div; D.dt_cnt += 1.f; }
}

// ---------------------------------------------------------------------------------------------------------
// Cartpole: Evaluator.evaluate_in_environment (scripts/evaluate_cartpole.py:78-262) with CartpoleWrapper
// (neural_control/controllers/network_wrapper.py:101-113), CartPoleEnv._step / is_upright
// (neural_control/environments/cartpole_env.py).  The policy input aliases the environment's float32 state and the
// network zeroes column 0 in place (models/simple_model.py:21), so from the second policy call on the cart position
// the ENVIRONMENT integrates starts from 0 every step; the first step still integrates from the start position.
// ---------------------------------------------------------------------------------------------------------
struct CartpoleEvalParams {
  int steps;            // max_steps
  int burn_in;          // |theta| enters the mean only for step index > burn_in
  float thresh_div;     // upright: -thresh_div < theta < thresh_div
};

struct CartpoleEvalDrone {
  float s[4];
  int alive, nsteps;
  float ang_sum, ang_cnt, vel_sum;
};

APG_HD void cartpole_eval_init(CartpoleEvalDrone& D, const float* init_state, int alive) {
  for (int j = 0; j < 4; ++j) D.s[j] = init_state[j];
  D.alive = alive; D.nsteps = 0; D.ang_sum = 0.f; D.ang_cnt = 0.f; D.vel_sum = 0.f;
}

// what the policy call of step i leaves in the environment state (the net's own input has column 0 zeroed anyway)
APG_HD void cartpole_eval_before_policy(CartpoleEvalDrone& D, int i) {
  if (i > 0) D.s[0] = 0.f;
}

// after the dynamics returned nxt (theta wrapped into (-pi, pi] by _step): bookkeeping and the upright test
APG_HD void cartpole_eval_post_step(CartpoleEvalDrone& D, float* nxt, int i, const CartpoleEvalParams& e) {
  const float PI_F = 3.14159265358979323846f;
  if (nxt[2] > PI_F) nxt[2] -= 2.f * PI_F;
  else if (nxt[2] <= -PI_F) nxt[2] += 2.f * PI_F;
  for (int j = 0; j < 4; ++j) D.s[j] = nxt[j];
  ++D.nsteps;
  D.vel_sum += nxt[1] < 0.f ? -nxt[1] : nxt[1];
  if (i > e.burn_in) { D.ang_sum += nxt[2] < 0.f ? -nxt[2] : nxt[2]; D.ang_cnt += 1.f; }
  if (!(nxt[2] > -e.thresh_div && nxt[2] < e.thresh_div)) D.alive = 0;
}

}  // namespace apg
