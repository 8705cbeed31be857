/* Plain-C caller of the boundary (include/apg_b200.h): one train-step evaluation of the cartpole rollout
 * (scripts/train_cartpole.py:118-155: Net(4, h) -> h dynamics steps -> cartpole_loss_mpc -> backward) from HOST
 * buffers, no Python and no torch in the process.
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_demo.c -o c_abi_demo -Lapg_trajectory_tracking_b200 -lapg_b200 -lm \
 *       -Wl,-rpath,$PWD/apg_trajectory_tracking_b200
 *   ./c_abi_demo            exit 0: loss printed, directional derivative of the analytic gradient checked
 *                           exit 3: no CUDA device (the library has no CPU path and says so)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "apg_b200.h"

static unsigned int rng_state = 12345u;
static float uniform_pm1(void) { /* LCG, fixed seed: the demo is deterministic */
  rng_state = rng_state * 1664525u + 1013904223u;
  return (float)((rng_state >> 8) & 0xffffff) / 8388608.0f - 1.0f;
}

static int value_and_grad(const apg_config* cfg, const float* params, const float* states, float* loss, float* grad) {
  /* cartpole: the policy input is the raw state, the reference is made from it inside (make_reference) */
  return apg_rollout_value_and_grad_host(cfg, params, states, states, NULL, NULL, NULL, loss, grad);
}

int main(void) {
  enum { N = 256, H = 5 };
  apg_config cfg;
  int i, np, rc;
  float *params, *grad, *states, *dir, *trial;
  float loss = 0.f, loss_p = 0.f, loss_m = 0.f, rel, slope = 0.f, fd;
  const float eps = 1e-2f;

  for (i = 0; i < (int)(sizeof cfg / sizeof(int)); ++i) ((int*)&cfg)[i] = 0;
  cfg.system = APG_SYS_CARTPOLE; cfg.mode = APG_MODE_CONCURRENT; cfg.net = APG_NET_SIMPLE;
  cfg.n_drones = N; cfg.horizon = H; cfg.state_feat = 4; cfg.ref_len = 0; cfg.ref_dim = 0; cfg.out_dim = H;
  cfg.dt = 0.05f;
  /* csrc/apg_math.cuh CartC: masscart, masspole, length, max_force_mag, friction (config_cartpole.json; friction 0.5 as in
   * cartpole_dynamics.py:34) */
  cfg.phys[0] = 1.0f; cfg.phys[1] = 0.1f; cfg.phys[2] = 0.5f; cfg.phys[3] = 30.0f; cfg.phys[4] = 0.5f;

  np = apg_num_params(&cfg);
  if (np <= 0) { fprintf(stderr, "bad config: %s\n", apg_error_string(np)); return 2; }
  params = (float*)malloc(sizeof(float) * np); grad = (float*)malloc(sizeof(float) * np);
  dir = (float*)malloc(sizeof(float) * np);    trial = (float*)malloc(sizeof(float) * np);
  states = (float*)malloc(sizeof(float) * N * 4);
  for (i = 0; i < np; ++i) { params[i] = 0.2f * uniform_pm1(); dir[i] = uniform_pm1(); }
  for (i = 0; i < N; ++i) {
    states[4 * i + 0] = 2.4f * uniform_pm1(); states[4 * i + 1] = 1.5f * uniform_pm1();
    states[4 * i + 2] = 0.5f * uniform_pm1(); states[4 * i + 3] = 1.5f * uniform_pm1();
  }

  rc = value_and_grad(&cfg, params, states, &loss, grad);
  if (rc == APG_ERR_NO_DEVICE) { fprintf(stderr, "apg_b200: %s\n", apg_error_string(rc)); return 3; }
  if (rc != 0) { fprintf(stderr, "apg_b200 error %d: %s\n", rc, apg_error_string(rc)); return 1; }

  /* directional derivative of the analytic gradient against a central difference of the loss */
  for (i = 0; i < np; ++i) slope += grad[i] * dir[i];
  for (i = 0; i < np; ++i) trial[i] = params[i] + eps * dir[i] / sqrtf((float)np);
  rc = value_and_grad(&cfg, trial, states, &loss_p, grad);
  for (i = 0; i < np; ++i) trial[i] = params[i] - eps * dir[i] / sqrtf((float)np);
  if (rc == 0) rc = value_and_grad(&cfg, trial, states, &loss_m, grad);
  if (rc != 0) { fprintf(stderr, "apg_b200 error %d: %s\n", rc, apg_error_string(rc)); return 1; }
  fd = (loss_p - loss_m) / (2.f * eps / sqrtf((float)np));
  rel = fabsf(fd - slope) / fmaxf(fabsf(slope), 1e-6f);
  printf("drones %d horizon %d params %d\nloss %.6f\nd loss along a random direction: analytic %.5f, central "
         "difference %.5f (rel. diff %.2e)\n", N, H, np, loss, slope, fd, rel);
  free(params); free(grad); free(dir); free(trial); free(states);
  return rel < 2e-2f ? 0 : 4;
}
