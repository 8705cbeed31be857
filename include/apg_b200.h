/* apg_b200 -- C ABI of the B200-native batched differentiable rollout (policy -> dynamics x horizon -> tracking
 * loss -> analytic policy gradient).
 *
 * The reference (lis-epfl/apg_trajectory_tracking) has no FFI: its "plugin interface" for this path is the Python
 * surface used by the trainers.  Each entry point below names the reference code it replaces:
 *
 *   apg_rollout_forward / apg_rollout_backward
 *        concurrent   : scripts/train_base.py:202-207 (net forward, sigmoid, reshape) +
 *                       scripts/train_drone.py:175-203 / scripts/train_fixed_wing.py:90-116 /
 *                       scripts/train_cartpole.py:127-155 (h dynamics steps, *_mpc_loss, loss.backward())
 *        recurrent    : scripts/train_drone.py:113-173 (autoregressive / LSTM loop)
 *   apg_dynamics_step / apg_dynamics_step_adjoint
 *        neural_control/dynamics/quad_dynamics_flightmare.py:125-216, fixed_wing_dynamics.py:95-267,
 *        cartpole_dynamics.py:50-119 (__call__) and what autograd records for them
 *   apg_rollout_value_and_grad_host
 *        the same train step called with HOST buffers (host<->device copies inside), i.e. what a CPU caller
 *        of the reference's trainer sees
 *
 * Conventions: all tensors fp32, row-major, contiguous, 16-byte aligned; `params` / `grad_params` are the tensors
 * of net.parameters() concatenated in order, each in its torch layout; device pointers unless the name ends in
 * _host; `stream` is a cudaStream_t; every function returns 0 on success, a negative apg error or a positive
 * cudaError_t otherwise (see apg_error_string).  No function allocates device memory except the *_host one.
 */
#ifndef APG_B200_H
#define APG_B200_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { APG_SYS_QUAD = 0, APG_SYS_WING = 1, APG_SYS_CARTPOLE = 2 };
enum { APG_MODE_CONCURRENT = 0, APG_MODE_AUTOREGRESSIVE = 1, APG_MODE_LSTM = 2 };
enum { APG_WINDOW_CUMULATIVE = 0, APG_WINDOW_RELATIVE = 1 };
enum { APG_NET_HUTTER_CONV = 0, APG_NET_HUTTER_LIN = 1, APG_NET_SIMPLE = 2, APG_NET_LSTM = 3 };
enum { APG_ERR_BAD_CONFIG = -1, APG_ERR_UNSUPPORTED = -2, APG_ERR_ALIGNMENT = -3, APG_ERR_NO_DEVICE = -4 };

#define APG_MAX_PHYS 48

typedef struct apg_config {
  int system;      /* APG_SYS_*                                                          */
  int mode;        /* APG_MODE_*                                                         */
  int window;      /* APG_WINDOW_* (recurrent modes; cumulative = the reference's forward) */
  int net;         /* APG_NET_*                                                          */
  int n_drones;    /* N: rows of every per-drone tensor                                  */
  int horizon;     /* h: dynamics steps per rollout                                      */
  int state_feat;  /* width of in_state rows (quad 15, wing 9, cartpole 4)               */
  int ref_len;     /* reference rows the policy sees per call (quad h, wing 1)           */
  int ref_dim;     /* width of one in_ref row (quad 9, wing 3)                           */
  int out_dim;     /* policy outputs per call (concurrent: A*h, recurrent: A)            */
  float dt;        /* integration step                                                   */
  float phys[APG_MAX_PHYS]; /* physical constants, layout of csrc/apg_math.cuh enums     */
} apg_config;

/* number of fp32 parameters of the policy described by cfg (== sum of net.parameters() sizes) */
int apg_num_params(const apg_config* cfg);
/* bytes of device workspace (packed weights, activation stash, per-CTA partials) the rollout calls need */
size_t apg_workspace_bytes(const apg_config* cfg);

/* Forward rollout.  in_state [N][state_feat], cur [N][S], in_ref [N][ref_len*ref_dim] (concurrent) or
 * [N][2h][ref_dim] (recurrent), ref [N][h][REFW] (concurrent; quad 9, wing 3, cartpole: NULL) or [N][2h][9],
 * h0c0 [2][N][8] (LSTM only).  Writes the scalar loss (sum over drones, horizon, components) to *loss (device)
 * and optionally states [N][h][S] / actions [N][h][A]. Leaves the activation stash in `workspace`. */
int apg_rollout_forward(const apg_config* cfg, const float* params, const float* in_state, const float* cur,
                        const float* in_ref, const float* ref, const float* h0c0, void* workspace, float* loss,
                        float* states_out, float* actions_out, void* stream);

/* Adjoint of the forward call that last used `workspace` (same cfg and inputs): writes
 * grad_params = grad_loss * d loss / d params  (n_params floats; entries of tensors the forward does not use
 * are 0, the Python wrapper leaves their .grad None like the reference). */
int apg_rollout_backward(const apg_config* cfg, const float* params, const float* in_state, const float* cur,
                         const float* in_ref, const float* ref, const float* h0c0, void* workspace,
                         float grad_loss, float* grad_params, void* stream);

/* apg_rollout_backward with the reference's optimizer step fused into the gradient reduction (optim.SGD(momentum=...),
 * scripts/train_base.py:139-143: buf = momentum * buf + g; p -= lr * buf).  `params_rw` (the vector the forward read) and
 * `momentum_buf` are updated in place; `grad_params` may be NULL.  Only for configurations served by the tcgen05 path
 * (apg_rollout_kernel_path(cfg) == 1); APG_ERR_UNSUPPORTED otherwise. */
int apg_rollout_backward_sgd(const apg_config* cfg, float* params_rw, const float* in_state, const float* cur,
                             const float* in_ref, const float* ref, const float* h0c0, void* workspace,
                             float grad_loss, float* grad_params, float* momentum_buf, float lr, float momentum,
                             void* stream);

/* One train-step evaluation with HOST buffers: H2D of params and inputs, forward, backward, D2H of the loss and
 * the gradient, synchronous.  Device buffers are cached inside the library between calls. */
int apg_rollout_value_and_grad_host(const apg_config* cfg, const float* params_host, const float* in_state_host,
                                    const float* cur_host, const float* in_ref_host, const float* ref_host,
                                    const float* h0c0_host, float* loss_host, float* grad_params_host);

/* Single dynamics step out = f(state, action) for N rows, and its vector-Jacobian product. */
int apg_dynamics_step(int system, const float* phys, const float* state, const float* action, float dt, int n,
                      float* out, void* stream);
int apg_dynamics_step_adjoint(int system, const float* phys, const float* state, const float* action, float dt,
                              int n, const float* grad_out, float* grad_state, float* grad_action, void* stream);
/* quadrotor featurizer (neural_control/dataset.py:207-220) and its vector-Jacobian product */
int apg_quad_features(const float* state, int n, float* feat, void* stream);
int apg_quad_features_adjoint(const float* state, const float* grad_feat, int n, float* grad_state, void* stream);

/* ---- input side of the path: the reference's dataset layouts produced on the device, so
 * that a train step only has to be handed the RAW samples.  Output pointers may be NULL (that tensor is skipped).
 *
 * apg_prepare_quad   neural_control/dataset.py:155-204 (QuadDataset.prepare_data): states [n][12], ref_states
 *                    [n][ref_rows][9] (pos, euler, vel) ->  in_state [n][15], cur_out [n][12] (position zeroed),
 *                    in_ref [n][ref_rows][9] (rel. pos, vel, vel - drone vel), ref_out [n][ref_rows][9] (rel. pos).
 *                    cur_out may alias states, ref_out may alias ref_states (the reference works in place too).
 * apg_prepare_wing   neural_control/dataset.py:309-350 (WingDataset.prepare_data): states [n][12], targets [n][3],
 *                    mean/std 12 HOST floats -> in_state [n][9], cur_out [n][12], in_ref [n][3], ref_out [n][h][3].
 * apg_sample_windows neural_control/environments/drone_env.py:232-269 (full_state_training_data): window gather
 *                    from one trajectory table [traj_rows][traj_cols >= 9]: states[i] = [traj[i*stride][0:9],0,0,0],
 *                    ref_states[i][k] = traj[i*stride + k + 1][0:9].
 * apg_poly_reference synthetic polynomial references: coef [n][3][6] (c0..c5 per axis) ->
 *                    ref_out [n][rows][9] = [p(t), 0 0 0, p'(t)], t = t_first + k*dt. */
int apg_prepare_quad(const float* states, const float* ref_states, int n, int ref_rows, float* in_state,
                     float* cur_out, float* in_ref, float* ref_out, void* stream);
int apg_prepare_wing(const float* states, const float* targets, const float* mean_host, const float* std_host,
                     float dt, int horizon, int n, float* in_state, float* cur_out, float* in_ref, float* ref_out,
                     void* stream);
int apg_sample_windows(const float* traj, int traj_rows, int traj_cols, int ref_rows, int stride, int n,
                       float* states, float* ref_states, void* stream);
int apg_poly_reference(const float* coef, int n, int rows, float t_first, float dt, float* ref_out, void* stream);
/* apg_reference_table  the table layout of load_prepare_trajectory (neural_control/trajectory/generate_trajectory.py:
 *                    566-603) + the z offset of Random.__init__ (trajectory/random_traj.py:35) from one raw
 *                    trajectory [traj_rows][traj_cols >= 10] = [pos, quaternion wxyz, vel, ...] (0.01 s steps):
 *                    table_out [table_rows][9], row k = raw row k*take_every_nth -> [pos (z + z_offset),
 *                    euler(q)*speed_factor, vel*speed_factor*2]; euler = q_funcs.quaternion_to_euler (:38-41,
 *                    pyquaternion yaw_pitch_roll restated, see csrc/prep_math.cuh). */
int apg_reference_table(const float* traj, int traj_rows, int traj_cols, int take_every_nth, float speed_factor,
                        float z_offset, int table_rows, float* table_out, void* stream);
/* apg_polynomial_points  the polynomial evaluation reference (neural_control/trajectory/polynomial.py:8-125):
 *                    Polynomial.random_polynomial's march along a fitted polynomial in steps of dist_points of arc
 *                    length, lifted to 3D ([x, 0, y] @ rot), shifted to `start` and padded with hover_steps copies of
 *                    its first / last point, for n trajectories (one thread each; the march runs in double).
 *                    coef [n][degree+1] highest power first (np.polyfit order), rot [n][9] row-major, start [n][3]
 *                    (may be NULL), all DOUBLE like the numpy arrays they come from, x from x_start to x_start + x_range.  points_out [n][max_rows][3];
 *                    ref_len_out [n] = rows of the full reference (rows beyond max_rows are not written). */
int apg_polynomial_points(const double* coef, int degree, const double* rot, const double* start, int n, double x_start,
                          double x_range, double dist_points, int hover_steps, int max_rows, float* points_out,
                          int* ref_len_out, void* stream);

/* ---- closed-loop evaluation on table references: QuadEvaluator.follow_trajectory("rand")
 * (scripts/evaluate_drone.py:81-194) with Random.get_ref_traj / project_on_ref / get_current_full_state
 * (neural_control/trajectory/random_traj.py:62-92), NetworkWrapper.predict_actions
 * (neural_control/controllers/network_wrapper.py:41-71) and QuadRotorEnvBase.step / get_is_stable
 * (neural_control/environments/drone_env.py:66-115), for cfg->n_drones independent drones in one launch, no gradient.
 * cfg: quadrotor hutter conv net, concurrent (out_dim 4h: the first predicted action is applied) or autoregressive
 * (out_dim 4); cfg->dt / cfg->phys are those of the EVALUATION dynamics.  tables [n_tables][table_rows][9] =
 * [pos, euler, vel] rows; table_index [N] (device, may be NULL: drone i walks table i); init_states [N][12].
 * Outputs (device, optional, caller zero-initialised: entries after a drone stopped are not written):
 * states_out [N][steps+1][12], div_out [N][steps], actions_out [N][steps][4], n_steps_out [N] (int).
 * workspace: apg_workspace_bytes(cfg). */
int apg_eval_rollout(const apg_config* cfg, const float* params, const float* tables, const int* table_index,
                     int n_tables, int table_rows, const float* init_states, int steps, float thresh_div,
                     float thresh_stable, int test_time, void* workspace, float* states_out, float* div_out,
                     float* actions_out, int* n_steps_out, void* stream);

/* The same closed loop with the LSTM policy (train_mode "LSTM": neural_control/models/rnn.py:8-50 LSTM_NEW; the applied
 * action is the net's 4 outputs, scripts/evaluate_drone.py:156-157).  cfg: quadrotor, APG_NET_LSTM, out_dim 4.
 * h0c0 [2][N][8] (device): hidden / cell state of every drone before its first policy call (the reference draws one
 * with torch.randn when the evaluator is constructed, evaluate_drone.py:55-57, and carries it through runs and drone
 * resets); hc_out [2][N][8] (optional): the state after the drone's last policy call. */
int apg_eval_rollout_lstm(const apg_config* cfg, const float* params, const float* h0c0, const float* tables,
                          const int* table_index, int n_tables, int table_rows, const float* init_states, int steps,
                          float thresh_div, float thresh_stable, int test_time, void* workspace, float* states_out,
                          float* div_out, float* actions_out, int* n_steps_out, float* hc_out, void* stream);

/* Fixed wing: FixedWingEvaluator.fly_to_point (scripts/evaluate_fixed_wing.py:46-130) with
 * FixedWingNetWrapper.predict_actions (controllers/network_wrapper.py:81-98), WingDataset.prepare_data,
 * SimpleWingEnv.step (environments/wing_env.py:44-58) and project_to_line (trajectory/q_funcs.py:6-18) for
 * cfg->n_drones drones in one launch.  cfg: wing, hutter linear-ref net, concurrent (out_dim 4h, the first predicted
 * action is applied); cfg->dt / cfg->phys: the evaluation environment; dt_data: the dataset's delta_t (reference line
 * spacing 12*dt_data); mean / std: 12 HOST floats of the dataset.  targets [N][n_targets][3], init_states [N][12].
 * Outputs (device, optional, caller zero-initialised): states_out [N][steps+1][12] (as returned by env.step),
 * div_linear_out [N][steps], actions_out [N][steps][4], n_steps_out [N] (int), div_target_sum_out / _cnt_out [N]
 * (sum and length of the evaluator's div_target list).  workspace: apg_workspace_bytes(cfg). */
int apg_eval_fly_to_points(const apg_config* cfg, const float* params, const float* targets, int n_targets,
                           const float* init_states, const float* mean_host, const float* std_host, float dt_data,
                           int steps, float thresh_div, float thresh_stable, int test_time, void* workspace,
                           float* states_out, float* div_linear_out, float* actions_out, int* n_steps_out,
                           float* div_target_sum_out, float* div_target_cnt_out, void* stream);

/* Cartpole: Evaluator.evaluate_in_environment (scripts/evaluate_cartpole.py:78-262) with CartpoleWrapper.predict_actions
 * (controllers/network_wrapper.py:101-113) and CartPoleEnv._step / is_upright (environments/cartpole_env.py) for
 * cfg->n_drones carts in one launch: how many steps the policy keeps |theta| < thresh_div.  cfg: cartpole, simple net,
 * concurrent (out_dim h: the first predicted action is applied); cfg->dt / cfg->phys: the evaluation environment.
 * Like the reference, the environment's cart position restarts from 0 at every step but the first (the network
 * zeroes column 0 of its input in place and the input aliases the environment state).  init_states [N][4].
 * Outputs (device, optional, caller zero-initialised): states_out [N][steps][4] (as returned by _step), actions_out
 * [N][steps], n_steps_out [N] (int; the reference's `success` = n_steps - 1), angle_sum_out / angle_cnt_out [N]
 * (sum and count of |theta| for step index > burn_in_steps), vel_sum_out [N] (sum of |x_dot| over the steps).
 * workspace: apg_workspace_bytes(cfg). */
int apg_eval_cartpole(const apg_config* cfg, const float* params, const float* init_states, int steps,
                      float thresh_div, int burn_in_steps, void* workspace, float* states_out, float* actions_out,
                      int* n_steps_out, float* angle_sum_out, float* angle_cnt_out, float* vel_sum_out, void* stream);

/* ---- learnt residual dynamics, system = APG_SYS_QUAD or APG_SYS_WING:
 *   quad  LearntDynamics.forward (neural_control/dynamics/quad_dynamics_trained.py:10-69) =
 *         simulate_quadrotor(linear_at @ action, state, dt) + linear_state_2(relu(linear_state_1([state, at])));
 *         flat parameters (named_parameters() order): linear_at 16 | mass 1 | torch_inertia_vector 3 |
 *         torch_kinv_vector 3 | linear_state_1.weight 1024 | .bias 64 | linear_state_2.weight 768 | .bias 12 = 1891;
 *         `phys`: the simulator constants of the construction-time parameters (the reference never refreshes its
 *         derived kinv / inertia matrices, :47-48).
 *   wing  LearntFixedWingDynamics.forward (neural_control/dynamics/fixed_wing_dynamics.py:270-326) =
 *         simulate_fixed_wing(state, action, dt) + the same residual MLP on [state, action]; every physical constant
 *         is a live parameter: I [3][3] | the 37 config scalars in SORTED key order (ParameterDict) | the MLP = 1914;
 *         `phys` is ignored.
 * and what autograd records for them: the vector-Jacobian products w.r.t. state, action and the flat parameter vector.
 * n rows; grad_state / grad_action / grad_params may be NULL; workspace: apg_learnt_workspace_bytes(system, n) bytes,
 * 16-byte aligned (per-block partial gradients, reduced in fixed order). */
int apg_learnt_num_params(int system);
size_t apg_learnt_workspace_bytes(int system, int n);
int apg_learnt_step(int system, const float* params, const float* phys, const float* state, const float* action,
                    float dt, int n, float* out, void* stream);
int apg_learnt_step_adjoint(int system, const float* params, const float* phys, const float* state, const float* action,
                            float dt, int n, const float* grad_out, float* grad_state, float* grad_action,
                            float* grad_params, void* workspace, void* stream);

/* The learnt quadrotor step INSIDE the fused rollout: apg_rollout_forward with every one of the h dynamics steps taken
 * by LearntDynamics.forward instead of the analytic model - the controller-training phase of run_dynamics
 * (scripts/train_base.py:334-375; scripts/train_drone.py:175-199 with self.train_dynamics = LearntDynamics, :260-278).
 * `learnt_params`: the 1891 floats above (device); cfg->phys: the construction-time constants of the learnt object.
 * Same inputs / outputs / raw-sample form as apg_rollout_forward; its adjoint is the ordinary apg_rollout_backward
 * (or _sgd / _p2p) on the same workspace: d loss / d logits through the learnt steps is produced by this call, the
 * gradient w.r.t. the LEARNT parameters is not (the reference's controller optimizer does not use it).
 * Configurations with apg_rollout_kernel_path(cfg) == 1 only; APG_ERR_UNSUPPORTED otherwise. */
int apg_rollout_forward_learnt(const apg_config* cfg, const float* params, const float* learnt_params,
                               const float* in_state, const float* cur, const float* in_ref, const float* ref,
                               void* workspace, float* loss, float* states_out, float* actions_out, void* stream);

/* ---- the path's one collective (sum of the flat weight gradient over the drone-axis shards) as this
 * library's own kernels over NVLink peer memory, instead of a library all-reduce after the adjoint pass:
 *   apg_rollout_backward_p2p   = apg_rollout_backward whose final gradient reduction stores its result straight into
 *                                slot `rank` of EVERY rank's receive set (peer-mapped symmetric memory) and raises
 *                                this rank's flag there (one kernel: reduction + exchange);
 *   apg_grad_gather_sgd_p2p    waits for all `world` flags of the local set, sums the slots in rank order (bitwise
 *                                identical on every rank) into grad_out (may be NULL) and, when `params` is given,
 *                                applies torch.optim.SGD's momentum update in the same pass
 *                                (scripts/train_base.py:139-143: buf = momentum*buf + g; p -= lr*buf).
 * Symmetric buffer of apg_grad_comm_bytes(world, n_params) bytes per rank, zero-initialised, mapped on all ranks
 * (e.g. torch.distributed._symmetric_memory); two sets used by alternate steps, set = epoch & 1, at the byte offsets
 * apg_grad_comm_offsets returns.  comm->slot_ptrs / flag_ptrs: DEVICE arrays of `world` pointers, entry q = rank q's
 * slots / flags of the current set; comm->epoch = step number 1, 2, ... (the same on every rank); comm->ticket: a
 * local device word, zero before the first call.  `local_set`: this rank's own current set.  The wait is bounded:
 * a missing peer yields a NaN gradient, not a hang. */
typedef struct apg_grad_comm {
  int rank, world;
  void* slot_ptrs;
  void* flag_ptrs;
  unsigned epoch;
  void* ticket;
} apg_grad_comm;
size_t apg_grad_comm_bytes(int world, int n_params);
int apg_grad_comm_offsets(int world, int n_params, int set, size_t* slots_offset_bytes, size_t* flags_offset_bytes);
int apg_rollout_backward_p2p(const apg_config* cfg, const float* params, const float* in_state, const float* cur,
                             const float* in_ref, const float* ref, const float* h0c0, void* workspace,
                             float grad_loss, const apg_grad_comm* comm, void* stream);
int apg_grad_gather_sgd_p2p(const apg_grad_comm* comm, const void* local_set, int n_params, float* grad_out,
                            float* params, float* momentum_buf, float lr, float momentum, void* stream);

int apg_sm_count(void);
int apg_version(void);
/* Optional per-kernel device timing of the tcgen05 path (CUDA events between the launches of the last forward +
 * backward): ms_out[7] = pack, forward chain, dynamics + reverse sweep, loss sum, dX chain, dW GEMM, gradient reduce. */
/* 1: apg_rollout_forward / backward run the tcgen05 / TMEM kernels for this configuration, 0: the mma.sync ones */
int apg_rollout_kernel_path(const apg_config* cfg);
int apg_debug_timing(int enable);
int apg_debug_kernel_times(float* ms_out);
const char* apg_error_string(int code);

#ifdef __cplusplus
}
#endif
#endif /* APG_B200_H */
