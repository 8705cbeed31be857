// Host-side launcher declarations (definitions in the *_kernels.cu files).
#pragma once
#ifndef APG_SIM
#include <cuda_runtime.h>
#endif
#include "layouts.h"
#include "rollout_args.h"
#include "eval_math.cuh"
#include "p2p_math.cuh"

namespace apg {

size_t hutter_fwd_smem_bytes(const HutterLayout& y);
size_t hutter_adj_smem_bytes(const HutterLayout& y);
cudaError_t launch_hutter_fwd(int system, const HutterLayout& y, const RolloutArgs& a, int grid, cudaStream_t st);
cudaError_t launch_hutter_adj(int system, const HutterLayout& y, const RolloutArgs& a, int grid, cudaStream_t st);

// tcgen05 / TMEM path of the quadrotor concurrent rollout (tq_kernels.cu, tq_dw_kernels.cu, tq_layout.cuh): forward,
// dynamics + reverse sweep, dX chain and streaming weight-gradient GEMM on operand-image stashes
bool tq_supported(const HutterLayout& y, int h);
size_t tq_blob_bytes();
size_t tq_tblob_bytes();
size_t tq_fstash_bytes(int n);
size_t tq_zstash_bytes(int n);
int tq_grid(int n, int sms);
int tq_dyn_grid(int n, int sms);
cudaError_t launch_tq_pack(const HutterLayout& y, const float* params, unsigned char* blob, unsigned char* tblob,
                           cudaStream_t st);
cudaError_t launch_tq_fwd(const unsigned char* blob, const RolloutArgs& a, unsigned char* fstash, int grid,
                          cudaStream_t st);
cudaError_t launch_tq_dyn(const RolloutArgs& a, unsigned char* fstash, unsigned char* zstash, float* loss_out,
                          unsigned* ticket, unsigned ticket0, int dyn_grid, cudaStream_t st);
cudaError_t launch_tq_dx(const unsigned char* tblob, const RolloutArgs& a, unsigned char* fstash,
                         unsigned char* zstash, const unsigned char* stamp, int want_stamp, int grid, cudaStream_t st);
cudaError_t launch_tq_dw(const HutterLayout& y, const RolloutArgs& a, const unsigned char* fstash,
                         const unsigned char* zstash, int grid, cudaStream_t st);
cudaError_t launch_reduce_grad4(const float* partials, int ncta, int n, float scale, float* grad, cudaStream_t st);
cudaError_t launch_reduce_grad4_sgd(const float* partials, int ncta, int n, float scale, float* grad, float* param,
                                    float* buf, float lr, float momentum, cudaStream_t st);

cudaError_t launch_rec_fwd(const HutterLayout& y, const RolloutArgs& a, int grid, cudaStream_t st);
cudaError_t launch_rec_adj(const HutterLayout& y, const RolloutArgs& a, int grid, cudaStream_t st);
cudaError_t launch_lstm_fwd(const LstmLayout& y, const RolloutArgs& a, int grid, cudaStream_t st);
cudaError_t launch_lstm_adj(const LstmLayout& y, const RolloutArgs& a, int grid, cudaStream_t st);
size_t simple_fwd_smem_bytes(const SimpleLayout& y);
size_t simple_adj_smem_bytes(const SimpleLayout& y);
cudaError_t launch_simple_fwd(const SimpleLayout& y, const RolloutArgs& a, int grid, cudaStream_t st);
cudaError_t launch_simple_adj(const SimpleLayout& y, const RolloutArgs& a, int grid, cudaStream_t st);

cudaError_t launch_pack(const PackTable& t, const float* params, float* wf, float* wb, cudaStream_t st);
cudaError_t launch_reduce_grad(const float* partials, int ncta, int n, float scale, float* grad, cudaStream_t st,
                               int pm_off = 0, int pm_k1 = 0, int pm_npos = 0);
// gradient reduction + exchange over NVLink peer memory (p2p_kernels.cu; layout and protocol: p2p_math.cuh)
cudaError_t launch_reduce_scatter_p2p(const float* partials, int ncta, int n, float scale, int pm_off, int pm_k1,
                                      int pm_npos, float* const* slots, unsigned* const* flags, int rank, int world,
                                      unsigned epoch, unsigned* ticket, cudaStream_t st);
cudaError_t launch_gather_sgd_p2p(const float* slots_local, const unsigned* flags_local, int world, int n,
                                  unsigned epoch, float* grad_out, float* param, float* momentum_buf, float lr,
                                  float momentum, cudaStream_t st);
cudaError_t launch_sum_loss(const float* partials, int ncta, float* loss, cudaStream_t st);
cudaError_t launch_step(int system, const PhysConsts& pc, const float* s, const float* a, float dt, int n, float* out,
                        cudaStream_t st);
cudaError_t launch_step_adj(int system, const PhysConsts& pc, const float* s, const float* a, float dt, int n,
                            const float* g, float* gs, float* ga, cudaStream_t st);
cudaError_t launch_features(const float* s, int n, float* f, cudaStream_t st);
cudaError_t launch_features_adj(const float* s, const float* gf, int n, float* gs, cudaStream_t st);

// closed-loop evaluation on table references (eval_kernels.cu)
cudaError_t launch_eval_rollout(const HutterLayout& y, const float* wf, const float* tables, const int* table_index,
                                const float* init_states, int n, float dt, const PhysConsts& pc, const EvalParams& ev,
                                float* states_out, float* div_out, float* actions_out, int* n_steps_out, int grid,
                                cudaStream_t st);

cudaError_t launch_eval_rollout_lstm(const LstmLayout& y, const float* wf, const float* h0c0, const float* tables,
                                     const int* table_index, const float* init_states, int n, float dt,
                                     const PhysConsts& pc, const EvalParams& ev, float* states_out, float* div_out,
                                     float* actions_out, int* n_steps_out, float* hc_out, int grid, cudaStream_t st);
cudaError_t launch_eval_wing(const HutterLayout& y, const float* wf, const float* targets, const float* init_states,
                             int n, float dt_env, const PhysConsts& pc, const float* mean_host, const float* std_host,
                             const WingEvalParams& ev, float* states_out, float* div_out, float* actions_out,
                             int* n_steps_out, float* dt_sum_out, float* dt_cnt_out, int grid, cudaStream_t st);

cudaError_t launch_eval_cartpole(const SimpleLayout& y, const float* wf, const float* init_states, int n, float dt,
                                 const PhysConsts& pc, const CartpoleEvalParams& ev, float* states_out,
                                 float* actions_out, int* n_steps_out, float* angle_sum_out, float* angle_cnt_out,
                                 float* vel_sum_out, int grid, cudaStream_t st);

// learnt residual dynamics, quadrotor / fixed wing (learnt_kernels.cu)
int learnt_num_params(int system);
int learnt_grid(int n, int sms);
size_t learnt_partials_floats(int system, int n, int sms);
cudaError_t launch_learnt_fwd(int system, const float* params, const PhysConsts& pc, const float* s, const float* a,
                              float dt, int n, float* out, int sms, cudaStream_t st);
cudaError_t launch_learnt_adj(int system, const float* params, const PhysConsts& pc, const float* s, const float* a,
                              float dt, int n, const float* g, float* gs, float* ga, float* grad_params,
                              float* partials, int sms, cudaStream_t st);

// data formats on the input side (prep_kernels.cu)
cudaError_t launch_prepare_quad(const float* states, const float* ref, int n, int L, float* in_state, float* cur_out,
                                float* in_ref, float* ref_out, cudaStream_t st);
cudaError_t launch_prepare_wing(const float* states, const float* targets, const float* mean_host,
                                const float* std_host, float dt, int h, int n, float* in_state, float* cur_out,
                                float* in_ref, float* ref_out, cudaStream_t st);
cudaError_t launch_poly_reference(const float* coef, int n, int L, float t_first, float dt, float* out,
                                  cudaStream_t st);
cudaError_t launch_polynomial_points(const double* coef, int degree, const double* rot, const double* start, int n,
                                     double x_start, double x_range, double dist_points, int hover, int max_rows,
                                     float* out, int* ref_len, cudaStream_t st);
cudaError_t launch_reference_table(const float* traj, int W, int nth, float speed, float z_offset, int rows,
                                   float* out, cudaStream_t st);
cudaError_t launch_sample_windows(const float* traj, int W, int L, int stride, int n, float* states, float* refs,
                                  cudaStream_t st);

}  // namespace apg
