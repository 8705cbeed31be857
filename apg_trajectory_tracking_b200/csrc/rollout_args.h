// Kernel argument block shared by the rollout kernels (passed by value).
#pragma once
#include "apg_math.cuh"

namespace apg {

struct RolloutArgs {
  // per-drone inputs (device, fp32, row-major contiguous, 16-byte aligned)
  const float* in_state;   // [N][F0]           policy state features (concurrent mode)
  const float* cur;        // [N][S]            start state
  const float* in_ref;     // [N][L*RD]         policy reference input (concurrent) / [N][2h][RD] (recurrent)
  const float* ref;        // [N][h][REFW]      loss reference (quad: 9 wide, wing: 3 wide, cartpole: none)
  const float* h0c0;       // [2][N][8]         LSTM initial hidden / cell state
  int N, h;
  int ref_rows;            // rows per drone in `ref` (h in concurrent mode, 2h in recurrent mode)
  int window;              // Window enum (recurrent modes)
  int raw_inputs;          // tcgen05 path: `cur` / `ref` are RAW samples (absolute positions), in_state / in_ref are
                           // derived in the kernels' prologue (QuadDataset.prepare_data, dataset.py:155-204)
  float dt;
  PhysConsts pc;
  // packed weights
  const float* wf;
  const float* wb;
  // activation stash, tile-major: [tile][rows][TMP]
  float* st_x1;
  float* st_h1;
  float* st_h2;
  float* st_h3;
  float* st_act;
  float* st_states;
  float* st_misc;          // recurrent modes: positions / lstm state per step
  // outputs
  float* loss_partials;    // [gridDim.x]
  float* grad_partials;    // [gridDim.x][n_params]   (adjoint kernel)
  float* states_out;       // optional [N][h][S]
  float* actions_out;      // optional [N][h][A]
  // learnt residual dynamics (quad_dynamics_trained.py:10-69) in place of the analytic step: flat parameter vector of
  // LearntDynamics (learnt_math.cuh layout, 1891 floats) or NULL.  tcgen05 path only (tq_dyn_kernel<true>).
  const float* learnt;
};

}  // namespace apg
