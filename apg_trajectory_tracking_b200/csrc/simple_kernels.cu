// Fused rollout kernels for the cartpole policy (models/simple_model.py Net: 4->32->64->64->32->h, tanh on every
// layer including the output, input column 0 zeroed) in concurrent mode: scripts/train_cartpole.py:118-155.
// Same structure as hutter_kernels.cu; the whole activation arena of a tile (192+Mo4 rows) is stashed / restored
// with a single bulk copy.
#include "dyn_phase.cuh"
#include "layouts.h"
#include "rollout_args.h"
#include "tile_engine.cuh"

namespace apg {

__device__ __forceinline__ void load_state_tile(float* dst, const float* __restrict__ src, int F0, int valid) {
  // drone-major tile with the first feature zeroed (simple_model.py:21: x[:, 0] *= 0)
  for (int i = threadIdx.x; i < TM * F0; i += NT) {
    const int d = i / F0, c = i - d * F0;
    dst[i] = (d < valid && c != 0) ? src[i] : 0.f;
  }
}

__global__ void __launch_bounds__(NT, 1) simple_fwd_kernel(const SimpleLayout y, const RolloutArgs g) {
  APG_DYNAMIC_SMEM_F32(smem);
  using Sys = Cartpole<float>;
  constexpr int S = Sys::S, A = Sys::A;
  float* s_w = smem;
  float* s_in = s_w + y.f_total;
  float* s_a = s_in + pad4(TM * y.F0);               // activation arena [rows_total][TMP]
  float* s_red = s_a + y.rows_total * TMP;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(s_red + 8);
  const Lane L;
  const int tid = threadIdx.x;
  const int ntiles = (g.N + TM - 1) / TM;
  if (tid == 0) {
    mbar_init(bar_w, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar_w, y.f_total * 4);
    bulk_g2s_chunked(s_w, g.wf, y.f_total * 4, bar_w);
  }
  mbar_wait(bar_w, 0);
  float cta_loss = 0.f;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int valid = min(TM, g.N - tile * TM);
    load_state_tile(s_in, g.in_state + (size_t)tile * TM * y.F0, y.F0, valid);
    __syncthreads();
    dense<SrcAoS, EPI_ACT>(L, SrcAoS{s_in, y.F0, 0}, y.din[0], s_w + y.f_w[0], y.ldf[0], s_w + y.f_b[0], y.ldf[0] / 4,
                           s_a, y.row[0], 1, ACT_TANH);
    __syncthreads();
    for (int l = 1; l < SIMPLE_NL; ++l) {
      dense<SrcT, EPI_ACT>(L, SrcT{s_a + y.row[l - 1] * TMP}, y.din[l], s_w + y.f_w[l], y.ldf[l], s_w + y.f_b[l],
                           y.ldf[l] / 4, s_a, y.row[l], 1, ACT_TANH);
      if (l == SIMPLE_NL - 1) fence_proxy_async();
      __syncthreads();
    }
    if (tid == 0) {
      bulk_s2g(g.st_x1 + (size_t)tile * y.rows_total * TMP, s_a, y.rows_total * TMP * 4);
      bulk_commit();
    }
    float my_loss = 0.f;
    if (tid < valid) {
      const size_t drone = (size_t)tile * TM + tid;
      my_loss = dyn_forward_conc<Cartpole>(s_a + y.row[SIMPLE_NL - 1] * TMP, tid, g.cur + drone * S, nullptr, g.h, g.dt,
                                           g.pc.v, g.st_states + (size_t)tile * g.h * S * TMP,
                                           g.states_out ? g.states_out + drone * g.h * S : nullptr,
                                           g.actions_out ? g.actions_out + drone * g.h * A : nullptr);
    }
    const float tl = block_sum(my_loss, s_red);
    if (tid == 0) {
      cta_loss += tl;
      bulk_wait_read<0>();
    }
    __syncthreads();
  }
  if (tid == 0) {
    g.loss_partials[blockIdx.x] = cta_loss;
    bulk_wait_all();
  }
}

__global__ void __launch_bounds__(NT, 1) simple_adj_kernel(const SimpleLayout y, const RolloutArgs g) {
  APG_DYNAMIC_SMEM_F32(smem);
  using Sys = Cartpole<float>;
  constexpr int S = Sys::S;
  float* s_w = smem;
  float* s_in = s_w + y.b_total;
  float* s_a = s_in + pad4(TM * y.F0);
  float* s_dlog = s_a + y.rows_total * TMP;          // [Mo4][TMP]
  float* s_red = s_dlog + y.Mo4 * TMP;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(s_red + 8);
  uint64_t* bar_a = bar_w + 1;
  const Lane L;
  const int tid = threadIdx.x;
  const int ntiles = (g.N + TM - 1) / TM;
  float* P = g.grad_partials + (size_t)blockIdx.x * y.n_params;
  for (int i = tid; i < y.n_params; i += NT) __stcg(P + i, 0.f);
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_a, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar_w, y.b_total * 4);
    bulk_g2s_chunked(s_w, g.wb, y.b_total * 4, bar_w);
  }
  mbar_wait(bar_w, 0);
  uint32_t ph = 0;
  const int LAST = SIMPLE_NL - 1;
  for (int tile = ntiles - 1 - (int)blockIdx.x; tile >= 0; tile -= gridDim.x) {
    const int valid = min(TM, g.N - tile * TM);
    if (tid == 0) {
      mbar_expect_tx(bar_a, y.rows_total * TMP * 4);
      bulk_g2s_chunked(s_a, g.st_x1 + (size_t)tile * y.rows_total * TMP, y.rows_total * TMP * 4, bar_a);
    }
    load_state_tile(s_in, g.in_state + (size_t)tile * TM * y.F0, y.F0, valid);
    if (tid < TM) {
      for (int r = 0; r < y.Mo4; ++r) s_dlog[r * TMP + tid] = 0.f;
      if (tid < valid) {
        const size_t drone = (size_t)tile * TM + tid;
        // actions come from the global stash (rows of the last layer); d loss/d pre-activation of the tanh output
        dyn_adjoint_conc<Cartpole>(g.st_x1 + ((size_t)tile * y.rows_total + y.row[LAST]) * TMP,
                                   g.st_states + (size_t)tile * g.h * S * TMP, tid, g.cur + drone * S, nullptr, g.h,
                                   g.dt, g.pc.v, s_dlog);
      }
    }
    mbar_wait(bar_a, ph);
    ph ^= 1;
    __syncthreads();
    // layers LAST .. 1: dW from (dz, input activation), then dz of the previous layer in place over its activation
    const float* dz = s_dlog;
    for (int l = LAST; l >= 1; --l) {
      float* xin = s_a + y.row[l - 1] * TMP;
      {
        const int K = y.din[l];
        if (K <= 32) dw_T<1>(L, dz, y.dout[l], xin, K, P + y.t_w[l], K, P + y.t_b[l]);
        else         dw_T<2>(L, dz, y.dout[l], xin, K, P + y.t_w[l], K, P + y.t_b[l]);
      }
      __syncthreads();
      dense<SrcT, EPI_DTANH>(L, SrcT{dz}, y.dout[l], s_w + y.b_w[l], y.ldb[l], nullptr, y.ldb[l] / 4, xin, 0, 1, 0);
      __syncthreads();
      dz = xin;
    }
    dw_AoS(L, dz, y.dout[0], s_in, y.F0, 0, y.F0, P + y.t_w[0], y.F0, P + y.t_b[0]);
    fence_proxy_async();
    __syncthreads();
  }
}

size_t simple_fwd_smem_bytes(const SimpleLayout& y) {
  return sizeof(float) * (size_t)(y.f_total + pad4(TM * y.F0) + y.rows_total * TMP + 8) + 32;
}
size_t simple_adj_smem_bytes(const SimpleLayout& y) {
  return sizeof(float) * (size_t)(y.b_total + pad4(TM * y.F0) + (y.rows_total + y.Mo4) * TMP + 8) + 32;
}

cudaError_t launch_simple_fwd(const SimpleLayout& y, const RolloutArgs& a, int grid, cudaStream_t st) {
  const size_t smem = simple_fwd_smem_bytes(y);
  cudaError_t e = cudaFuncSetAttribute(simple_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  APG_LAUNCH(grid, NT, smem, st, simple_fwd_kernel)(y, a);
  return cudaGetLastError();
}

cudaError_t launch_simple_adj(const SimpleLayout& y, const RolloutArgs& a, int grid, cudaStream_t st) {
  const size_t smem = simple_adj_smem_bytes(y);
  cudaError_t e = cudaFuncSetAttribute(simple_adj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  APG_LAUNCH(grid, NT, smem, st, simple_adj_kernel)(y, a);
  return cudaGetLastError();
}

}  // namespace apg
