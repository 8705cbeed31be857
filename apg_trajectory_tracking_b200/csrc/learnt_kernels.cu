// Learnt residual dynamics: one step for N rows and its adjoint w.r.t. state, action and all
// parameters, for the quadrotor (reference quad_dynamics_trained.py:10-69, 1891 parameters) and the fixed wing
// (fixed_wing_dynamics.py:270-326, 1914 parameters).  One kernel template, two per-drone models (learnt_math.cuh,
// learnt_wing_math.cuh).
//
//   forward  one thread per drone, parameters in shared memory, hidden activations in a shared-memory column of
//            the thread (stride LP -> conflict-free), 1.8 kMAC per drone.
//   adjoint  persistent blocks of 128 threads walk tiles of 128 drones: (1) each thread recomputes the forward of its
//            drone and back-propagates it, leaving the per-drone factors (g, h, dh, x, the cotangents of the physical
//            parameters, 1) in shared-memory rows of LP = 129 floats; (2) the parameter-gradient entries are spread
//            over the threads, each entry a dot product of two factor rows over the tile's drones (rows padded to 129
//            so that both the per-drone column accesses of phase 1 and the per-entry row walks of phase 2 are
//            bank-conflict free), accumulated in registers over all tiles of the block; (3) one partial vector per
//            block, summed in fixed order by apg_reduce_kernel -> bitwise reproducible.
#include "learnt_math.cuh"
#include "learnt_wing_math.cuh"
#include "kernels.h"
#ifdef APG_SIM
#define APG_LEARNT_DYNAMIC_SMEM(name) float* name = reinterpret_cast<float*>(::simte::dynamic_smem())
#else
#define APG_LEARNT_DYNAMIC_SMEM(name) extern __shared__ __align__(16) float name[]
#endif

namespace apg {

namespace {
constexpr int LT = 128;            // threads per block = drones per tile
constexpr int LP = LT + 1;         // padded factor-row length
}  // namespace

template <class M>
__global__ void __launch_bounds__(LT) learnt_fwd_kernel(const float* __restrict__ params, const PhysConsts pc,
                                                        const float* __restrict__ s, const float* __restrict__ a,
                                                        float dt, int n, float* __restrict__ out) {
  using RW = LearntRows<M::NPH>;
  APG_LEARNT_DYNAMIC_SMEM(sm);
  float* sP = sm;                        // [NP]
  float* sH = sm + RW::NP + 1;           // [64][LP]
  for (int i = threadIdx.x; i < RW::NP; i += LT) sP[i] = params[i];
  __syncthreads();
  for (int base = blockIdx.x * LT; base < n; base += gridDim.x * LT) {
    const int d = base + threadIdx.x;
    if (d < n) {
      float si[12], ai[4], x[16], o[12];
#pragma unroll
      for (int j = 0; j < 12; ++j) si[j] = s[(size_t)d * 12 + j];
#pragma unroll
      for (int j = 0; j < 4; ++j) ai[j] = a[(size_t)d * 4 + j];
      M::fwd(sP, pc.v, si, ai, dt, o, x, sH + threadIdx.x, LP);
#pragma unroll
      for (int j = 0; j < 12; ++j) out[(size_t)d * 12 + j] = o[j];
    }
  }
}

template <class M>
__global__ void __launch_bounds__(LT) learnt_adj_kernel(const float* __restrict__ params, const PhysConsts pc,
                                                        const float* __restrict__ s, const float* __restrict__ a,
                                                        float dt, int n, const float* __restrict__ g,
                                                        float* __restrict__ gs, float* __restrict__ ga,
                                                        float* __restrict__ partials) {
  using RW = LearntRows<M::NPH>;
  constexpr int EPT = (RW::NP + LT - 1) / LT;      // entries per thread: 15
  APG_LEARNT_DYNAMIC_SMEM(sm);
  float* sP = sm;                        // [NP]
  float* sF = sm + RW::NP + 1;           // [R_TOTAL][LP] factor rows
  const int t = threadIdx.x;
  for (int i = t; i < RW::NP; i += LT) sP[i] = params[i];
  float acc[EPT];
#pragma unroll
  for (int q = 0; q < EPT; ++q) acc[q] = 0.f;
  __syncthreads();
  for (int base = blockIdx.x * LT; base < n; base += gridDim.x * LT) {
    const int d = base + t;
    const int valid = min(LT, n - base);
    // ---- phase 1: per-drone forward + adjoint, factors into the shared rows (column t)
    if (d < n) {
      float si[12], ai[4], x[16], o[12], gi[12], gso[12], gao[4], dph[M::NPH];
#pragma unroll
      for (int j = 0; j < 12; ++j) { si[j] = s[(size_t)d * 12 + j]; gi[j] = g[(size_t)d * 12 + j]; }
#pragma unroll
      for (int j = 0; j < 4; ++j) ai[j] = a[(size_t)d * 4 + j];
      M::fwd(sP, pc.v, si, ai, dt, o, x, sF + RW::R_H * LP + t, LP);
      M::adj(sP, pc.v, si, ai, x, sF + RW::R_H * LP + t, LP, dt, gi, gso, gao, sF + RW::R_DH * LP + t, dph);
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        if (gs) gs[(size_t)d * 12 + j] = gso[j];
        sF[(RW::R_G + j) * LP + t] = gi[j];
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) sF[(RW::R_X + j) * LP + t] = x[j];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (ga) ga[(size_t)d * 4 + j] = gao[j];
      }
#pragma unroll
      for (int j = 0; j < M::NPH; ++j) sF[(RW::R_DP + j) * LP + t] = dph[j];
      sF[RW::R_ONE * LP + t] = 1.f;
    }
    __syncthreads();
    // ---- phase 2: parameter-gradient entries t, t + 128, ... += dot of two factor rows over the tile's drones
#pragma unroll
    for (int q = 0; q < EPT; ++q) {
      const int e = t + q * LT;
      if (e < RW::NP) {
        int ra, rb;
        RW::entry_rows(e, &ra, &rb);
        const float* pa = sF + ra * LP;
        const float* pb = sF + rb * LP;
        float v = 0.f;
        for (int k = 0; k < valid; ++k) v = fmaf(pa[k], pb[k], v);
        acc[q] += v;
      }
    }
    __syncthreads();
  }
  if (partials) {
#pragma unroll
    for (int q = 0; q < EPT; ++q) {
      const int e = t + q * LT;
      if (e < RW::NP) partials[(size_t)blockIdx.x * RW::NP + e] = acc[q];
    }
  }
}

int learnt_num_params(int system) {
  return system == SYS_WING ? LearntRows<LearntWing<float>::NPH>::NP : LearntRows<LearntQuad<float>::NPH>::NP;
}
int learnt_grid(int n, int sms) {
  const int tiles = (n + LT - 1) / LT;
  const int cap = (sms > 0 ? sms : 148) * 2;
  return tiles < cap ? (tiles > 0 ? tiles : 1) : cap;
}
size_t learnt_partials_floats(int system, int n, int sms) { return (size_t)learnt_grid(n, sms) * learnt_num_params(system); }

template <class M>
static cudaError_t launch_fwd_t(const float* params, const PhysConsts& pc, const float* s, const float* a, float dt,
                                int n, float* out, int sms, cudaStream_t st) {
  using RW = LearntRows<M::NPH>;
  const size_t smem = sizeof(float) * (RW::NP + 1 + (size_t)RW::HD * LP);
  cudaError_t e = cudaFuncSetAttribute(learnt_fwd_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  APG_LAUNCH(learnt_grid(n, sms), LT, smem, st, learnt_fwd_kernel<M>)(params, pc, s, a, dt, n, out);
  return cudaGetLastError();
}

template <class M>
static cudaError_t launch_adj_t(const float* params, const PhysConsts& pc, const float* s, const float* a, float dt,
                                int n, const float* g, float* gs, float* ga, float* grad_params, float* partials,
                                int sms, cudaStream_t st) {
  using RW = LearntRows<M::NPH>;
  const size_t smem = sizeof(float) * (RW::NP + 1 + (size_t)RW::R_TOTAL * LP);
  cudaError_t e = cudaFuncSetAttribute(learnt_adj_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int grid = learnt_grid(n, sms);
  APG_LAUNCH(grid, LT, smem, st, learnt_adj_kernel<M>)(params, pc, s, a, dt, n, g, gs, ga, grad_params ? partials : nullptr);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (grad_params) return launch_reduce_grad(partials, grid, RW::NP, 1.0f, grad_params, st);
  return cudaSuccess;
}

cudaError_t launch_learnt_fwd(int system, const float* params, const PhysConsts& pc, const float* s, const float* a,
                              float dt, int n, float* out, int sms, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if (system == SYS_QUAD) return launch_fwd_t<LearntQuad<float>>(params, pc, s, a, dt, n, out, sms, st);
  if (system == SYS_WING) return launch_fwd_t<LearntWing<float>>(params, pc, s, a, dt, n, out, sms, st);
  return cudaErrorInvalidValue;
}

cudaError_t launch_learnt_adj(int system, const float* params, const PhysConsts& pc, const float* s, const float* a,
                              float dt, int n, const float* g, float* gs, float* ga, float* grad_params,
                              float* partials, int sms, cudaStream_t st) {
  if (n <= 0)                                   // empty batch: the parameter gradient is zero, not "untouched"
    return grad_params ? cudaMemsetAsync(grad_params, 0, sizeof(float) * learnt_num_params(system), st) : cudaSuccess;
  if (system == SYS_QUAD)
    return launch_adj_t<LearntQuad<float>>(params, pc, s, a, dt, n, g, gs, ga, grad_params, partials, sms, st);
  if (system == SYS_WING)
    return launch_adj_t<LearntWing<float>>(params, pc, s, a, dt, n, g, gs, ga, grad_params, partials, sms, st);
  return cudaErrorInvalidValue;
}


}  // namespace apg
