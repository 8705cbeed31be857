// Helpers shared by the recurrent rollout kernels (autoregressive MLP and LSTM): the per-step reference window,
// the input gradients through the conv encoder and the fold of feature / window cotangents into the state cotangent.
#pragma once
#include "apg_math.cuh"
#include "layouts.h"
#include "tile_engine.cuh"

namespace apg {

// all threads: s_win[d][r*RD + c] = window row r of drone d at step k (drone-major, like the concurrent in_ref tile)
__device__ __forceinline__ void build_window(float* __restrict__ s_win, const float* __restrict__ in_ref0_tile,
                                             const float* __restrict__ s_P, const float* __restrict__ s_pos, int k,
                                             int h, int RD, int valid, int window) {
  const int LR = h * RD;
  for (int idx = threadIdx.x; idx < TM * LR; idx += NT) {
    const int d = idx / LR, e = idx - d * LR;
    const int r = e / RD, c = e - r * RD;
    float v = 0.f;
    if (d < valid) {
      v = in_ref0_tile[(size_t)d * 2 * LR + (k + r) * RD + c];
      if (c < 3) {
        float sub;
        if (window == WINDOW_RELATIVE) {
          sub = s_pos[c * TMP + d];
        } else {
          sub = s_P[(k * 3 + c) * TMP + d];
          const int m = k + r - h;
          if (m >= 0) sub -= s_P[(m * 3 + c) * TMP + d];
        }
        v -= sub;
      }
    }
    s_win[idx] = v;
  }
}

// d loss / d window[r][c], c < 3 (position columns), through the conv encoder:
//   dwin[r][c] = sum_{j<3, t=r-j in [0,npos)} sum_ch dconv[ch*npos+t] * Wc[ch][c][j]
// dzr = rows [64, 64+20*npos) of dX1 (already multiplied by relu'), wb = packed [20][ld_bwr] (kk = j*RD + c)
__device__ __forceinline__ void conv_dx_pos(const HutterLayout& y, const float* __restrict__ dzr,
                                            const float* __restrict__ wb, float* __restrict__ dwin) {
  for (int idx = threadIdx.x; idx < TM * y.L; idx += NT) {
    const int r = idx / TM, d = idx - r * TM;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int t = r - j;
      if (t >= 0 && t < y.npos) {
        for (int ch = 0; ch < CONV_CH; ++ch) {
          const float z = dzr[(ch * y.conv_cs + t * y.conv_ts) * TMP + d];
          const float* w = wb + ch * y.ld_bwr + j * y.RD;
          a0 = fmaf(z, w[0], a0); a1 = fmaf(z, w[1], a1); a2 = fmaf(z, w[2], a2);
        }
      }
    }
    dwin[(r * 3 + 0) * TMP + d] = a0;
    dwin[(r * 3 + 1) * TMP + d] = a1;
    dwin[(r * 3 + 2) * TMP + d] = a2;
  }
}

// thread d (< TM): fold d loss/d features and d loss/d window into the state cotangent g (12) of step k
__device__ __forceinline__ void state_input_adjoint(const float* sk, const float* __restrict__ din,
                                                    const float* __restrict__ dwin, float* __restrict__ s_dP,
                                                    float* srun, float* g, int k, int h, int window, int d) {
  float gf[15];
#pragma unroll
  for (int i = 0; i < 15; ++i) gf[i] = din[i * TMP + d];
  Quad<float>::features_adj(sk, gf, g);
  float G[3] = {0.f, 0.f, 0.f};
  for (int r = 0; r < h; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = dwin[(r * 3 + c) * TMP + d];
      G[c] += v;
      if (window == WINDOW_CUMULATIVE) {
        const int m = k + r - h;
        if (m >= 0) s_dP[(m * 3 + c) * TMP + d] += v;          // + P_{k+r-h}
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (window == WINDOW_CUMULATIVE) {
      srun[c] += s_dP[(k * 3 + c) * TMP + d] - G[c];            // d/dP_k complete: every later step has contributed
      g[c] += srun[c];                                           // d/dpos_k = sum_{m>=k} d/dP_m
    } else {
      g[c] -= G[c];
    }
  }
}

}  // namespace apg
