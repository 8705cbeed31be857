// Device functions shared by the concurrent (hutter_kernels.cu) and recurrent (rec_kernels.cu) rollout kernels:
// the layers of the hutter policy (models/hutter_model.py:32-49) on a 64-drone tile.
#pragma once
#include "layouts.h"
#include "tile_engine.cuh"

namespace apg {

__device__ __forceinline__ void load_tile_manual(float* dst, const float* __restrict__ src, int row_floats, int valid) {
  const int nv = valid * row_floats;
  for (int i = threadIdx.x; i < TM * row_floats; i += NT) dst[i] = i < nv ? src[i] : 0.f;
}


// First layer: s = tanh(states_in(state)) -> X1 rows [0,64);  reference branch -> X1 rows [64, K1):
//   CONV : Conv1d(RD -> 20, k=3, valid) over the L reference rows == npos dense layers on shifted 3*RD windows,
//          relu, output index channel-major c*npos + t (hutter_model.py:36-40)
//   !CONV: tanh(ref_in(ref))
// s_ins / s_inr are the drone-major input tiles [TM][F0] / [TM][L*RD].
// Conv1d(RD -> 20, k=3, valid) + relu on the drone-major reference tile s_inr [TM][LR]; output rows row0 + c*npos + t
__device__ __forceinline__ void conv_layer_fwd(const Lane& L, int npos, int RD, int LR, int KC, const float* Wc,
                                               const float* bc, const float* s_inr, float* Y, int row0) {
  const int ncg = CONV_CH / 4;
  for (int vg = L.og0; vg < npos * ncg; vg += 16) {
    const int t = vg / ncg, cg = vg - t * ncg;
    float acc[4][4] = {};
    mac_tile(acc, SrcAoS{s_inr, LR, RD * t}, KC, Wc + 4 * cg, CONV_CH, L.dg);
    store_tile<EPI_ACT>(acc, bc, cg, Y, row0 + t, npos, ACT_RELU, L.dg);
  }
}

template <bool CONV>
__device__ __forceinline__ void hutter_first_layer(const Lane& L, const HutterLayout& y, const float* s_w,
                                                   const float* s_ins, const float* s_inr, float* s_x1) {
  dense<SrcAoS, EPI_ACT>(L, SrcAoS{s_ins, y.F0, 0}, y.F0, s_w + y.f_ws, HID, s_w + y.f_bs, HID / 4, s_x1, 0, 1,
                         ACT_TANH);
  if (CONV) {
    conv_layer_fwd(L, y.npos, y.RD, y.LR, y.KC, s_w + y.f_wr, s_w + y.f_br, s_inr, s_x1, HID);
  } else {
    dense<SrcAoS, EPI_ACT>(L, SrcAoS{s_inr, y.LR, 0}, y.LR, s_w + y.f_wr, HID, s_w + y.f_br, HID / 4, s_x1, HID, 1,
                           ACT_TANH);
  }
}

// fc1 .. fc_out on a tile whose X1 is complete in s_x1 AND whose X1 stash store has been committed as the most
// recent bulk group by thread 0.  Leaves sigmoid(logits) in s_x1 rows [64, 64+Mo4) and every hidden activation in
// the global stash.  Ends with a __syncthreads (actions visible to all threads).
__device__ __forceinline__ void hutter_trunk(const Lane& L, const HutterLayout& y, const float* s_w, float* s_x1,
                                             float* s_h, float* st_h1, float* st_h2, float* st_h3, float* st_act) {
  const int tid = threadIdx.x;
  dense_auto<EPI_ACT>(L, s_x1, y.K1, s_w + y.f_w1, HID, mma_sw(HID), s_w + y.f_b1, HID, s_h, 0, ACT_TANH);
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    bulk_s2g(st_h1, s_h, HID * TMP * 4);
    bulk_commit();
    bulk_wait_read<1>();      // the X1 store has finished reading s_x1
  }
  __syncthreads();
  // fc2 : s_h -> s_x1 rows [0,64)
  dense_auto<EPI_ACT>(L, s_h, HID, s_w + y.f_w2, HID, mma_sw(HID), s_w + y.f_b2, HID, s_x1, 0, ACT_TANH);
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    bulk_s2g(st_h2, s_x1, HID * TMP * 4);
    bulk_commit();
    bulk_wait_read<1>();      // the h1 store has finished reading s_h
  }
  __syncthreads();
  // fc3 : s_x1 rows [0,64) -> s_h
  dense_auto<EPI_ACT>(L, s_x1, HID, s_w + y.f_w3, HID, mma_sw(HID), s_w + y.f_b3, HID, s_h, 0, ACT_TANH);
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    bulk_s2g(st_h3, s_h, HID * TMP * 4);
    bulk_commit();
  }
  // fc_out + sigmoid : s_h -> s_x1 rows [64, 64+Mo4)
  float* s_act = s_x1 + HID * TMP;
  dense_auto<EPI_ACT>(L, s_h, HID, s_w + y.f_wo, y.ld_fwo, mma_sw(y.ld_fwo), s_w + y.f_bo, y.Mo4, s_act, 0,
                      ACT_SIGMOID);
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    bulk_s2g(st_act, s_act, y.Mo4 * TMP * 4);
    bulk_commit();
  }
}

// conv_ref weight gradient: dWc[c][d][j] = sum_t sum_drone dz[c*npos+t][drone] * in_ref[drone][(t+j)*RD + d].
// One warp per position t, lanes over the 3*RD window entries, 20 channel accumulators per lane; the 8 per-warp
// partials are combined in a fixed order through shared-memory scratch.
__device__ __forceinline__ void conv_dw(const Lane& L, const HutterLayout& y, const float* __restrict__ dzr,
                                        const float* __restrict__ s_inr, float* __restrict__ scratch,
                                        float* __restrict__ P) {
  float acc[CONV_CH], accb[CONV_CH];
#pragma unroll
  for (int c = 0; c < CONV_CH; ++c) acc[c] = accb[c] = 0.f;
  const bool kin = L.lane < y.KC;
  for (int t = L.warp; t < y.npos; t += NWARP) {
    for (int d4 = 0; d4 < TM / 4; ++d4) {
      const float* xp = s_inr + (4 * d4) * y.LR + y.RD * t + L.lane;
      const float4 xv = kin ? make_float4(xp[0], xp[y.LR], xp[2 * y.LR], xp[3 * y.LR])
                            : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int c = 0; c < CONV_CH; ++c) {
        const float4 z = *reinterpret_cast<const float4*>(dzr + (c * y.npos + t) * TMP + 4 * d4);
        acc[c] = dot4(z, xv, acc[c]);
        accb[c] += (z.x + z.y) + (z.z + z.w);
      }
    }
  }
  const int stride = CONV_CH * y.KC + CONV_CH;
  float* my = scratch + L.warp * stride;
#pragma unroll
  for (int c = 0; c < CONV_CH; ++c) {
    if (kin) my[c * y.KC + L.lane] = acc[c];
    if (L.lane == 0) my[CONV_CH * y.KC + c] = accb[c];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < stride; idx += NT) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) s += scratch[w * stride + idx];
    if (idx < CONV_CH * y.KC) {
      const int c = idx / y.KC, kk = idx - c * y.KC;
      const int j = kk / y.RD, dch = kk - j * y.RD;
      red_add(P + y.t_wc + c * y.KC + dch * 3 + j, s);      // torch layout [c][d][j]
    } else {
      red_add(P + y.t_bc + idx - CONV_CH * y.KC, s);
    }
  }
}

}  // namespace apg
