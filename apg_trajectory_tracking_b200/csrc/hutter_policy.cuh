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
// ---- first-layer GEMMs on the tensor path.  Their A operand is a drone-major (AoS) input tile: A[m = drone][k]
//      = tile[drone * lda + off + k]; k beyond K reads as 0 (the packed weights are zero there as well).
__device__ __forceinline__ void load_a_aos(const float* __restrict__ tile, int lda, int off, int m0, int g, int t,
                                           int k0, int K, uint32_t (&ah)[4], uint32_t (&al)[4]) {
  const float* p0 = tile + (m0 + g) * lda + off + k0 + t;
  const float* p1 = p0 + 8 * lda;
  const bool v0 = k0 + t < K, v1 = k0 + t + 4 < K;
  split_tf32(v0 ? p0[0] : 0.f, ah[0], al[0]);
  split_tf32(v0 ? p1[0] : 0.f, ah[1], al[1]);
  split_tf32(v1 ? p0[4] : 0.f, ah[2], al[2]);
  split_tf32(v1 ? p1[4] : 0.f, ah[3], al[3]);
}

// Y[row0 + n*row_stride][d] = act(bias[n] + sum_k A[d][k] W[k][n]), n < Nreal, for ONE 16-drone block m0 and NTC
// 8-wide column tiles starting at column tile nt0.  W [Kp][ldw] (rows >= K zero), Kp = K rounded up to 8.
template <int NTC>
__device__ __forceinline__ void aos_mma_block(const Lane& L, const float* __restrict__ tile, int lda, int off, int K,
                                              const float* __restrict__ W, int ldw, int sw,
                                              const float* __restrict__ bias, int Nreal, int m0, int nt0, float* Y,
                                              int row0, int row_stride, int act) {
  const int g = L.lane >> 2, t = L.lane & 3;
  const int xs = sw ? (t << 3) : 0;
  float acc[NTC][4], acc1[NTC][4], acc2[NTC][4];
#pragma unroll
  for (int j = 0; j < NTC; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = acc1[j][e] = acc2[j][e] = 0.f;
  const float* wp = W + t * ldw;
  for (int k0 = 0; k0 < K; k0 += 8) {
    uint32_t ah[4], al[4];
    load_a_aos(tile, lda, off, m0, g, t, k0, K, ah, al);
    uint32_t bh[NTC][2], bl[NTC][2];
#pragma unroll
    for (int j = 0; j < NTC; ++j) {
      const int n = ((nt0 + j) * 8 + g) ^ xs;
      split_tf32(wp[k0 * ldw + n], bh[j][0], bl[j][0]);
      split_tf32(wp[(k0 + 4) * ldw + n], bh[j][1], bl[j][1]);
    }
#pragma unroll
    for (int j = 0; j < NTC; ++j) mma_tf32(acc1[j], al, bh[j][0], bh[j][1]);
#pragma unroll
    for (int j = 0; j < NTC; ++j) mma_tf32(acc2[j], ah, bl[j][0], bl[j][1]);
#pragma unroll
    for (int j = 0; j < NTC; ++j) mma_tf32(acc[j], ah, bh[j][0], bh[j][1]);
  }
#pragma unroll
  for (int j = 0; j < NTC; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] += acc1[j][e] + acc2[j][e];
#pragma unroll
  for (int j = 0; j < NTC; ++j) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int n = (nt0 + j) * 8 + 2 * t + (e & 1);
      const int d = m0 + g + ((e >> 1) << 3);
      if (n < Nreal) Y[(row0 + n * row_stride) * TMP + d] = act_apply(acc[j][e] + bias[n], act);
    }
  }
}

// Conv1d(RD -> 20, k=3, valid) + relu on the drone-major reference tile s_inr [TM][LR]; output rows row0 + c*npos + t.
// Wc packed [pad8(KC)][24] (rows >= KC and columns >= 20 zero).
constexpr int CONV_LD = 24;
__device__ __forceinline__ void conv_layer_fwd(const Lane& L, int npos, int RD, int LR, int KC, const float* Wc,
                                               const float* bc, const float* s_inr, float* Y, int row0, int cs,
                                               int ts) {
  const int m0 = (L.warp & 3) * 16;
  for (int t = (L.warp >> 2); t < npos; t += 2)
    aos_mma_block<3>(L, s_inr, LR, RD * t, KC, Wc, CONV_LD, 0, bc, CONV_CH, m0, 0, Y, row0 + t * ts, cs, ACT_RELU);
}

// tanh(Linear(K -> 64)) of a drone-major input tile -> rows [row0, row0+64)
__device__ __forceinline__ void aos_linear64_fwd(const Lane& L, const float* s_in, int lda, int K, const float* W,
                                                 const float* b, float* Y, int row0) {
  aos_mma_block<4>(L, s_in, lda, 0, K, W, HID, mma_sw(HID), b, HID, (L.warp & 3) * 16, (L.warp >> 2) * 4, Y, row0, 1,
                   ACT_TANH);
}

// First layer: s = tanh(states_in(state)) -> X1 rows [0,64);  reference branch -> X1 rows [64, K1):
//   CONV : Conv1d + relu, output index channel-major c*npos + t (hutter_model.py:36-40)
//   !CONV: tanh(ref_in(ref))
// s_ins / s_inr are the drone-major input tiles [TM][F0] / [TM][L*RD].
template <bool CONV>
__device__ __forceinline__ void hutter_first_layer(const Lane& L, const HutterLayout& y, const float* s_w,
                                                   const float* s_ins, const float* s_inr, float* s_x1) {
  aos_linear64_fwd(L, s_ins, y.F0, y.F0, s_w + y.f_ws, s_w + y.f_bs, s_x1, 0);
  if (CONV)
    conv_layer_fwd(L, y.npos, y.RD, y.LR, y.KC, s_w + y.f_wr, s_w + y.f_br, s_inr, s_x1, HID, y.conv_cs, y.conv_ts);
  else
    aos_linear64_fwd(L, s_inr, y.LR, y.LR, s_w + y.f_wr, s_w + y.f_br, s_x1, HID);
}

// fc1 .. fc_out on a tile whose X1 is complete in s_x1 AND whose X1 stash store has been committed as the most
// recent bulk group by thread 0.  Leaves sigmoid(logits) in s_act (rows [0, Mo4)) and every hidden activation in the
// global stash.  Ends with a group barrier (actions visible to the GEMM group).  WS: warp-specialised caller (the
// GEMM group is threads 0..255 and syncs on named barrier 1); act_wait: optional mbarrier to wait on (parity
// act_parity) before s_act is overwritten (the dynamics warps have finished with the previous tile's actions).
template <bool WS>
__device__ __forceinline__ void hutter_trunk(const Lane& L, const HutterLayout& y, const float* s_w, float* s_x1,
                                             float* s_h, float* s_act, float* st_h1, float* st_h2, float* st_h3,
                                             float* st_act, uint64_t* act_wait = nullptr, uint32_t act_parity = 0) {
  const int tid = threadIdx.x;
  dense_auto<EPI_ACT>(L, s_x1, y.K1, s_w + y.f_w1, HID, mma_sw(HID), s_w + y.f_b1, HID, s_h, 0, ACT_TANH);
  fence_proxy_async();
  gsync<WS>();
  if (tid == 0) {
    bulk_s2g(st_h1, s_h, HID * TMP * 4);
    bulk_commit();
    bulk_wait_read<1>();      // the X1 store has finished reading s_x1
  }
  gsync<WS>();
  // fc2 : s_h -> s_x1 rows [0,64)
  dense_auto<EPI_ACT>(L, s_h, HID, s_w + y.f_w2, HID, mma_sw(HID), s_w + y.f_b2, HID, s_x1, 0, ACT_TANH);
  fence_proxy_async();
  gsync<WS>();
  if (tid == 0) {
    bulk_s2g(st_h2, s_x1, HID * TMP * 4);
    bulk_commit();
    bulk_wait_read<1>();      // the h1 store has finished reading s_h
  }
  gsync<WS>();
  // fc3 : s_x1 rows [0,64) -> s_h
  dense_auto<EPI_ACT>(L, s_x1, HID, s_w + y.f_w3, HID, mma_sw(HID), s_w + y.f_b3, HID, s_h, 0, ACT_TANH);
  fence_proxy_async();
  gsync<WS>();
  if (tid == 0) {
    bulk_s2g(st_h3, s_h, HID * TMP * 4);
    bulk_commit();
  }
  // fc_out + sigmoid : s_h -> s_act
  if (act_wait) mbar_wait(act_wait, act_parity);
  dense_auto<EPI_ACT>(L, s_h, HID, s_w + y.f_wo, y.ld_fwo, mma_sw(y.ld_fwo), s_w + y.f_bo, y.Mo4, s_act, 0,
                      ACT_SIGMOID);
  fence_proxy_async();
  gsync<WS>();
  if (tid == 0) {
    bulk_s2g(st_act, s_act, y.Mo4 * TMP * 4);
    bulk_commit();
  }
}

// ---- first-layer weight gradients on the tensor path.  One 16x8 output tile per warp, reduction over the drones
//      (and over the conv positions).  B operand = the drone-major input tile: B[k = drone][n] = tile[drone*lda+off+n].
__device__ __forceinline__ void load_b_aos(const float* __restrict__ tile, int lda, int off, int d0, int t, int n,
                                           bool nvalid, uint32_t (&bh)[2], uint32_t (&bl)[2]) {
  const float* p = tile + (d0 + t) * lda + off + n;
  split_tf32(nvalid ? p[0] : 0.f, bh[0], bl[0]);
  split_tf32(nvalid ? p[4 * lda] : 0.f, bh[1], bl[1]);
}

// dW[j][k] += sum_d dZ[j][d] * in[d][k]  for a Linear(K -> 64) whose input is the drone-major tile (K <= 16 per pass)
__device__ __forceinline__ void aos_linear64_dw(const Lane& L, const float* __restrict__ dz, const float* __restrict__ s_in,
                                                int lda, int K, float* __restrict__ P, float* __restrict__ Pb) {
  const int g = L.lane >> 2, t = L.lane & 3;
  bias_grad(dz, HID, Pb);
  const int j0 = (L.warp & 3) * 16;
  for (int nt = (L.warp >> 2); nt * 8 < K; nt += 2) {
    // three independent accumulator chains (lo*hi, hi*lo, hi*hi), summed at the end
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
    const float* zp = dz + (j0 + g) * TMP + t;
    const int n = nt * 8 + g;
#pragma unroll 2
    for (int d0 = 0; d0 < TM; d0 += 8) {
      uint32_t ah[4], al[4], bh[2], bl[2];
      split_tf32(zp[d0], ah[0], al[0]);
      split_tf32(zp[8 * TMP + d0], ah[1], al[1]);
      split_tf32(zp[d0 + 4], ah[2], al[2]);
      split_tf32(zp[8 * TMP + d0 + 4], ah[3], al[3]);
      load_b_aos(s_in, lda, 0, d0, t, n, n < K, bh, bl);
      mma_tf32(acc1, al, bh[0], bh[1]);
      mma_tf32(acc2, ah, bl[0], bl[1]);
      mma_tf32(acc, ah, bh[0], bh[1]);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e] += acc1[e] + acc2[e];
    const int k = nt * 8 + 2 * t;
    if (k < K) { red_add(P + (j0 + g) * K + k, acc[0]); red_add(P + (j0 + g + 8) * K + k, acc[2]); }
    if (k + 1 < K) { red_add(P + (j0 + g) * K + k + 1, acc[1]); red_add(P + (j0 + g + 8) * K + k + 1, acc[3]); }
  }
}

// conv_ref weight gradient: dWc[c][d][j] = sum_t sum_drone dz[c*npos+t][drone] * in_ref[drone][(t+j)*RD + d]
// (kk = j*RD + d).  8 output tiles (2 row tiles of channels x 4 column tiles of kk) <-> 8 warps.
__device__ __forceinline__ void conv_dw(const Lane& L, const HutterLayout& y, const float* __restrict__ dzr,
                                        const float* __restrict__ s_inr, float* __restrict__ /*scratch*/,
                                        float* __restrict__ P) {
  const int g = L.lane >> 2, t = L.lane & 3;
  // bias: db[c] = sum_t sum_d dz[c*npos+t][d]
  if (threadIdx.x < CONV_CH) {
    float s = 0.f;
    for (int tt = 0; tt < y.npos; ++tt) {
      const float* zr = dzr + (threadIdx.x * y.conv_cs + tt * y.conv_ts) * TMP;
#pragma unroll
      for (int d4 = 0; d4 < TM / 4; ++d4) {
        const float4 z = *reinterpret_cast<const float4*>(zr + 4 * d4);
        s += (z.x + z.y) + (z.z + z.w);
      }
    }
    red_add(P + y.t_bc + threadIdx.x, s);
  }
  const int c0 = (L.warp & 1) * 16;
  const int ca = min(c0 + g, CONV_CH - 1), cb = min(c0 + g + 8, CONV_CH - 1);      // clamped rows (discarded)
  for (int nt = (L.warp >> 1); nt * 8 < y.KC; nt += 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
    const int n = nt * 8 + g;
    for (int tt = 0; tt < y.npos; ++tt) {
      const float* za = dzr + (ca * y.conv_cs + tt * y.conv_ts) * TMP + t;
      const float* zb = dzr + (cb * y.conv_cs + tt * y.conv_ts) * TMP + t;
#pragma unroll 2
      for (int d0 = 0; d0 < TM; d0 += 8) {
        uint32_t ah[4], al[4], bh[2], bl[2];
        split_tf32(za[d0], ah[0], al[0]);
        split_tf32(zb[d0], ah[1], al[1]);
        split_tf32(za[d0 + 4], ah[2], al[2]);
        split_tf32(zb[d0 + 4], ah[3], al[3]);
        load_b_aos(s_inr, y.LR, y.RD * tt, d0, t, n, n < y.KC, bh, bl);
        mma_tf32(acc1, al, bh[0], bh[1]);
        mma_tf32(acc2, ah, bl[0], bl[1]);
        mma_tf32(acc, ah, bh[0], bh[1]);
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e] += acc1[e] + acc2[e];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = c0 + g + ((e >> 1) << 3), kk = nt * 8 + 2 * t + (e & 1);
      if (c < CONV_CH && kk < y.KC) {
        const int j = kk / y.RD, dch = kk - j * y.RD;
        red_add(P + y.t_wc + c * y.KC + dch * 3 + j, acc[e]);      // torch layout [c][d][j]
      }
    }
  }
}

}  // namespace apg
