// Test-only harness for the tcgen05 forward's host-checkable parts (csrc/tc_layout.cuh): the pack-kernel body run
// over its flat index, and an emulation of the kernel's op list FROM THE PACKED IMAGES (fp64 accumulation), with the
// activations written through the same stash addressing the kernel uses.  Not part of the product.
#include <math.h>
#include <string.h>
#include <vector>
#ifndef __CUDACC__
#define __host__      // csrc/layouts.h marks its layout builders with the CUDA execution-space keywords
#define __device__
#endif
#include "tc_layout.cuh"

using namespace apg;
using namespace apg::tc;

extern "C" int hc_tc_blob_bytes() { return BLOB_BYTES; }

extern "C" void hc_tc_pack(const float* params, unsigned char* blob) {
  const HutterLayout y = make_hutter_layout(F0, H, RD, MO, 1);
  for (int e = 0; e < PAIRS_TOTAL + B_TOTAL; ++e) pack_body(e, params, y, blob);
}

// n drones (any n): emulates tiles of 128, returns the stash arrays exactly as the kernel addresses them
// (st_x1 [ntiles64][224][TMP], st_h1/2/3 [ntiles64][64][TMP], st_act [ntiles64][40][TMP])
extern "C" void hc_tc_emulate(const unsigned char* blob, const float* in_state, const float* in_ref, int n,
                              float* st_x1, float* st_h1, float* st_h2, float* st_h3, float* st_act) {
  auto W = [&](const Op& op, int r, int k) {
    float h, l;
    memcpy(&h, blob + op.img_off + kmajor_off(r, k, op.K), 4);
    memcpy(&l, blob + op.img_off + img_bytes(op.rows, op.K) + kmajor_off(r, k, op.K), 4);
    return (double)h + (double)l;
  };
  const float* bias = (const float*)(blob + IMG_TOTAL);
  const int ntiles = (n + TMT - 1) / TMT, ntiles64 = (n + TM - 1) / TM;
  for (int tile = 0; tile < ntiles; ++tile)
    for (int row = 0; row < TMT; ++row) {
      const long d = (long)tile * TMT + row;
      const bool live = d < n;
      if (!(tile * 2 + (row >> 6) < ntiles64)) continue;
      double A[64] = {0}, Dm[64] = {0}, Dc[48] = {0};
      auto run = [&](int i) {
        const Op op = op_of(i);
        double* D = op.d_col == C_DMAIN ? Dm : Dc;
        for (int r = 0; r < op.N; ++r) {
          double a = op.clear ? 0.0 : D[r];
          for (int k = 0; k < op.K; ++k) a += A[k] * W(op, r, k);
          D[r] = a;
        }
      };
      for (int k = 0; k < 16; ++k) A[k] = (live && k < F0) ? in_state[d * F0 + k] : 0.0;
      run(0);
      for (int k = 0; k < 64; ++k) {
        A[k] = tanh(Dm[k] + bias[B_S + k]);
        st_x1[stash_index(tile, row, K1, k)] = (float)A[k];
      }
      run(1);
      for (int g = 0; g < 4; ++g) {
        for (int k = 0; k < 40; ++k) A[k] = (live && k < 36) ? in_ref[d * REFW + 18 * g + k] : 0.0;
        run(2 + 2 * g);
        for (int k = 0; k < 40; ++k) {
          const double v = Dc[k] + bias[B_C + k];
          A[k] = v > 0 ? v : 0;
          st_x1[stash_index(tile, row, K1, x1_row_of_conv(g, k))] = (float)A[k];
        }
        run(3 + 2 * g);
      }
      const int bo[3] = {B_1, B_2, B_3};
      float* st[3] = {st_h1, st_h2, st_h3};
      for (int l = 0; l < 3; ++l) {
        for (int k = 0; k < 64; ++k) {
          A[k] = tanh(Dm[k] + bias[bo[l] + k]);
          st[l][stash_index(tile, row, HID, k)] = (float)A[k];
        }
        run(10 + l);
      }
      for (int o = 0; o < MO; ++o)
        st_act[stash_index(tile, row, MO, o)] = (float)(1.0 / (1.0 + exp(-(Dm[o] + bias[B_O + o]))));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// streaming weight-gradient GEMM (csrc/adj_dw_layout.cuh): the loaders' image fill and the issuer's 3xTF32 products
// emulated from the images (fp64 accumulation), then the accumulator -> gradient map of the kernel's epilogue
// ---------------------------------------------------------------------------------------------------------------
#include "adj_dw_layout.cuh"
#include "tc_emu.h"

extern "C" int hc_dw_num_params() { return make_hutter_layout(F0, H, RD, MO, 1).n_params; }

extern "C" void hc_dw_emulate(const float* st_x1, const float* st_h1, const float* st_h2, const float* st_h3,
                              const float* in_state, const float* in_ref, const float* dzo, const float* dz3,
                              const float* dz2, const float* dz1, const float* dzx, int n, float* P) {
  using dw::AM; using dw::KD; using dw::C_TOTAL; using dw::STAGE_BYTES; using dw::A_IMG_BYTES;
  using dw::B_IMG_BYTES; using dw::B_ROWS; using dw::chunk_off; using dw::chunk_of_item; using dw::a_chunk;
  using dw::b_chunk; using dw::grad_index; using dw::conv_weight_from_block; using dw::conv_bias_from_block;
  using dw::C_WO; using dw::C_W3; using dw::C_W2; using dw::C_W1A; using dw::C_W1B; using dw::C_WS; using dw::C_WT;
  const HutterLayout y = make_hutter_layout(F0, H, RD, MO, 1);
  dw::Sources S;
  S.h3 = st_h3; S.h2 = st_h2; S.h1 = st_h1; S.x1 = st_x1; S.in_state = in_state; S.in_ref = in_ref;
  S.dzo = dzo; S.dz3 = dz3; S.dz2 = dz2; S.dz1 = dz1; S.dzx = dzx;
  std::vector<double> D((size_t)AM * C_TOTAL, 0.0);                       // [lane = A row][column]
  std::vector<unsigned char> stage(STAGE_BYTES);
  unsigned char *a_hi = stage.data(), *a_lo = a_hi + A_IMG_BYTES, *b_hi = a_lo + A_IMG_BYTES, *b_lo = b_hi + B_IMG_BYTES;
  auto put = [&](unsigned char* hi, unsigned char* lo, uint32_t off, const float* x) {
    for (int c = 0; c < 4; ++c) {
      float h, l;
      split_hi_lo(x[c], &h, &l);
      memcpy(hi + off + 4 * c, &h, 4);
      memcpy(lo + off + 4 * c, &l, 4);
    }
  };
  auto at = [&](const unsigned char* img, int r, int k) {
    float v;
    memcpy(&v, img + chunk_off(r, k >> 2) + (k & 3) * 4, 4);
    return (double)v;
  };
  const int ntiles = (n + TM - 1) / TM;
  for (int tile = 0; tile < ntiles; ++tile) {
    const int valid = n - tile * TM < TM ? n - tile * TM : TM;
    for (int i = 0; i < dw::NOPS; ++i) {
      const dw::Op op = dw::op_of(i);
      for (int q = 0; q < AM * (KD / 4); ++q) {
        int r, d4; float x[4];
        chunk_of_item(q, &r, &d4);
        a_chunk(op, S, tile, valid, r, d4, x);
        put(a_hi, a_lo, chunk_off(r, d4), x);
      }
      for (int q = 0; q < B_ROWS * (KD / 4); ++q) {
        int r, d4; float x[4];
        chunk_of_item(q, &r, &d4);
        b_chunk(op, S, tile, r, d4, x);
        put(b_hi, b_lo, chunk_off(r, d4), x);
      }
      const bool clear = tile == 0 && op.first;
      for (int r = 0; r < AM; ++r)
        for (int c = 0; c < op.N; ++c) {
          double acc = clear ? 0.0 : D[(size_t)r * C_TOTAL + op.d_col + c];
          for (int k = 0; k < KD; ++k)
            acc += at(a_lo, r, k) * at(b_hi, c, k) + at(a_hi, r, k) * at(b_lo, c, k) + at(a_hi, r, k) * at(b_hi, c, k);
          D[(size_t)r * C_TOTAL + op.d_col + c] = acc;
        }
    }
  }
  for (int i = 0; i < y.n_params; ++i) P[i] = 0.f;
  const int col0[6] = {C_WO, C_W3, C_W2, C_W1A, C_W1B, C_WS}, ncol[6] = {48, 64, 64, 64, 64, 64};
  std::vector<int> written(y.n_params, 0);
  for (int reg = 0; reg < 6; ++reg)
    for (int r = 0; r < AM; ++r)
      for (int c = 0; c < ncol[reg]; ++c) {
        const int idx = grad_index(y, reg, r, c);
        if (idx >= 0) { P[idx] = (float)D[(size_t)r * C_TOTAL + col0[reg] + c]; ++written[idx]; }
      }
  std::vector<float> T(37 * 48);
  for (int r = 0; r <= 4 * RD; ++r)
    for (int c = 0; c < 48; ++c) T[r * 48 + c] = (float)D[(size_t)r * C_TOTAL + C_WT + c];
  for (int i = 0; i < NC * RD * 3 + NC; ++i) {
    if (i < NC * RD * 3) P[y.t_wc + i] = conv_weight_from_block(T.data(), 48, i / (RD * 3), (i / 3) % RD, i % 3);
    else P[y.t_wc + i] = conv_bias_from_block(T.data(), 48, i - NC * RD * 3);
    ++written[y.t_wc + i];
  }
  // every entry outside ref_in.* must have been written exactly once (the kernel relies on it: no zero-fill there)
  for (int i = 0; i < y.n_params; ++i) {
    const bool ref_in = i >= y.t_wr && i < y.t_br + HID;
    if (written[i] != (ref_in ? 0 : 1)) P[i] = NAN;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Descriptor-level emulation: a software model of tcgen05.mma.kind::tf32 that DECODES the shared-memory and
// instruction descriptors the kernels build (tc_layout.cuh: kmajor_desc / mnmajor_desc / idesc_tf32) and gathers the
// operands from a byte image of shared memory by the canonical unswizzled forms of cute/atom/mma_traits_sm100.hpp
//   Major-K  : ((8,n),2):((1,SBO),LBO)         [16-byte units]  element (i, k): (i%8)*16 + (i/8)*SBO + (k/4)*LBO + (k%4)*4
//   Major-MN : ((1,n),(8,k)):((X,SBO),(1,LBO))                  element (i, k): (i/4)*SBO + (k%8)*16 + (k/8)*LBO + (i%4)*4
// Operands are truncated to TF32 (top 19 bits) like the tensor core input path.  This checks the descriptor arithmetic
// (image offsets, k-step advance, LBO / SBO, M / N fields, MN-major reuse of the forward images) against the
// documented layouts; the hardware itself is checked by the GPU tests.
// ---------------------------------------------------------------------------------------------------------------
namespace emu {
struct Tmem { std::vector<float> v; Tmem() : v(128 * 512, 0.f) {} float& at(int lane, int col) { return v[lane * 512 + col]; } };
static bool mma(Tmem& T, const std::vector<unsigned char>& sm, int d_col, int a_col, uint64_t a_desc, uint64_t b_desc,
                uint32_t idesc, bool acc) {
  return mma(T.v.data(), sm.data(), sm.size(), d_col, a_col, a_desc, b_desc, idesc, acc);
}
static void st_split(Tmem& T, int lane, int ahi, int alo, int col, float x) {
  float h, l; split_hi_lo(x, &h, &l); T.at(lane, ahi + col) = h; T.at(lane, alo + col) = l;
}
}  // namespace emu

// forward of ONE 128-drone tile through the issuer's descriptor arithmetic (slot 0); actions [128][40]
extern "C" int hc_tc_forward_desc(const unsigned char* blob, const float* in_state, const float* in_ref, int n,
                                  float* actions) {
  const uint32_t base = 2048;                                   // some 1024-aligned shared-memory address
  std::vector<unsigned char> sm(base + BLOB_BYTES);
  memcpy(sm.data() + base, blob, BLOB_BYTES);
  const float* bias = (const float*)(blob + IMG_TOTAL);
  emu::Tmem T;
  auto issue = [&](int i) {                                     // == the issuer loop of hutter_fwd_tc_kernel
    const Op op = op_of(i);
    const uint32_t idesc = idesc_tf32(TMT, op.N);
    const uint32_t whi = base + op.img_off, wlo = whi + img_bytes(op.rows, op.K);
    bool ok = true;
    for (int ks = 0; ks < op.K / 8; ++ks) {
      const uint64_t bh = kmajor_desc(whi, ks, op.K), bl = kmajor_desc(wlo, ks, op.K);
      ok &= emu::mma(T, sm, op.d_col, C_ALO + ks * 8, 0, bh, idesc, ks > 0 || !op.clear);
      ok &= emu::mma(T, sm, op.d_col, C_AHI + ks * 8, 0, bl, idesc, true);
      ok &= emu::mma(T, sm, op.d_col, C_AHI + ks * 8, 0, bh, idesc, true);
    }
    return ok;
  };
  bool ok = true;
  for (int r = 0; r < TMT; ++r)
    for (int k = 0; k < 16; ++k) emu::st_split(T, r, C_AHI, C_ALO, k, (r < n && k < F0) ? in_state[r * F0 + k] : 0.f);
  ok &= issue(0);
  for (int r = 0; r < TMT; ++r)
    for (int c = 0; c < 64; ++c) emu::st_split(T, r, C_AHI, C_ALO, c, tanhf(T.at(r, C_DMAIN + c) + bias[B_S + c]));
  ok &= issue(1);
  for (int g = 0; g < 4; ++g) {
    for (int r = 0; r < TMT; ++r)
      for (int k = 0; k < 40; ++k) emu::st_split(T, r, C_AHI, C_ALO, k, (r < n && k < 36) ? in_ref[r * REFW + 18 * g + k] : 0.f);
    ok &= issue(2 + 2 * g);
    for (int r = 0; r < TMT; ++r)
      for (int c = 0; c < 40; ++c) emu::st_split(T, r, C_AHI, C_ALO, c, fmaxf(T.at(r, C_DCONV + c) + bias[B_C + c], 0.f));
    ok &= issue(3 + 2 * g);
  }
  const int bo[3] = {B_1, B_2, B_3};
  for (int l = 0; l < 3; ++l) {
    for (int r = 0; r < TMT; ++r)
      for (int c = 0; c < 64; ++c) emu::st_split(T, r, C_AHI, C_ALO, c, tanhf(T.at(r, C_DMAIN + c) + bias[bo[l] + c]));
    ok &= issue(10 + l);
  }
  for (int r = 0; r < TMT; ++r)
    for (int o = 0; o < MO; ++o) actions[r * MO + o] = 1.f / (1.f + expf(-(T.at(r, C_DMAIN + o) + bias[B_O + o])));
  return ok ? 1 : 0;
}

// dX chain of ONE tile through the MN-major descriptors of hutter_adj_dx_tc_kernel.  Inputs row-major [128][..]:
// dlog (40), h3, h2, h1 (64), x1 (224, position-major).  Outputs dz3, dz2, dz1 [128][64], dzx [128][224].
extern "C" int hc_tc_dx_desc(const unsigned char* blob, const float* dlog, const float* h3, const float* h2,
                             const float* h1, const float* x1, float* dz3, float* dz2, float* dz1, float* dzx) {
  const uint32_t base = 3072;
  std::vector<unsigned char> sm(base + BLOB_BYTES);
  memcpy(sm.data() + base, blob, BLOB_BYTES);
  emu::Tmem T;
  auto issue = [&](int i) {                                     // == the issuer loop of hutter_adj_dx_tc_kernel
    const ROp op = rop_of(i);
    const uint32_t idesc = idesc_tf32(TMT, op.N, 1);
    const uint32_t whi = base + op.img_off, wlo = whi + img_bytes(op.rows, op.Kf);
    bool ok = true;
    for (int ks = 0; ks < op.K / 8; ++ks) {
      const uint64_t bh = mnmajor_desc(whi, ks, op.Kf), bl = mnmajor_desc(wlo, ks, op.Kf);
      ok &= emu::mma(T, sm, op.d_col, C_ALO + ks * 8, 0, bh, idesc, ks > 0);
      ok &= emu::mma(T, sm, op.d_col, C_AHI + ks * 8, 0, bl, idesc, true);
      ok &= emu::mma(T, sm, op.d_col, C_AHI + ks * 8, 0, bh, idesc, true);
    }
    return ok;
  };
  bool ok = true;
  for (int r = 0; r < TMT; ++r)
    for (int c = 0; c < MO; ++c) emu::st_split(T, r, C_AHI, C_ALO, c, dlog[r * MO + c]);
  const float* xs[3] = {h3, h2, h1};
  float* zs[3] = {dz3, dz2, dz1};
  for (int l = 0; l < 3; ++l) {
    ok &= issue(l);
    for (int r = 0; r < TMT; ++r)
      for (int c = 0; c < 64; ++c) {
        const float yv = xs[l][r * 64 + c], v = T.at(r, C_DMAIN + c) * (1.f - yv * yv);
        zs[l][r * 64 + c] = v;
        emu::st_split(T, r, C_AHI, C_ALO, c, v);
      }
  }
  ok &= issue(3);
  for (int r = 0; r < TMT; ++r)
    for (int c = 0; c < 64; ++c) { const float yv = x1[r * K1 + c]; dzx[r * K1 + c] = T.at(r, C_DMAIN + c) * (1.f - yv * yv); }
  for (int g = 0; g < 4; ++g) {
    ok &= issue(4 + g);
    for (int r = 0; r < TMT; ++r)
      for (int c = 0; c < 40; ++c) {
        const int xr = x1_row_of_conv(g, c);
        dzx[r * K1 + xr] = x1[r * K1 + xr] > 0.f ? T.at(r, C_DCONV + c) : 0.f;
      }
  }
  return ok ? 1 : 0;
}

// one (tile, op) of the streaming dW GEMM through the SS descriptors of adj_dw_tc_kernel: returns D[128][N]
extern "C" int hc_dw_op_desc(const float* a_img_src /*[128][64] A rows x drones*/, const float* b_img_src /*[64][64]*/,
                             int N, float* D) {
  using dw::KD; using dw::AM; using dw::A_IMG_BYTES; using dw::B_IMG_BYTES; using dw::STAGE_BYTES; using dw::chunk_off;
  const uint32_t base = 1024 + STAGE_BYTES;                      // stage 1
  std::vector<unsigned char> sm(base + STAGE_BYTES, 0);
  unsigned char *a_hi = sm.data() + base, *a_lo = a_hi + A_IMG_BYTES, *b_hi = a_lo + A_IMG_BYTES, *b_lo = b_hi + B_IMG_BYTES;
  auto put = [&](unsigned char* hi, unsigned char* lo, int r, int d, float x) {
    float h, l; split_hi_lo(x, &h, &l);
    memcpy(hi + chunk_off(r, d >> 2) + (d & 3) * 4, &h, 4); memcpy(lo + chunk_off(r, d >> 2) + (d & 3) * 4, &l, 4);
  };
  for (int r = 0; r < AM; ++r) for (int d = 0; d < KD; ++d) put(a_hi, a_lo, r, d, a_img_src[r * KD + d]);
  for (int r = 0; r < 64; ++r) for (int d = 0; d < KD; ++d) put(b_hi, b_lo, r, d, b_img_src[r * KD + d]);
  emu::Tmem T;
  const uint32_t A_hi = base, A_lo = A_hi + A_IMG_BYTES, B_hi = A_lo + A_IMG_BYTES, B_lo = B_hi + B_IMG_BYTES;
  const uint32_t idesc = idesc_tf32(AM, N);
  bool ok = true;
  for (int ks = 0; ks < KD / 8; ++ks) {                          // == the issuer loop of adj_dw_tc_kernel
    const uint64_t ah = kmajor_desc(A_hi, ks, KD), al = kmajor_desc(A_lo, ks, KD);
    const uint64_t bh = kmajor_desc(B_hi, ks, KD), bl = kmajor_desc(B_lo, ks, KD);
    ok &= emu::mma(T, sm, 16, -1, al, bh, idesc, ks > 0);
    ok &= emu::mma(T, sm, 16, -1, ah, bl, idesc, true);
    ok &= emu::mma(T, sm, 16, -1, ah, bh, idesc, true);
  }
  for (int r = 0; r < AM; ++r) for (int c = 0; c < N; ++c) D[r * N + c] = T.at(r, 16 + c);
  return ok ? 1 : 0;
}

// the sliced reduction of the per-CTA partials exactly as apg_reduce4_kernel computes it
extern "C" void hc_reduce4(const float* partials, int ncta, int n, float scale, float* grad) {
  for (int p = 0; p < n; ++p) {
    float part[dw::RED_SLICES];
    for (int s = 0; s < dw::RED_SLICES; ++s) {
      int c0, c1;
      dw::reduce_slice_bounds(ncta, s, &c0, &c1);
      part[s] = dw::reduce_slice_sum(partials, n, p, c0, c1);
    }
    grad[p] = scale * ((part[0] + part[1]) + (part[2] + part[3]));
  }
}
