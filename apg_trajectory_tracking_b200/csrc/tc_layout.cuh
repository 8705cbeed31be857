// tcgen05 / TMEM forward of the quadrotor concurrent policy Net(15, 10, 9, 40, conv=True): shared-memory weight
// images and op list, shared by the device code (tq_kernels.cu) and the CPU checks (tests/hostcheck).  Everything
// here is plain `__host__ __device__` index arithmetic.  Stash format, dX chain and dW GEMM: tq_layout.cuh.
//
// Design (measured prototype: tools/micro/tcgen05_policy.cu):
//   * tile = 128 drones = 128 TMEM lanes; epilogue thread r owns drone r of the tile.
//   * every weight matrix is resident in shared memory as a (hi, lo) pair of K-major, unswizzled TF32 images
//     (core matrix = 8 rows x 16 B); 3xTF32: D = A_lo W_hi + A_hi W_lo + A_hi W_hi, fp32 accumulation in TMEM.
//   * activations stay in TMEM (tcgen05.ld -> bias / activation / split -> tcgen05.st into the A columns).
//   * conv1d(9 -> 20, k = 3) is evaluated two output positions at a time: a 4-row window of in_ref (36 values,
//     padded to 40) times a 40 x 36 Toeplitz block; fc1 is accumulated in five pieces (the state block and one
//     40-wide block per position pair), its columns permuted to the conv block's (position, channel) order.
#pragma once
#include <stdint.h>
#include "apg_math.cuh"
#include "layouts.h"

namespace apg {
namespace tc {

constexpr int TMT = 128;                      // drones per tcgen05 tile
constexpr int F0 = 15, H = 10, RD = 9, NC = 20, NPOS = 8, MO = 40;
constexpr int REFW = H * RD;                  // floats of in_ref per drone
constexpr int K1 = HID + NC * NPOS;           // 224

struct Img { int off, rows, K; };
APG_HD constexpr int img_bytes(int rows, int K) { return rows * K * 4; }
constexpr Img I_WS{0, 64, 16};
constexpr Img I_W1S{I_WS.off + 2 * img_bytes(64, 16), 64, 64};
constexpr Img I_WT{I_W1S.off + 2 * img_bytes(64, 64), 48, 40};
constexpr Img I_W1G{I_WT.off + 2 * img_bytes(48, 40), 64, 40};     // 4 consecutive (hi, lo) pairs, one per position pair
constexpr Img I_W2{I_W1G.off + 4 * 2 * img_bytes(64, 40), 64, 64};
constexpr Img I_W3{I_W2.off + 2 * img_bytes(64, 64), 64, 64};
constexpr Img I_WO{I_W3.off + 2 * img_bytes(64, 64), 48, 64};
constexpr int IMG_TOTAL = I_WO.off + 2 * img_bytes(48, 64);        // 228352 B
// biases (floats) after the images: bs 64 | bc 48 (two positions x 20 channels, 8 pad) | b1 64 | b2 64 | b3 64 | bo 48
constexpr int B_S = 0, B_C = 64, B_1 = 112, B_2 = 176, B_3 = 240, B_O = 304, B_TOTAL = 352;
constexpr int BLOB_BYTES = IMG_TOTAL + B_TOTAL * 4;                // what the pack kernel writes / the CTA loads
constexpr int NUM_IMAGES = 10;                                     // WS, W1S, WT, W1G x4, W2, W3, WO

// TMEM columns inside one 256-column slot (two slots = two tiles in flight)
constexpr int C_DMAIN = 0, C_DCONV = 64, C_AHI = 112, C_ALO = 176, SLOT_COLS = 256;

// byte offset of element (r, k) inside a K-major unswizzled image with K columns
APG_HD constexpr uint32_t kmajor_off(int r, int k, int K) {
  return (uint32_t)((r >> 3) * ((K >> 2) * 128) + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4);
}

// image i of the blob: offset of its hi part, rows, K
APG_HD Img image_of(int i) {
  if (i == 0) return I_WS;
  if (i == 1) return I_W1S;
  if (i == 2) return I_WT;
  if (i < 7) return Img{I_W1G.off + (i - 3) * 2 * img_bytes(64, 40), 64, 40};
  if (i == 7) return I_W2;
  if (i == 8) return I_W3;
  return I_WO;
}

// value of element (r, k) of image i, read from the torch-flat parameter vector (layout y: make_hutter_layout(15,
// 10, 9, 40, conv)); zero in the padding.  models/hutter_model.py:12-30 parameter layouts.
APG_HD float image_value(const float* P, const HutterLayout& y, int i, int r, int k) {
  if (i == 0) return k < F0 ? P[y.t_ws + r * F0 + k] : 0.f;                              // states_in.weight [64][15]
  if (i == 1) return P[y.t_w1 + r * K1 + k];                                             // fc1 columns of the s block
  if (i == 2) {                                                                           // Toeplitz block of conv_ref
    if (r >= 2 * NC || k >= 4 * RD) return 0.f;
    const int tl = r / NC, c = r - tl * NC, tr = k / RD, ci = k - tr * RD, jj = tr - tl;
    return (jj >= 0 && jj < 3) ? P[y.t_wc + (c * RD + ci) * 3 + jj] : 0.f;               // weight [c][ci][j]
  }
  if (i < 7) {                                                                            // fc1 columns of pair g
    const int g = i - 3, tl = k / NC, c = k - tl * NC;
    return P[y.t_w1 + r * K1 + HID + c * NPOS + 2 * g + tl];                             // torch: channel-major
  }
  if (i == 7) return P[y.t_w2 + r * HID + k];
  if (i == 8) return P[y.t_w3 + r * HID + k];
  return r < MO ? P[y.t_wo + r * HID + k] : 0.f;                                          // fc_out.weight [40][64]
}

APG_HD float bias_value(const float* P, const HutterLayout& y, int j) {
  if (j < B_C) return P[y.t_bs + j];
  if (j < B_1) { const int q = j - B_C; return q < 2 * NC ? P[y.t_bc + q % NC] : 0.f; }
  if (j < B_2) return P[y.t_b1 + (j - B_1)];
  if (j < B_3) return P[y.t_b2 + (j - B_2)];
  if (j < B_O) return P[y.t_b3 + (j - B_3)];
  return (j - B_O) < MO ? P[y.t_bo + (j - B_O)] : 0.f;
}

APG_HD void split_hi_lo(float x, float* hi, float* lo) {
#if defined(__CUDA_ARCH__)
  const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
#else
  union { float f; uint32_t u; } v;
  v.f = x; v.u &= 0xffffe000u;
  const float h = v.f;
#endif
  *hi = h; *lo = x - h;
}

// Pack-kernel body: flat index e over all (hi, lo) element pairs of the images, then the biases.
constexpr int PAIRS_TOTAL = IMG_TOTAL / 8;
APG_HD void pack_body(int e, const float* P, const HutterLayout& y, unsigned char* blob) {
  if (e >= PAIRS_TOTAL) {
    const int j = e - PAIRS_TOTAL;
    if (j < B_TOTAL) reinterpret_cast<float*>(blob + IMG_TOTAL)[j] = bias_value(P, y, j);
    return;
  }
  int i = 0, base = 0;
  for (; i < NUM_IMAGES; ++i) {
    const Img im = image_of(i);
    const int cnt = im.rows * im.K;
    if (e < base + cnt) {
      const int q = e - base, r = q / im.K, k = q - r * im.K;
      float hi, lo;
      split_hi_lo(image_value(P, y, i, r, k), &hi, &lo);
      *reinterpret_cast<float*>(blob + im.off + kmajor_off(r, k, im.K)) = hi;
      *reinterpret_cast<float*>(blob + im.off + img_bytes(im.rows, im.K) + kmajor_off(r, k, im.K)) = lo;
      return;
    }
    base += cnt;
  }
}

// ---- tcgen05 descriptors (cute/arch/mma_sm100_desc.hpp).  Pure integer functions of a 32-bit shared-memory address so
//      that the CPU check can decode them again (tests/hostcheck/hostcheck_tc.cpp emulates the MMAs from the
//      descriptors, following the canonical unswizzled forms of cute/atom/mma_traits_sm100.hpp).
// shared-memory matrix descriptor: start address [0,14), leading byte offset [16,30), stride byte offset [32,46)
// (all >> 4), version 1 at bit 46, layout type 0 (no swizzle)
APG_HD uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// K-major unswizzled image with K columns, k-step ks (8 tf32 = two 16-byte chunks): K-adjacent core matrices 128 B
// apart (LBO), 8-row groups (K/4)*128 B apart (SBO)
APG_HD uint64_t kmajor_desc(uint32_t base, int ks, int K) { return smem_desc(base + ks * 256, 128, (K >> 2) * 128); }
// instruction descriptor, kind::tf32: c_format F32 (bit 4), a/b format TF32 (bits 7, 10), N >> 3 at bit 17, M >> 4 at
// bit 24; both operands K-major (bits 15 / 16 = 0: an MN-major tf32 operand yields zeros on B200, measured with
// tools/micro/tcgen05_probe.cu)
APG_HD uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// one GEMM of the op list: D[d_col, +N) (=|+=) A[0, K) * W^T
struct Op { int img_off, rows, K, d_col, N, clear; };
constexpr int NOPS = 13;
APG_HD Op op_of(int i) {
  if (i == 0) return {I_WS.off, 64, 16, C_DMAIN, 64, 1};
  if (i == 1) return {I_W1S.off, 64, 64, C_DMAIN, 64, 1};
  if (i < 10) {
    const int g = (i - 2) >> 1;
    if ((i & 1) == 0) return {I_WT.off, 48, 40, C_DCONV, 48, 1};
    return {I_W1G.off + g * 2 * img_bytes(64, 40), 64, 40, C_DMAIN, 64, 0};
  }
  if (i == 10) return {I_W2.off, 64, 64, C_DMAIN, 64, 1};
  if (i == 11) return {I_W3.off, 64, 64, C_DMAIN, 64, 1};
  return {I_WO.off, 48, 64, C_DMAIN, 48, 1};
}

}  // namespace tc
}  // namespace apg
