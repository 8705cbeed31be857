// Streaming weight-gradient GEMM (tq_dw_kernels.cu): op list, accumulator columns and the map from accumulator
// entries to the torch-flat gradient.  Plain `__host__ __device__` index arithmetic, shared with the CPU checks.
// Quadrotor concurrent net Net(15,10,9,40,conv), see tc_layout.cuh; operand sources (stash sets): tq_layout.cuh.
//
//   dW_l[out][in] = sum over drones dZ_l[drone][out] * X_l[drone][in]
// as  D[M = in (+ a row of ones -> bias gradient)][N = out] += A[in][K = drone] * B[out][K = drone]^T  per 32-drone
// panel; the accumulators stay in TMEM for a whole pass over the CTA's tiles.
#pragma once
#include "tc_layout.cuh"

namespace apg {
namespace dw {

using tc::F0; using tc::H; using tc::RD; using tc::NC; using tc::NPOS; using tc::MO; using tc::K1; using tc::REFW;

enum ASrc { A_H3 = 0, A_H2, A_H1, A_X1_LO, A_X1_HI, A_INSTATE, A_WINDOW };
enum BSrc { B_DZO = 0, B_DZ3, B_DZ2, B_DZ1, B_DZS, B_DZC };

// one GEMM per (tile, op): A rows [0, a_rows) real, row a_rows = ones (if ones >= 0), the rest zero
struct Op { int a_src, a_row0, a_rows, ones, b_src, b_row0, b_rows, N, d_col, first; };
constexpr int NOPS = 10;
// Two PASSES over the CTA's tiles, so that the accumulators of one pass leave TMEM columns for the ring the A operand
// is fed through (tq_dw_kernels.cu): pass 0 = every op but fc1 (288 columns), pass 1 = the two fc1 ops (128 columns).
// Every operand panel is still read exactly once.  Accumulator columns, pass 0: fc_out [0,48) | fc3 [48,112) | fc2
// [112,176) | states_in [176,240) | conv Toeplitz block [240,288); pass 1: fc1 rows 0..127 [0,64) | fc1 rows 128..223
// [64,128).  The A ring starts at C_ARING in both passes.
constexpr int C_WO = 0, C_W3 = 48, C_W2 = 112, C_WS = 176, C_WT = 240, C_W1A = 0, C_W1B = 64, C_ARING = 288;
constexpr int NPASS = 2;
APG_HD constexpr int pass_nops(int pass) { return pass == 0 ? 8 : 2; }
APG_HD constexpr int pass_op(int pass, int k) { return pass == 0 ? (k < 3 ? k : k + 2) : 3 + k; }
APG_HD Op op_of(int i) {
  if (i == 0) return {A_H3, 0, 64, 64, B_DZO, 0, MO, 48, C_WO, 1};
  if (i == 1) return {A_H2, 0, 64, 64, B_DZ3, 0, 64, 64, C_W3, 1};
  if (i == 2) return {A_H1, 0, 64, 64, B_DZ2, 0, 64, 64, C_W2, 1};
  if (i == 3) return {A_X1_LO, 0, 128, -1, B_DZ1, 0, 64, 64, C_W1A, 1};
  if (i == 4) return {A_X1_HI, 128, K1 - 128, K1 - 128, B_DZ1, 0, 64, 64, C_W1B, 1};
  if (i == 5) return {A_INSTATE, 0, F0, F0, B_DZS, 0, 64, 64, C_WS, 1};
  const int g = i - 6;                                   // conv position pair g: window rows 2g .. 2g+3 of in_ref
  return {A_WINDOW, 18 * g, 4 * RD, 4 * RD, B_DZC, HID + 2 * NC * g, 2 * NC, 48, C_WT, g == 0};
}

// x1 row (position-major: 64 + t*20 + c) -> torch fc1 column (channel-major: 64 + c*8 + t)
APG_HD constexpr int fc1_col_of_x1_row(int r) { return r < HID ? r : HID + ((r - HID) % NC) * NPOS + (r - HID) / NC; }

// Where accumulator entry (op region, A row r, column n) goes in the torch-flat gradient; -1: nowhere (padding).
// The conv block is folded separately (conv_entry_sources).
APG_HD int grad_index(const HutterLayout& y, int region, int r, int n) {
  switch (region) {
    case 0: if (n >= MO) return -1; return r < 64 ? y.t_wo + n * HID + r : (r == 64 ? y.t_bo + n : -1);
    case 1: return r < 64 ? y.t_w3 + n * HID + r : (r == 64 ? y.t_b3 + n : -1);
    case 2: return r < 64 ? y.t_w2 + n * HID + r : (r == 64 ? y.t_b2 + n : -1);
    case 3: return y.t_w1 + n * K1 + fc1_col_of_x1_row(r);
    case 4: return r < K1 - 128 ? y.t_w1 + n * K1 + fc1_col_of_x1_row(128 + r) : (r == K1 - 128 ? y.t_b1 + n : -1);
    case 5: return r < F0 ? y.t_ws + n * F0 + r : (r == F0 ? y.t_bs + n : -1);
    default: return -1;
  }
}
// conv_ref.weight[c][ci][j] = sum over tl in {0,1} of T[(tl + j)*9 + ci][tl*20 + c];  bias[c] = sum_tl T[36][tl*20 + c]
// where T = the Toeplitz accumulator block (rows = window element / ones row, columns = conv block output)
APG_HD float conv_weight_from_block(const float* T, int ldt, int c, int ci, int j) {
  return T[((0 + j) * RD + ci) * ldt + c] + T[((1 + j) * RD + ci) * ldt + NC + c];
}
APG_HD float conv_bias_from_block(const float* T, int ldt, int c) { return T[4 * RD * ldt + c] + T[4 * RD * ldt + NC + c]; }

// ---- reduction of the per-CTA partials (torch order, no permutation) with four CTA slices per parameter: thread
//      (slice, p) sums partials[c][p] for the CTAs of its slice in fixed order; the slice sums are added in fixed
//      order -> bitwise reproducible, 4x the memory-level parallelism of apg_reduce_kernel's one thread per parameter.
constexpr int RED_SLICES = 4;
APG_HD void reduce_slice_bounds(int ncta, int slice, int* c0, int* c1) {
  const int per = (ncta + RED_SLICES - 1) / RED_SLICES;
  *c0 = slice * per < ncta ? slice * per : ncta;
  *c1 = (slice + 1) * per < ncta ? (slice + 1) * per : ncta;
}
APG_HD float reduce_slice_sum(const float* partials, int n, int p, int c0, int c1) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int c = c0;
  for (; c + 3 < c1; c += 4) {
    s0 += partials[(size_t)(c + 0) * n + p];
    s1 += partials[(size_t)(c + 1) * n + p];
    s2 += partials[(size_t)(c + 2) * n + p];
    s3 += partials[(size_t)(c + 3) * n + p];
  }
  for (; c < c1; ++c) s0 += partials[(size_t)c * n + p];
  return (s0 + s1) + (s2 + s3);
}

}  // namespace dw
}  // namespace apg
