// Streaming weight-gradient GEMM of the split adjoint (adj_dw_tc_kernels.cu): op list, operand sources and the map
// from accumulator entries to the torch-flat gradient.  Plain `__host__ __device__` index arithmetic, shared with the
// CPU check (tests/hostcheck/hostcheck_tc.cpp).  Quadrotor concurrent net Net(15,10,9,40,conv), see tc_layout.cuh.
//
//   dW_l[out][in] = sum over drones dZ_l[drone][out] * X_l[drone][in]
// as  D[M = in (+ a row of ones -> bias gradient)][N = out] += A[in][K = drone] * B[out][K = drone]^T  per 64-drone
// stash tile (K = 64, eight tcgen05 k-steps), both operands K-major unswizzled (hi, lo) images filled by the loader
// warps from the feature-major stash tiles [rows][TMP]; the accumulators stay in TMEM for the whole launch.
#pragma once
#include "tc_layout.cuh"

namespace apg {
namespace dw {

using tc::F0; using tc::H; using tc::RD; using tc::NC; using tc::NPOS; using tc::MO; using tc::K1; using tc::REFW;

constexpr int KD = 64;                         // drones per K block = one stash tile
constexpr int AM = 128;                        // rows of every A image (M of the MMA)
constexpr int A_IMG_BYTES = AM * KD * 4;       // 32 KiB per hi or lo image
constexpr int B_ROWS = 64;
constexpr int B_IMG_BYTES = B_ROWS * KD * 4;   // 16 KiB
constexpr int STAGE_BYTES = 2 * A_IMG_BYTES + 2 * B_IMG_BYTES;     // 96 KiB
constexpr int NSTAGE = 2;

enum ASrc { A_H3 = 0, A_H2, A_H1, A_X1_LO, A_X1_HI, A_INSTATE, A_WINDOW };
enum BSrc { B_DZO = 0, B_DZ3, B_DZ2, B_DZ1, B_DZS, B_DZC };

// one GEMM per (tile, op): A rows [0, a_rows) real, row a_rows = ones (if ones >= 0), the rest zero
struct Op { int a_src, a_row0, a_rows, ones, b_src, b_row0, b_rows, N, d_col, first; };
constexpr int NOPS = 10;
// accumulator columns: fc_out [0,48) | fc3 [48,112) | fc2 [112,176) | fc1 rows 0..127 [176,240) | fc1 rows 128..223
// [240,304) | states_in [304,368) | conv Toeplitz block [368,416)
constexpr int C_WO = 0, C_W3 = 48, C_W2 = 112, C_W1A = 176, C_W1B = 240, C_WS = 304, C_WT = 368, C_TOTAL = 416;
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

// byte offset of the 16-byte chunk (row r, drones 4*d4 .. 4*d4+3) inside a K-major unswizzled image with K = 64
APG_HD constexpr uint32_t chunk_off(int r, int d4) { return (uint32_t)((r >> 3) * 2048 + d4 * 128 + (r & 7) * 16); }

// loader work item q -> (image row r, chunk d4): a warp's 32 items cover 8 rows x 4 chunks = 512 contiguous bytes of
// the image (conflict-free st.shared.v4) and 8 x 64 B of the stash tile
APG_HD void chunk_of_item(int q, int* r, int* d4) { *r = (q & 7) + ((q >> 7) << 3); *d4 = (q >> 3) & 15; }

// base pointers of everything the GEMM streams (tile-major stashes [tile][rows][TMP], drone-major policy inputs)
struct Sources {
  const float *h3, *h2, *h1, *x1, *in_state, *in_ref;     // X_l
  const float *dzo, *dz3, *dz2, *dz1, *dzx;                 // dZ_l
};

APG_HD void load4(const float* p, float* o) {
#if defined(__CUDA_ARCH__)
  const float4 v = *reinterpret_cast<const float4*>(p);
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
#else
  o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[3] = p[3];
#endif
}

// the four values (row r, drones 4*d4 .. 4*d4+3) of the A image of op for stash tile `tile` (`valid` live drones)
APG_HD void a_chunk(const Op& op, const Sources& S, int tile, int valid, int r, int d4, float* o) {
  o[0] = o[1] = o[2] = o[3] = 0.f;
  if (r < op.a_rows) {
    const float* t = nullptr;
    if (op.a_src == A_H3) t = S.h3 + (size_t)tile * HID * TMP;
    else if (op.a_src == A_H2) t = S.h2 + (size_t)tile * HID * TMP;
    else if (op.a_src == A_H1) t = S.h1 + (size_t)tile * HID * TMP;
    else if (op.a_src == A_X1_LO || op.a_src == A_X1_HI) t = S.x1 + ((size_t)tile * K1 + op.a_row0) * TMP;
    if (t) {
      load4(t + r * TMP + 4 * d4, o);
    } else {
      for (int c = 0; c < 4; ++c) {
        const int dd = 4 * d4 + c;
        const size_t drone = (size_t)tile * TM + dd;
        if (dd < valid)
          o[c] = op.a_src == A_INSTATE ? S.in_state[drone * F0 + r] : S.in_ref[drone * REFW + op.a_row0 + r];
      }
    }
  } else if (r == op.ones) {
    o[0] = o[1] = o[2] = o[3] = 1.f;
  }
}
// the same for the B image (dZ rows [b_row0, b_row0 + b_rows), zero above)
APG_HD void b_chunk(const Op& op, const Sources& S, int tile, int r, int d4, float* o) {
  o[0] = o[1] = o[2] = o[3] = 0.f;
  if (r >= op.b_rows) return;
  const float* t = op.b_src == B_DZO   ? S.dzo + (size_t)tile * MO * TMP
                   : op.b_src == B_DZ3 ? S.dz3 + (size_t)tile * HID * TMP
                   : op.b_src == B_DZ2 ? S.dz2 + (size_t)tile * HID * TMP
                   : op.b_src == B_DZ1 ? S.dz1 + (size_t)tile * HID * TMP
                                       : S.dzx + ((size_t)tile * K1 + op.b_row0) * TMP;
  load4(t + r * TMP + 4 * d4, o);
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
