// tcgen05 / TMEM kernels of the quadrotor CONCURRENT rollout, second generation ("tq": tq_kernels.cu forward + dX
// chain, tq_dw_kernels.cu streaming weight-gradient GEMM): stash format, transposed weight images, op lists.
// Plain `__host__ __device__` index arithmetic shared by the device code and the CPU checks (tests/hostcheck).
//
// What the first hardware runs of the round-1 tcgen05 kernels showed (profiles/r2_tcgen05_first_runs.md) and what
// this layout answers:
//   * kind::tf32 with an MN-major operand returns ZEROS on B200 (address-function probe tools/micro/tcgen05_probe.cu),
//     K-major unswizzled and K-major 128B-swizzled operands decode exactly as CUTLASS's canonical forms.  So the dX
//     chain gets its own TRANSPOSED K-major weight images (packed next to the forward images every call) instead of
//     re-reading the forward images MN-major.
//   * the kernels were latency bound (9 warps per SM, dependent global loads, scalar stash traffic), not tensor bound.
//     So: the stash is written in the exact shared-memory image a tcgen05 operand needs (K-major, 128B swizzle, K =
//     drone axis in 32-drone panels), every warp store / load of it is one full 128-byte line, and the weight-gradient
//     GEMM fetches whole operand panels with 1-D bulk copies (no loader arithmetic, no tensor maps).
//
// Stash of one 128-drone tile = a list of SETS; set with R rows (R % 8 == 0): [panel 0..3][row 0..R-1][32 drones] fp32,
// the 16-byte chunk index of a row XOR-ed with (row & 7) (UMMA SWIZZLE_128B, K-major):
//     byte(row, d) = (d >> 5) * R * 128 + row * 128 + ((((d & 31) >> 2) ^ (row & 7)) << 4) + (d & 3) * 4
// A panel of any row range [r0, r0 + m) with r0 % 8 == 0 is m * 128 contiguous bytes that can be bulk-copied to a
// 1024-byte aligned shared-memory address and used as an SS operand (start address + 32 * kstep, SBO = 1024).
#pragma once
#include "tc_layout.cuh"
#include "adj_dw_layout.cuh"

// ---- optional per-role cycle accounting (build with -DAPG_PROFILE; tools/tq_profile.py): kernel k, CTA, counter
#ifdef APG_PROFILE
#define TQ_NPROF 20
__device__ __forceinline__ long long tq_globaltimer() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TQP_DECL long long tqp_t_ = clock64(); long long tqp_a_[6] = {0, 0, 0, 0, 0, 0};
#define TQP(i) do { const long long t_ = clock64(); tqp_a_[i] += t_ - tqp_t_; tqp_t_ = t_; } while (0)
#define TQP_FLUSH(k, base, n) do { if (blockIdx.x < 148) for (int i_ = 0; i_ < (n); ++i_) TQ_PROF_ARRAY[k][blockIdx.x][(base) + i_] = tqp_a_[i_]; } while (0)
#else
#define TQP_DECL
#define TQP(i)
#define TQP_FLUSH(k, base, n)
#endif

namespace apg {
namespace tq {

using tc::F0; using tc::H; using tc::RD; using tc::NC; using tc::NPOS; using tc::MO; using tc::K1; using tc::REFW;
using tc::TMT;

constexpr int NPANEL = TMT / 32;
// ---- forward stash sets (row offsets inside the tile block, rows)
constexpr int R_XS = 16;          // in_state (15) + a row of ones (live drones)            A of dW states_in
constexpr int R_WIN = 40;         // per position pair g: in_ref rows 2g..2g+3 (36) + ones + 3 zero   A of dW conv
constexpr int R_X1 = K1;          // 224: s (64) | conv outputs, position-major (64 + 20 t + c)
constexpr int R_H = HID;          // 64
constexpr int R_ACT = MO;         // 40 sigmoid outputs
constexpr int O_XS = 0, O_WIN = O_XS + R_XS, O_X1 = O_WIN + 4 * R_WIN, O_H1 = O_X1 + R_X1, O_H2 = O_H1 + R_H,
              O_H3 = O_H2 + R_H, O_ACT = O_H3 + R_H, F_ROWS = O_ACT + R_ACT;                        // 632
// ---- dZ stash sets
constexpr int O_ZO = 0, O_Z3 = O_ZO + MO, O_Z2 = O_Z3 + HID, O_Z1 = O_Z2 + HID, O_ZX = O_Z1 + HID,
              Z_ROWS = O_ZX + K1;                                                                   // 456
constexpr size_t ROW_BYTES = (size_t)TMT * 4;                  // bytes of one row over the whole tile (4 panels)
constexpr size_t F_TILE_BYTES = (size_t)F_ROWS * ROW_BYTES;    // 323,584
constexpr size_t Z_TILE_BYTES = (size_t)Z_ROWS * ROW_BYTES;    // 233,472

// byte offset of (row, drone-in-tile d) inside a set with R rows
APG_HD constexpr uint32_t set_off(int R, int row, int d) {
  return (uint32_t)((d >> 5) * (R * 128) + row * 128 + ((((d & 31) >> 2) ^ (row & 7)) << 4) + (d & 3) * 4);
}
// byte offset of a set (first row `o_rows` of the tile block)
APG_HD constexpr size_t set_base(int o_rows) { return (size_t)o_rows * ROW_BYTES; }

// ---- shared-memory descriptor of a 128B-swizzled K-major operand panel (32 k = 128 bytes per row; 8-row groups
//      1024 bytes apart); k-step ks in 0..3 advances the start address by 32 bytes (cute: Layout_K_SW128_Atom)
APG_HD uint64_t sw128_desc(uint32_t base, int ks) {
  uint64_t d = 0;
  d |= (uint64_t)(((base + (uint32_t)ks * 32u) >> 4) & 0x3fffu);
  d |= (uint64_t)(16u >> 4) << 16;                 // LBO: unused for swizzled K-major
  d |= (uint64_t)(1024u >> 4) << 32;               // SBO
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}

// ---- transposed weight images of the dX chain: B[n = input feature][k = output feature], K-major unswizzled
//      (hi, lo) pairs like the forward images (tc::kmajor_off / tc::kmajor_desc)
struct TImg { int off, rows, K; };
constexpr TImg T_WO{0, 64, MO};                                            // dH3 = dZo Wo      (K = 40)
constexpr TImg T_W3{T_WO.off + 2 * tc::img_bytes(64, MO), 64, 64};
constexpr TImg T_W2{T_W3.off + 2 * tc::img_bytes(64, 64), 64, 64};
constexpr int T_W1_ROWS = K1 + 8;                                          // 8 zero rows: the last N = 48 op reads them
constexpr TImg T_W1{T_W2.off + 2 * tc::img_bytes(64, 64), T_W1_ROWS, 64};  // rows = x1 rows (position-major conv)
constexpr int TBLOB_BYTES = T_W1.off + 2 * tc::img_bytes(T_W1_ROWS, 64);   // 204,800
constexpr int NUM_TIMAGES = 4;
APG_HD TImg timage_of(int i) {
  if (i == 0) return TImg{T_WO.off, T_WO.rows, T_WO.K};
  if (i == 1) return TImg{T_W3.off, T_W3.rows, T_W3.K};
  if (i == 2) return TImg{T_W2.off, T_W2.rows, T_W2.K};
  return TImg{T_W1.off, T_W1.rows, T_W1.K};
}
APG_HD float timage_value(const float* P, const HutterLayout& y, int i, int n, int k) {
  if (i == 0) return P[y.t_wo + k * HID + n];                              // fc_out.weight [40][64]
  if (i == 1) return P[y.t_w3 + k * HID + n];
  if (i == 2) return P[y.t_w2 + k * HID + n];
  return n < K1 ? P[y.t_w1 + k * K1 + dw::fc1_col_of_x1_row(n)] : 0.f;     // fc1.weight [64][224], torch column order
}
constexpr int TPAIRS_TOTAL = TBLOB_BYTES / 8;
APG_HD void pack_t_body(int e, const float* P, const HutterLayout& y, unsigned char* blob) {
  int base = 0;
  for (int i = 0; i < NUM_TIMAGES; ++i) {
    const TImg im = timage_of(i);
    const int cnt = im.rows * im.K;
    if (e < base + cnt) {
      const int q = e - base, n = q / im.K, k = q - n * im.K;
      float hi, lo;
      tc::split_hi_lo(timage_value(P, y, i, n, k), &hi, &lo);
      *reinterpret_cast<float*>(blob + im.off + tc::kmajor_off(n, k, im.K)) = hi;
      *reinterpret_cast<float*>(blob + im.off + tc::img_bytes(im.rows, im.K) + tc::kmajor_off(n, k, im.K)) = lo;
      return;
    }
    base += cnt;
  }
}

// a 40-column piece is split [0,24) / [24,40) between the two column halves of a slot's epilogue threads
constexpr int P40_SPLIT = 24;

// ---- TMEM columns of one 256-column slot
// forward (as tc_layout.cuh): D_main [0,64) | D_conv [64,112) | A_hi [112,176) | A_lo [176,240)
// dX chain:                   D [0,128) | A_hi [128,192) | A_lo [192,256)
constexpr int XC_D = 0, XC_AHI = 128, XC_ALO = 192, SLOT_COLS = 256;

// one MMA series of the dX chain: D[d_col, +N) = A[0, K) * B^T, B = rows [row0, row0 + N) of image `img`
struct XOp { int img, row0, K, N, d_col; };
// hand-off h (one commit each) issues series [xh_first(h), xh_first(h + 1))
constexpr int NXH = 5, NXS = 7;
APG_HD int xh_first(int h) { return h < 3 ? h : (h == 3 ? 3 : (h == 4 ? 5 : 7)); }
APG_HD XOp xop_of(int i) {
  if (i == 0) return {0, 0, MO, 64, 0};                       // dH3  = dZo Wo
  if (i == 1) return {1, 0, 64, 64, 0};                       // dH2  = dZ3 W3
  if (i == 2) return {2, 0, 64, 64, 0};                       // dH1  = dZ2 W2
  if (i == 3) return {3, 0, 64, 64, 0};                       // ds   = dZ1 W1[:, s block]
  if (i == 4) return {3, 64, 64, 48, 64};                     // dconv, position pair 0 (8 spill columns ignored)
  if (i == 5) return {3, 64 + 40, 64, 80, 0};                 // pairs 1 and 2
  return {3, 64 + 120, 64, 48, 80};                           // pair 3 (reads the 8 zero rows)
}

// ---- streaming weight-gradient GEMM: the op list, accumulator columns and gradient map are dw::op_of / dw::C_* /
//      dw::grad_index (adj_dw_layout.cuh); what changes is where the operands come from: stash sets, one 32-drone
//      panel per pipeline stage.
struct DwSrc { int a_set, a_row0, a_rows, a_R, ones, b_set, b_row0, b_rows, b_R; };   // *_set: first row of the set
APG_HD DwSrc dw_src(int i) {
  if (i == 0) return {O_H3, 0, 64, R_H, 64, O_ZO, 0, MO, MO};
  if (i == 1) return {O_H2, 0, 64, R_H, 64, O_Z3, 0, 64, HID};
  if (i == 2) return {O_H1, 0, 64, R_H, 64, O_Z2, 0, 64, HID};
  if (i == 3) return {O_X1, 0, 128, R_X1, -1, O_Z1, 0, 64, HID};
  if (i == 4) return {O_X1, 128, K1 - 128, R_X1, K1 - 128, O_Z1, 0, 64, HID};
  if (i == 5) return {O_XS, 0, R_XS, R_XS, -1, O_ZX, 0, 64, K1};             // the ones row is part of the set
  const int g = i - 6;
  return {O_WIN + R_WIN * g, 0, R_WIN, R_WIN, -1, O_ZX, HID + 2 * NC * g, 2 * NC, K1};
}
// raw operand ring of the dW kernel: the stage geometry follows the pass (adj_dw_layout.cuh) - pass 0 (every op but
// fc1): A <= 64 rows, B <= 64 rows -> 16 KiB stages, 12 of them; pass 1 (fc1): A <= 128 rows -> 24 KiB stages, 7.
// Operand SLOTS (3): what a unit needs besides its raw stage - 8 KiB of shared memory for the B lo image and 64 TMEM
// columns for the A panel (32 drones x (raw, lo)).
constexpr int DW_B_BYTES = 64 * 128;
APG_HD constexpr int dw_stage_bytes(int pass) { return pass == 0 ? 16384 : 24576; }
APG_HD constexpr int dw_b_offset(int pass) { return pass == 0 ? 8192 : 16384; }
APG_HD constexpr int dw_nraw(int pass) { return pass == 0 ? 12 : 7; }
constexpr int DW_NRAW_MAX = 12, DW_RAW_BYTES = 12 * 16384;
constexpr int DW_T_OFFSET = 7 * 24576;            // conv Toeplitz block: raw-ring bytes that pass 1 never touches
constexpr int DW_NSLOT = 3;
static_assert(dw_nraw(0) * dw_stage_bytes(0) <= DW_RAW_BYTES && dw_nraw(1) * dw_stage_bytes(1) <= DW_T_OFFSET, "raw ring");

}  // namespace tq
}  // namespace apg
