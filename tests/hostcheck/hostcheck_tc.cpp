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
