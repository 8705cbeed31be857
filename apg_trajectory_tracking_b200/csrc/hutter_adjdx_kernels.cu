// Split adjoint of the quadrotor concurrent rollout, first half (OPTIONAL PATH, APG_TC_DW=1; default off):
//   hutter_adj_dx_kernel = hutter_adj_kernel (hutter_kernels.cu) WITHOUT its weight-gradient GEMMs: the reverse
//   dynamics sweep and the dX chain dZ_l -> dZ_l W_l (.) act'(X_l) on the same tile engine (mma.sync 3xTF32, [out][in]
//   weights resident), and every dZ_l tile is written to HBM next to the activation it belongs to.  The second half,
//   adj_dw_tc_kernel (adj_dw_tc_kernels.cu), turns the (X_l, dZ_l) stashes into the weight gradient as ONE streaming
//   tcgen05 GEMM over the drone axis.  Rationale and budget: DESIGN.md 8.1.
#include "dyn_phase.cuh"
#include "layouts.h"
#include "rollout_args.h"
#include "tile_engine.cuh"
#include "hutter_policy.cuh"
#include "kernels.h"

namespace apg {

namespace {
constexpr int NTH_DX = NT + TM;     // 8 GEMM warps + 2 dynamics warps, as in hutter_adj_kernel
}

template <template <typename> class SysT, bool CONV>
__global__ void __launch_bounds__(NTH_DX, 1) hutter_adj_dx_kernel(const HutterLayout y, const RolloutArgs g,
                                                                  const DzStash z) {
  APG_DYNAMIC_SMEM_F32(smem);
  using Sys = SysT<float>;
  constexpr int S = Sys::S, R = Sys::REFW;
  const int wb_floats = y.b_ws;              // no first-layer dX: the policy inputs need no gradient
  float* s_w = smem;
  float* bufA = s_w + wb_floats;
  float* bufB = bufA + y.K1 * TMP;
  float* bufD = bufB + HID * TMP;
  float* bufC = bufD + HID * TMP;
  float* s_dlog = bufC + HID * TMP;          // [Mo4][TMP]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_dlog + y.Mo4 * TMP + 8);
  uint64_t *bar_w = bars, *bar_A = bars + 1, *bar_B = bars + 2, *bar_D = bars + 3, *bar_C = bars + 4,
           *dlog_full = bars + 6, *dlog_empty = bars + 7;

  const int tid = threadIdx.x;
  const int ntiles = (g.N + TM - 1) / TM;
  if (tid == 0) {
    for (int b = 0; b < 6; ++b) mbar_init(bars + b, 1);
    mbar_init(dlog_full, TM);
    mbar_init(dlog_empty, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int first = ntiles - 1 - (int)blockIdx.x;     // reverse tile order: the freshest stash is still in L2

  if (tid < NT) {
    // ===================================================== GEMM group
    const Lane L;
    const uint32_t hbytes = HID * TMP * 4;
    auto issue_stage_loads = [&](int tile) {    // thread 0
      mbar_expect_tx(bar_A, y.K1 * TMP * 4);
      bulk_g2s_chunked(bufA, g.st_x1 + (size_t)tile * y.K1 * TMP, y.K1 * TMP * 4, bar_A);
      mbar_expect_tx(bar_B, hbytes);
      bulk_g2s(bufB, g.st_h3 + (size_t)tile * HID * TMP, hbytes, bar_B);
      mbar_expect_tx(bar_D, hbytes);
      bulk_g2s(bufD, g.st_h2 + (size_t)tile * HID * TMP, hbytes, bar_D);
      mbar_expect_tx(bar_C, hbytes);
      bulk_g2s(bufC, g.st_h1 + (size_t)tile * HID * TMP, hbytes, bar_C);
    };
    if (tid == 0) {
      mbar_expect_tx(bar_w, wb_floats * 4);
      bulk_g2s_chunked(s_w, g.wb, wb_floats * 4, bar_w);
      if (first >= 0) issue_stage_loads(first);
    }
    mbar_wait(bar_w, 0);
    uint32_t ph = 0;
    for (int tile = first; tile >= 0; tile -= gridDim.x) {
      // ---- d loss / d logits of this tile (dZ of fc_out) comes from the dynamics warps (generic-proxy writes)
      mbar_wait(dlog_full, ph);
      fence_proxy_async();
      gsync<true>();
      if (tid == 0) {
        bulk_s2g(z.o + (size_t)tile * y.Mo4 * TMP, s_dlog, y.Mo4 * TMP * 4);
        bulk_commit();
      }
      // ---- fc_out: dz3 = (dlog Wo) (.) (1 - h3^2), in place over h3
      mbar_wait(bar_B, ph);
      dense_auto<EPI_DTANH>(L, s_dlog, y.Mo, s_w + y.b_wo, HID, mma_sw(HID), nullptr, HID, bufB, 0, 0);
      fence_proxy_async();
      gsync<true>();
      if (tid == 0) {
        bulk_wait_read<0>();                     // the dlog store has finished reading s_dlog
        mbar_arrive(dlog_empty);                 // the dynamics warps may sweep the next tile into it
        bulk_s2g(z.z3 + (size_t)tile * HID * TMP, bufB, hbytes);
        bulk_commit();
      }
      // ---- fc3: dz2 over h2
      mbar_wait(bar_D, ph);
      dense_auto<EPI_DTANH>(L, bufB, HID, s_w + y.b_w3, HID, mma_sw(HID), nullptr, HID, bufD, 0, 0);
      fence_proxy_async();
      gsync<true>();
      if (tid == 0) {
        bulk_s2g(z.z2 + (size_t)tile * HID * TMP, bufD, hbytes);
        bulk_commit();
      }
      // ---- fc2: dz1 over h1
      mbar_wait(bar_C, ph);
      dense_auto<EPI_DTANH>(L, bufD, HID, s_w + y.b_w2, HID, mma_sw(HID), nullptr, HID, bufC, 0, 0);
      fence_proxy_async();
      gsync<true>();
      if (tid == 0) {
        bulk_s2g(z.z1 + (size_t)tile * HID * TMP, bufC, hbytes);
        bulk_commit();
      }
      // ---- fc1: d(pre-activation) of the first layer over X1 = [s | r]
      mbar_wait(bar_A, ph);
      dense_auto<EPI_DTANH>(L, bufC, HID, s_w + y.b_w1, y.ld_bw1, mma_sw(y.ld_bw1), nullptr, HID, bufA, 0, 0);
      if (CONV)
        dense_auto<EPI_DRELU>(L, bufC, HID, s_w + y.b_w1, y.ld_bw1, mma_sw(y.ld_bw1), nullptr, y.NRtot, bufA, HID, 0, HID);
      else
        dense_auto<EPI_DTANH>(L, bufC, HID, s_w + y.b_w1, y.ld_bw1, mma_sw(y.ld_bw1), nullptr, y.NRtot, bufA, HID, 0, HID);
      fence_proxy_async();
      gsync<true>();
      ph ^= 1;
      if (tid == 0) {
        bulk_s2g(z.x + (size_t)tile * y.K1 * TMP, bufA, y.K1 * TMP * 4);
        bulk_commit();
        bulk_wait_read<0>();                     // every dZ store has finished reading its buffer
        const int next = tile - gridDim.x;
        if (next >= 0) issue_stage_loads(next);
      }
    }
    if (tid == 0) bulk_wait_all();
  } else {
    // ===================================================== dynamics warps: reverse sweep, one thread per drone
    const int d = tid - NT;
    uint32_t it = 0;
    for (int tile = first; tile >= 0; tile -= gridDim.x, ++it) {
      const int valid = min(TM, g.N - tile * TM);
      if (it > 0) mbar_wait(dlog_empty, (it - 1) & 1);
      if (d < valid) {
        const size_t drone = (size_t)tile * TM + d;
        dyn_adjoint_conc<SysT>(g.st_act + (size_t)tile * y.Mo4 * TMP, g.st_states + (size_t)tile * g.h * S * TMP, d,
                               g.cur + drone * S, g.ref + drone * g.ref_rows * R, g.h, g.dt, g.pc.v, s_dlog);
      } else {
        for (int r = 0; r < y.Mo4; ++r) s_dlog[r * TMP + d] = 0.f;
      }
      mbar_arrive(dlog_full);
    }
  }
}

size_t hutter_adj_dx_smem_bytes(const HutterLayout& y) {
  return sizeof(float) * (size_t)(y.b_ws + y.K1 * TMP + 3 * HID * TMP + y.Mo4 * TMP + 8) + 80;
}

cudaError_t launch_hutter_adj_dx(int system, const HutterLayout& y, const RolloutArgs& a, const DzStash& z, int grid,
                                 cudaStream_t st) {
  if (!(system == SYS_QUAD && y.conv)) return cudaErrorInvalidValue;
  const size_t smem = hutter_adj_dx_smem_bytes(y);
  cudaError_t e = cudaFuncSetAttribute(hutter_adj_dx_kernel<Quad, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return e;
  APG_LAUNCH(grid, NTH_DX, smem, st, hutter_adj_dx_kernel<Quad, true>)(y, a, z);
  return cudaGetLastError();
}


}  // namespace apg
