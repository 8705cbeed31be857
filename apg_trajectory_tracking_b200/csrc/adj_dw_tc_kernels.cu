// Split adjoint of the quadrotor concurrent rollout, second half (OPTIONAL PATH, APG_TC_DW=1; default off):
// the weight gradient of every used tensor as ONE streaming tcgen05 GEMM over the drone axis,
//     dW_l[out][in] = sum over drones dZ_l[drone][out] * X_l[drone][in],
// with the accumulators of all layers resident in TMEM (416 columns) for the whole launch.  Per 64-drone stash tile
// and layer the loader warps turn the feature-major tiles [rows][TMP] of X_l and dZ_l (HBM, written by the forward
// kernel and by hutter_adj_dx_kernel) into K-major (hi, lo) shared-memory images; one thread issues 3xTF32
// tcgen05.mma (M = 128 input features incl. a row of ones for the bias, N = outputs, K = 64 drones); two stages of
// 96 KiB.  No epilogue until the CTA has streamed all its tiles; then the accumulators are mapped to the torch-flat
// gradient partial of the CTA (fc1 column permutation, Toeplitz fold of the conv block; every entry written once ->
// the fixed-order reduction over CTAs stays bitwise reproducible).  The kernel is HBM-bound by construction:
// 3.6 KB per drone.  Layout / index arithmetic: adj_dw_layout.cuh (host-checked).
#include "adj_dw_layout.cuh"
#include "tc_prims.cuh"
#include "rollout_args.h"
#ifndef APG_TC_SIM
#include "tile_engine.cuh"
#endif
#include "kernels.h"

namespace apg {

using namespace dw;

namespace {

constexpr int DW_LOADERS = 256;
constexpr int DW_THREADS = DW_LOADERS + 32;
constexpr int DW_SMEM_BYTES = 1024 + NSTAGE * STAGE_BYTES;

__device__ __forceinline__ void dw_mma_ss(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  tcp::mma_ss(d_tmem, a, b, idesc, acc);
}
__device__ __forceinline__ void dw_commit(uint32_t bar) { tcp::commit(bar); }
__device__ __forceinline__ void dw_mbar_init(uint32_t bar, int count) { tcp::mbar_init(bar, count); }
__device__ __forceinline__ void dw_mbar_arrive(uint32_t bar) { tcp::mbar_arrive(bar); }
// clock-bounded wait (a protocol error ends the launch with a NaN gradient instead of hanging the GPU)
__device__ __forceinline__ void dw_mbar_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag) {
  const long long t0 = tcp::clock_now();
  for (int spin = 0;; ++spin) {
    if (tcp::mbar_try_wait(bar, parity)) return;
    if ((spin & 63) == 63) {
      if (*abort_flag) return;
      if (tcp::clock_now() - t0 > 2000000000LL) { *abort_flag = 1; return; }
    }
  }
}
__device__ __forceinline__ void dw_tmem_ld8(uint32_t addr, float* v) {
  uint32_t r[8];
  tcp::tmem_ld8(addr, r);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}
// (hi, lo) split of four values, one 16-byte chunk into each image
__device__ __forceinline__ void store_split4(unsigned char* img_hi, unsigned char* img_lo, uint32_t off, float4 x) {
  float4 h, l;
  h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); l.x = x.x - h.x;
  h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u); l.y = x.y - h.y;
  h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); l.z = x.z - h.z;
  h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u); l.w = x.w - h.w;
  *reinterpret_cast<float4*>(img_hi + off) = h;
  *reinterpret_cast<float4*>(img_lo + off) = l;
}

struct DwBars {
  unsigned long long full[NSTAGE];
  unsigned long long empty[NSTAGE];
  unsigned long long done;
};

}  // namespace

__global__ void __launch_bounds__(DW_THREADS, 1)
    adj_dw_tc_kernel(const HutterLayout y, const RolloutArgs g, const DzStash z) {
  APG_TC_DYNAMIC_SMEM(smem_raw);
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) DwBars s_bars;
  __shared__ uint32_t s_tmem;
  __shared__ int s_abort;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = g.N;
  const int ntiles = (n + TM - 1) / TM;
  const int my_tiles = (ntiles > (int)blockIdx.x) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  float* P = g.grad_partials + (size_t)blockIdx.x * y.n_params;
  volatile int* abort_flag = &s_abort;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      dw_mbar_init(smem_u32(&s_bars.full[s]), DW_LOADERS);
      dw_mbar_init(smem_u32(&s_bars.empty[s]), 1);
    }
    dw_mbar_init(smem_u32(&s_bars.done), 1);
    s_abort = 0;
    tcp::fence_mbar_init();
  }
  if (warp == 8) {
    tcp::tmem_alloc512(&s_tmem);
  }
  tcp::fence_before_thread_sync();
  __syncthreads();
  tcp::fence_after_thread_sync();
  const uint32_t tmem = s_tmem;

  if (warp == 8) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      uint32_t full_par[NSTAGE] = {0, 0};
      int it = 0;
      for (int j = 0; j < my_tiles; ++j)
        for (int i = 0; i < NOPS; ++i, ++it) {
          const int s = it % NSTAGE;
          const Op op = op_of(i);
          dw_mbar_wait(smem_u32(&s_bars.full[s]), full_par[s], abort_flag);
          full_par[s] ^= 1;
          tcp::fence_after_thread_sync();
          unsigned char* st = base + s * STAGE_BYTES;
          const uint32_t a_hi = smem_u32(st), a_lo = a_hi + A_IMG_BYTES, b_hi = a_lo + A_IMG_BYTES,
                         b_lo = b_hi + B_IMG_BYTES;
          const uint32_t idesc = tc::idesc_tf32(AM, op.N);
          const uint32_t d = tmem + op.d_col;
          const bool clear = (j == 0) && op.first;
#pragma unroll
          for (int ks = 0; ks < KD / 8; ++ks) {
            const uint64_t ah = tc::kmajor_desc(a_hi, ks, KD), al = tc::kmajor_desc(a_lo, ks, KD);
            const uint64_t bh = tc::kmajor_desc(b_hi, ks, KD), bl = tc::kmajor_desc(b_lo, ks, KD);
            dw_mma_ss(d, al, bh, idesc, (ks > 0 || !clear) ? 1u : 0u);
            dw_mma_ss(d, ah, bl, idesc, 1u);
            dw_mma_ss(d, ah, bh, idesc, 1u);
          }
          dw_commit(smem_u32(&s_bars.empty[s]));        // the stage is free once these MMAs have read it
        }
      dw_commit(smem_u32(&s_bars.done));                // all accumulators final
    }
  } else {
    // ===================================================== loader warps: stash tiles -> (hi, lo) K-major images
    uint32_t empty_par[NSTAGE] = {0, 0};
    int it = 0;
    Sources S;
    S.h3 = g.st_h3; S.h2 = g.st_h2; S.h1 = g.st_h1; S.x1 = g.st_x1; S.in_state = g.in_state; S.in_ref = g.in_ref;
    S.dzo = z.o; S.dz3 = z.z3; S.dz2 = z.z2; S.dz1 = z.z1; S.dzx = z.x;
    for (int j = 0; j < my_tiles; ++j) {
      const int tile = (int)blockIdx.x + j * (int)gridDim.x;
      const int valid = min(TM, n - tile * TM);
      for (int i = 0; i < NOPS; ++i, ++it) {
        const int s = it % NSTAGE;
        const Op op = op_of(i);
        if (it >= NSTAGE) {                               // wait until the MMAs of the previous use have drained
          dw_mbar_wait(smem_u32(&s_bars.empty[s]), empty_par[s], abort_flag);
          empty_par[s] ^= 1;
        }
        unsigned char* st = base + s * STAGE_BYTES;
        unsigned char *a_hi = st, *a_lo = st + A_IMG_BYTES, *b_hi = st + 2 * A_IMG_BYTES,
                      *b_lo = st + 2 * A_IMG_BYTES + B_IMG_BYTES;
        // ---- A image (AM rows) and B image (dZ rows, zero up to 64): 16-byte chunks of four drones, (hi, lo) split
        for (int q = tid; q < AM * (KD / 4); q += DW_LOADERS) {
          int r, d4;
          float x[4];
          chunk_of_item(q, &r, &d4);
          a_chunk(op, S, tile, valid, r, d4, x);
          store_split4(a_hi, a_lo, chunk_off(r, d4), make_float4(x[0], x[1], x[2], x[3]));
        }
        for (int q = tid; q < B_ROWS * (KD / 4); q += DW_LOADERS) {
          int r, d4;
          float x[4];
          chunk_of_item(q, &r, &d4);
          b_chunk(op, S, tile, r, d4, x);
          store_split4(b_hi, b_lo, chunk_off(r, d4), make_float4(x[0], x[1], x[2], x[3]));
        }
        tcp::fence_proxy_async_smem();       // generic writes -> tensor core reads
        dw_mbar_arrive(smem_u32(&s_bars.full[s]));
      }
    }
  }

  // ===================================================== epilogue: accumulators -> this CTA's gradient partial
  // unused tensors (ref_in.*) and padding stay zero; every entry of the partial is written exactly once
  for (int i = tid; i < y.n_params; i += DW_THREADS) {
    const bool conv_w = i >= y.t_wc && i < y.t_wc + NC * RD * 3 + NC;    // conv_ref weight + bias: written below
    if (!conv_w && (my_tiles == 0 || (i >= y.t_wr && i < y.t_br + HID))) P[i] = 0.f;
  }
  if (my_tiles > 0 && warp < 4) {
    dw_mbar_wait(smem_u32(&s_bars.done), 0, abort_flag);
    tcp::fence_after_thread_sync();
    const int r = warp * 32 + lane;                            // TMEM lane = A row
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    const float poison = __int_as_float(0x7fc00000);
    float* s_T = reinterpret_cast<float*>(base);               // [37][48] conv Toeplitz block (stage memory is free)
    const int col0[6] = {C_WO, C_W3, C_W2, C_W1A, C_W1B, C_WS};
    const int ncol[6] = {48, 64, 64, 64, 64, 64};
#pragma unroll
    for (int reg = 0; reg < 6; ++reg) {
      for (int c0 = 0; c0 < ncol[reg]; c0 += 8) {
        float v[8];
        dw_tmem_ld8(lane_addr + col0[reg] + c0, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int idx = grad_index(y, reg, r, c0 + q);
          if (idx >= 0) P[idx] = *abort_flag ? poison : v[q];
        }
      }
    }
    for (int c0 = 0; c0 < 48; c0 += 8) {
      float v[8];
      dw_tmem_ld8(lane_addr + C_WT + c0, v);
      if (r <= 4 * RD) {
#pragma unroll
        for (int q = 0; q < 8; ++q) s_T[r * 48 + c0 + q] = v[q];
      }
    }
  }
  tcp::fence_before_thread_sync();
  __syncthreads();
  if (my_tiles > 0) {
    const float* s_T = reinterpret_cast<const float*>(base);
    for (int i = tid; i < NC * RD * 3 + NC; i += DW_THREADS) {
      float v;
      if (i < NC * RD * 3) {
        const int c = i / (RD * 3), ci = (i / 3) % RD, jj = i % 3;
        v = conv_weight_from_block(s_T, 48, c, ci, jj);
      } else {
        v = conv_bias_from_block(s_T, 48, i - NC * RD * 3);
      }
      P[y.t_wc + i] = *abort_flag ? __int_as_float(0x7fc00000) : v;
    }
  } else {
    for (int i = tid; i < NC * RD * 3 + NC; i += DW_THREADS) P[y.t_wc + i] = 0.f;
  }
  if (warp == 8) {
    tcp::tmem_dealloc512(tmem);
  }
}

// grad[p] = scale * sum over CTAs of partials[c][p]: 32 parameters x 4 CTA slices per block of 128 threads
__global__ void __launch_bounds__(128) apg_reduce4_kernel(const float* __restrict__ partials, int ncta, int n,
                                                          float scale, float* __restrict__ grad) {
  __shared__ float s_part[RED_SLICES][32];
  const int pl = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int p = blockIdx.x * 32 + pl;
  int c0, c1;
  reduce_slice_bounds(ncta, slice, &c0, &c1);
  s_part[slice][pl] = p < n ? reduce_slice_sum(partials, n, p, c0, c1) : 0.f;
  __syncthreads();
  if (slice == 0 && p < n) grad[p] = scale * ((s_part[0][pl] + s_part[1][pl]) + (s_part[2][pl] + s_part[3][pl]));
}

cudaError_t launch_reduce_grad4(const float* partials, int ncta, int n, float scale, float* grad, cudaStream_t st) {
  APG_LAUNCH((n + 31) / 32, 128, 0, st, apg_reduce4_kernel)(partials, ncta, n, scale, grad);
  return cudaGetLastError();
}

bool adj_dw_tc_supported(const HutterLayout& y, int h) {
  return y.conv && y.F0 == F0 && y.L == H && y.RD == RD && y.Mo == MO && h == H;
}

cudaError_t launch_adj_dw_tc(const HutterLayout& y, const RolloutArgs& a, const DzStash& z, int grid, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(adj_dw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  APG_LAUNCH(grid, DW_THREADS, DW_SMEM_BYTES, st, adj_dw_tc_kernel)(y, a, z);
  return cudaGetLastError();
}


}  // namespace apg
