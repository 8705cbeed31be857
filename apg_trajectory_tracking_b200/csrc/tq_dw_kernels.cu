// Weight gradient of the quadrotor concurrent rollout as ONE streaming tcgen05 GEMM over the drone axis,
//     dW_l[out][in] = sum over drones dZ_l[drone][out] * X_l[drone][in]          (loss.backward() of train_base.py:205),
// on the two operand-image stashes written by tq_fwd_kernel (X_l) and tq_dx_kernel (dZ_l): tq_layout.cuh.
// Accumulators of all layers resident in TMEM (416 columns) for the whole launch: D[M = in (+ ones row -> bias)][N = out].
//
// Pipeline (unit = one 32-drone panel of one op; a ring of 7 RAW stages of 24 KiB that the bulk copies land in, and a
// ring of 2 LO stages that only live from conversion to the MMAs - the copies in flight are what hides the HBM latency,
// so the raw ring is deep and the lo images do not take shared memory away from it):
//   warp 0 lane 0   producer  : two 1-D bulk copies per unit (A panel, B panel) straight from the stash - the stash IS
//                               the 128B-swizzled K-major image, so there is no loader arithmetic and no tensor map
//   warps 2-9       converters: lo image = x - tf32(x) of each raw image (the tensor core truncates the raw fp32
//                               image itself: that is the hi part), the constant ones rows of the bias gradients
//   warp 1 lane 0   MMA issuer: per unit 4 k-steps x 3 MMAs (3xTF32), two tcgen05.commit free the raw and the lo stage
// HBM-bound by construction: 4.2 KB per drone.  Op list / accumulator columns / gradient map: adj_dw_layout.cuh.
#include "tq_layout.cuh"
#include "tc_prims.cuh"
#include "rollout_args.h"
#ifndef APG_TC_SIM
#include "tile_engine.cuh"
#endif
#include "kernels.h"

#ifdef APG_PROFILE
__device__ long long g_tq_prof_dw[3][148][TQ_NPROF];
#define TQ_PROF_ARRAY g_tq_prof_dw
extern "C" __attribute__((visibility("default"))) int apg_debug_profile_tq_dw(long long* out_host) {
  return (int)cudaMemcpyFromSymbol(out_host, g_tq_prof_dw, sizeof(long long) * 3 * 148 * TQ_NPROF);
}
#endif

namespace apg {

namespace {

constexpr int DWQ_CONV = 256;                                 // converter threads (warps 2..9)
constexpr int DWQ_THREADS = 64 + DWQ_CONV;
constexpr int NR = tq::DW_NRAW, NL = tq::DW_NLO;
constexpr int STAGE = tq::DW_A_BYTES + tq::DW_B_BYTES;        // one raw or lo stage: A panel (128 rows) | B panel (64 rows)
constexpr int DWQ_SMEM = 1024 + (NR + NL) * STAGE;
static_assert(DWQ_SMEM <= 232448, "stage rings do not fit in shared memory");

struct DwqBars {
  unsigned long long full[NR];                   // bulk copies landed (1 arrival + bytes)
  unsigned long long rfree[NR];                  // MMAs that read the raw stage are complete (tcgen05.commit)
  unsigned long long lo_ready[NL];               // lo images written (256 arrivals)
  unsigned long long lo_free[NL];                // MMAs that read the lo stage are complete (tcgen05.commit)
  unsigned long long done;
};

__device__ __forceinline__ void dwq_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag) {
  if (tcp::mbar_try_wait(bar, parity)) return;
  const long long t0 = tcp::clock_now();
  for (int spin = 0;; ++spin) {
    if (tcp::mbar_try_wait(bar, parity)) return;
    if ((spin & 63) == 63) {
      if (*abort_flag) return;
      if (tcp::clock_now() - t0 > 2000000000LL) {
#ifdef APG_TC_SIM
        if (getenv("APG_SIM_FAST_TIMEOUT")) fprintf(stderr, "dwq_wait timeout: thread %u bar %x parity %u\n", threadIdx.x, bar, parity);
#endif
        *abort_flag = 1;
        return;
      }
    }
  }
}
__device__ __forceinline__ float4 lo_of(float4 x) {
  float4 l;
  l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
  l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
  l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
  l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
  return l;
}

}  // namespace

__global__ void __launch_bounds__(DWQ_THREADS, 1)
    tq_dw_kernel(const HutterLayout y, const RolloutArgs g, const unsigned char* __restrict__ fstash,
                 const unsigned char* __restrict__ zstash) {
  APG_TC_DYNAMIC_SMEM(smem_raw);
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) DwqBars s_bars;
  __shared__ uint32_t s_tmem;
  __shared__ int s_abort;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef APG_PROFILE
  const long long tqp_k0_ = clock64();
  long long tqp_k1_ = 0, tqp_k2_ = 0;
#endif
  const int n = g.N;
  const int ntiles = (n + tc::TMT - 1) / tc::TMT;
  const int my_tiles = (ntiles > (int)blockIdx.x) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  float* P = g.grad_partials + (size_t)blockIdx.x * y.n_params;
  volatile int* abort_flag = &s_abort;
  unsigned char* lo_base = base + NR * STAGE;

  if (tid == 0) {
    for (int s = 0; s < NR; ++s) {
      tcp::mbar_init(smem_u32(&s_bars.full[s]), 1);
      tcp::mbar_init(smem_u32(&s_bars.rfree[s]), 1);
    }
    for (int s = 0; s < NL; ++s) {
      tcp::mbar_init(smem_u32(&s_bars.lo_ready[s]), DWQ_CONV);
      tcp::mbar_init(smem_u32(&s_bars.lo_free[s]), 1);
    }
    tcp::mbar_init(smem_u32(&s_bars.done), 1);
    s_abort = 0;
    tcp::fence_mbar_init();
  }
  if (warp == 0) tcp::tmem_alloc512(&s_tmem);
  tcp::fence_before_thread_sync();
  __syncthreads();
  tcp::fence_after_thread_sync();
  const uint32_t tmem = s_tmem;
#ifdef APG_PROFILE
  tqp_k1_ = clock64();
#endif

  if (warp == 0) {
    // ===================================================== producer (warp-uniform loop, the elected lane issues)
    {
      const bool leader = tcp::elect_one();
      TQP_DECL
      int u = 0;
      for (int j = 0; j < my_tiles; ++j) {
        const int tile = (int)blockIdx.x + j * (int)gridDim.x;
        const unsigned char* fb = fstash + (size_t)tile * tq::F_TILE_BYTES;
        const unsigned char* zb = zstash + (size_t)tile * tq::Z_TILE_BYTES;
        for (int i = 0; i < dw::NOPS; ++i) {
          const tq::DwSrc src = tq::dw_src(i);
          const uint32_t a_bytes = (uint32_t)src.a_rows * 128u, b_bytes = (uint32_t)src.b_rows * 128u;
          for (int p = 0; p < tq::NPANEL; ++p, ++u) {
            const int r = u % NR;
            if (u >= NR) dwq_wait(smem_u32(&s_bars.rfree[r]), (uint32_t)(u / NR - 1) & 1u, abort_flag);
            TQP(0);
            unsigned char* st = base + r * STAGE;
            const uint32_t bar = smem_u32(&s_bars.full[r]);
            if (leader) {
            tcp::mbar_expect_tx(bar, a_bytes + b_bytes);
            tcp::bulk_g2s(smem_u32(st), fb + tq::set_base(src.a_set) + (size_t)p * (size_t)(src.a_R * 128) +
                                            (size_t)src.a_row0 * 128, a_bytes, bar);
            tcp::bulk_g2s(smem_u32(st + tq::DW_A_BYTES), zb + tq::set_base(src.b_set) +
                                                             (size_t)p * (size_t)(src.b_R * 128) +
                                                             (size_t)src.b_row0 * 128, b_bytes, bar);
            }
            __syncwarp();          // lanes stay within one unit of each other (parity waits alias with period 2)
            TQP(1);
          }
        }
      }
      if (leader) TQP_FLUSH(2, 0, 2);
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (warp-uniform loop, the elected lane issues)
    {
      const bool leader = tcp::elect_one();
      TQP_DECL
      const uint32_t raw0 = smem_u32(base), lo0 = smem_u32(lo_base);
      const uint64_t DESC_HI = ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29))) << 32;
      int u = 0;
      for (int j = 0; j < my_tiles; ++j)
        for (int i = 0; i < dw::NOPS; ++i) {
          const dw::Op op = dw::op_of(i);
          const uint32_t idesc = tc::idesc_tf32(128, op.N);
          const uint32_t d = tmem + op.d_col;
          for (int p = 0; p < tq::NPANEL; ++p, ++u) {
            const int r = u % NR, l = u % NL;
            // only the issuing lane polls: lo_ready[l] can complete AGAIN (unit u + NL) as soon as this unit's
            // MMAs are committed, so a lane that looked late would see the parity it is waiting for already gone
            if (leader) dwq_wait(smem_u32(&s_bars.lo_ready[l]), (uint32_t)(u / NL) & 1u, abort_flag);
            __syncwarp();
            TQP(0);
            tcp::fence_after_thread_sync();
            // descriptors: constant high word (SBO 1024, version, SWIZZLE_128B), low word = address >> 4 | LBO field;
            // k-step ks adds 2 (32 bytes >> 4)
            uint32_t ar = ((raw0 + (uint32_t)r * STAGE) >> 4) | (1u << 16), br = ar + (tq::DW_A_BYTES >> 4);
            uint32_t al = ((lo0 + (uint32_t)l * STAGE) >> 4) | (1u << 16), bl = al + (tq::DW_A_BYTES >> 4);
            const bool clear = (j == 0) && op.first && (p == 0);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks, ar += 2, br += 2, al += 2, bl += 2) {
              const uint64_t dah = DESC_HI | ar, dal = DESC_HI | al, dbh = DESC_HI | br, dbl = DESC_HI | bl;
              if (leader) {
                tcp::mma_ss(d, dal, dbh, idesc, (ks > 0 || !clear) ? 1u : 0u);
                tcp::mma_ss(d, dah, dbl, idesc, 1u);
                tcp::mma_ss(d, dah, dbh, idesc, 1u);
              }
            }
            if (leader) {
              tcp::commit(smem_u32(&s_bars.rfree[r]));    // both stages are free once these MMAs have read them
              tcp::commit(smem_u32(&s_bars.lo_free[l]));
            }
            __syncwarp();          // lanes stay within one unit of each other (parity waits alias with period 2)
            TQP(1);
          }
        }
      if (leader) TQP_FLUSH(2, 2, 2);
      if (leader) tcp::commit(smem_u32(&s_bars.done));    // all accumulators final
    }
  } else {
    // ===================================================== converters: lo images + constant ones rows
    const int ct = tid - 64;
    TQP_DECL
    int u = 0;
    for (int j = 0; j < my_tiles; ++j)
      for (int i = 0; i < dw::NOPS; ++i) {
        const tq::DwSrc src = tq::dw_src(i);
        const int na = src.a_rows * 8, nb = src.b_rows * 8;  // 16-byte chunks
        for (int p = 0; p < tq::NPANEL; ++p, ++u) {
          const int r = u % NR, l = u % NL;
          dwq_wait(smem_u32(&s_bars.full[r]), (uint32_t)(u / NR) & 1u, abort_flag);
          TQP(0);
          if (u >= NL) dwq_wait(smem_u32(&s_bars.lo_free[l]), (uint32_t)(u / NL - 1) & 1u, abort_flag);
          TQP(1);
          float4* a_raw = reinterpret_cast<float4*>(base + r * STAGE);
          float4* b_raw = reinterpret_cast<float4*>(base + r * STAGE + tq::DW_A_BYTES);
          float4* a_lo = reinterpret_cast<float4*>(lo_base + l * STAGE);
          float4* b_lo = reinterpret_cast<float4*>(lo_base + l * STAGE + tq::DW_A_BYTES);
          {
            // all loads first (up to 4 + 2 chunks of 16 bytes per thread), then split and store: one exposed
            // shared-memory latency per unit instead of six
            float4 va[4], vb[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) if (ct + i * DWQ_CONV < na) va[i] = a_raw[ct + i * DWQ_CONV];
#pragma unroll
            for (int i = 0; i < 2; ++i) if (ct + i * DWQ_CONV < nb) vb[i] = b_raw[ct + i * DWQ_CONV];
#pragma unroll
            for (int i = 0; i < 4; ++i) if (ct + i * DWQ_CONV < na) a_lo[ct + i * DWQ_CONV] = lo_of(va[i]);
#pragma unroll
            for (int i = 0; i < 2; ++i) if (ct + i * DWQ_CONV < nb) b_lo[ct + i * DWQ_CONV] = lo_of(vb[i]);
          }
          if (src.ones >= 0 && ct < 64) {                    // 8-row group [ones, ones + 8): row `ones` = 1, rest 0
            const float v = (ct >> 3) == 0 ? 1.f : 0.f;
            a_raw[src.ones * 8 + ct] = make_float4(v, v, v, v);
            a_lo[src.ones * 8 + ct] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          tcp::fence_proxy_async_smem();                     // generic writes -> tensor core reads
          tcp::mbar_arrive(smem_u32(&s_bars.lo_ready[l]));
          TQP(2);
        }
      }
    if (ct == 0) TQP_FLUSH(2, 4, 3);
  }
#ifdef APG_PROFILE
  tqp_k2_ = clock64();
#endif

  // ===================================================== epilogue: accumulators -> this CTA's gradient partial
  // unused tensors (ref_in.*) and padding stay zero; every entry of the partial is written exactly once
  for (int i = tid; i < y.n_params; i += DWQ_THREADS) {
    const bool conv_w = i >= y.t_wc && i < y.t_wc + tc::NC * tc::RD * 3 + tc::NC;    // conv_ref weight + bias: below
    if (!conv_w && (my_tiles == 0 || (i >= y.t_wr && i < y.t_br + HID))) P[i] = 0.f;
  }
  if (my_tiles > 0 && warp >= 2) {
    // the eight converter warps: warp w reads TMEM lanes 32 (w & 3) .. +31 (= A rows) and one half of every region's
    // columns.  The row -> gradient map (base + column * stride) is worked out ONCE per thread and region - measured:
    // with the index arithmetic (divisions by 20, the switch of grad_index) inside the element loop on four warps this
    // epilogue took 73 k cycles, a third of the kernel.
    dwq_wait(smem_u32(&s_bars.done), 0, abort_flag);
    tcp::fence_after_thread_sync();
    const int r = (warp & 3) * 32 + lane, half = (warp - 2) >> 2;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float* s_T = reinterpret_cast<float*>(base);               // [37][48] conv Toeplitz block (stage memory is free)
    const int col0[6] = {dw::C_WO, dw::C_W3, dw::C_W2, dw::C_W1A, dw::C_W1B, dw::C_WS};
    const int ncol[6] = {48, 64, 64, 64, 64, 64};
#pragma unroll
    for (int reg = 0; reg < 6; ++reg) {
      // entry (r, n) -> P[i0 + n * stride]; i0 < 0: this row is padding in this region
      const int i0 = dw::grad_index(y, reg, r, 0);
      const int stride = i0 < 0 ? 0 : dw::grad_index(y, reg, r, 1) - i0;
      const int nvalid = reg == 0 ? tc::MO : ncol[reg];
      const int c_lo = half * (ncol[reg] / 2), c_hi = c_lo + ncol[reg] / 2;
      for (int c0 = c_lo; c0 < c_hi; c0 += 8) {
        uint32_t vb[8];
        tcp::tmem_ld8(lane_addr + col0[reg] + c0, vb);
        if (i0 >= 0) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (c0 + q < nvalid) P[i0 + (c0 + q) * stride] = __uint_as_float(vb[q]);
        }
      }
    }
    for (int c0 = half * 24; c0 < half * 24 + 24; c0 += 8) {
      uint32_t vb[8];
      tcp::tmem_ld8(lane_addr + dw::C_WT + c0, vb);
      if (r <= 4 * tc::RD) {
#pragma unroll
        for (int q = 0; q < 8; ++q) s_T[r * 48 + c0 + q] = __uint_as_float(vb[q]);
      }
    }
  }
  tcp::fence_before_thread_sync();
  __syncthreads();
  if (my_tiles > 0) {
    const float* s_T = reinterpret_cast<const float*>(base);
    for (int i = tid; i < tc::NC * tc::RD * 3 + tc::NC; i += DWQ_THREADS) {
      float v;
      if (i < tc::NC * tc::RD * 3) {
        const int c = i / (tc::RD * 3), ci = (i / 3) % tc::RD, jj = i % 3;
        v = dw::conv_weight_from_block(s_T, 48, c, ci, jj);
      } else {
        v = dw::conv_bias_from_block(s_T, 48, i - tc::NC * tc::RD * 3);
      }
      P[y.t_wc + i] = v;
    }
    if (tid == 0 && *abort_flag) P[0] = __int_as_float(0x7fc00000);       // protocol timeout: poison the gradient
  } else {
    for (int i = tid; i < tc::NC * tc::RD * 3 + tc::NC; i += DWQ_THREADS) P[y.t_wc + i] = 0.f;
  }
  if (warp == 0) tcp::tmem_dealloc512(tmem);
#ifdef APG_PROFILE
  if (tid == 64 && blockIdx.x < 148) {                       // setup | main loop (converter 0) | epilogue
    TQ_PROF_ARRAY[2][blockIdx.x][8] = tqp_k1_ - tqp_k0_;
    TQ_PROF_ARRAY[2][blockIdx.x][9] = tqp_k2_ - tqp_k1_;
    TQ_PROF_ARRAY[2][blockIdx.x][10] = clock64() - tqp_k2_;
  }
#endif
}

// grad[p] = scale * sum over CTAs of partials[c][p]: 32 parameters x 4 CTA slices per block of 128 threads (fixed
// order -> bitwise reproducible; the partials are already in torch order)
__global__ void __launch_bounds__(128) apg_reduce4_kernel(const float* __restrict__ partials, int ncta, int n,
                                                          float scale, float* __restrict__ grad) {
  __shared__ float s_part[dw::RED_SLICES][32];
  const int pl = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int p = blockIdx.x * 32 + pl;
  int c0, c1;
  dw::reduce_slice_bounds(ncta, slice, &c0, &c1);
  s_part[slice][pl] = p < n ? dw::reduce_slice_sum(partials, n, p, c0, c1) : 0.f;
  __syncthreads();
  if (slice == 0 && p < n) grad[p] = scale * ((s_part[0][pl] + s_part[1][pl]) + (s_part[2][pl] + s_part[3][pl]));
}

// the same reduction with the optimizer step of the reference fused in (optim.SGD(momentum), train_base.py:139-143:
// buf = momentum * buf + g; p -= lr * buf) - one launch instead of the reduction + two element-wise passes
__global__ void __launch_bounds__(128) apg_reduce4_sgd_kernel(const float* __restrict__ partials, int ncta, int n,
                                                              float scale, float* __restrict__ grad,
                                                              float* __restrict__ param, float* __restrict__ buf,
                                                              float lr, float momentum) {
  __shared__ float s_part[dw::RED_SLICES][32];
  const int pl = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int p = blockIdx.x * 32 + pl;
  int c0, c1;
  dw::reduce_slice_bounds(ncta, slice, &c0, &c1);
  s_part[slice][pl] = p < n ? dw::reduce_slice_sum(partials, n, p, c0, c1) : 0.f;
  __syncthreads();
  if (slice == 0 && p < n) {
    const float gsum = scale * ((s_part[0][pl] + s_part[1][pl]) + (s_part[2][pl] + s_part[3][pl]));
    if (grad) grad[p] = gsum;
    const float b = momentum * buf[p] + gsum;
    buf[p] = b;
    param[p] -= lr * b;
  }
}

cudaError_t launch_reduce_grad4_sgd(const float* partials, int ncta, int n, float scale, float* grad, float* param,
                                    float* buf, float lr, float momentum, cudaStream_t st) {
  APG_LAUNCH((n + 31) / 32, 128, 0, st, apg_reduce4_sgd_kernel)(partials, ncta, n, scale, grad, param, buf, lr, momentum);
  return cudaGetLastError();
}

cudaError_t launch_reduce_grad4(const float* partials, int ncta, int n, float scale, float* grad, cudaStream_t st) {
  APG_LAUNCH((n + 31) / 32, 128, 0, st, apg_reduce4_kernel)(partials, ncta, n, scale, grad);
  return cudaGetLastError();
}

cudaError_t launch_tq_dw(const HutterLayout& y, const RolloutArgs& a, const unsigned char* fstash,
                         const unsigned char* zstash, int grid, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(tq_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DWQ_SMEM);
  if (e != cudaSuccess) return e;
  APG_LAUNCH(grid, DWQ_THREADS, DWQ_SMEM, st, tq_dw_kernel)(y, a, fstash, zstash);
  return cudaGetLastError();
}

}  // namespace apg
