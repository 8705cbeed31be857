// tcgen05 / TMEM kernels of the quadrotor CONCURRENT rollout (Net(15,10,9,40,conv), h = 10), second generation:
//   tq_fwd_kernel : policy forward on the 5th-generation tensor cores (A operands and accumulators in TMEM, weight
//                   images resident in shared memory, 3xTF32) -> sigmoid -> h dynamics steps + tracking loss by the
//                   thread that owns the drone; every activation goes to the stash in UMMA operand-image format.
//   tq_dx_kernel  : reverse sweep through dynamics / loss by the owning thread -> d loss / d logits -> dX chain on the
//                   tensor cores against TRANSPOSED K-major weight images -> dZ of every layer to the dZ stash.
// The weight gradient is tq_dw_kernels.cu (streaming GEMM over the drone axis on the two stashes).
// Reference path: scripts/train_base.py:188-218 + scripts/train_drone.py:175-203 (forward), loss.backward() (adjoint).
//
// Roles (544 threads): warps 0-15 = four epilogue GROUPS of 128 threads (thread r of a group owns TMEM lane r = drone r
// of the group's tile); warp 16 lane 0 loads the weight images (one bulk copy) and issues every tcgen05.mma.
// Two TMEM slots of 256 columns = two tiles in the tensor chain at any time; group g works in slot g & 1, so each
// slot is shared by two groups that alternate: while one group runs the thread-per-drone dynamics phase of its tile
// (no TMEM needed), the other one runs the GEMM chain of the next tile in the same slot.  Hand-off by mbarriers:
// a_ready[s] (128 arrivals: the A operand of the next op is in TMEM), d_ready[s] (tcgen05.commit), slot_free[s] (128
// arrivals: the group has read the last accumulator of its tile, the slot belongs to the other group now).
#include "tq_layout.cuh"
#include "tc_prims.cuh"
#include "rollout_args.h"
#ifndef APG_TC_SIM
#include "tile_engine.cuh"
#endif
#include "kernels.h"

namespace apg {

using namespace tc;

namespace {

constexpr int TQ_EPI_WARPS = 16;
constexpr int TQ_THREADS = (TQ_EPI_WARPS + 1) * 32;          // 544
constexpr int TQ_FWD_SMEM = 1024 + BLOB_BYTES;
constexpr int TQ_DX_SMEM = 1024 + tq::TBLOB_BYTES;
constexpr int BULK_CHUNK = 32768;
static_assert(TQ_FWD_SMEM <= 232448 - 1024, "forward weight images do not fit in shared memory");
static_assert(BLOB_BYTES % 16 == 0 && tq::TBLOB_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");

struct TqBars {
  unsigned long long a_ready[2];
  unsigned long long d_ready[2];
  unsigned long long slot_free[2];
  unsigned long long w_ready;
};

// bounded wait: a protocol error must end the launch (the caller sees a NaN loss / gradient), never hang the GPU
__device__ __forceinline__ void tq_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag) {
  if (tcp::mbar_try_wait(bar, parity)) return;
  const long long t0 = tcp::clock_now();
  for (int spin = 0;; ++spin) {
    if (tcp::mbar_try_wait(bar, parity)) return;
    if ((spin & 63) == 63) {
      if (*abort_flag) return;
      if (tcp::clock_now() - t0 > 2000000000LL) {
#ifdef APG_TC_SIM
        if (getenv("APG_SIM_FAST_TIMEOUT")) fprintf(stderr, "tq_wait timeout: thread %u bar %x parity %u\n", threadIdx.x, bar, parity);
#endif
        *abort_flag = 1;
        return;
      }
    }
  }
}

#ifdef APG_TC_SIM
inline float tq_tanh(float x) { return tanhf(x); }
inline float tq_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }
#else
__device__ __forceinline__ float tq_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float tq_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// branch-free tanh: odd Taylor polynomial below 0.25 (truncation error 2e-9), 1 - 2 / (exp(2x) + 1) on the SFU above
// (absolute error <= 3e-7); saturates correctly for large |x|
__device__ __forceinline__ float tq_tanh(float x) {
  const float x2 = x * x;
  float p = fmaf(x2, 0.0218694885f, -0.0539682540f);
  p = fmaf(x2, p, 0.133333333f);
  p = fmaf(x2, p, -0.333333333f);
  const float small = fmaf(x * x2, p, x);
  const float big = fmaf(-2.f, tq_rcp(tq_ex2(x * 2.885390082f) + 1.f), 1.f);
  return fabsf(x) < 0.25f ? small : big;
}
__device__ __forceinline__ float tq_sigmoid(float x) { return tq_rcp(1.f + tq_ex2(x * -1.442695041f)); }
#endif

// the eight row-phase pointers of one stash set for the thread that owns drone `row` of the tile:
// element (set row r) lives at p[r & 7] + r * 128
struct SetPtr { unsigned char* p[8]; };
__device__ __forceinline__ SetPtr set_ptr(unsigned char* tile_block, int o_rows, int R, int row) {
  SetPtr sp;
  unsigned char* sb = tile_block + tq::set_base(o_rows) + (size_t)(row >> 5) * (size_t)(R * 128) + (row & 3) * 4;
  const uint32_t c = (uint32_t)(row & 31) >> 2;
#pragma unroll
  for (int k = 0; k < 8; ++k) sp.p[k] = sb + ((c ^ (uint32_t)k) << 4);
  return sp;
}
__device__ __forceinline__ void set_store(const SetPtr& sp, int r, float v) {
  *reinterpret_cast<float*>(sp.p[r & 7] + r * 128) = v;
}
__device__ __forceinline__ float set_load(const SetPtr& sp, int r) {
  return *reinterpret_cast<const float*>(sp.p[r & 7] + r * 128);
}

__device__ __forceinline__ void split_bits(float y, uint32_t* hi, uint32_t* lo) {
  const uint32_t h = __float_as_uint(y) & 0xffffe000u;
  *hi = h;
  *lo = __float_as_uint(y - __uint_as_float(h));
}
__device__ __forceinline__ void a_operand_ready(uint32_t bar) {
  tcp::wait_st();
  tcp::fence_before_thread_sync();
  tcp::mbar_arrive(bar);
}
// values x[0, 8*n8) -> (hi, lo) A-operand columns
template <int N8>
__device__ __forceinline__ void a_store(uint32_t ahi, uint32_t alo, const float* x) {
#pragma unroll
  for (int c0 = 0; c0 < N8 * 8; c0 += 8) {
    uint32_t h[8], l[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) split_bits(x[c0 + q], &h[q], &l[q]);
    tcp::tmem_st8(ahi + c0, h);
    tcp::tmem_st8(alo + c0, l);
  }
}

// common prologue: barriers, TMEM, one bulk copy of the weight images; returns the TMEM base
__device__ __forceinline__ uint32_t tq_setup(TqBars& bars, uint32_t* s_tmem, int* s_abort, unsigned char* base,
                                             const unsigned char* blob, int blob_bytes) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      tcp::mbar_init(smem_u32(&bars.a_ready[s]), 128);
      tcp::mbar_init(smem_u32(&bars.d_ready[s]), 1);
      tcp::mbar_init(smem_u32(&bars.slot_free[s]), 128);
    }
    tcp::mbar_init(smem_u32(&bars.w_ready), 1);
    *s_abort = 0;
    tcp::fence_mbar_init();
  }
  if (warp == TQ_EPI_WARPS) tcp::tmem_alloc512(s_tmem);
  tcp::fence_before_thread_sync();
  __syncthreads();
  tcp::fence_after_thread_sync();
  if (warp == TQ_EPI_WARPS && lane == 0) {
    const uint32_t bar = smem_u32(&bars.w_ready);
    tcp::mbar_expect_tx(bar, (uint32_t)blob_bytes);
    for (int off = 0; off < blob_bytes; off += BULK_CHUNK)
      tcp::bulk_g2s(smem_u32(base + off), blob + off, (uint32_t)min(BULK_CHUNK, blob_bytes - off), bar);
  }
  return *s_tmem;
}

}  // namespace

// weights (torch-flat) -> forward images + biases (tc_layout.cuh) and transposed images of the dX chain
__global__ void tq_pack_kernel(const float* __restrict__ params, const HutterLayout y, unsigned char* __restrict__ blob,
                               unsigned char* __restrict__ tblob) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < PAIRS_TOTAL + B_TOTAL) pack_body(e, params, y, blob);
  else if (e - (PAIRS_TOTAL + B_TOTAL) < tq::TPAIRS_TOTAL) tq::pack_t_body(e - (PAIRS_TOTAL + B_TOTAL), params, y, tblob);
}

__global__ void __launch_bounds__(TQ_THREADS, 1)
    tq_fwd_kernel(const unsigned char* __restrict__ blob, const RolloutArgs g, unsigned char* __restrict__ fstash) {
  APG_TC_DYNAMIC_SMEM(smem_raw);
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const float* s_bias = (const float*)(base + IMG_TOTAL);
  __shared__ __align__(8) TqBars s_bars;
  __shared__ uint32_t s_tmem;
  __shared__ int s_abort;
  __shared__ float s_red[TQ_EPI_WARPS];
  using Sys = Quad<float>;
  constexpr int S = Sys::S, A = Sys::A, R = Sys::REFW;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t tmem = tq_setup(s_bars, &s_tmem, &s_abort, base, blob, BLOB_BYTES);
  const int n = g.N;
  const int ntiles = (n + TMT - 1) / TMT;
  const int my_tiles = (ntiles > (int)blockIdx.x) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  volatile int* abort_flag = &s_abort;
  float my_loss = 0.f;
  tq_wait(smem_u32(&s_bars.w_ready), 0, abort_flag);          // weight images + biases have landed

  if (warp == TQ_EPI_WARPS) {
    if (lane == 0) {
      // the two slots advance independently: whichever has its next A operand ready gets its next op issued
      uint32_t par[2] = {0, 0};
      int op_i[2] = {0, 0}, tile_j[2] = {0, 1};
      int remaining = my_tiles * NOPS;
      long long t_idle = tcp::clock_now();
      while (remaining > 0) {
        bool progressed = false;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (tile_j[s] >= my_tiles) continue;
          if (!tcp::mbar_test_wait(smem_u32(&s_bars.a_ready[s]), par[s])) continue;
          par[s] ^= 1;
          tcp::fence_after_thread_sync();
          const Op op = op_of(op_i[s]);
          const uint32_t idesc = idesc_tf32(TMT, op.N);
          const uint32_t whi = smem_u32(base + op.img_off), wlo = whi + img_bytes(op.rows, op.K);
          const uint32_t slot = tmem + s * SLOT_COLS;
          const uint32_t d = slot + op.d_col, ahi = slot + C_AHI, alo = slot + C_ALO;
          for (int ks = 0; ks < op.K / 8; ++ks) {
            const uint64_t bh = kmajor_desc(whi, ks, op.K), bl = kmajor_desc(wlo, ks, op.K);
            tcp::mma_ts(d, alo + ks * 8, bh, idesc, (ks > 0 || !op.clear) ? 1u : 0u);
            tcp::mma_ts(d, ahi + ks * 8, bl, idesc, 1u);
            tcp::mma_ts(d, ahi + ks * 8, bh, idesc, 1u);
          }
          tcp::commit(smem_u32(&s_bars.d_ready[s]));
          if (++op_i[s] == NOPS) { op_i[s] = 0; tile_j[s] += 2; }
          --remaining;
          progressed = true;
        }
        if (progressed) {
          t_idle = tcp::clock_now();
        } else if (*abort_flag || tcp::clock_now() - t_idle > 2000000000LL) {
          *abort_flag = 1;
          break;
        }
      }
    }
  } else {
    const int grp = warp >> 2, s = grp & 1;
    const int row = (warp & 3) * 32 + lane;                  // TMEM lane = drone of the tile
    const uint32_t slot = tmem + s * SLOT_COLS + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t d_main = slot + C_DMAIN, d_conv = slot + C_DCONV, ahi = slot + C_AHI, alo = slot + C_ALO;
    const uint32_t bar_a = smem_u32(&s_bars.a_ready[s]), bar_d = smem_u32(&s_bars.d_ready[s]),
                   bar_f = smem_u32(&s_bars.slot_free[s]);
    for (int j = grp; j < my_tiles; j += 4) {
      const int tile = (int)blockIdx.x + j * (int)gridDim.x;
      const size_t drone = (size_t)tile * TMT + row;
      const bool live = drone < (size_t)n;
      const int t = j >> 1;                                   // this tile is the t-th one of its slot
      uint32_t dcnt = (uint32_t)t * NOPS;                     // commits of the slot before this tile
      unsigned char* tb = fstash + (size_t)tile * tq::F_TILE_BYTES;
      auto wait_d = [&]() {
        tq_wait(bar_d, dcnt & 1u, abort_flag);
        ++dcnt;
        tcp::fence_after_thread_sync();
      };
      // D_main (64 columns) -> tanh(x + b) -> A operand (hi, lo) + stash rows [row0, row0 + 64) of set `sp`
      auto dense_epilogue = [&](const float* b, const SetPtr& sp, int row0) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t v[32], l[32];
          tcp::tmem_ld32(d_main + hf * 32, v);
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const float yv = tq_tanh(__uint_as_float(v[q]) + b[hf * 32 + q]);
            set_store(sp, row0 + hf * 32 + q, yv);
            split_bits(yv, &v[q], &l[q]);
          }
          tcp::tmem_st32(ahi + hf * 32, v);
          tcp::tmem_st32(alo + hf * 32, l);
        }
        a_operand_ready(bar_a);
      };
      // ---- op 0 operand: in_state (15) + 1 (the ones row of the states_in weight gradient; its image column is 0)
      float x0[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) x0[k] = (live && k < F0) ? g.in_state[drone * F0 + k] : 0.f;
      x0[F0] = live ? 1.f : 0.f;
      // the slot's previous tile (the other group's) has left TMEM.  Parity waits alias with period 2, so a thread
      // must first see the hand-over of its OWN previous tile complete (all 128 arrivals, not just its own) before it
      // may ask for the next one.
      if (t >= 2) tq_wait(bar_f, (uint32_t)(t - 2) & 1u, abort_flag);
      if (t >= 2) tq_wait(bar_f, (uint32_t)(t - 2) & 1u, abort_flag);     // see tq_fwd_kernel
      if (t >= 1) {
        tq_wait(bar_f, (uint32_t)(t - 1) & 1u, abort_flag);
        tcp::fence_after_thread_sync();
      }
      a_store<2>(ahi, alo, x0);
      a_operand_ready(bar_a);
      {
        const SetPtr sp = set_ptr(tb, tq::O_XS, tq::R_XS, row);
#pragma unroll
        for (int k = 0; k < 16; ++k) set_store(sp, k, x0[k]);
      }
      const SetPtr sp_x1 = set_ptr(tb, tq::O_X1, tq::R_X1, row);
      wait_d();                                               // op 0: states_in
      dense_epilogue(s_bias + B_S, sp_x1, 0);                 // s -> X1 rows [0, 64), operand of op 1
      const float* rr = g.in_ref + drone * REFW;
#pragma unroll 1
      for (int gq = 0; gq < 4; ++gq) {
        float x[40];
#pragma unroll
        for (int k = 0; k < 36; k += 2) {
          const float2 tt = live ? *(const float2*)(rr + 18 * gq + k) : make_float2(0.f, 0.f);
          x[k] = tt.x;
          x[k + 1] = tt.y;
        }
        x[36] = live ? 1.f : 0.f;                             // ones row of the conv weight gradient (image column 0)
        x[37] = x[38] = x[39] = 0.f;
        wait_d();        // op 1 (gq = 0) or the fc1 piece of the previous pair: the A columns are free again
        a_store<5>(ahi, alo, x);
        a_operand_ready(bar_a);
        {
          const SetPtr sp = set_ptr(tb, tq::O_WIN + tq::R_WIN * gq, tq::R_WIN, row);
#pragma unroll
          for (int k = 0; k < 40; ++k) set_store(sp, k, x[k]);
        }
        wait_d();        // conv of this position pair
        {
          const float* b = s_bias + B_C;
          uint32_t v[40], l[40];
          tcp::tmem_ld32(d_conv, v);
          tcp::tmem_ld8(d_conv + 32, v + 32);
#pragma unroll
          for (int q = 0; q < 40; ++q) {
            const float yv = fmaxf(__uint_as_float(v[q]) + b[q], 0.f);
            // position-major x1 row of (pair gq, output q): 64 + 40 gq + q; the row phase only depends on q
            *reinterpret_cast<float*>(sp_x1.p[q & 7] + (HID + q) * 128 + gq * (40 * 128)) = yv;
            split_bits(yv, &v[q], &l[q]);
          }
          tcp::tmem_st32(ahi, v);
          tcp::tmem_st8(ahi + 32, v + 32);
          tcp::tmem_st32(alo, l);
          tcp::tmem_st8(alo + 32, l + 32);
          a_operand_ready(bar_a);
        }
      }
      wait_d();                                               // last fc1 piece
      dense_epilogue(s_bias + B_1, set_ptr(tb, tq::O_H1, tq::R_H, row), 0);
      wait_d();                                               // fc2
      dense_epilogue(s_bias + B_2, set_ptr(tb, tq::O_H2, tq::R_H, row), 0);
      wait_d();                                               // fc3
      dense_epilogue(s_bias + B_3, set_ptr(tb, tq::O_H3, tq::R_H, row), 0);
      wait_d();                                               // fc_out
      float act[MO];
      {
        const float* b = s_bias + B_O;
        const SetPtr sp = set_ptr(tb, tq::O_ACT, tq::R_ACT, row);
        uint32_t v[40];
        tcp::tmem_ld32(d_main, v);
        tcp::tmem_ld8(d_main + 32, v + 32);
#pragma unroll
        for (int q = 0; q < MO; ++q) {
          act[q] = tq_sigmoid(__uint_as_float(v[q]) + b[q]);                 // train_base.py:203
          set_store(sp, q, act[q]);
        }
      }
      // every tcgen05.ld of this tile has completed: the slot belongs to the other group of this slot now
      tcp::fence_before_thread_sync();
      tcp::mbar_arrive(bar_f);
      // ---- h dynamics steps + tracking loss of this drone (train_drone.py:175-203), states to the stash
      if (live) {
        const SetPtr sp = set_ptr(tb, tq::O_ST, tq::R_ST, row);
        float sc[S], s0[S], sn[S], rf[R];
        const float* cur_g = g.cur + drone * S;
        const float* ref_g = g.ref + drone * g.ref_rows * R;
#pragma unroll
        for (int q = 0; q < S; ++q) s0[q] = sc[q] = cur_g[q];
#pragma unroll
        for (int k = 0; k < H; ++k) {
          const float* a = act + k * A;
#pragma unroll
          for (int c = 0; c < R; ++c) rf[c] = ref_g[k * R + c];
          Sys::step(sc, a, g.dt, g.pc.v, sn);
          my_loss += Sys::loss(sn, rf, a, s0, k, H);
#pragma unroll
          for (int q = 0; q < S; ++q) {
            sc[q] = sn[q];
            set_store(sp, k * S + q, sn[q]);
          }
          if (g.states_out) {
#pragma unroll
            for (int q = 0; q < S; ++q) g.states_out[(drone * H + k) * S + q] = sn[q];
          }
          if (g.actions_out) {
#pragma unroll
            for (int c = 0; c < A; ++c) g.actions_out[(drone * H + k) * A + c] = a[c];
          }
        }
      }
    }
  }
  // ---- loss of this CTA: fixed-order sum over the epilogue threads
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) my_loss += __shfl_xor_sync(0xffffffffu, my_loss, o);
  if (lane == 0 && warp < TQ_EPI_WARPS) s_red[warp] = my_loss;
  tcp::fence_before_thread_sync();
  __syncthreads();
  if (tid == 0) {
    float tsum = 0.f;
#pragma unroll
    for (int w = 0; w < TQ_EPI_WARPS; ++w) tsum += s_red[w];
    // a protocol timeout poisons the loss on purpose: the caller must never take such a launch for a result
    g.loss_partials[blockIdx.x] = s_abort ? __int_as_float(0x7fc00000) : tsum;
  }
  if (warp == TQ_EPI_WARPS) tcp::tmem_dealloc512(tmem);
}

// =========================================================================================================
// dX chain: d loss / d logits from the reverse sweep of the owning thread, then
//   dZ3 = (dZo Wo) (.) (1 - h3^2), dZ2, dZ1 likewise, ds = (dZ1 W1[:, :64]) (.) (1 - s^2), dconv = (dZ1 W1[:, 64:]) (.) relu'
// with the B operands = transposed K-major weight images (tq_layout.cuh T_*).  Five hand-offs per tile; the first
// layer's 224 columns come out of two of them (accumulator columns [0,128) of the slot).
// =========================================================================================================
__global__ void __launch_bounds__(TQ_THREADS, 1)
    tq_dx_kernel(const unsigned char* __restrict__ tblob, const RolloutArgs g, unsigned char* __restrict__ fstash,
                 unsigned char* __restrict__ zstash, const unsigned char* __restrict__ stamp, int want_stamp) {
  APG_TC_DYNAMIC_SMEM(smem_raw);
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) TqBars s_bars;
  __shared__ uint32_t s_tmem;
  __shared__ int s_abort;
  using Sys = Quad<float>;
  constexpr int S = Sys::S, A = Sys::A, R = Sys::REFW;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t tmem = tq_setup(s_bars, &s_tmem, &s_abort, base, tblob, tq::TBLOB_BYTES);
  const int n = g.N;
  const int ntiles = (n + TMT - 1) / TMT;
  const int my_tiles = (ntiles > (int)blockIdx.x) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  volatile int* abort_flag = &s_abort;

  if (warp == TQ_EPI_WARPS) {
    if (lane == 0) {
      tq_wait(smem_u32(&s_bars.w_ready), 0, abort_flag);
      uint32_t par[2] = {0, 0};
      int h_i[2] = {0, 0}, tile_j[2] = {0, 1};
      int remaining = my_tiles * tq::NXH;
      long long t_idle = tcp::clock_now();
      while (remaining > 0) {
        bool progressed = false;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (tile_j[s] >= my_tiles) continue;
          if (!tcp::mbar_test_wait(smem_u32(&s_bars.a_ready[s]), par[s])) continue;
          par[s] ^= 1;
          tcp::fence_after_thread_sync();
          const uint32_t slot = tmem + s * tq::SLOT_COLS;
          const uint32_t ahi = slot + tq::XC_AHI, alo = slot + tq::XC_ALO;
          for (int i = tq::xh_first(h_i[s]); i < tq::xh_first(h_i[s] + 1); ++i) {
            const tq::XOp op = tq::xop_of(i);
            const tq::TImg im = tq::timage_of(op.img);
            const uint32_t whi = smem_u32(base + im.off) + (uint32_t)(op.row0 >> 3) * (uint32_t)((im.K >> 2) * 128);
            const uint32_t wlo = whi + img_bytes(im.rows, im.K);
            const uint32_t idesc = idesc_tf32(TMT, op.N);
            const uint32_t d = slot + tq::XC_D + op.d_col;
            for (int ks = 0; ks < op.K / 8; ++ks) {
              const uint64_t bh = kmajor_desc(whi, ks, im.K), bl = kmajor_desc(wlo, ks, im.K);
              tcp::mma_ts(d, alo + ks * 8, bh, idesc, ks > 0 ? 1u : 0u);
              tcp::mma_ts(d, ahi + ks * 8, bl, idesc, 1u);
              tcp::mma_ts(d, ahi + ks * 8, bh, idesc, 1u);
            }
          }
          tcp::commit(smem_u32(&s_bars.d_ready[s]));
          if (++h_i[s] == tq::NXH) { h_i[s] = 0; tile_j[s] += 2; }
          --remaining;
          progressed = true;
        }
        if (progressed) {
          t_idle = tcp::clock_now();
        } else if (*abort_flag || tcp::clock_now() - t_idle > 2000000000LL) {
          *abort_flag = 1;
          break;
        }
      }
    }
  } else {
    const int grp = warp >> 2, s = grp & 1;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t slot = tmem + s * tq::SLOT_COLS + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t d0 = slot + tq::XC_D, ahi = slot + tq::XC_AHI, alo = slot + tq::XC_ALO;
    const uint32_t bar_a = smem_u32(&s_bars.a_ready[s]), bar_d = smem_u32(&s_bars.d_ready[s]),
                   bar_f = smem_u32(&s_bars.slot_free[s]);
    for (int j = grp; j < my_tiles; j += 4) {
      const int tile = (int)blockIdx.x + j * (int)gridDim.x;
      const size_t drone = (size_t)tile * TMT + row;
      const bool live = drone < (size_t)n;
      const int t = j >> 1;
      uint32_t dcnt = (uint32_t)t * tq::NXH;
      unsigned char* tb = fstash + (size_t)tile * tq::F_TILE_BYTES;
      unsigned char* zb = zstash + (size_t)tile * tq::Z_TILE_BYTES;
      auto wait_d = [&]() {
        tq_wait(bar_d, dcnt & 1u, abort_flag);
        ++dcnt;
        tcp::fence_after_thread_sync();
      };
      // ---- reverse dynamics sweep of this drone (dyn_phase.cuh dyn_adjoint_conc on the stash sets): dlog[40]
      float dlog[MO];
#pragma unroll
      for (int q = 0; q < MO; ++q) dlog[q] = 0.f;
      if (live) {
        const SetPtr sp_st = set_ptr(tb, tq::O_ST, tq::R_ST, row);
        const SetPtr sp_act = set_ptr(tb, tq::O_ACT, tq::R_ACT, row);
        float s0[S], sk[S], sn[S], a[A], rf[R], gq[S], gs[S], ga[A], ga2[A];
        const float* cur_g = g.cur + drone * S;
        const float* ref_g = g.ref + drone * g.ref_rows * R;
#pragma unroll
        for (int q = 0; q < S; ++q) { s0[q] = cur_g[q]; gq[q] = 0.f; }
#pragma unroll
        for (int q = 0; q < S; ++q) sn[q] = set_load(sp_st, (H - 1) * S + q);
#pragma unroll
        for (int k = H - 1; k >= 0; --k) {
#pragma unroll
          for (int c = 0; c < A; ++c) { a[c] = set_load(sp_act, k * A + c); ga[c] = 0.f; }
#pragma unroll
          for (int c = 0; c < R; ++c) rf[c] = ref_g[k * R + c];
          if (k > 0) {
#pragma unroll
            for (int q = 0; q < S; ++q) sk[q] = set_load(sp_st, (k - 1) * S + q);
          } else {
#pragma unroll
            for (int q = 0; q < S; ++q) sk[q] = s0[q];
          }
          Sys::loss_grad(sn, rf, a, s0, k, H, gq, ga);
          Sys::step_adj(sk, a, g.dt, g.pc.v, gq, gs, ga2);
#pragma unroll
          for (int c = 0; c < A; ++c) dlog[k * A + c] = (ga[c] + ga2[c]) * a[c] * (1.f - a[c]);      // sigmoid'
#pragma unroll
          for (int q = 0; q < S; ++q) { gq[q] = gs[q]; sn[q] = sk[q]; }
        }
      }
      if (t >= 2) tq_wait(bar_f, (uint32_t)(t - 2) & 1u, abort_flag);     // see tq_fwd_kernel
      if (t >= 1) {
        tq_wait(bar_f, (uint32_t)(t - 1) & 1u, abort_flag);
        tcp::fence_after_thread_sync();
      }
      a_store<5>(ahi, alo, dlog);
      a_operand_ready(bar_a);
      {
        const SetPtr sp = set_ptr(zb, tq::O_ZO, MO, row);
#pragma unroll
        for (int q = 0; q < MO; ++q) set_store(sp, q, dlog[q]);
      }
      // ---- dZ_l = (dZ_{l+1} W_{l+1}) (.) (1 - X_l^2) for h3, h2, h1: D -> A operand + dZ stash
#pragma unroll 1
      for (int l = 0; l < 3; ++l) {
        const SetPtr sp_y = set_ptr(tb, l == 0 ? tq::O_H3 : (l == 1 ? tq::O_H2 : tq::O_H1), tq::R_H, row);
        const SetPtr sp_z = set_ptr(zb, l == 0 ? tq::O_Z3 : (l == 1 ? tq::O_Z2 : tq::O_Z1), HID, row);
        float yv[64];                                          // issued before the wait: hides the stash latency
#pragma unroll
        for (int q = 0; q < 64; ++q) yv[q] = set_load(sp_y, q);
        wait_d();
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 16) {
          uint32_t v[16], lo[16];
          tcp::tmem_ld16(d0 + c0, v);
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const float yy = yv[c0 + q];
            const float z = __uint_as_float(v[q]) * (1.f - yy * yy);
            set_store(sp_z, c0 + q, z);
            split_bits(z, &v[q], &lo[q]);
          }
          tcp::tmem_st16(ahi + c0, v);
          tcp::tmem_st16(alo + c0, lo);
        }
        a_operand_ready(bar_a);
      }
      // ---- first layer (A operand = dZ1 stays in TMEM): hand-off 3 = ds [0,64) + pair 0 [64,104); hand-off 4 =
      //      pairs 1, 2 [0,80) + pair 3 [80,120)
      const SetPtr sp_x1 = set_ptr(tb, tq::O_X1, tq::R_X1, row);
      const SetPtr sp_zx = set_ptr(zb, tq::O_ZX, K1, row);
      {
        float yv[64];
#pragma unroll
        for (int q = 0; q < 64; ++q) yv[q] = set_load(sp_x1, q);
        wait_d();
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 16) {
          uint32_t v[16];
          tcp::tmem_ld16(d0 + c0, v);
#pragma unroll
          for (int q = 0; q < 16; ++q)
            set_store(sp_zx, c0 + q, __uint_as_float(v[q]) * (1.f - yv[c0 + q] * yv[c0 + q]));
        }
      }
      {
        uint32_t v[40];
        tcp::tmem_ld32(d0 + 64, v);
        tcp::tmem_ld8(d0 + 96, v + 32);
#pragma unroll
        for (int q = 0; q < 40; ++q) {
          const float yy = set_load(sp_x1, HID + q);
          set_store(sp_zx, HID + q, yy > 0.f ? __uint_as_float(v[q]) : 0.f);             // relu'
        }
      }
      tcp::fence_before_thread_sync();
      tcp::mbar_arrive(bar_a);                                 // D has been read: go on with the other three pairs
      wait_d();
#pragma unroll 1
      for (int gp = 1; gp < 4; ++gp) {
        uint32_t v[40];
        const uint32_t dc = d0 + (gp - 1) * 40;
        tcp::tmem_ld32(dc, v);
        tcp::tmem_ld8(dc + 32, v + 32);
#pragma unroll
        for (int q = 0; q < 40; ++q) {
          // x1 / zx row 64 + 40 gp + q: the row phase only depends on q
          const float yy = *reinterpret_cast<const float*>(sp_x1.p[q & 7] + (HID + q) * 128 + gp * (40 * 128));
          *reinterpret_cast<float*>(sp_zx.p[q & 7] + (HID + q) * 128 + gp * (40 * 128)) =
              yy > 0.f ? __uint_as_float(v[q]) : 0.f;
        }
      }
      tcp::fence_before_thread_sync();
      tcp::mbar_arrive(bar_f);                                 // the slot belongs to the other group now
    }
  }
  tcp::fence_before_thread_sync();
  __syncthreads();
  // the stash / weight images must come from the forward of THIS path (workspace stamp, capi.cu)
  if (tid == 0 && stamp && (int)stamp[0] != want_stamp) s_abort = 1;
  if (tid == 0 && s_abort && my_tiles > 0)                     // poison the gradient: never a silent wrong result
    *reinterpret_cast<float*>(zstash + (size_t)blockIdx.x * tq::Z_TILE_BYTES) = __int_as_float(0x7fc00000);
  if (warp == TQ_EPI_WARPS) tcp::tmem_dealloc512(tmem);
}

size_t tq_blob_bytes() { return (size_t)BLOB_BYTES; }
size_t tq_tblob_bytes() { return (size_t)tq::TBLOB_BYTES; }
size_t tq_fstash_bytes(int n) { return (size_t)((n + TMT - 1) / TMT) * tq::F_TILE_BYTES; }
size_t tq_zstash_bytes(int n) { return (size_t)((n + TMT - 1) / TMT) * tq::Z_TILE_BYTES; }
int tq_grid(int n, int sms) { const int nt = (n + TMT - 1) / TMT; return nt < sms ? nt : sms; }

bool tq_supported(const HutterLayout& y, int h) {
  return y.conv && y.F0 == F0 && y.L == H && y.RD == RD && y.Mo == MO && h == H;
}

cudaError_t launch_tq_fwd(const HutterLayout& y, const float* params, unsigned char* blob, unsigned char* tblob,
                          const RolloutArgs& a, unsigned char* fstash, int grid, cudaStream_t st) {
  const int items = PAIRS_TOTAL + B_TOTAL + tq::TPAIRS_TOTAL;
  APG_LAUNCH((items + 255) / 256, 256, 0, st, tq_pack_kernel)(params, y, blob, tblob);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(tq_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TQ_FWD_SMEM);
  if (e != cudaSuccess) return e;
  APG_LAUNCH(grid, TQ_THREADS, TQ_FWD_SMEM, st, tq_fwd_kernel)(blob, a, fstash);
  return cudaGetLastError();
}

cudaError_t launch_tq_dx(const unsigned char* tblob, const RolloutArgs& a, unsigned char* fstash,
                         unsigned char* zstash, const unsigned char* stamp, int want_stamp, int grid, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(tq_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TQ_DX_SMEM);
  if (e != cudaSuccess) return e;
  APG_LAUNCH(grid, TQ_THREADS, TQ_DX_SMEM, st, tq_dx_kernel)(tblob, a, fstash, zstash, stamp, want_stamp);
  return cudaGetLastError();
}

}  // namespace apg
