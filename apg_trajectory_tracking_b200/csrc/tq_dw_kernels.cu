// Weight gradient of the quadrotor concurrent rollout as ONE streaming tcgen05 GEMM over the drone axis,
//     dW_l[out][in] = sum over drones dZ_l[drone][out] * X_l[drone][in]          (loss.backward() of train_base.py:205),
// on the two operand-image stashes written by tq_fwd_kernel (X_l) and tq_dx_kernel (dZ_l): tq_layout.cuh.
// Accumulators resident in TMEM, D[M = in (+ ones row -> bias)][N = out], in two passes over the CTA's tiles (pass 0:
// every layer but fc1, 288 columns; pass 1: fc1, 128 columns - adj_dw_layout.cuh) so that 224 columns are left for the
// ring the A operand is fed through.
//
// Pipeline (unit = one 32-drone panel of one op):
//   warp 0       producer  : two 1-D bulk copies per unit (A panel, B panel) straight from the stash into a ring of
//                            12 (pass 0) / 7 (pass 1) raw stages - the stash IS the 128B-swizzled K-major image, so there
//                            is no loader arithmetic and no tensor map; the copies in flight hide the HBM latency
//   warps 2-13   A feeders : three groups of four warps, one per operand slot; thread = A row (= TMEM lane).  Reads its 128-byte row of the raw panel from shared memory
//                            ONCE (conflict free: the swizzle spreads 8 rows over the 8 bank groups), splits it into
//                            (raw, lo = x - tf32(x)) and writes both with tcgen05.st into the unit's slot of TMEM columns.
//                            Rows past the op's A rows are neither read nor written: row m of A only ever reaches row m
//                            of D, and the flush ignores those rows
//   warps 14-16  B feeders : one warp per operand slot: lo image of the B panel into the slot's shared memory
//   warp 1       MMA issuer: per unit ONE wait (slot ready: 5 warp arrivals), 4 k-steps x 3 MMAs (3xTF32) with A FROM
//                            TMEM and B from shared memory, two commits (slot free, raw stage free)
// Measured on the way here (profiles/r2):
//   * both operands from shared memory: every MMA re-read a full 128-row A panel (also for the 16- and 37-row ops) and
//     the feeders wrote a second lo image - ~144 KiB of shared-memory traffic per unit at 128 B/clk, 48 cycles per MMA
//     against 32 with A from TMEM (tools/micro/tcgen05_rate.cu);
//   * an mbarrier wait costs the waiting warp ~150 cycles even when the phase is already complete
//     (tools/micro/mbar_pingpong.cu: 374 cycles per wait + arrive pair, whatever the ring depth): with separate
//     barriers for the B lo image and each half of the A columns the issuer spent 1200 cycles per unit, 336 of them on
//     tensor work.  Hence ONE ready barrier per unit - and, for the same reason, feeder warps that each own WHOLE units
//     (a feeder's two waits + fence + arrive cost it ~900 cycles per unit, whatever the amount of data): with every
//     feeder warp taking part in every unit the feeders, not the tensor core or HBM, set the pace.
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

// warps: producer, MMA issuer, then one TEAM per operand slot - four A feeder warps and one B feeder warp that own
// every third unit (unit u -> slot u % 3 -> team u % 3).  A team observes every phase of its slot's barriers; that is
// what makes waiting by parity safe (a waiter that skipped a phase could take the phase before for the one it needs).
constexpr int DWQ_TEAMS = tq::DW_NSLOT;
constexpr int DWQ_THREADS = 32 * (2 + 5 * DWQ_TEAMS);         // 544
constexpr int NRMAX = tq::DW_NRAW_MAX, NS = tq::DW_NSLOT;
constexpr int DWQ_T_FLOATS = (4 * tc::RD + 1) * 48;           // conv Toeplitz block between its flush and the fold
constexpr int DWQ_SMEM = 1024 + tq::DW_RAW_BYTES + NS * tq::DW_B_BYTES;
static_assert(DWQ_SMEM <= 232448, "stage rings do not fit in shared memory");
static_assert(tq::DW_T_OFFSET + DWQ_T_FLOATS * 4 <= tq::DW_RAW_BYTES, "conv block does not fit behind the pass-1 stages");
static_assert(dw::C_ARING + NS * 64 <= 512, "accumulators + A slots do not fit in TMEM");

// Cursor of one role over the raw stages of the current pass.  The parity a wait uses is kept per stage in a bit mask
// (flipped at every use) instead of being derived from a unit counter: the number of stages changes with the pass.
// Consumers start at 0 (wait for the first completion), the producer at all-ones (a fresh barrier counts as "the phase
// of parity 1 has completed": the first wait on every stage passes at once).
struct RawCursor {
  int r;
  uint32_t par;
  __device__ __forceinline__ uint32_t parity() const { return (par >> r) & 1u; }
  __device__ __forceinline__ void next(int nr) { par ^= 1u << r; r = (r + 1 == nr) ? 0 : r + 1; }
};

struct DwqBars {
  unsigned long long full[NRMAX];                   // bulk copies landed (1 arrival + bytes)
  unsigned long long rfree[NRMAX];                  // MMAs of the unit are complete (tcgen05.commit): raw stage reusable
  unsigned long long ready[NS];                  // A columns + B lo image of the slot written (5 arrivals: lane 0 of
                                                 // the four A feeder warps of the unit's group and of its B feeder warp)
  unsigned long long sfree[NS];                  // MMAs that read the slot are complete (tcgen05.commit)
  unsigned long long done[dw::NPASS];            // accumulators of the pass are final (tcgen05.commit)
  unsigned long long flushed;                    // pass-0 accumulators are in the partial (12 arrivals): columns reusable
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

__device__ __forceinline__ uint32_t lo_bits(uint32_t x) {
  return __float_as_uint(__uint_as_float(x) - __uint_as_float(x & 0xffffe000u));
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
  if (threadIdx.x == 0 && blockIdx.x < 148) TQ_PROF_ARRAY[2][blockIdx.x][16] = tq_globaltimer();
  long long tqp_k1_ = 0, tqp_k2_ = 0;
#endif
  const int n = g.N;
  const int ntiles = (n + tc::TMT - 1) / tc::TMT;
  // The drone axis is the K dimension of this GEMM, so the work splits at PANEL granularity (32 drones), not by tile:
  // CTA b takes the contiguous items [i0, i1) of the ntiles * NPANEL (tile, panel) pairs - 2048 panels over 148 CTAs
  // are 13 or 14 each (99 % balanced) where 512 whole tiles were 3 or 4 (86 %).  Its first and last tile may be partial:
  // tile t0 + j contributes the panels [p_lo(j), p_hi(j)).  Every role below walks (tile, op, panel) in this order.
  const int n_items = ntiles * tq::NPANEL;
  const int i0 = (int)(((long long)blockIdx.x * n_items) / (int)gridDim.x);
  const int i1 = (int)(((long long)(blockIdx.x + 1) * n_items) / (int)gridDim.x);
  const int t0 = i0 / tq::NPANEL;
  const int my_tiles = i1 > i0 ? (i1 - 1) / tq::NPANEL - t0 + 1 : 0;
  // first panel of the first tile | (end panel of the last tile) << 4, in one register (the feeder warps are at 96)
  const int p_ends = (i0 - t0 * tq::NPANEL) | ((i1 - (t0 + my_tiles - 1) * tq::NPANEL) << 4);
  auto p_lo = [&](int j) { return j == 0 ? (p_ends & 15) : 0; };
  auto p_hi = [&](int j) { return j == my_tiles - 1 ? (p_ends >> 4) : tq::NPANEL; };
  float* P = g.grad_partials + (size_t)blockIdx.x * y.n_params;
  volatile int* abort_flag = &s_abort;
  unsigned char* lo_base = base + tq::DW_RAW_BYTES;
  float* s_T = reinterpret_cast<float*>(base + tq::DW_T_OFFSET);            // [37][48] conv Toeplitz block

  if (tid == 0) {
    for (int s = 0; s < NRMAX; ++s) {
      tcp::mbar_init(smem_u32(&s_bars.full[s]), 1);
      tcp::mbar_init(smem_u32(&s_bars.rfree[s]), 1);
    }
    for (int s = 0; s < NS; ++s) {
      tcp::mbar_init(smem_u32(&s_bars.ready[s]), 5);
      tcp::mbar_init(smem_u32(&s_bars.sfree[s]), 1);
    }
    for (int s = 0; s < dw::NPASS; ++s) tcp::mbar_init(smem_u32(&s_bars.done[s]), 1);
    tcp::mbar_init(smem_u32(&s_bars.flushed), 4 * DWQ_TEAMS);
    s_abort = 0;
    tcp::fence_mbar_init();
  }
  if (warp == 0) tcp::tmem_alloc512(&s_tmem);
  tcp::fence_before_thread_sync();
  __syncthreads();
  tcp::fence_after_thread_sync();
  const uint32_t tmem = s_tmem;
  // launched with programmatic serialization behind the dX chain: the set-up above overlaps its tail
  tcp::griddep_wait();
  tcp::griddep_launch();
#ifdef APG_PROFILE
  tqp_k1_ = clock64();
#endif

  // ---- accumulators of a pass -> this CTA's gradient partial, by the eight feeder warps: warp w reads TMEM lanes
  // 32 (w & 3) .. +31 (= A rows) and one half of every region's columns.  The row -> gradient map (base + column *
  // stride) is worked out ONCE per thread and region - measured: with the index arithmetic (divisions by 20, the switch
  // of grad_index) inside the element loop on four warps this took 73 k cycles, a third of the kernel.
  // Every entry of the partial is written exactly once per launch; unused tensors (ref_in.*) and padding are zeroed.
  auto flush = [&](int pass) {
    dwq_wait(smem_u32(&s_bars.done[pass]), 0, abort_flag);
    tcp::fence_after_thread_sync();
    const int r = (warp & 3) * 32 + lane, grp = (warp - 2) >> 2;      // the twelve A feeder warps
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int col0[6] = {dw::C_WO, dw::C_W3, dw::C_W2, dw::C_W1A, dw::C_W1B, dw::C_WS};
    const int ncol[6] = {48, 64, 64, 64, 64, 64};
    if (grp < 2) {                                             // groups 0 and 1: one half of every region's columns
#pragma unroll
      for (int reg = 0; reg < 6; ++reg) {
        if ((reg == 3 || reg == 4) != (pass == 1)) continue;    // fc1 regions belong to pass 1
        // entry (r, n) -> P[i0 + n * stride]; i0 < 0: this row is padding in this region
        const int i0 = dw::grad_index(y, reg, r, 0);
        const int stride = i0 < 0 ? 0 : dw::grad_index(y, reg, r, 1) - i0;
        const int nvalid = reg == 0 ? tc::MO : ncol[reg];
        const int c_lo = grp * (ncol[reg] / 2), c_hi = c_lo + ncol[reg] / 2;
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
    } else if (pass == 0) {                                    // group 2: the conv Toeplitz block
      for (int c0 = 0; c0 < 48; c0 += 8) {
        uint32_t vb[8];
        tcp::tmem_ld8(lane_addr + dw::C_WT + c0, vb);
        if (r <= 4 * tc::RD) {
#pragma unroll
          for (int q = 0; q < 8; ++q) s_T[r * 48 + c0 + q] = __uint_as_float(vb[q]);   // folded after the last pass
        }
      }
    }
    if (pass == 0) {
      // every tcgen05.ld above has completed (tmem_ld8 waits): the issuer may overwrite the columns
      tcp::fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) tcp::mbar_arrive(smem_u32(&s_bars.flushed));
    }
  };
  for (int i = tid; i < y.n_params; i += DWQ_THREADS) {
    const bool conv_w = i >= y.t_wc && i < y.t_wc + tc::NC * tc::RD * 3 + tc::NC;    // conv_ref weight + bias: at the end
    if (!conv_w && (my_tiles == 0 || (i >= y.t_wr && i < y.t_br + HID))) P[i] = 0.f;
  }

  if (warp == 0) {
    // ===================================================== producer (warp-uniform loop, the elected lane issues)
    {
      const bool leader = tcp::elect_one();
      TQP_DECL
      RawCursor rc{0, 0xffffffffu};
      for (int pass = 0; pass < dw::NPASS; ++pass) {
      rc.r = 0;
      const int nr = tq::dw_nraw(pass), sbytes = tq::dw_stage_bytes(pass), boff = tq::dw_b_offset(pass);
      // the stages of this pass lie across the stages of the one before (different size): every MMA of that pass
      // must have read its operands before the first copy of this one lands.  (The bubble overlaps the flush.)
      if (pass > 0 && my_tiles > 0) dwq_wait(smem_u32(&s_bars.done[pass - 1]), 0, abort_flag);
      for (int j = 0; j < my_tiles; ++j) {
        const int tile = t0 + j;
        const unsigned char* fb = fstash + (size_t)tile * tq::F_TILE_BYTES;
        const unsigned char* zb = zstash + (size_t)tile * tq::Z_TILE_BYTES;
        for (int k = 0; k < dw::pass_nops(pass); ++k) {
          const tq::DwSrc src = tq::dw_src(dw::pass_op(pass, k));
          const uint32_t a_bytes = (uint32_t)src.a_rows * 128u, b_bytes = (uint32_t)src.b_rows * 128u;
          for (int p = p_lo(j); p < p_hi(j); ++p) {
            const int r = rc.r;
            dwq_wait(smem_u32(&s_bars.rfree[r]), rc.parity(), abort_flag);
            rc.next(nr);
            TQP(0);
            unsigned char* st = base + r * sbytes;
            const uint32_t bar = smem_u32(&s_bars.full[r]);
            if (leader) {
            tcp::mbar_expect_tx(bar, a_bytes + b_bytes);
            tcp::bulk_g2s(smem_u32(st), fb + tq::set_base(src.a_set) + (size_t)p * (size_t)(src.a_R * 128) +
                                            (size_t)src.a_row0 * 128, a_bytes, bar);
            tcp::bulk_g2s(smem_u32(st + boff), zb + tq::set_base(src.b_set) +
                                                             (size_t)p * (size_t)(src.b_R * 128) +
                                                             (size_t)src.b_row0 * 128, b_bytes, bar);
            }
            __syncwarp();          // lanes stay within one unit of each other (parity waits alias with period 2)
            TQP(1);
          }
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
      RawCursor rc{0, 0u};                                  // (only the stage index: this warp does not wait on `full`)
      for (int pass = 0; pass < dw::NPASS; ++pass) {
      rc.r = 0;
      const int nr = tq::dw_nraw(pass), sbytes = tq::dw_stage_bytes(pass), boff = tq::dw_b_offset(pass);
      if (pass > 0 && my_tiles > 0) {                      // the columns of pass 0 must have been flushed
        if (leader) dwq_wait(smem_u32(&s_bars.flushed), 0, abort_flag);
        __syncwarp();
        tcp::fence_after_thread_sync();
      }
      for (int j = 0; j < my_tiles; ++j)
        for (int k = 0; k < dw::pass_nops(pass); ++k) {
          const dw::Op op = dw::op_of(dw::pass_op(pass, k));
          const uint32_t idesc = tc::idesc_tf32(128, op.N);
          const uint32_t d = tmem + op.d_col;
          for (int p = p_lo(j); p < p_hi(j); ++p, ++u) {
            const int r = rc.r, sl = u % NS;
            rc.next(nr);
            // only the issuing lane polls: a barrier of a short ring can complete AGAIN as soon as this unit's MMAs
            // are committed, so a lane that looked late would see the parity it is waiting for already gone
            if (leader) dwq_wait(smem_u32(&s_bars.ready[sl]), (uint32_t)(u / NS) & 1u, abort_flag);
            __syncwarp();
            TQP(0);
            tcp::fence_after_thread_sync();
            // B descriptors: constant high word (SBO 1024, version, SWIZZLE_128B), low word = address >> 4 | LBO
            // field; k-step ks adds 2 (32 bytes >> 4)
            uint32_t br = ((raw0 + (uint32_t)(r * sbytes + boff)) >> 4) | (1u << 16);
            uint32_t bl = ((lo0 + (uint32_t)sl * tq::DW_B_BYTES) >> 4) | (1u << 16);
            const uint32_t a_hi = tmem + dw::C_ARING + sl * 64, a_lo = a_hi + 32;
            const bool clear = (j == 0) && op.first && (p == p_lo(0));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks, br += 2, bl += 2) {
              const uint64_t dbh = DESC_HI | br, dbl = DESC_HI | bl;
#ifdef DWQ_PROBE_NO_MMA      // timing experiment: how fast can the operands be delivered at all (wrong results)
              if (false) {
#else
              if (leader) {
#endif
                tcp::mma_ts(d, a_lo + ks * 8, dbh, idesc, (ks > 0 || !clear) ? 1u : 0u);
                tcp::mma_ts(d, a_hi + ks * 8, dbl, idesc, 1u);
                tcp::mma_ts(d, a_hi + ks * 8, dbh, idesc, 1u);
              }
            }
            if (leader) {
              tcp::commit(smem_u32(&s_bars.sfree[sl]));   // slot and raw stage are free once these MMAs have read them
              tcp::commit(smem_u32(&s_bars.rfree[r]));
            }
            __syncwarp();          // lanes stay within one unit of each other (parity waits alias with period 2)
            TQP(1);
          }
        }
      if (leader) tcp::commit(smem_u32(&s_bars.done[pass]));    // the accumulators of this pass are final
      __syncwarp();
      }
      if (leader) TQP_FLUSH(2, 2, 2);
    }
  } else if (warp < 2 + 4 * DWQ_TEAMS) {
    // ===================================================== A feeders: raw panel row -> (raw, lo) TMEM columns
    const int q = warp & 3, row = q * 32 + lane, ph = row & 7, grp = (warp - 2) >> 2;
    const uint32_t a_cols = tmem + ((uint32_t)(q * 32) << 16) + dw::C_ARING;
    TQP_DECL
    int u = 0;
    RawCursor rc{0, 0u};
    for (int pass = 0; pass < dw::NPASS; ++pass) {
    rc.r = 0;
    const int nr = tq::dw_nraw(pass), sbytes = tq::dw_stage_bytes(pass);
    for (int j = 0; j < my_tiles; ++j)
      for (int k = 0; k < dw::pass_nops(pass); ++k) {
        const tq::DwSrc src = tq::dw_src(dw::pass_op(pass, k));
        // rows of this warp that exist in the op (the ones row included); a warp without any neither reads nor writes
        const bool warp_active = q * 32 < src.a_rows + (src.ones >= 0 ? 1 : 0);
        const bool has_row = row < src.a_rows;
        const uint32_t fill = (row == src.ones) ? 0x3f800000u : 0u;       // constant ones row of a bias gradient
        for (int p = p_lo(j); p < p_hi(j); ++p, ++u) {
          const int r = rc.r, sl = u % NS;
          const uint32_t full_parity = rc.parity();
          rc.next(nr);
          if (sl != grp) continue;                            // another team's unit
          uint4 v[8];
#ifndef DWQ_PROBE_NO_FEED
          if (warp_active) {
            dwq_wait(smem_u32(&s_bars.full[r]), full_parity, abort_flag);
            TQP(0);
            // logical 16-byte chunk c (drones 4c .. 4c+3) of row `row` sits at chunk c ^ (row & 7) of its 128 bytes
            const unsigned char* a_row = base + r * sbytes + row * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              v[c] = has_row ? *reinterpret_cast<const uint4*>(a_row + ((c ^ ph) << 4)) : make_uint4(fill, fill, fill, fill);
          }
#endif
          if (u >= NS) dwq_wait(smem_u32(&s_bars.sfree[sl]), (uint32_t)(u / NS - 1) & 1u, abort_flag);
          TQP(1);
          tcp::fence_after_thread_sync();
#ifndef DWQ_PROBE_NO_FEED
          if (warp_active) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const uint4 x = v[4 * hh + c];
                hi[4 * c + 0] = x.x; hi[4 * c + 1] = x.y; hi[4 * c + 2] = x.z; hi[4 * c + 3] = x.w;
                lo[4 * c + 0] = has_row ? lo_bits(x.x) : 0u; lo[4 * c + 1] = has_row ? lo_bits(x.y) : 0u;
                lo[4 * c + 2] = has_row ? lo_bits(x.z) : 0u; lo[4 * c + 3] = has_row ? lo_bits(x.w) : 0u;
              }
              tcp::tmem_st16(a_cols + sl * 64 + hh * 16, hi);        // the tensor core truncates the raw image: the hi part
              tcp::tmem_st16(a_cols + sl * 64 + 32 + hh * 16, lo);
            }
            tcp::wait_st();
          }
#endif
          tcp::fence_before_thread_sync();
          __syncwarp();
          if (lane == 0) tcp::mbar_arrive(smem_u32(&s_bars.ready[sl]));
          TQP(2);
        }
      }
    if (my_tiles > 0) flush(pass);
    TQP(3);
    }
    if (tid == 64) TQP_FLUSH(2, 4, 4);
  } else {
    // ===================================================== B feeders: lo image of the B panel, one warp per unit
    const int bw = warp - (2 + 4 * DWQ_TEAMS);
    TQP_DECL
    int u = 0;
    RawCursor rc{0, 0u};
    for (int pass = 0; pass < dw::NPASS; ++pass) {
    rc.r = 0;
    const int nr = tq::dw_nraw(pass), sbytes = tq::dw_stage_bytes(pass), boff = tq::dw_b_offset(pass);
    // A wait by parity is only safe if the phase BEFORE the awaited one is known to be complete.  Within a pass that
    // follows from the order of the MMAs (see DWQ_TEAMS); across the pass boundary this warp - unlike the A feeders,
    // which flush - would otherwise reach the first `full` barrier of the new pass while the last copies of the old
    // pass, made by other teams into other stages that share these barriers, may still be in flight.
    if (pass > 0 && my_tiles > 0) dwq_wait(smem_u32(&s_bars.done[pass - 1]), 0, abort_flag);
    for (int j = 0; j < my_tiles; ++j)
      for (int k = 0; k < dw::pass_nops(pass); ++k) {
        const tq::DwSrc src = tq::dw_src(dw::pass_op(pass, k));
        const int nb = src.b_rows * 8;                       // 16-byte chunks (320 or 512: a multiple of 32)
        for (int p = p_lo(j); p < p_hi(j); ++p, ++u) {
          const int r = rc.r, sl = u % NS;
          const uint32_t full_parity = rc.parity();
          rc.next(nr);
          if (sl != bw) continue;                             // another team's unit
          dwq_wait(smem_u32(&s_bars.full[r]), full_parity, abort_flag);
          TQP(0);
          const float4* b_raw = reinterpret_cast<const float4*>(base + r * sbytes + boff);
          float4* b_lo = reinterpret_cast<float4*>(lo_base + sl * tq::DW_B_BYTES);
          float4 vb[16];
#ifndef DWQ_PROBE_NO_FEED
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) if (lane + kk * 32 < nb) vb[kk] = b_raw[lane + kk * 32];
#endif
          if (u >= NS) dwq_wait(smem_u32(&s_bars.sfree[sl]), (uint32_t)(u / NS - 1) & 1u, abort_flag);
          TQP(1);
#ifndef DWQ_PROBE_NO_FEED
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) if (lane + kk * 32 < nb) b_lo[lane + kk * 32] = lo_of(vb[kk]);
#endif
          tcp::fence_proxy_async_smem();                     // generic writes -> tensor core reads
          __syncwarp();
          if (lane == 0) tcp::mbar_arrive(smem_u32(&s_bars.ready[sl]));
          TQP(2);
        }
      }
    }
    if (lane == 0 && bw == 0) TQP_FLUSH(2, 11, 3);
  }
#ifdef APG_PROFILE
  tqp_k2_ = clock64();
#endif

  // ===================================================== tail: fold the conv Toeplitz block
  tcp::fence_before_thread_sync();
  __syncthreads();
  if (my_tiles > 0) {
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
    TQ_PROF_ARRAY[2][blockIdx.x][17] = tq_globaltimer();
    TQ_PROF_ARRAY[2][blockIdx.x][18] = clock64() - tqp_k0_;
  }
#endif
}

// grad[p] = scale * sum over CTAs of partials[c][p]: 32 parameters x 4 CTA slices per block of 128 threads (fixed
// order -> bitwise reproducible; the partials are already in torch order)
__global__ void __launch_bounds__(128) apg_reduce4_kernel(const float* __restrict__ partials, int ncta, int n,
                                                          float scale, float* __restrict__ grad) {
  tcp::griddep_wait();                                        // the partials of the kernel before are complete
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
  tcp::griddep_wait();                                        // the partials of the kernel before are complete
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
  APG_LAUNCH_PDL((n + 31) / 32, 128, 0, st, apg_reduce4_sgd_kernel)(partials, ncta, n, scale, grad, param, buf, lr, momentum);
  return cudaGetLastError();
}

cudaError_t launch_reduce_grad4(const float* partials, int ncta, int n, float scale, float* grad, cudaStream_t st) {
  APG_LAUNCH_PDL((n + 31) / 32, 128, 0, st, apg_reduce4_kernel)(partials, ncta, n, scale, grad);
  return cudaGetLastError();
}

cudaError_t launch_tq_dw(const HutterLayout& y, const RolloutArgs& a, const unsigned char* fstash,
                         const unsigned char* zstash, int grid, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(tq_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DWQ_SMEM);
  if (e != cudaSuccess) return e;
  APG_LAUNCH_PDL(grid, DWQ_THREADS, DWQ_SMEM, st, tq_dw_kernel)(y, a, fstash, zstash);
  return cudaGetLastError();
}

}  // namespace apg
