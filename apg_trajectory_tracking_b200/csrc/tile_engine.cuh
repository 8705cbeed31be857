// Tile engine: the building blocks every rollout kernel is made of.
//
// One CTA (256 threads, one per SM, persistent) works on a tile of TM = 64 drones.  Activations of a tile live in
// shared memory "feature-major": row = feature, TMP = 68 floats per row (64 drones + 4 pad so that lanes that
// walk consecutive rows hit distinct 16-byte bank groups).  Input tiles arrive in their natural drone-major
// (AoS) layout by bulk async copies (TMA 1-D, cp.async.bulk + mbarrier) and are consumed in place.
//
//  tensor path (contractions with K and N multiples of 8 -- all the large layers):
//    dense_mma()  Y[n][d] = epi( b[n] + sum_k A[k][d] * W[k][n] )      mma.sync m16n8k8 TF32, 3xTF32 split (fp32-level
//    dw_mma()     dW[j][k] += sum_d dZ[j][d] * X[k][d]                 accuracy), warp tile 16 drones x 32 outputs
//  FFMA path (odd-sized layers: cartpole net, LSTM cell, AR fc_out):
//    dense()      same contraction, register tile 4 drones x 4 outputs per thread
//    dw_T() / dw_AoS()  same dW, 8 rows x NKI*32 columns per warp (X feature-major / drone-major)
//  dense_auto() / dw_auto() pick the path.  Forward layers read W packed [in][out]; dX reads the [out][in] copy with
//  A = dZ and multiplies by the stored activation's derivative in place.  fp32 parity (1e-5 on the loss) is required,
//  hence 3xTF32 and not plain TF32; see DESIGN.md section 3.
#pragma once
#ifndef APG_SIM
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include "layouts.h"
// -DAPG_SIM (tests only): every PTX helper below calls the software model of tests/hostcheck/te_sim.h instead, so
// that the unchanged kernel sources built on this header run on the CPU (one OS thread per GPU thread).

// ---- optional per-phase cycle accounting (build with -DAPG_PROFILE; tools/phase_profile.py) ----------------
#ifdef APG_PROFILE
#define APG_NPROF 24
#define PROF_DECL long long prof_t0_ = clock64(); long long prof_acc_[APG_NPROF]; \
  for (int i_ = 0; i_ < APG_NPROF; ++i_) prof_acc_[i_] = 0;
#define PROF(i) do { if (threadIdx.x == 0) { const long long t_ = clock64(); prof_acc_[i] += t_ - prof_t0_; prof_t0_ = t_; } } while (0)
#define PROF_FLUSH(k) do { if (threadIdx.x == 0 && blockIdx.x < 148) for (int i_ = 0; i_ < APG_NPROF; ++i_) g_apg_prof[k][blockIdx.x][i_] = prof_acc_[i_]; } while (0)
#else
#define PROF_DECL
#define PROF(i) do { } while (0)
#define PROF_FLUSH(k) do { } while (0)
#endif

// the CTA's dynamic shared memory as a float array (the simulator hands out its own buffer)
#ifdef APG_SIM
#define APG_DYNAMIC_SMEM_F32(name) float* name = reinterpret_cast<float*>(::simte::dynamic_smem())
#else
#define APG_DYNAMIC_SMEM_F32(name) extern __shared__ __align__(128) float name[]
#endif

namespace apg {


enum Act { ACT_NONE = 0, ACT_TANH = 1, ACT_RELU = 2, ACT_SIGMOID = 3 };
enum Epi { EPI_ACT = 0, EPI_DTANH = 1, EPI_DRELU = 2, EPI_DNONE = 3, EPI_DSIGMOID = 4 };

// ------------------------------------------------------------------------------------------------------------
// mbarrier + bulk async copy (TMA 1-D) helpers
// ------------------------------------------------------------------------------------------------------------
#ifdef APG_SIM
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { simte::mbar_init(bar, count); }
__device__ __forceinline__ void fence_mbar_init() {}
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { simte::mbar_arrive(bar); }
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#endif
// barrier of the two dynamics warps (64 threads) of the warp-specialised kernels
__device__ __forceinline__ void dyn_group_sync() {
#ifdef APG_SIM
  simte::named_barrier_64();
#else
  asm volatile("bar.sync 2, 64;" ::: "memory");
#endif
}
// barrier of the GEMM warp group (threads 0..255) in the warp-specialised kernels; plain __syncthreads otherwise
template <bool WS>
__device__ __forceinline__ void gsync() {
#ifdef APG_SIM
  if (WS) simte::named_barrier_256();
#else
  if (WS) asm volatile("bar.sync 1, 256;" ::: "memory");
#endif
  else __syncthreads();
}
#ifdef APG_SIM
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { simte::mbar_expect_tx(bar, bytes); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { simte::mbar_wait(bar, parity); }
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  simte::bulk_g2s(dst_smem, src_gmem, bytes, bar);
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  simte::bulk_s2g(dst_gmem, src_smem, bytes);
}
__device__ __forceinline__ void bulk_commit() { simte::bulk_commit(); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { simte::bulk_wait(N); }
__device__ __forceinline__ void bulk_wait_all() { simte::bulk_wait(0); }
#else
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// Spin on the phase parity.  A bounded spin that traps instead of hanging the GPU if a copy never lands.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
// global -> shared, completion signalled on the mbarrier (bytes multiple of 16, both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// shared -> global (bulk group completion)
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
#endif

// Copy `bytes` (multiple of 16) global -> shared in chunks, all signalled on one barrier. Call from ONE thread,
// after mbar_expect_tx(bar, bytes).
__device__ __forceinline__ void bulk_g2s_chunked(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  constexpr uint32_t CH = 32768;
  for (uint32_t o = 0; o < bytes; o += CH) {
    const uint32_t n = bytes - o < CH ? bytes - o : CH;
    bulk_g2s(static_cast<char*>(dst) + o, static_cast<const char*>(src) + o, n, bar);
  }
}

// Accumulate into the CTA's gradient partial without waiting for the old value: red.global.add.f32.  Every address is
// owned by exactly one thread of one CTA for the whole launch (same-address reductions of one thread retire in
// program order), so the result is deterministic although the instruction is an atomic.
__device__ __forceinline__ void red_add(float* addr, float v) {
#ifdef APG_SIM
  simte::red_add(addr, v);
#else
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
#endif
}

// Destination column of gradient column k.  Conv activations are kept position-major (64 + t*20 + c) in shared
// memory / the stash while the reference's fc1 weight is channel-major (64 + c*npos + t): perm_npos > 0 maps back.
// two adjacent floats (8-byte aligned) in one reduction: the (col 2t, 2t+1) pair of an mma C fragment
__device__ __forceinline__ void red_add2(float* addr, float a, float b) {
#ifdef APG_SIM
  simte::red_add(addr, a);
  simte::red_add(addr + 1, b);
#else
  asm volatile("red.global.v2.f32.add [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
#endif
}

__device__ __forceinline__ int perm_col(int k, int perm_npos) {
  if (perm_npos <= 0 || k < 64) return k;
  const int tt = (k - 64) / 20, c = (k - 64) - tt * 20;
  return 64 + c * perm_npos + tt;
}

// ------------------------------------------------------------------------------------------------------------
// thread mapping of the 4x4 register-tile GEMM: a warp spans 4 drone groups x 8 output groups
// ------------------------------------------------------------------------------------------------------------
struct Lane {
  int lane, warp, dg, og0;
  __device__ __forceinline__ Lane() {
    lane = threadIdx.x & 31;
    warp = threadIdx.x >> 5;
    dg = (lane >> 3) + 4 * (warp & 3);     // 0..15 : drones 4*dg .. 4*dg+3
    og0 = (lane & 7) + 8 * (warp >> 2);    // 0..15 : outputs 4*og .. 4*og+3
  }
};

// A operand sources: value of feature k for the 4 drones of drone-group dg
struct SrcT {   // feature-major smem tile [K][TMP]
  const float* base;
  __device__ __forceinline__ float4 ld(int k, int dg) const {
    return *reinterpret_cast<const float4*>(base + k * TMP + 4 * dg);
  }
};
struct SrcAoS {  // drone-major smem tile [TM][ld], features start at column `off`
  const float* base;
  int ld_, off;
  __device__ __forceinline__ float4 ld(int k, int dg) const {
    const float* p = base + (4 * dg) * ld_ + off + k;
    return make_float4(p[0], p[ld_], p[2 * ld_], p[3 * ld_]);
  }
};

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == ACT_TANH) return tanhf(v);
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  return v;
}

// acc[i][j] += A[k][drone i] * W[k][4*og + j]  over k in [0,K)
// `Wcol` points at column 4*og of row 0; with sw != 0 the matrix is stored XOR-swizzled (col ^ ((row&3)<<3)), which
// keeps every aligned group of 4 columns contiguous.
template <class Src>
__device__ __forceinline__ void mac_tile(float (&acc)[4][4], const Src& A, int K, const float* __restrict__ Wcol,
                                         int ldw, int dg, int sw = 0, int col0 = 0) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float4 a = A.ld(k, dg);
    const int cs = sw ? ((col0 ^ ((k & 3) << 3)) - col0) : 0;
    const float4 w = *reinterpret_cast<const float4*>(Wcol + k * ldw + cs);
    acc[0][0] = fmaf(a.x, w.x, acc[0][0]); acc[0][1] = fmaf(a.x, w.y, acc[0][1]);
    acc[0][2] = fmaf(a.x, w.z, acc[0][2]); acc[0][3] = fmaf(a.x, w.w, acc[0][3]);
    acc[1][0] = fmaf(a.y, w.x, acc[1][0]); acc[1][1] = fmaf(a.y, w.y, acc[1][1]);
    acc[1][2] = fmaf(a.y, w.z, acc[1][2]); acc[1][3] = fmaf(a.y, w.w, acc[1][3]);
    acc[2][0] = fmaf(a.z, w.x, acc[2][0]); acc[2][1] = fmaf(a.z, w.y, acc[2][1]);
    acc[2][2] = fmaf(a.z, w.z, acc[2][2]); acc[2][3] = fmaf(a.z, w.w, acc[2][3]);
    acc[3][0] = fmaf(a.w, w.x, acc[3][0]); acc[3][1] = fmaf(a.w, w.y, acc[3][1]);
    acc[3][2] = fmaf(a.w, w.z, acc[3][2]); acc[3][3] = fmaf(a.w, w.w, acc[3][3]);
  }
}

// Epilogue of one 4x4 tile: outputs o = 4*og + j go to row (row0 + o*row_stride) of the feature-major tile Y.
//   EPI_ACT    : Y = act(acc + bias)
//   EPI_D*     : Y = acc * act'(Y_old)   (in-place backward through the activation whose OUTPUT is stored in Y)
template <int EPI>
__device__ __forceinline__ void store_tile(const float (&acc)[4][4], const float* __restrict__ bias, int og,
                                           float* Y, int row0, int row_stride, int act, int dg) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int o = 4 * og + j;
    float* yp = Y + (row0 + o * row_stride) * TMP + 4 * dg;
    float4 v = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
    if (EPI == EPI_ACT) {
      const float b = bias ? bias[o] : 0.f;
      v.x = act_apply(v.x + b, act); v.y = act_apply(v.y + b, act);
      v.z = act_apply(v.z + b, act); v.w = act_apply(v.w + b, act);
    } else if (EPI == EPI_DTANH) {
      const float4 y = *reinterpret_cast<const float4*>(yp);
      v.x *= 1.f - y.x * y.x; v.y *= 1.f - y.y * y.y; v.z *= 1.f - y.z * y.z; v.w *= 1.f - y.w * y.w;
    } else if (EPI == EPI_DRELU) {
      const float4 y = *reinterpret_cast<const float4*>(yp);
      v.x = y.x > 0.f ? v.x : 0.f; v.y = y.y > 0.f ? v.y : 0.f;
      v.z = y.z > 0.f ? v.z : 0.f; v.w = y.w > 0.f ? v.w : 0.f;
    } else if (EPI == EPI_DSIGMOID) {
      const float4 y = *reinterpret_cast<const float4*>(yp);
      v.x *= y.x * (1.f - y.x); v.y *= y.y * (1.f - y.y); v.z *= y.z * (1.f - y.z); v.w *= y.w * (1.f - y.w);
    }
    *reinterpret_cast<float4*>(yp) = v;
  }
}

// Generic dense layer on a tile.  W is [K][ldw] in shared memory with ldw >= 4*M4 (zero-padded columns),
// M4 = number of 4-wide output groups.  All 256 threads call it; no internal barrier.
template <class Src, int EPI>
__device__ __forceinline__ void dense(const Lane& L, const Src& A, int K, const float* __restrict__ W, int ldw,
                                      const float* __restrict__ bias, int M4, float* Y, int row0, int row_stride,
                                      int act, int sw = 0, int wcol0 = 0) {
  for (int og = L.og0; og < M4; og += 16) {
    float acc[4][4] = {};
    mac_tile(acc, A, K, W + wcol0 + 4 * og, ldw, L.dg, sw, wcol0 + 4 * og);
    store_tile<EPI>(acc, bias, og, Y, row0, row_stride, act, L.dg);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Tensor-core variants: mma.sync m16n8k8 TF32 with the 3xTF32 split (x = hi + lo, hi = top 19 bits;
// a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, fp32 accumulate) -> fp32-level accuracy (~2^-21 relative) at three
// tensor instructions per tile, and a quarter of the shared-memory operand traffic of the FFMA tiles.
// Fragment coordinates (g = lane>>2, t = lane&3): A(16x8): a0 (g,t) a1 (g+8,t) a2 (g,t+4) a3 (g+8,t+4);
// B(8x8): b0 (k=t,n=g) b1 (k=t+4,n=g); C(16x8): c0 (g,2t) c1 (g,2t+1) c2 (g+8,2t) c3 (g+8,2t+1).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
#ifdef APG_SIM
  simte::mma_m16n8k8_tf32(c, a, b0, b1);
#else
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
#endif
}
// Y[row0 + n][d] = epi(bias[n] + sum_k A[k][d] * W[k][n]) for the 64 drones of the tile; A feature-major [K][TMP],
// W [K][ldw] (swizzled when sw), K % 8 == 0, N % 8 == 0.  Warp w: drones 16*(w&3) .., groups of 4 n-tiles
// alternate between the two warp halves.
template <int EPI>
__device__ __forceinline__ void dense_mma(const Lane& L, const float* __restrict__ A, int K,
                                          const float* __restrict__ W, int ldw, int sw, const float* __restrict__ bias,
                                          int N, float* Y, int row0, int act, int wcol0 = 0) {
  const int g = L.lane >> 2, t = L.lane & 3;
  const int m0 = (L.warp & 3) * 16;
  const int nt8 = N >> 3;
  const int xs = sw ? (t << 3) : 0;
  for (int nc = (L.warp >> 2); nc * 4 < nt8; nc += 2) {
    const int nt0 = nc * 4;
    const int ntc = nt8 - nt0 < 4 ? nt8 - nt0 : 4;
    // one accumulator chain per 3xTF32 term (lo*hi, hi*lo, hi*hi): a dependent tensor instruction only every 12th
    float acc[4][4], acc1[4][4], acc2[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[j][e] = acc1[j][e] = acc2[j][e] = 0.f;
    const float* ap = A + t * TMP + m0 + g;
    const float* wp = W + t * ldw;
    // column of B fragment j; tiles beyond ntc are clamped onto a valid tile (computed, then discarded): predicated
    // mma.sync would cost a WARPSYNC + NOP per instruction
    int ncol[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) ncol[j] = (wcol0 + (nt0 + (j < ntc ? j : 0)) * 8 + g) ^ xs;
#pragma unroll 2
    for (int k0 = 0; k0 < K; k0 += 8) {
      uint32_t ah[4], al[4];
      split_tf32(ap[k0 * TMP], ah[0], al[0]);
      split_tf32(ap[k0 * TMP + 8], ah[1], al[1]);
      split_tf32(ap[(k0 + 4) * TMP], ah[2], al[2]);
      split_tf32(ap[(k0 + 4) * TMP + 8], ah[3], al[3]);
      uint32_t bh[4][2], bl[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        split_tf32(wp[k0 * ldw + ncol[j]], bh[j][0], bl[j][0]);
        split_tf32(wp[(k0 + 4) * ldw + ncol[j]], bh[j][1], bl[j][1]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) mma_tf32(acc1[j], al, bh[j][0], bh[j][1]);
#pragma unroll
      for (int j = 0; j < 4; ++j) mma_tf32(acc2[j], ah, bl[j][0], bl[j][1]);
#pragma unroll
      for (int j = 0; j < 4; ++j) mma_tf32(acc[j], ah, bh[j][0], bh[j][1]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[j][e] += acc1[j][e] + acc2[j][e];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < ntc) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int n = (nt0 + j) * 8 + 2 * t + (e & 1);
          const int d = m0 + g + ((e >> 1) << 3);
          float* yp = Y + (row0 + n) * TMP + d;
          float v = acc[j][e];
          if (EPI == EPI_ACT) {
            v = act_apply(v + (bias ? bias[n] : 0.f), act);
          } else if (EPI == EPI_DTANH) {
            const float y = *yp;
            v *= 1.f - y * y;
          } else if (EPI == EPI_DRELU) {
            v = *yp > 0.f ? v : 0.f;
          } else if (EPI == EPI_DSIGMOID) {
            const float y = *yp;
            v *= y * (1.f - y);
          }
          *yp = v;
        }
      }
    }
  }
}

// db[j] += sum_d dZ[j][d]  (thread j < M)
__device__ __forceinline__ void bias_grad(const float* __restrict__ dz, int M, float* __restrict__ Pb) {
  for (int j = threadIdx.x; j < M; j += NT) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < TM / 4; ++d4) {
      const float4 z = *reinterpret_cast<const float4*>(dz + j * TMP + 4 * d4);
      s0 += z.x + z.y;
      s1 += z.z + z.w;
    }
    red_add(Pb + j, s0 + s1);
  }
}

// dW[j][k] += sum_d dZ[j][d] X[k][d]  (both feature-major, reduction over the 64 drones), M rows, K % 8 == 0 columns.
// Warp w: 16-row tile (w & 3) (+4, ...), the 8-column tiles split between the two warp halves; NT n-tiles per pass.
template <int NTP>
__device__ __forceinline__ void dw_mma(const Lane& L, const float* __restrict__ dz, int M, const float* __restrict__ x,
                                       int K, float* __restrict__ P, int ldp_, int perm_npos = 0) {
  const int g = L.lane >> 2, t = L.lane & 3;
  const int nt8 = K >> 3;
  const int per_half = (nt8 + 1) >> 1;
  for (int mt = (L.warp & 3); mt * 16 < M; mt += 4) {
    const int j0 = mt * 16;
    const int nb = (L.warp >> 2) * per_half;
    const int ne = nb + per_half < nt8 ? nb + per_half : nt8;
    for (int n0 = nb; n0 < ne; n0 += NTP) {
      float acc[NTP][4], acc1[NTP][4], acc2[NTP][4];
#pragma unroll
      for (int j = 0; j < NTP; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][e] = acc1[j][e] = acc2[j][e] = 0.f;
      // rows beyond M (last 16-row tile of e.g. M = 40) and column tiles beyond `ne` are clamped onto valid data and
      // their results discarded below: every mma.sync stays unconditional
      const float* zp = dz + (j0 + g < M ? j0 + g : M - 1) * TMP + t;
      const float* zq = dz + (j0 + g + 8 < M ? j0 + g + 8 : M - 1) * TMP + t;
      const float* xp[NTP];
#pragma unroll
      for (int j = 0; j < NTP; ++j) xp[j] = x + ((n0 + j < ne ? n0 + j : n0) * 8 + g) * TMP + t;
#pragma unroll 2
      for (int d0 = 0; d0 < TM; d0 += 8) {
        uint32_t ah[4], al[4];
        split_tf32(zp[d0], ah[0], al[0]);
        split_tf32(zq[d0], ah[1], al[1]);
        split_tf32(zp[d0 + 4], ah[2], al[2]);
        split_tf32(zq[d0 + 4], ah[3], al[3]);
        uint32_t bh[NTP][2], bl[NTP][2];
#pragma unroll
        for (int j = 0; j < NTP; ++j) {
          split_tf32(xp[j][d0], bh[j][0], bl[j][0]);
          split_tf32(xp[j][d0 + 4], bh[j][1], bl[j][1]);
        }
#pragma unroll
        for (int j = 0; j < NTP; ++j) mma_tf32(acc1[j], al, bh[j][0], bh[j][1]);
#pragma unroll
        for (int j = 0; j < NTP; ++j) mma_tf32(acc2[j], ah, bl[j][0], bl[j][1]);
#pragma unroll
        for (int j = 0; j < NTP; ++j) mma_tf32(acc[j], ah, bh[j][0], bh[j][1]);
      }
      // The CTA's partial keeps the KERNEL's column order (fragment pairs adjacent -> one 8-byte vector reduction, a
      // warp instruction touches 8 sectors instead of 32); apg_reduce_kernel maps to the torch order.
      const bool vec = (ldp_ & 1) == 0;
#pragma unroll
      for (int j = 0; j < NTP; ++j) {
        if (n0 + j < ne) {
          const int k = (n0 + j) * 8 + 2 * t;
          const float c0 = acc[j][0] + (acc1[j][0] + acc2[j][0]), c1 = acc[j][1] + (acc1[j][1] + acc2[j][1]);
          const float c2 = acc[j][2] + (acc1[j][2] + acc2[j][2]), c3 = acc[j][3] + (acc1[j][3] + acc2[j][3]);
          if (j0 + g < M) {
            float* a0 = P + (j0 + g) * ldp_ + k;
            if (vec) red_add2(a0, c0, c1); else { red_add(a0, c0); red_add(a0 + 1, c1); }
          }
          if (j0 + g + 8 < M) {
            float* a1 = P + (j0 + g + 8) * ldp_ + k;
            if (vec) red_add2(a1, c2, c3); else { red_add(a1, c2); red_add(a1 + 1, c3); }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// weight-gradient contractions.  P points at the CTA's partial-gradient matrix in global memory (torch layout
// [M][ldp]); each (j,k) entry is owned by exactly one thread -> plain read-modify-write, deterministic.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float dot4(const float4& a, const float4& b, float c) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, c))));
}

// dW[j][k] += sum_d dZ[j][d] X[k][d], X feature-major [K][TMP]; K <= 32*NKI.  Also db[j] += sum_d dZ[j][d].
template <int NKI>
__device__ __forceinline__ void dw_T(const Lane& L, const float* __restrict__ dz, int M, const float* __restrict__ x,
                                     int K, float* __restrict__ P, int ldp, float* __restrict__ Pb, int perm_npos = 0,
                                     int col0 = 0) {
  for (int j0 = 8 * L.warp; j0 < M; j0 += 8 * NWARP) {
    float acc[8][NKI];
    float accb[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      accb[jj] = 0.f;
#pragma unroll
      for (int i = 0; i < NKI; ++i) acc[jj][i] = 0.f;
    }
#pragma unroll 2
    for (int d4 = 0; d4 < TM / 4; ++d4) {
      float4 xv[NKI];
#pragma unroll
      for (int i = 0; i < NKI; ++i) {
        const int k = L.lane + 32 * i;
        xv[i] = k < K ? *reinterpret_cast<const float4*>(x + k * TMP + 4 * d4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float4 z = (j0 + jj < M) ? *reinterpret_cast<const float4*>(dz + (j0 + jj) * TMP + 4 * d4)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
        accb[jj] += (z.x + z.y) + (z.z + z.w);
#pragma unroll
        for (int i = 0; i < NKI; ++i) acc[jj][i] = dot4(z, xv[i], acc[jj][i]);
      }
    }
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int j = j0 + jj;
      if (j < M) {
#pragma unroll
        for (int i = 0; i < NKI; ++i) {
          const int k = L.lane + 32 * i;
          if (k < K) red_add(P + j * ldp + col0 + k, acc[jj][i]);
        }
        if (Pb && L.lane == 0) red_add(Pb + j, accb[jj]);
      }
    }
  }
}

// Same with X drone-major: X[k][d] = xa[d*ld + off + k], K <= 32.
__device__ __forceinline__ void dw_AoS(const Lane& L, const float* __restrict__ dz, int M, const float* __restrict__ xa,
                                       int ld, int off, int K, float* __restrict__ P, int ldp,
                                       float* __restrict__ Pb) {
  for (int j0 = 8 * L.warp; j0 < M; j0 += 8 * NWARP) {
    float acc[8], accb[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) acc[jj] = accb[jj] = 0.f;
    const bool kin = L.lane < K;
    for (int d4 = 0; d4 < TM / 4; ++d4) {
      const float* xp = xa + (4 * d4) * ld + off + L.lane;
      const float4 xv = kin ? make_float4(xp[0], xp[ld], xp[2 * ld], xp[3 * ld]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float4 z = (j0 + jj < M) ? *reinterpret_cast<const float4*>(dz + (j0 + jj) * TMP + 4 * d4)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
        accb[jj] += (z.x + z.y) + (z.z + z.w);
        acc[jj] = dot4(z, xv, acc[jj]);
      }
    }
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int j = j0 + jj;
      if (j < M) {
        if (kin) red_add(P + j * ldp + L.lane, acc[jj]);
        if (Pb && L.lane == 0) red_add(Pb + j, accb[jj]);
      }
    }
  }
}

// Dispatchers: tensor-core path when the contraction is 8-aligned, FFMA tiles otherwise (small / odd layers).
template <int EPI>
__device__ __forceinline__ void dense_auto(const Lane& L, const float* __restrict__ A, int K,
                                           const float* __restrict__ W, int ldw, int sw, const float* __restrict__ bias,
                                           int N, float* Y, int row0, int act, int wcol0 = 0) {
  if ((((K | N) & 7) == 0))
    dense_mma<EPI>(L, A, K, W, ldw, sw, bias, N, Y, row0, act, wcol0);
  else
    dense<SrcT, EPI>(L, SrcT{A}, K, W, ldw, bias, N / 4, Y, row0, 1, act, sw, wcol0);
}


__device__ __forceinline__ void dw_auto(const Lane& L, const float* __restrict__ dz, int M, const float* __restrict__ x,
                                        int K, float* __restrict__ P, int ldp, float* __restrict__ Pb,
                                        int perm_npos = 0) {
  if ((K & 7) == 0 && M >= 16) {
    bias_grad(dz, M, Pb);
    dw_mma<4>(L, dz, M, x, K, P, ldp, perm_npos);
    return;
  }
  // FFMA fallback in column blocks of 64 (keeps the register footprint of the unaligned path small)
  for (int c0 = 0; c0 < K; c0 += 64) {
    const int kc = K - c0 < 64 ? K - c0 : 64;
    if (kc <= 32) dw_T<1>(L, dz, M, x + c0 * TMP, kc, P, ldp, c0 == 0 ? Pb : nullptr, perm_npos, c0);
    else dw_T<2>(L, dz, M, x + c0 * TMP, kc, P, ldp, c0 == 0 ? Pb : nullptr, perm_npos, c0);
  }
}

// block-wide sum of one float per thread (result valid in thread 0); `scratch` = NWARP floats of shared memory
__device__ __forceinline__ float block_sum(float v, float* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < NWARP; ++w) r += scratch[w];
  }
  __syncthreads();
  return r;
}

}  // namespace apg
