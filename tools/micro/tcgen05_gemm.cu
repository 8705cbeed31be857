// tcgen05 / TMEM 3xTF32 GEMM microbenchmark + encoding check (B200, sm_100a). NOT part of the product.
//
// One CTA computes D[M x N] = A[M x K] * B[N x K]^T with tcgen05.mma.kind::tf32, accumulator in TMEM,
// read back with tcgen05.ld, and checks it against a CPU fp64 reference. It exists to pin down, on the
// real hardware and before the product kernels depend on them:
//   * the shared-memory matrix descriptor encodings (K-major no-swizzle, K-major 128B swizzle,
//     MN-major 128B swizzle) and the per-K-step start-address advance,
//   * the instruction descriptor for kind::tf32 (a/b format, major bits, M/N fields),
//   * the A-from-TMEM ("TS") form fed by tcgen05.st,
//   * the accumulator lane mapping for M = 64,
//   * the error of 1xTF32 and of the 3xTF32 split (hi = x & 0xffffe000, lo = x - hi),
//   * issue-rate (cycles per MMA) and the MMA -> commit -> mbarrier -> tcgen05.ld round-trip latency.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tcgen05_gemm tcgen05_gemm.cu
// run:   ./tcgen05_gemm <variant> [M=128] [N=64] [K=64] [reps=64] [swap_lbo_sbo=0]
//   variant 0: SS, A/B K-major, no swizzle, 1xTF32
//   variant 1: SS, A/B K-major, no swizzle, 3xTF32
//   variant 2: SS, A/B K-major, 128B swizzle, 3xTF32      (what a TMA box load of row-major fp32 produces)
//   variant 3: SS, A/B MN-major, 128B swizzle, 3xTF32     (the dW = dY^T X shape: K runs over the drones)
//   variant 4: TS, A in TMEM (tcgen05.st), B K-major no swizzle, 3xTF32
//   variant 5: SS, K-major no swizzle, 1xTF32 on RAW fp32 images (low 13 bits not cleared): does the tensor core
//              truncate or round fp32 -> tf32?  (compare the two "err_vs_*_inputs" fields)
//   variant 7: SS, A K-major no swizzle, B MN-major NO swizzle, 3xTF32: B's image is the K-major image of B^T, i.e. a
//              forward weight image W[out][in] serves dX = dZ W without a transposed copy (needs b_major = MN)
//   variant 8: the same with A in TMEM (the dX chain of the adjoint: dZ written by tcgen05.st)
//   variant 6: SS, K-major no swizzle, "2.5-term" split with a 16-bit correction: tf32(A_raw, B_raw) + tf32(A_lo, B_raw)
//              + bf16(A_bf16, B_lo_bf16), all three accumulating into the same fp32 TMEM tile (mixed kinds);
//              the B side then costs 4 + 2 bytes per weight instead of 4 + 4
// Each variant runs in its own process (a bad descriptor kills the context); tools/micro/run_tcgen05.sh loops.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

enum { LAY_K_NONE = 0, LAY_K_SW128 = 1, LAY_MN_SW128 = 2, LAY_MN_NONE = 3 };

struct Params {
  int M, N, K;
  int variant;
  int layA, layB;
  int split;       // 1 or 3 MMAs per K-step
  int a_in_tmem;   // TS form
  int reps;        // timing repetitions of the whole K loop
  int swap;        // swap the LBO and SBO fields (debug aid)
  int raw;         // hi images hold the unmasked fp32 value
  int mixed;       // variant 6
};

// byte offset of element (r, k) of an R x K operand inside its shared-memory image
__host__ __device__ inline uint32_t lay_off(int lay, int r, int k, int R, int K) {
  if (lay == LAY_K_NONE) {
    // core matrix = 8 rows x 16 B, stored contiguously (128 B); K-adjacent core matrices 128 B apart (LBO),
    // 8-row groups (K/4)*128 B apart (SBO)
    return (uint32_t)((r >> 3) * ((K >> 2) * 128) + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4);
  }
  if (lay == LAY_K_SW128) {
    // panels of 32 k-elements (128 B rows); inside a panel rows are 128 B apart, 8-row groups 1024 B apart (SBO);
    // 16-byte chunk index ^= row % 8
    uint32_t o = (uint32_t)((k >> 5) * (R * 128) + (r >> 3) * 1024 + (r & 7) * 128 + (k & 31) * 4);
    return o ^ (((o >> 7) & 7u) << 4);
  }
  if (lay == LAY_MN_NONE) {
    // MN-major, no swizzle (cute: ((1,n),(8,k)):((X,SBO),(1,LBO)) in 16-byte units): core matrix = 8 k-rows x 16 B
    // (4 mn elements); mn-adjacent core matrices 128 B apart (SBO), 8-k groups (R/4)*128 B apart (LBO).  This is
    // byte for byte the K-major unswizzled image of the TRANSPOSED operand ([k][mn] with mn contiguous), i.e. a
    // forward weight image W[out][in] read as the B operand of dX = dZ W (n = in, k = out) without a second copy.
    return (uint32_t)((k >> 3) * ((R >> 2) * 128) + (r >> 2) * 128 + (k & 7) * 16 + (r & 3) * 4);
  }
  // LAY_MN_SW128: panels of 32 mn-elements; inside a panel the k rows are 128 B apart, 8-k groups 1024 B apart (SBO),
  // panels K*128 B apart (LBO); 16-byte chunk index ^= k % 8
  uint32_t o = (uint32_t)((r >> 5) * (K * 128) + (k >> 3) * 1024 + (k & 7) * 128 + (r & 31) * 4);
  return o ^ (((o >> 7) & 7u) << 4);
}

__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp: SmemDescriptor)
__device__ inline uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, int layout_type, int swap) {
  if (swap) { uint32_t t = lbo_bytes; lbo_bytes = sbo_bytes; sbo_bytes = t; }
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
  d |= (uint64_t)(layout_type & 7) << 61;    // 0 none, 2 128B, 4 64B, 6 32B
  return d;
}

__device__ inline uint64_t operand_desc(int lay, uint32_t base, int ks, int R, int K, int swap) {
  if (lay == LAY_K_NONE) return make_desc(base + ks * 256, 128, (K >> 2) * 128, 0, swap);
  if (lay == LAY_K_SW128) return make_desc(base + (ks >> 2) * (R * 128) + (ks & 3) * 32, 16, 1024, 2, swap);
  if (lay == LAY_MN_NONE) return make_desc(base + ks * ((R >> 2) * 128), (R >> 2) * 128, 128, 0, swap);
  return make_desc(base + ks * 1024, K * 128, 1024, 2, swap);
}

// instruction descriptor (cute/arch/mma_sm100_desc.hpp: InstrDescriptor), kind::tf32, fp32 accumulate
__host__ __device__ inline uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                       // c_format = F32
  d |= 2u << 7;                       // a_format = TF32
  d |= 2u << 10;                      // b_format = TF32
  d |= (uint32_t)(a_mn_major & 1) << 15;
  d |= (uint32_t)(b_mn_major & 1) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

// kind::f16 with bf16 operands, fp32 accumulate; K = 16 per instruction
__host__ __device__ inline uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// 16-bit K-major no-swizzle image: core matrix 8 rows x 16 B (8 elements)
__host__ __device__ inline uint32_t lay_off16(int r, int k, int K) {
  return (uint32_t)((r >> 3) * ((K >> 3) * 128) + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}
__device__ inline void mma_ss_f16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ inline uint16_t f2bf(float x) {  // round to nearest even
  uint32_t u = __float_as_uint(x);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

__device__ inline void mma_ss(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ inline void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ inline void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ inline void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ inline bool mbar_wait(uint32_t bar, uint32_t parity) {
  for (int spin = 0; spin < (1 << 26); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ inline void tmem_ld8(uint32_t addr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(addr)
               : "memory");
}
__device__ inline void tmem_st8(uint32_t addr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(addr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

constexpr int TMEM_COLS = 512;  // whole TMEM: D in [0,256), A hi in [256,384), A lo in [384,512)

// timing slots: [0] issue+complete cycles for reps*K-loop, [1] single-MMA round trip, [2] tcgen05.ld of the tile,
// [3] mbarrier timeouts
__global__ void __launch_bounds__(128, 1)
    tc_gemm_kernel(Params p, const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                   long long* __restrict__ timing) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve: 1024-aligned operand images
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int a_bytes = 128 * p.K * 4;  // always a full 128-row image (rows >= M are zero)
  const int b_bytes = ((p.N + 31) / 32 * 32) * p.K * 4;
  unsigned char* sAhi = base;
  unsigned char* sAlo = sAhi + a_bytes;
  unsigned char* sBhi = sAlo + a_bytes;
  unsigned char* sBlo = sBhi + b_bytes;
  unsigned char* sA16 = sBlo + b_bytes;            // variant 6: bf16(A), 128 x K
  unsigned char* sB16 = sA16 + 128 * p.K * 2;      // variant 6: bf16(B - trunc(B)), RB x K
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar = smem_u32(&s_bar);

  // zero then fill the operand images (generic proxy), hi/lo split on the fly
  const int img_bytes = 2 * a_bytes + 2 * b_bytes + (p.mixed ? (128 + (p.N + 31) / 32 * 32) * p.K * 2 : 0);
  for (int i = tid; i < img_bytes / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0u;
  __syncthreads();
  for (int i = tid; i < p.M * p.K; i += blockDim.x) {
    const int r = i / p.K, k = i % p.K;
    const float x = A[i];
    const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    const uint32_t o = lay_off(p.layA, r, k, 128, p.K);
    *(float*)(sAhi + o) = p.raw ? x : hi;
    *(float*)(sAlo + o) = x - hi;
    if (p.mixed) *(uint16_t*)(sA16 + lay_off16(r, k, p.K)) = f2bf(x);
  }
  for (int i = tid; i < p.N * p.K; i += blockDim.x) {
    const int r = i / p.K, k = i % p.K;
    const float x = B[i];
    const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    const uint32_t o = lay_off(p.layB, r, k, (p.N + 31) / 32 * 32, p.K);
    *(float*)(sBhi + o) = p.raw ? x : hi;
    *(float*)(sBlo + o) = x - hi;
    if (p.mixed) *(uint16_t*)(sB16 + lay_off16(r, k, p.K)) = f2bf(x - hi);
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  // generic-proxy writes -> visible to the async proxy (tcgen05.mma reads smem through it)
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = s_tmem;
  const uint32_t tmem_d = tmem;  // columns [0, N)
  const uint32_t tmem_ahi = tmem + 256, tmem_alo = tmem + 384;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

  if (p.a_in_tmem) {
    // thread = row: write this row's K values (hi and lo) into TMEM columns
    const int r = warp * 32 + lane;
    for (int k0 = 0; k0 < p.K; k0 += 8) {
      uint32_t vh[8], vl[8];
      for (int j = 0; j < 8; ++j) {
        const float x = (r < p.M) ? A[r * p.K + k0 + j] : 0.f;
        const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
        vh[j] = __float_as_uint(hi);
        vl[j] = __float_as_uint(x - hi);
      }
      tmem_st8(tmem_ahi + lane_base + k0, vh);
      tmem_st8(tmem_alo + lane_base + k0, vl);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  }

  const uint32_t idesc = make_idesc(p.M, p.N, p.layA == LAY_MN_SW128 || p.layA == LAY_MN_NONE,
                                    p.layB == LAY_MN_SW128 || p.layB == LAY_MN_NONE);
  const int ksteps = p.K / 8;
  const int RB = (p.N + 31) / 32 * 32;
  uint32_t parity = 0;
  long long t_loop = 0, t_one = 0, t_ld = 0, timeouts = 0;

  auto issue_kloop = [&](bool first_clears) {
    if (p.mixed) {
      const uint32_t idesc16 = make_idesc_bf16(p.M, p.N);
      for (int ks = 0; ks < p.K / 16; ++ks) {  // 16-bit correction term first (smallest magnitude)
        const uint32_t sbo16 = (p.K >> 3) * 128;
        mma_ss_f16(tmem_d, make_desc(smem_u32(sA16) + ks * 256, 128, sbo16, 0, p.swap),
                   make_desc(smem_u32(sB16) + ks * 256, 128, sbo16, 0, p.swap), idesc16,
                   (ks > 0 || !first_clears) ? 1u : 0u);
      }
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t bh = operand_desc(p.layB, smem_u32(sBhi), ks, RB, p.K, p.swap);
        mma_ss(tmem_d, operand_desc(p.layA, smem_u32(sAlo), ks, 128, p.K, p.swap), bh, idesc, 1u);
        mma_ss(tmem_d, operand_desc(p.layA, smem_u32(sAhi), ks, 128, p.K, p.swap), bh, idesc, 1u);
      }
      return;
    }
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint64_t bh = operand_desc(p.layB, smem_u32(sBhi), ks, RB, p.K, p.swap);
      const uint64_t bl = operand_desc(p.layB, smem_u32(sBlo), ks, RB, p.K, p.swap);
      const uint32_t acc0 = (ks > 0 || !first_clears) ? 1u : 0u;
      if (p.a_in_tmem) {
        if (p.split == 3) {
          mma_ts(tmem_d, tmem_alo + ks * 8, bh, idesc, acc0);
          mma_ts(tmem_d, tmem_ahi + ks * 8, bl, idesc, 1u);
          mma_ts(tmem_d, tmem_ahi + ks * 8, bh, idesc, 1u);
        } else {
          mma_ts(tmem_d, tmem_ahi + ks * 8, bh, idesc, acc0);
        }
      } else {
        const uint64_t ah = operand_desc(p.layA, smem_u32(sAhi), ks, 128, p.K, p.swap);
        const uint64_t al = operand_desc(p.layA, smem_u32(sAlo), ks, 128, p.K, p.swap);
        if (p.split == 3) {
          mma_ss(tmem_d, al, bh, idesc, acc0);
          mma_ss(tmem_d, ah, bl, idesc, 1u);
          mma_ss(tmem_d, ah, bh, idesc, 1u);
        } else {
          mma_ss(tmem_d, ah, bh, idesc, acc0);
        }
      }
    }
  };

  if (tid == 0) {
    // (1) single-MMA round trip: issue, commit, wait
    long long t0 = clock64();
    {
      const uint64_t bh = operand_desc(p.layB, smem_u32(sBhi), 0, RB, p.K, p.swap);
      if (p.a_in_tmem) mma_ts(tmem_d, tmem_ahi, bh, idesc, 0u);
      else mma_ss(tmem_d, operand_desc(p.layA, smem_u32(sAhi), 0, 128, p.K, p.swap), bh, idesc, 0u);
    }
    mma_commit(bar);
    if (!mbar_wait(bar, parity)) ++timeouts;
    parity ^= 1;
    t_one = clock64() - t0;
    // (2) throughput: reps back-to-back K loops, one commit at the end
    t0 = clock64();
    for (int rep = 0; rep < p.reps; ++rep) issue_kloop(false);
    mma_commit(bar);
    if (!mbar_wait(bar, parity)) ++timeouts;
    parity ^= 1;
    t_loop = clock64() - t0;
    // (3) the checked product
    issue_kloop(true);
    mma_commit(bar);
    if (!mbar_wait(bar, parity)) ++timeouts;
    parity ^= 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

  // epilogue: thread = TMEM lane; dump all 128 lanes x N columns (host sorts out the M = 64 lane mapping)
  {
    const long long t0 = clock64();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < p.N; c0 += 8) {
      uint32_t v[8];
      tmem_ld8(tmem_d + lane_base + c0, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      for (int j = 0; j < 8; ++j) D[row * p.N + c0 + j] = __uint_as_float(v[j]);
    }
    if (tid == 0) t_ld = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
  if (tid == 0) {
    timing[0] = t_loop;
    timing[1] = t_one;
    timing[2] = t_ld;
    timing[3] = timeouts;
  }
}

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);           \
      return 2;                                                                                 \
    }                                                                                           \
  } while (0)

int main(int argc, char** argv) {
  Params p{};
  p.variant = argc > 1 ? atoi(argv[1]) : 0;
  p.M = argc > 2 ? atoi(argv[2]) : 128;
  p.N = argc > 3 ? atoi(argv[3]) : 64;
  p.K = argc > 4 ? atoi(argv[4]) : 64;
  p.reps = argc > 5 ? atoi(argv[5]) : 64;
  p.swap = argc > 6 ? atoi(argv[6]) : 0;
  p.split = (p.variant == 0 || p.variant == 5) ? 1 : 3;
  p.raw = (p.variant == 5 || p.variant == 6);
  p.mixed = p.variant == 6;
  p.a_in_tmem = p.variant == 4;
  p.layA = p.layB = LAY_K_NONE;
  if (p.variant == 2) p.layA = p.layB = LAY_K_SW128;
  if (p.variant == 3) p.layA = p.layB = LAY_MN_SW128;
  if (p.variant == 7 || p.variant == 8) p.layB = LAY_MN_NONE;      // A K-major (smem / TMEM), B MN-major unswizzled
  if (p.variant == 8) p.a_in_tmem = 1;
  if (!(p.M == 64 || p.M == 128) || p.N % 16 || p.N < 16 || p.N > 256 || p.K % 32 || p.K < 32 || p.K > 128) {
    printf("unsupported shape: M in {64,128}, N %% 16 == 0 <= 256, K %% 32 == 0 <= 128\n");
    return 1;
  }
  const int RB = (p.N + 31) / 32 * 32;
  const size_t smem = 1024 + 2 * (size_t)128 * p.K * 4 + 2 * (size_t)RB * p.K * 4 + (size_t)(128 + RB) * p.K * 2;
  if (smem > 227 * 1024) { printf("operands do not fit in shared memory (%zu B)\n", smem); return 1; }

  std::vector<float> hA((size_t)p.M * p.K), hB((size_t)p.N * p.K), hD((size_t)128 * p.N, 0.f);
  uint32_t s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 32768.0f - 1.0f; };
  for (auto& x : hA) x = rnd();
  for (auto& x : hB) x = rnd();

  float *dA, *dB, *dD; long long* dT;
  CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dB, hB.size() * 4)); CK(cudaMalloc(&dD, hD.size() * 4));
  CK(cudaMalloc(&dT, 4 * sizeof(long long)));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, hD.size() * 4));
  CK(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_gemm_kernel<<<1, 128, smem>>>(p, dA, dB, dD, dT);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  long long hT[4];
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hT, dT, sizeof(hT), cudaMemcpyDeviceToHost));

  // CPU references: exact fp64 product, and the product of the tf32-truncated operands
  std::vector<double> ref((size_t)p.M * p.N), ref_t((size_t)p.M * p.N), ref_r((size_t)p.M * p.N);
  auto rne = [](float x) {
    uint32_t u; memcpy(&u, &x, 4); u += 0xfffu + ((u >> 13) & 1u); u &= 0xffffe000u; float y; memcpy(&y, &u, 4); return y;
  };
  auto trunc = [](float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; float y; memcpy(&y, &u, 4); return y; };
  for (int m = 0; m < p.M; ++m)
    for (int n = 0; n < p.N; ++n) {
      double a = 0, b = 0, c = 0;
      for (int k = 0; k < p.K; ++k) {
        a += (double)hA[m * p.K + k] * (double)hB[n * p.K + k];
        b += (double)trunc(hA[m * p.K + k]) * (double)trunc(hB[n * p.K + k]);
        c += (double)rne(hA[m * p.K + k]) * (double)rne(hB[n * p.K + k]);
      }
      ref[(size_t)m * p.N + n] = a;
      ref_t[(size_t)m * p.N + n] = b;
      ref_r[(size_t)m * p.N + n] = c;
    }
  // candidate accumulator lane mappings (row m -> TMEM lane)
  struct Map { const char* name; int (*f)(int); };
  const Map maps[] = {{"lane=m", [](int m) { return m; }},
                      {"lane=(m/16)*32+m%16", [](int m) { return (m / 16) * 32 + m % 16; }},
                      {"lane=(m/32)*64+m%32", [](int m) { return (m / 32) * 64 + m % 32; }}};
  double best = 1e30; const char* best_name = "?"; double best_t = 0, best_r = 0;
  for (const Map& mp : maps) {
    if (p.M == 128 && mp.f(127) != 127) continue;
    double e = 0, et = 0, er = 0;
    for (int m = 0; m < p.M; ++m)
      for (int n = 0; n < p.N; ++n) {
        const double d = hD[(size_t)mp.f(m) * p.N + n];
        e = std::max(e, std::fabs(d - ref[(size_t)m * p.N + n]));
        et = std::max(et, std::fabs(d - ref_t[(size_t)m * p.N + n]));
        er = std::max(er, std::fabs(d - ref_r[(size_t)m * p.N + n]));
      }
    if (!(e >= 0)) e = 1e30;  // NaN
    if (e < best) { best = e; best_name = mp.name; best_t = et; best_r = er; }
  }
  const int mmas = p.mixed ? (p.K / 8) * 2 + p.K / 16 : (p.K / 8) * p.split;
  const double cyc_per_mma = (double)hT[0] / ((double)mmas * p.reps);
  const double mac_per_clk = (double)p.M * p.N * 8 / cyc_per_mma;
  // expectation: |err| ~ K * 2^-11 for 1xTF32, ~ K * 2^-21 for 3xTF32 (inputs in [-1,1))
  const double tol = (p.split == 3 ? 4e-6 : 4e-3) * p.K / 64.0 * 4;
  const bool ok = best < tol && hT[3] == 0;
  printf("{\"variant\": %d, \"M\": %d, \"N\": %d, \"K\": %d, \"split\": %d, \"a_in_tmem\": %d, \"swap\": %d, "
         "\"max_abs_err_vs_fp64\": %.3e, \"max_abs_err_vs_truncated_inputs\": %.3e, "
         "\"max_abs_err_vs_rne_inputs\": %.3e, \"lane_map\": \"%s\", "
         "\"cycles_per_mma\": %.1f, \"mac_per_clk_sm\": %.0f, \"single_mma_roundtrip_cycles\": %lld, "
         "\"tile_tmem_ld_cycles\": %lld, \"mbarrier_timeouts\": %lld, \"ok\": %s}\n",
         p.variant, p.M, p.N, p.K, p.split, p.a_in_tmem, p.swap, best, best_t, best_r, best_name, cyc_per_mma, mac_per_clk,
         hT[1], hT[2], hT[3], ok ? "true" : "false");
  return ok ? 0 : 3;
}
