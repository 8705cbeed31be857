// tcgen05 / TMEM forward pass of the quadrotor concurrent policy Net(15, 10, 9, 40, conv=True) (B200, sm_100a).
// Standalone prototype of the round-2 forward kernel's policy phase; NOT part of the product.
//
//   s      = tanh(states_in(in_state))                  15 -> 64
//   c      = relu(conv1d(in_ref^T, 9 -> 20, k = 3))      (10 x 9) -> 20 x 8, flattened channel-major (c*8 + t)
//   h1..h3 = tanh(fc1([s, c])), tanh(fc2), tanh(fc3)     224 -> 64 -> 64 -> 64
//   a      = sigmoid(fc_out(h3))                         64 -> 40
// (reference: neural_control/models/hutter_model.py:12-49, sigmoid by the caller train_drone.py:151)
//
// Mapping
//   * tile = 128 drones = 128 TMEM lanes; epilogue thread r of a group owns drone r of the tile.
//   * every weight matrix lives in shared memory for the whole kernel as a (hi, lo) pair of K-major, unswizzled
//     TF32 images (223 KiB in total); 3xTF32: D = A_lo W_hi + A_hi W_lo + A_hi W_hi.
//   * activations never touch shared memory: tcgen05.ld -> bias/activation/split in registers -> tcgen05.st into
//     the A-operand columns of the next tcgen05.mma (A from TMEM).
//   * the conv is two output positions at a time: a 4-row window of in_ref (36 values) times a 40 x 36 Toeplitz
//     block (the same block for every pair of positions); fc1 is accumulated in five pieces (the s block and one
//     40-wide block per position pair, fc1's columns being permuted on the host to position-major order).
//   * warps 0-3 / 4-7: epilogue groups of TMEM slot 0 / 1 (two tiles in flight); warp 8 lane 0 issues every MMA.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tcgen05_policy tcgen05_policy.cu
// run:   ./tcgen05_policy [N=65536] [iters=20] [ctas=0 (= #SMs)]
//        ./tcgen05_policy 64 1 1 selftest      (no GPU: emulates the op list from the packed images on the host)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

constexpr int TM = 128;
constexpr int NTHREADS = 288;
constexpr int F0 = 15, H = 10, RD = 9, NC = 20, NPOS = 8, MO = 40, HID = 64;
constexpr int REFW = H * RD;  // 90 floats of in_ref per drone

// ---- shared-memory image table (bytes); every image is hi followed by lo
struct Img { int off, rows, K; };
__host__ __device__ constexpr int img_bytes(int rows, int K) { return rows * K * 4; }
constexpr Img I_WS{0, 64, 16};
constexpr Img I_W1S{I_WS.off + 2 * img_bytes(64, 16), 64, 64};
constexpr Img I_WT{I_W1S.off + 2 * img_bytes(64, 64), 48, 40};
constexpr Img I_W1G{I_WT.off + 2 * img_bytes(48, 40), 64, 40};  // 4 consecutive images (one per position pair)
constexpr Img I_W2{I_W1G.off + 4 * 2 * img_bytes(64, 40), 64, 64};
constexpr Img I_W3{I_W2.off + 2 * img_bytes(64, 64), 64, 64};
constexpr Img I_WO{I_W3.off + 2 * img_bytes(64, 64), 48, 64};
constexpr int IMG_TOTAL = I_WO.off + 2 * img_bytes(48, 64);
// biases (floats) after the images: bs 64 | bc 48 | b1 64 | b2 64 | b3 64 | bo 48
constexpr int B_S = 0, B_C = 64, B_1 = 112, B_2 = 176, B_3 = 240, B_O = 304, B_TOTAL = 352;
constexpr int SMEM_BYTES = 1024 + IMG_TOTAL + B_TOTAL * 4;
static_assert(SMEM_BYTES <= 232448 - 256, "weight images do not fit in shared memory");

// ---- TMEM columns inside one 256-column slot
constexpr int C_DMAIN = 0, C_DCONV = 64, C_AHI = 112, C_ALO = 176, SLOT_COLS = 256;

__host__ __device__ inline uint32_t kmajor_off(int r, int k, int K) {
  return (uint32_t)((r >> 3) * ((K >> 2) * 128) + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4);
}
__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ inline uint64_t kmajor_desc(uint32_t base, int ks, int K) {
  const uint32_t addr = base + ks * 256, lbo = 128, sbo = (K >> 2) * 128;
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__host__ __device__ inline uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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
__device__ inline void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
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
__device__ inline void tmem_ld8(uint32_t addr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(addr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}
// split 8 values into (hi, lo) and store them into the A-operand columns [col, col + 8)
__device__ inline void tmem_st8_split(uint32_t a_hi, uint32_t a_lo, const float* x) {
  uint32_t h[8], l[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    h[j] = __float_as_uint(x[j]) & 0xffffe000u;
    l[j] = __float_as_uint(x[j] - __uint_as_float(h[j]));
  }
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(a_hi), "r"(h[0]),
               "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7])
               : "memory");
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(a_lo), "r"(l[0]),
               "r"(l[1]), "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7])
               : "memory");
}
__device__ inline void a_operand_ready(uint32_t bar) {
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  mbar_arrive(bar);
}

struct Bars {
  unsigned long long a_ready[2];
  unsigned long long d_ready[2];
};

// one GEMM of the op list: D[d_col, +N) (=|+=) A[0, K) * W^T
struct Op { int img_off, rows, K, d_col, N, clear; };
constexpr int NOPS = 13;
__host__ __device__ inline Op op_of(int i) {
  if (i == 0) return {I_WS.off, 64, 16, C_DMAIN, 64, 1};
  if (i == 1) return {I_W1S.off, 64, 64, C_DMAIN, 64, 1};
  if (i < 10) {
    const int g = (i - 2) >> 1;
    if ((i & 1) == 0) return {I_WT.off, 48, 40, C_DCONV, 48, 1};
    return {I_W1G.off + g * 2 * img_bytes(64, 40), 64, 40, C_DMAIN, 64, 0};
  }
  if (i == 10) return {I_W2.off, 64, 64, C_DMAIN, 64, 1};
  if (i == 11) return {I_W3.off, 64, 64, C_DMAIN, 64, 1};
  return {I_WO.off, 48, 64, C_DMAIN, 48, 1};
}

// timing (per CTA 0 only): [0] MMA-thread cycles, [1] mbarrier timeouts (all CTAs), [2] MMA-thread wait cycles
__global__ void __launch_bounds__(NTHREADS, 1)
    policy_fwd_kernel(const unsigned char* __restrict__ images, const float* __restrict__ in_state,
                      const float* __restrict__ in_ref, float* __restrict__ actions, int n, long long* __restrict__ timing) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const float* s_bias = (const float*)(base + IMG_TOTAL);
  __shared__ __align__(8) Bars s_bars;
  __shared__ uint32_t s_tmem;
  __shared__ int s_timeouts;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < (IMG_TOTAL + B_TOTAL * 4) / 16; i += blockDim.x)
    ((uint4*)base)[i] = ((const uint4*)images)[i];
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&s_bars.a_ready[s]), 128);
      mbar_init(smem_u32(&s_bars.d_ready[s]), 1);
    }
    s_timeouts = 0;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)),
                 "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = s_tmem;
  const int ntiles = (n + TM - 1) / TM;
  // this CTA's tiles: blockIdx.x + j * gridDim.x, j = 0, 1, ...; tile j runs in slot j & 1
  const int my_tiles = (ntiles > (int)blockIdx.x) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 8) {
    if (lane == 0) {
      uint32_t par[2] = {0, 0};
      long long wait_cycles = 0;
      const long long t0 = clock64();
      for (int j0 = 0; j0 < my_tiles; j0 += 2)
        for (int i = 0; i < NOPS; ++i) {
          const Op op = op_of(i);
          const uint32_t idesc = idesc_tf32(TM, op.N);
          const uint32_t whi = smem_u32(base + op.img_off), wlo = whi + img_bytes(op.rows, op.K);
          for (int s = 0; s < 2; ++s) {
            if (j0 + s >= my_tiles) continue;
            const long long w0 = clock64();
            if (!mbar_wait(smem_u32(&s_bars.a_ready[s]), par[s])) atomicAdd(&s_timeouts, 1);
            par[s] ^= 1;
            wait_cycles += clock64() - w0;
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t slot = tmem + s * SLOT_COLS;
            const uint32_t d = slot + op.d_col, ahi = slot + C_AHI, alo = slot + C_ALO;
            for (int ks = 0; ks < op.K / 8; ++ks) {
              const uint64_t bh = kmajor_desc(whi, ks, op.K), bl = kmajor_desc(wlo, ks, op.K);
              mma_ts(d, alo + ks * 8, bh, idesc, (ks > 0 || !op.clear) ? 1u : 0u);
              mma_ts(d, ahi + ks * 8, bl, idesc, 1u);
              mma_ts(d, ahi + ks * 8, bh, idesc, 1u);
            }
            mma_commit(smem_u32(&s_bars.d_ready[s]));
          }
        }
      if (blockIdx.x == 0) {
        timing[0] = clock64() - t0;
        timing[2] = wait_cycles;
      }
    }
  } else {
    const int s = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t slot = tmem + s * SLOT_COLS + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t d_main = slot + C_DMAIN, d_conv = slot + C_DCONV, ahi = slot + C_AHI, alo = slot + C_ALO;
    const uint32_t bar_a = smem_u32(&s_bars.a_ready[s]), bar_d = smem_u32(&s_bars.d_ready[s]);
    uint32_t par = 0;
    auto wait_d = [&]() {
      if (!mbar_wait(bar_d, par)) atomicAdd(&s_timeouts, 1);
      par ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    };
    // D_main (64 columns) -> tanh(x + b) -> A operand
    auto dense_epilogue = [&](const float* b) {
#pragma unroll
      for (int c0 = 0; c0 < HID; c0 += 8) {
        float v[8];
        tmem_ld8(d_main + c0, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = tanhf(v[j] + b[c0 + j]);
        tmem_st8_split(ahi + c0, alo + c0, v);
      }
      a_operand_ready(bar_a);
    };
    for (int j = s; j < my_tiles; j += 2) {
      const int tile = (int)blockIdx.x + j * (int)gridDim.x;
      const int drone = tile * TM + row;
      const bool live = drone < n;
      // op 0 operand: in_state, K = 16 (column 15 zero)
      {
        float x[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = (live && k < F0) ? in_state[(size_t)drone * F0 + k] : 0.f;
        tmem_st8_split(ahi, alo, x);
        tmem_st8_split(ahi + 8, alo + 8, x + 8);
        a_operand_ready(bar_a);
      }
      wait_d();                      // op 0: states_in
      dense_epilogue(s_bias + B_S);  // -> operand of op 1 (fc1, s block)
      const float* rr = in_ref + (size_t)drone * REFW;
      for (int g = 0; g < 4; ++g) {
        wait_d();  // op 1 (g = 0) or the fc1 piece of the previous pair: the A columns are free again
        {
          float x[40];
#pragma unroll
          for (int k = 0; k < 36; k += 2) {
            const float2 t = live ? *(const float2*)(rr + 18 * g + k) : make_float2(0.f, 0.f);
            x[k] = t.x;
            x[k + 1] = t.y;
          }
          x[36] = x[37] = x[38] = x[39] = 0.f;
#pragma unroll
          for (int c0 = 0; c0 < 40; c0 += 8) tmem_st8_split(ahi + c0, alo + c0, x + c0);
          a_operand_ready(bar_a);
        }
        wait_d();  // conv of this position pair
        {
          const float* b = s_bias + B_C;
#pragma unroll
          for (int c0 = 0; c0 < 40; c0 += 8) {
            float v[8];
            tmem_ld8(d_conv + c0, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j] + b[c0 + j], 0.f);
            tmem_st8_split(ahi + c0, alo + c0, v);
          }
          a_operand_ready(bar_a);
        }
      }
      wait_d();  // last fc1 piece
      dense_epilogue(s_bias + B_1);
      wait_d();  // fc2
      dense_epilogue(s_bias + B_2);
      wait_d();  // fc3
      dense_epilogue(s_bias + B_3);
      wait_d();  // fc_out
      {
        const float* b = s_bias + B_O;
#pragma unroll
        for (int c0 = 0; c0 < MO; c0 += 8) {
          float v[8];
          tmem_ld8(d_main + c0, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = 1.f / (1.f + __expf(-(v[j] + b[c0 + j])));
          if (live) {
            float4* o = (float4*)(actions + (size_t)drone * MO + c0);
            o[0] = make_float4(v[0], v[1], v[2], v[3]);
            o[1] = make_float4(v[4], v[5], v[6], v[7]);
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512) : "memory");
  }
  if (tid == 0 && s_timeouts) atomicAdd((unsigned long long*)&timing[1], (unsigned long long)s_timeouts);
}

// ---------------------------------------------------------------- host
struct Net {  // torch layouts
  std::vector<float> ws, bs, wc, bc, w1, b1, w2, b2, w3, b3, wo, bo;
};

static void put_image(std::vector<unsigned char>& buf, Img im, int index, int rows_used, int k_used,
                      const std::vector<float>& dense /* rows_used x k_used */) {
  unsigned char* hi = buf.data() + im.off + index * 2 * img_bytes(im.rows, im.K);
  unsigned char* lo = hi + img_bytes(im.rows, im.K);
  for (int r = 0; r < rows_used; ++r)
    for (int k = 0; k < k_used; ++k) {
      const float x = dense[(size_t)r * k_used + k];
      uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u;
      float h; memcpy(&h, &u, 4);
      const float l = x - h;
      memcpy(hi + kmajor_off(r, k, im.K), &h, 4);
      memcpy(lo + kmajor_off(r, k, im.K), &l, 4);
    }
}

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 2;                                                                       \
    }                                                                                 \
  } while (0)

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 65536;
  const int iters = argc > 2 ? atoi(argv[2]) : 20;
  int ctas = argc > 3 ? atoi(argv[3]) : 0;
  uint32_t seed = 2024u;
  auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return ((seed >> 8) & 0xffff) / 32768.0f - 1.0f; };
  auto fill = [&](std::vector<float>& v, size_t cnt, float scale) { v.resize(cnt); for (auto& x : v) x = rnd() * scale; };
  Net net;
  fill(net.ws, 64 * 15, 0.258f); fill(net.bs, 64, 0.258f);        // ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in))
  fill(net.wc, 20 * 9 * 3, 0.192f); fill(net.bc, 20, 0.192f);
  fill(net.w1, 64 * 224, 0.0668f); fill(net.b1, 64, 0.0668f);
  fill(net.w2, 64 * 64, 0.125f); fill(net.b2, 64, 0.125f);
  fill(net.w3, 64 * 64, 0.125f); fill(net.b3, 64, 0.125f);
  fill(net.wo, 40 * 64, 0.125f); fill(net.bo, 40, 0.125f);
  std::vector<float> h_state((size_t)n * F0), h_ref((size_t)n * REFW), h_act((size_t)n * MO, 0.f);
  for (auto& x : h_state) x = rnd() * 2.f;
  for (auto& x : h_ref) x = rnd();

  // ---- pack the images
  std::vector<unsigned char> img(IMG_TOTAL + B_TOTAL * 4, 0);
  put_image(img, I_WS, 0, 64, 15, net.ws);
  {
    std::vector<float> w1s(64 * 64);
    for (int o = 0; o < 64; ++o) for (int k = 0; k < 64; ++k) w1s[o * 64 + k] = net.w1[o * 224 + k];
    put_image(img, I_W1S, 0, 64, 64, w1s);
  }
  {
    // Toeplitz block: row n = tl*20 + c, column k = tr*9 + ci, value w[c][ci][tr - tl]
    std::vector<float> wt(40 * 36, 0.f);
    for (int tl = 0; tl < 2; ++tl) for (int c = 0; c < NC; ++c) for (int ci = 0; ci < RD; ++ci) for (int jj = 0; jj < 3; ++jj)
      wt[(tl * 20 + c) * 36 + (tl + jj) * 9 + ci] = net.wc[(c * RD + ci) * 3 + jj];
    put_image(img, I_WT, 0, 40, 36, wt);
  }
  for (int g = 0; g < 4; ++g) {
    // fc1 columns of positions 2g, 2g+1 in the order of the conv block's outputs: k = tl*20 + c <-> 64 + c*8 + (2g + tl)
    std::vector<float> w1g(64 * 40);
    for (int o = 0; o < 64; ++o) for (int tl = 0; tl < 2; ++tl) for (int c = 0; c < NC; ++c)
      w1g[o * 40 + tl * 20 + c] = net.w1[o * 224 + 64 + c * NPOS + 2 * g + tl];
    put_image(img, I_W1G, g, 64, 40, w1g);
  }
  put_image(img, I_W2, 0, 64, 64, net.w2);
  put_image(img, I_W3, 0, 64, 64, net.w3);
  put_image(img, I_WO, 0, 40, 64, net.wo);
  {
    float* b = (float*)(img.data() + IMG_TOTAL);
    for (int i = 0; i < 64; ++i) { b[B_S + i] = net.bs[i]; b[B_1 + i] = net.b1[i]; b[B_2 + i] = net.b2[i]; b[B_3 + i] = net.b3[i]; }
    for (int tl = 0; tl < 2; ++tl) for (int c = 0; c < NC; ++c) b[B_C + tl * 20 + c] = net.bc[c];
    for (int i = 0; i < MO; ++i) b[B_O + i] = net.bo[i];
  }

  if (argc > 4 && !strcmp(argv[4], "selftest")) {
    // host emulation of the kernel's data flow from the PACKED images (validates offsets, Toeplitz block, fc1
    // permutation and the op list; needs no GPU)
    auto W = [&](const Op& op, int r, int k) {
      float h, l;
      memcpy(&h, img.data() + op.img_off + kmajor_off(r, k, op.K), 4);
      memcpy(&l, img.data() + op.img_off + img_bytes(op.rows, op.K) + kmajor_off(r, k, op.K), 4);
      return (double)h + (double)l;
    };
    const float* bias = (const float*)(img.data() + IMG_TOTAL);
    for (int d = 0; d < std::min(n, 64); ++d) {
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
      for (int k = 0; k < 16; ++k) A[k] = k < F0 ? h_state[(size_t)d * F0 + k] : 0.0;
      run(0);
      for (int k = 0; k < 64; ++k) A[k] = std::tanh(Dm[k] + bias[B_S + k]);
      run(1);
      for (int g = 0; g < 4; ++g) {
        for (int k = 0; k < 40; ++k) A[k] = k < 36 ? h_ref[(size_t)d * REFW + 18 * g + k] : 0.0;
        run(2 + 2 * g);
        for (int k = 0; k < 40; ++k) A[k] = std::max(Dc[k] + bias[B_C + k], 0.0);
        run(3 + 2 * g);
      }
      const int bo[3] = {B_1, B_2, B_3};
      for (int l = 0; l < 3; ++l) {
        for (int k = 0; k < 64; ++k) A[k] = std::tanh(Dm[k] + bias[bo[l] + k]);
        run(10 + l);
      }
      for (int o = 0; o < MO; ++o) h_act[(size_t)d * MO + o] = (float)(1.0 / (1.0 + std::exp(-(Dm[o] + bias[B_O + o]))));
    }
    // compare with the direct fp64 network below by falling through with n clipped
    printf("selftest: emulated %d drones from the packed images\n", std::min(n, 64));
  }
  const bool selftest = argc > 4 && !strcmp(argv[4], "selftest");
  cudaDeviceProp prop;
  if (!selftest) CK(cudaGetDeviceProperties(&prop, 0));
  if (ctas <= 0) ctas = selftest ? 1 : prop.multiProcessorCount;
  long long hT[4] = {0, 0, 0, 0};
  float ms = 1.f;
  if (!selftest) {
  unsigned char* d_img; float *d_state, *d_ref, *d_act; long long* d_t;
  CK(cudaMalloc(&d_img, img.size())); CK(cudaMalloc(&d_state, h_state.size() * 4)); CK(cudaMalloc(&d_ref, h_ref.size() * 4));
  CK(cudaMalloc(&d_act, h_act.size() * 4)); CK(cudaMalloc(&d_t, 4 * sizeof(long long)));
  CK(cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_state, h_state.data(), h_state.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_ref, h_ref.data(), h_ref.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_act, 0xff, h_act.size() * 4));
  CK(cudaMemset(d_t, 0, 4 * sizeof(long long)));
  CK(cudaFuncSetAttribute(policy_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  policy_fwd_kernel<<<ctas, NTHREADS, SMEM_BYTES>>>(d_img, d_state, d_ref, d_act, n, d_t);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h_act.data(), d_act, h_act.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hT, d_t, sizeof(hT), cudaMemcpyDeviceToHost));
  // ---- whole-launch timing with CUDA events
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) policy_fwd_kernel<<<ctas, NTHREADS, SMEM_BYTES>>>(d_img, d_state, d_ref, d_act, n, d_t);
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) policy_fwd_kernel<<<ctas, NTHREADS, SMEM_BYTES>>>(d_img, d_state, d_ref, d_act, n, d_t);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= iters;
  CK(cudaMemcpy(hT, d_t, sizeof(hT), cudaMemcpyDeviceToHost));
  }

  // ---- fp64 reference on a sample of drones (all of them up to 4096, then every 37th)
  double err = 0;
  long checked = 0;
  const int n_check = selftest ? std::min(n, 64) : n;
  for (int d = 0; d < n_check; d += (d < 4096 ? 1 : 37)) {
    double s[64], c[160], x[224], h1[64], h2[64], h3[64];
    for (int o = 0; o < 64; ++o) {
      double a = net.bs[o];
      for (int k = 0; k < F0; ++k) a += (double)net.ws[o * F0 + k] * h_state[(size_t)d * F0 + k];
      s[o] = std::tanh(a);
    }
    for (int ch = 0; ch < NC; ++ch)
      for (int t = 0; t < NPOS; ++t) {
        double a = net.bc[ch];
        for (int ci = 0; ci < RD; ++ci)
          for (int jj = 0; jj < 3; ++jj) a += (double)net.wc[(ch * RD + ci) * 3 + jj] * h_ref[(size_t)d * REFW + (t + jj) * RD + ci];
        c[ch * NPOS + t] = a > 0 ? a : 0;
      }
    for (int k = 0; k < 64; ++k) x[k] = s[k];
    for (int k = 0; k < 160; ++k) x[64 + k] = c[k];
    auto layer = [&](const std::vector<float>& w, const std::vector<float>& b, const double* in, int K, double* out) {
      for (int o = 0; o < 64; ++o) {
        double a = b[o];
        for (int k = 0; k < K; ++k) a += (double)w[(size_t)o * K + k] * in[k];
        out[o] = std::tanh(a);
      }
    };
    layer(net.w1, net.b1, x, 224, h1);
    layer(net.w2, net.b2, h1, 64, h2);
    layer(net.w3, net.b3, h2, 64, h3);
    for (int o = 0; o < MO; ++o) {
      double a = net.bo[o];
      for (int k = 0; k < 64; ++k) a += (double)net.wo[o * 64 + k] * h3[k];
      const double ref = 1.0 / (1.0 + std::exp(-a));
      const double e = std::fabs((double)h_act[(size_t)d * MO + o] - ref);
      err = (e >= 0) ? std::max(err, e) : 1e30;
    }
    ++checked;
  }

  const int ntiles = (n + TM - 1) / TM, tiles_cta0 = (ntiles - 1) / ctas + 1;
  const bool ok = err < 2e-6 && hT[1] == 0;
  printf("{\"prog\": \"policy_fwd\", \"n\": %d, \"ctas\": %d, \"checked_drones\": %ld, \"max_abs_err_vs_fp64\": %.3e, "
         "\"ms_per_launch\": %.4f, \"drones_per_s\": %.3e, \"useful_tflops\": %.2f, \"cycles_per_tile_cta0\": %.0f, "
         "\"mma_thread_wait_frac\": %.3f, \"mbarrier_timeouts\": %lld, \"smem_bytes\": %d, \"ok\": %s}\n",
         n, ctas, checked, err, ms, n / (ms * 1e-3), 2.0 * 30368 * n / (ms * 1e-3) / 1e12,
         (double)hT[0] / tiles_cta0, hT[0] ? (double)hT[2] / (double)hT[0] : 0.0, hT[1], SMEM_BYTES, ok ? "true" : "false");
  return ok ? 0 : 3;
}
