// Rate probe for tcgen05.mma kind::tf32 (B200, sm_100a).  NOT part of the product.
// One thread issues COUNT MMAs (M = 128, K = 8, given N) back to back and commits; cycles from the first issue to the
// completion of the commit / COUNT = sustained cost of one MMA.  Parameters: A from TMEM or shared memory, how many
// DISTINCT accumulators the MMAs rotate over (1 = every MMA depends on the previous one), swizzled or plain B.
// usage: tcgen05_rate <ts 0|1> <N> <ndist> <count> <sw128 0|1> [grid=148]
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tcgen05_rate tcgen05_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ inline uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, int layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}
__host__ __device__ inline uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
struct P { int ts, N, ndist, count, sw; };

template <int TS>
__device__ __forceinline__ void mma1(uint32_t dcol, uint32_t a_t, uint64_t da, uint64_t db, uint32_t idesc) {
  if (TS) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, q;\n\t}\n" ::"r"(dcol), "r"(a_t), "l"(db),
                 "r"(idesc), "r"(1u) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, q;\n\t}\n" ::"r"(dcol), "l"(da), "l"(db),
                 "r"(idesc), "r"(1u) : "memory");
  }
}

template <int TS, int N, int NDIST>
__global__ void __launch_bounds__(128, 1) rate_kernel(P p, long long* out) {
  extern __shared__ __align__(1024) unsigned char raw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar = smem_u32(&s_bar);
  for (int i = tid; i < 65536 / 4; i += blockDim.x) ((float*)base)[i] = 0.f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = s_tmem;
  if (warp == 0) {
    // warp-uniform loop, one elected lane issues (the form the product kernels use)
    uint32_t leader;
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}\n" : "=r"(leader));
    const uint32_t idesc = make_idesc(128, N);
    const uint64_t da = p.sw ? make_desc(smem_u32(base), 16, 1024, 2) : make_desc(smem_u32(base), 128, 256, 0);
    const uint64_t db = p.sw ? make_desc(smem_u32(base + 32768), 16, 1024, 2) : make_desc(smem_u32(base + 32768), 128, 256, 0);
    const uint32_t a_t = tmem + 480;
    const long long t0 = clock64();
    for (int i = 0; i < p.count; i += 16) {
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (leader) mma1<TS>(tmem + (k % NDIST) * N, a_t + (k & 3) * 8, da + 2 * (k & 3), db + 2 * (k & 3), idesc);
    }
    const long long t1 = clock64();
    if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
    int ok = 0;
    for (int spin = 0; spin < (1 << 26) && !ok; ++spin) {
      asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                   : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
    }
    const long long t2 = clock64();
    if (tid == 0) {
      out[blockIdx.x * 2] = t1 - t0;
      out[blockIdx.x * 2 + 1] = ok ? t2 - t0 : -1;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512) : "memory");
}

template <int TS, int N, int NDIST>
void run(P p, long long* d, int grid, int smem) {
  cudaFuncSetAttribute(rate_kernel<TS, N, NDIST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 3; ++rep) rate_kernel<TS, N, NDIST><<<grid, 128, smem>>>(p, d);
}
template <int TS>
bool dispatch(P p, long long* d, int grid, int smem) {
#define CASE(n, nd) if (p.N == n && p.ndist == nd) { run<TS, n, nd>(p, d, grid, smem); return true; }
  CASE(16, 1) CASE(16, 4) CASE(32, 1) CASE(32, 4) CASE(48, 1) CASE(48, 2) CASE(48, 4) CASE(64, 1) CASE(64, 2) CASE(64, 4)
  CASE(128, 1) CASE(128, 2) CASE(256, 1)
#undef CASE
  return false;
}

int main(int argc, char** argv) {
  if (argc < 6) { printf("usage: %s ts N ndist count sw128 [grid]\n", argv[0]); return 1; }
  P p{atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5])};
  const int grid = argc > 6 ? atoi(argv[6]) : 148;
  if (p.ndist * p.N > 480) { printf("{\"error\": \"ndist * N > 480\"}\n"); return 1; }
  long long* d;
  cudaMalloc(&d, grid * 16);
  const int smem = 1024 + 65536;
  if (!(p.ts ? dispatch<1>(p, d, grid, smem) : dispatch<0>(p, d, grid, smem))) { printf("{\"error\": \"no such case\"}\n"); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(grid * 2);
  cudaMemcpy(h.data(), d, grid * 16, cudaMemcpyDeviceToHost);
  double issue = 0, total = 0;
  for (int i = 0; i < grid; ++i) { issue += h[2 * i]; total += h[2 * i + 1]; }
  printf("{\"ts\": %d, \"N\": %d, \"ndist\": %d, \"count\": %d, \"sw128\": %d, \"grid\": %d, \"cuda\": \"%s\", "
         "\"issue_cycles_per_mma\": %.1f, \"total_cycles_per_mma\": %.1f}\n",
         p.ts, p.N, p.ndist, p.count, p.sw, grid, cudaGetErrorString(e), issue / grid / p.count, total / grid / p.count);
  return 0;
}
