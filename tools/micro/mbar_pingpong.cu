// Round-trip latency of mbarrier hand-offs between two warps of one CTA (B200, sm_100a).  NOT part of the product.
// Warp 0 ("issuer"): waits a[i % D], then signals b[i % D] - by mbarrier.arrive (mode 0) or by tcgen05.commit (mode 1).
// Warp 1 ("feeder"): waits b[(i - D) % D] (ring of depth D), then lane 0 arrives on a[i % D].
// Reports cycles per iteration.  Extra warps (nbusy) spin on a barrier that never completes, like idle pipeline roles.
// usage: mbar_pingpong <mode 0|1> <depth> <iters> <nbusy> <wait 0 try_wait | 1 test_wait>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ inline void mbar_init(uint32_t bar, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(c) : "memory"); }
__device__ inline void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory"); }
__device__ inline bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ inline bool test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred q;\n\tmbarrier.test_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
template <int W> __device__ inline void wait(uint32_t bar, uint32_t parity) {
  if (W == 0) { while (!try_wait(bar, parity)) {} } else { while (!test_wait(bar, parity)) {} }
}
struct P { int mode, depth, iters, nbusy, wait; };

template <int W>
__global__ void __launch_bounds__(1024, 1) pp_kernel(P p, long long* out) {
  __shared__ __align__(8) unsigned long long a[16], b[16], never;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 16; ++i) { mbar_init(smem_u32(&a[i]), 1); mbar_init(smem_u32(&b[i]), 1); }
    mbar_init(smem_u32(&never), 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "n"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  __syncthreads();
  const int D = p.depth;
  if (warp == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < p.iters; ++i) {
      const int s = i % D;
      if (lane == 0) wait<W>(smem_u32(&a[s]), (uint32_t)(i / D) & 1u);
      __syncwarp();
      if (lane == 0) {
        if (p.mode == 0) mbar_arrive(smem_u32(&b[s]));
        else asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&b[s])) : "memory");
      }
      __syncwarp();
    }
    if (lane == 0) { out[blockIdx.x] = clock64() - t0; *(volatile int*)&never; mbar_arrive(smem_u32(&never)); }
  } else if (warp == 1) {
    for (int i = 0; i < p.iters; ++i) {
      const int s = i % D;
      if (i >= D) wait<W>(smem_u32(&b[s]), (uint32_t)(i / D - 1) & 1u);
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&a[s]));
    }
  } else if (warp < 2 + p.nbusy) {
    wait<W>(smem_u32(&never), 0);                     // all 32 lanes poll until the issuer is done
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(s_tmem), "n"(32) : "memory");
}

int main(int argc, char** argv) {
  if (argc < 6) { printf("usage\n"); return 1; }
  P p{atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5])};
  long long* d; cudaMalloc(&d, 148 * 8);
  for (int rep = 0; rep < 2; ++rep) {
    if (p.wait == 0) pp_kernel<0><<<148, 32 * (2 + p.nbusy)>>>(p, d); else pp_kernel<1><<<148, 32 * (2 + p.nbusy)>>>(p, d);
  }
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
  printf("{\"mode\": \"%s\", \"depth\": %d, \"nbusy\": %d, \"wait\": \"%s\", \"cuda\": \"%s\", \"cycles_per_iter\": %.1f}\n",
         p.mode ? "tcgen05.commit" : "mbarrier.arrive", p.depth, p.nbusy, p.wait ? "test_wait" : "try_wait", cudaGetErrorString(e), s / 148 / p.iters);
  return 0;
}
