// Address-function probe for tcgen05.mma shared-memory descriptors (B200, sm_100a).  NOT part of the product.
//
// One MMA (kind::tf32, M = 128, K = 8).  The probed operand's shared-memory region is filled with its own WORD INDEX
// (two passes: idx & 1023 and idx >> 10, both exact in TF32); the other operand is a unit matrix (e[k] in row k), so
//     probe B:  D[m][n] = B(n, k = m)   for m < 8     -> the word the hardware fetched for element (n, k)
//     probe A:  D[m][n] = A(m, k = n)   for n < 8
// i.e. the output IS the address function of the descriptor under test.  The host prints it as byte offsets and
// fits  off(mn, k) = c0*(mn%4)... against the candidate canonical forms.
//
// usage: tcgen05_probe <which: A|B> <major: 0 K / 1 MN> <layout_type 0|2|4|6> <lbo_bytes> <sbo_bytes> [N=64]
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tcgen05_probe tcgen05_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
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
__host__ __device__ inline uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a_mn & 1) << 15) | ((uint32_t)(b_mn & 1) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct P { int probeA, major, lt, lbo, sbo, N; };
constexpr int REGION = 65536;   // bytes per operand region

__global__ void __launch_bounds__(128, 1) probe_kernel(P p, float* D, int* status) {
  extern __shared__ __align__(1024) unsigned char raw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sP = base;              // probed operand region
  unsigned char* sU = base + REGION;     // unit operand, K-major unswizzled, K = 8: (r>>3)*256 + (k>>2)*128 + (r&7)*16 + (k&3)*4
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar = smem_u32(&s_bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = s_tmem;
  uint32_t parity = 0;
  const int rowsU = p.probeA ? p.N : 128;
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = tid; i < REGION / 4; i += blockDim.x) ((float*)sP)[i] = (float)(pass == 0 ? (i & 1023) : (i >> 10));
    for (int i = tid; i < rowsU * 8; i += blockDim.x) {
      const int r = i / 8, k = i % 8;
      *(float*)(sU + (r >> 3) * 256 + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4) = (r == k) ? 1.f : 0.f;
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      const uint64_t dP = make_desc(smem_u32(sP), p.lbo, p.sbo, p.lt);
      const uint64_t dU = make_desc(smem_u32(sU), 128, 256, 0);
      const uint32_t idesc = p.probeA ? make_idesc(128, p.N, p.major, 0) : make_idesc(128, p.N, 0, p.major);
      const uint64_t a = p.probeA ? dP : dU, b = p.probeA ? dU : dP;
      asm volatile(
          "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, q;\n\t}\n" ::"r"(tmem + pass * 256),
          "l"(a), "l"(b), "r"(idesc), "r"(0u)
          : "memory");
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
      int ok = 0;
      for (int spin = 0; spin < (1 << 24) && !ok; ++spin) {
        asm volatile(
            "{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
      }
      if (!ok) status[0] = 1;
      parity ^= 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  }
  const int row = warp * 32 + lane;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  for (int pass = 0; pass < 2; ++pass)
    for (int c0 = 0; c0 < p.N; c0 += 8) {
      uint32_t v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(tmem + pass * 256 + lane_base + c0)
                   : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      for (int j = 0; j < 8; ++j) D[(pass * 128 + row) * 256 + c0 + j] = __uint_as_float(v[j]);
    }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512) : "memory");
}

int main(int argc, char** argv) {
  if (argc < 6) { printf("usage: %s A|B major lt lbo sbo [N]\n", argv[0]); return 1; }
  P p{};
  p.probeA = argv[1][0] == 'A';
  p.major = atoi(argv[2]); p.lt = atoi(argv[3]); p.lbo = atoi(argv[4]); p.sbo = atoi(argv[5]);
  p.N = argc > 6 ? atoi(argv[6]) : 64;
  float* dD; int* dS;
  cudaMalloc(&dD, 2 * 128 * 256 * 4); cudaMalloc(&dS, 4);
  cudaMemset(dD, 0xff, 2 * 128 * 256 * 4); cudaMemset(dS, 0, 4);
  const int smem = 1024 + 2 * REGION;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem>>>(p, dD, dS);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> h(2 * 128 * 256);
  int st = 0;
  cudaMemcpy(h.data(), dD, h.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
  printf("{\"probe\": \"%s\", \"major\": %d, \"layout_type\": %d, \"lbo\": %d, \"sbo\": %d, \"N\": %d, \"cuda\": \"%s\", "
         "\"timeout\": %d, \"byte_off\": [",
         p.probeA ? "A" : "B", p.major, p.lt, p.lbo, p.sbo, p.N, cudaGetErrorString(e), st);
  // element (mn, k): probe B -> D[k][mn]; probe A -> D[mn][k]
  const int MN = p.probeA ? 128 : p.N;
  for (int mn = 0; mn < MN; ++mn) {
    printf("%s[", mn ? ", " : "");
    for (int k = 0; k < 8; ++k) {
      const int r = p.probeA ? mn : k, c = p.probeA ? k : mn;
      const float lo = h[(0 * 128 + r) * 256 + c], hi = h[(1 * 128 + r) * 256 + c];
      long off = (lo == lo && hi == hi) ? ((long)hi * 1024 + (long)lo) * 4 : -1;
      printf("%s%ld", k ? "," : "", off);
    }
    printf("]");
  }
  printf("]}\n");
  return 0;
}
