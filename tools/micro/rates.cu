// microbenchmarks: FFMA issue rate, mma.sync tf32 m16n8k8 rate, LDS.128 / LDS.32 broadcast wavefront cost (B200)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ffma(float* out, int iters) {
  float a[16], x = threadIdx.x * 1e-3f, y = 1.0001f;
  for (int i = 0; i < 16; ++i) a[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], y, x);
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mma(float* out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a0 = 0x3f800000u + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 0x3f000000u, b1 = b0 + 5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mma_bf16(float* out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a0 = 0x3f803f80u, a1 = a0, a2 = a0, a3 = a0, b0 = 0x3f003f00u, b1 = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
__global__ void k_lds(float* out, int iters) {
  __shared__ __align__(16) float sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  float s = 0;
  const int lane = threadIdx.x & 31;
  int idx;
  if (MODE == 0) idx = 0;                    // LDS.128, all lanes same address
  else if (MODE == 1) idx = (lane & 7) * 4;  // LDS.128, 8 distinct chunks repeated over quarter-warps
  else if (MODE == 2) idx = lane * 4;        // LDS.128, 32 distinct chunks
  else if (MODE == 3) idx = 0;               // LDS.32 broadcast
  else idx = lane;                           // LDS.32 distinct
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int o = (idx + u * 128 + it * 4) & 4095 & ~3;
      if (MODE <= 2) { float4 v = *reinterpret_cast<const float4*>(sm + o); s += v.x + v.w; }
      else { s += sm[(idx + u * 128 + it) & 4095]; }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 1024);
  const int iters = 20000;
  for (int wps = 1; wps <= 4; wps *= 2) {   // warps per SMSP
    const int threads = 128 * wps;
    float ms = timeit([&] { k_ffma<<<sms, threads>>>(out, iters); });
    double cyc = ms * 1e-3 * khz * 1e3;
    printf("FFMA  %d warps/SMSP: %.1f FMA/clk/SM (%.2f cyc per warp-FFMA per SMSP)\n", wps, (double)iters * 16 * threads / cyc, cyc / ((double)iters * 16 * wps));
    ms = timeit([&] { k_mma<<<sms, threads>>>(out, iters / 4); });
    cyc = ms * 1e-3 * khz * 1e3;
    printf("MMA tf32 m16n8k8 %d warps/SMSP: %.0f MAC/clk/SM (%.2f cyc per mma per SMSP)\n", wps, (double)(iters / 4) * 8 * 1024 * (threads / 32) / cyc, cyc / ((double)(iters / 4) * 8 * wps));
    ms = timeit([&] { k_mma_bf16<<<sms, threads>>>(out, iters / 4); });
    cyc = ms * 1e-3 * khz * 1e3;
    printf("MMA bf16 m16n8k16 %d warps/SMSP: %.0f MAC/clk/SM (%.2f cyc per mma per SMSP)\n", wps, (double)(iters / 4) * 8 * 2048 * (threads / 32) / cyc, cyc / ((double)(iters / 4) * 8 * wps));
  }
  const char* names[5] = {"LDS.128 uniform", "LDS.128 8 chunks x4", "LDS.128 32 distinct", "LDS.32 uniform", "LDS.32 distinct"};
  for (int m = 0; m < 5; ++m) {
    float ms;
    if (m == 0) ms = timeit([&] { k_lds<0><<<sms, 256>>>(out, iters / 4); });
    if (m == 1) ms = timeit([&] { k_lds<1><<<sms, 256>>>(out, iters / 4); });
    if (m == 2) ms = timeit([&] { k_lds<2><<<sms, 256>>>(out, iters / 4); });
    if (m == 3) ms = timeit([&] { k_lds<3><<<sms, 256>>>(out, iters / 4); });
    if (m == 4) ms = timeit([&] { k_lds<4><<<sms, 256>>>(out, iters / 4); });
    double cyc = ms * 1e-3 * khz * 1e3;
    printf("%-22s: %.2f cycles per warp-load per SM\n", names[m], cyc / ((double)(iters / 4) * 8 * 8));
  }
  printf("clock %d kHz, %d SMs\n", khz, sms);
  return 0;
}
