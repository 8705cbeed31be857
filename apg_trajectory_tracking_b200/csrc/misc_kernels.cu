// Small helper kernels: weight packing, deterministic partial reductions, single dynamics steps.
#include "apg_math.cuh"
#include "kernels.h"

namespace apg {

__device__ __forceinline__ int pack_perm(int k, int npos) {     // position-major fc1 input index -> torch column
  if (npos <= 0 || k < 64) return k;
  const int tt = (k - 64) / 20, c = (k - 64) - tt * 20;
  return 64 + c * npos + tt;
}

// dst = layout transform of src (see PackMode); zero-fills padding.  One block per segment chunk.
__global__ void apg_pack_kernel(const PackTable t, const float* __restrict__ params, float* __restrict__ wf,
                                float* __restrict__ wb) {
  for (int s = blockIdx.y; s < t.n; s += gridDim.y) {
    const PackSeg g = t.seg[s];
    const float* src = params + g.src;
    float* dst = (g.which ? wb : wf) + g.dst;
    int total;
    if (g.mode == PK_COPY_PAD) total = g.rows * g.wcols;
    else if (g.mode == PK_CONV_BWD) total = g.rows * g.ldd;
    else if (g.mode == PK_TRANSPOSE) total = g.cols * g.wcols;
    else total = g.cols * g.ldd;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      float v = 0.f;
      int di = i;
      if (g.mode == PK_COPY_PAD) {                 // dst[r][c], c < wcols
        const int r = i / g.wcols, c = i - r * g.wcols;
        if (c < g.cols) v = src[r * g.sld + pack_perm(c, g.perm)];
        di = r * g.ldd + (g.sw ? (c ^ ((r & 3) << 3)) : c);
      } else if (g.mode == PK_TRANSPOSE) {         // dst[c][r], r < wcols  (src [rows][cols])
        const int c = i / g.wcols, r = i - c * g.wcols;
        if (r < g.rows) v = src[r * g.sld + pack_perm(c, g.perm)];
        di = c * g.ldd + (g.sw ? (r ^ ((c & 3) << 3)) : r);
      } else if (g.mode == PK_CONV_FWD) {          // src [C=rows][RD][3] -> dst[kk = j*RD + d][c], cols = 3*RD
        const int kk = i / g.ldd, c = i - kk * g.ldd;
        const int rd = g.cols / 3, j = kk / rd, d = kk - j * rd;
        if (c < g.rows) v = src[c * g.cols + d * 3 + j];
      } else {                                     // PK_CONV_BWD: dst[c][kk], kk < ldd
        const int c = i / g.ldd, kk = i - c * g.ldd;
        const int rd = g.cols / 3;
        if (kk < g.cols) { const int j = kk / rd, d = kk - j * rd; v = src[c * g.cols + d * 3 + j]; }
      }
      dst[di] = v;
    }
  }
}

cudaError_t launch_pack(const PackTable& t, const float* params, float* wf, float* wb, cudaStream_t st) {
  dim3 grid(8, t.n);
  APG_LAUNCH(grid, 256, 0, st, apg_pack_kernel)(t, params, wf, wb);
  return cudaGetLastError();
}

// grad[p] = scale * sum_c partials[c][p], fixed summation order -> bitwise reproducible
// The kernels keep the fc1 weight-gradient block [64][K1] of the hutter conv nets in their position-major column
// order (64 + t*20 + c); pm_* describe that block so that the torch (channel-major, 64 + c*npos + t) entry p reads
// the right partial column.  pm_npos == 0: no permutation.
__global__ void apg_reduce_kernel(const float* __restrict__ partials, int ncta, int n, float scale,
                                  float* __restrict__ grad, int pm_off, int pm_k1, int pm_npos) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  int q = p;
  if (pm_npos > 0 && p >= pm_off && p < pm_off + 64 * pm_k1) {
    const int j = (p - pm_off) / pm_k1, k = (p - pm_off) - j * pm_k1;
    if (k >= 64) {
      const int c = (k - 64) / pm_npos, tt = (k - 64) - c * pm_npos;
      q = pm_off + j * pm_k1 + 64 + tt * 20 + c;
    }
  }
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int c = 0;
  for (; c + 3 < ncta; c += 4) {
    s0 += partials[(size_t)(c + 0) * n + q];
    s1 += partials[(size_t)(c + 1) * n + q];
    s2 += partials[(size_t)(c + 2) * n + q];
    s3 += partials[(size_t)(c + 3) * n + q];
  }
  for (; c < ncta; ++c) s0 += partials[(size_t)c * n + q];
  grad[p] = scale * ((s0 + s1) + (s2 + s3));
}

cudaError_t launch_reduce_grad(const float* partials, int ncta, int n, float scale, float* grad, cudaStream_t st,
                               int pm_off, int pm_k1, int pm_npos) {
  APG_LAUNCH((n + 127) / 128, 128, 0, st, apg_reduce_kernel)(partials, ncta, n, scale, grad, pm_off, pm_k1, pm_npos);
  return cudaGetLastError();
}

__global__ void apg_sum_loss_kernel(const float* __restrict__ partials, int ncta, float* __restrict__ loss) {
  // one warp, fixed order; accumulate in double (148 partials of ~1e5 magnitude)
  double s = 0.0;
  for (int c = threadIdx.x; c < ncta; c += 32) s += (double)partials[c];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) *loss = (float)s;
}

cudaError_t launch_sum_loss(const float* partials, int ncta, float* loss, cudaStream_t st) {
  APG_LAUNCH(1, 32, 0, st, apg_sum_loss_kernel)(partials, ncta, loss);
  return cudaGetLastError();
}

// ---- single dynamics steps (un-fused callers: Dynamics.__call__)
template <template <typename> class SysT>
__global__ void apg_step_kernel(const PhysConsts pc, const float* __restrict__ s, const float* __restrict__ a, float dt,
                                int n, float* __restrict__ out) {
  using Sys = SysT<float>;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float si[Sys::S], ai[Sys::A], o[Sys::S];
#pragma unroll
  for (int j = 0; j < Sys::S; ++j) si[j] = s[(size_t)i * Sys::S + j];
#pragma unroll
  for (int j = 0; j < Sys::A; ++j) ai[j] = a[(size_t)i * Sys::A + j];
  Sys::step(si, ai, dt, pc.v, o);
#pragma unroll
  for (int j = 0; j < Sys::S; ++j) out[(size_t)i * Sys::S + j] = o[j];
}

template <template <typename> class SysT>
__global__ void apg_step_adj_kernel(const PhysConsts pc, const float* __restrict__ s, const float* __restrict__ a,
                                    float dt, int n, const float* __restrict__ g, float* __restrict__ gs,
                                    float* __restrict__ ga) {
  using Sys = SysT<float>;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float si[Sys::S], ai[Sys::A], gi[Sys::S], gso[Sys::S], gao[Sys::A];
#pragma unroll
  for (int j = 0; j < Sys::S; ++j) { si[j] = s[(size_t)i * Sys::S + j]; gi[j] = g[(size_t)i * Sys::S + j]; }
#pragma unroll
  for (int j = 0; j < Sys::A; ++j) ai[j] = a[(size_t)i * Sys::A + j];
  Sys::step_adj(si, ai, dt, pc.v, gi, gso, gao);
#pragma unroll
  for (int j = 0; j < Sys::S; ++j) gs[(size_t)i * Sys::S + j] = gso[j];
#pragma unroll
  for (int j = 0; j < Sys::A; ++j) ga[(size_t)i * Sys::A + j] = gao[j];
}

cudaError_t launch_step(int system, const PhysConsts& pc, const float* s, const float* a, float dt, int n, float* out,
                        cudaStream_t st) {
  const int b = 128, gr = (n + b - 1) / b;
  if (n <= 0) return cudaSuccess;
  if (system == SYS_QUAD) APG_LAUNCH(gr, b, 0, st, apg_step_kernel<Quad>)(pc, s, a, dt, n, out);
  else if (system == SYS_WING) APG_LAUNCH(gr, b, 0, st, apg_step_kernel<Wing>)(pc, s, a, dt, n, out);
  else if (system == SYS_CARTPOLE) APG_LAUNCH(gr, b, 0, st, apg_step_kernel<Cartpole>)(pc, s, a, dt, n, out);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_step_adj(int system, const PhysConsts& pc, const float* s, const float* a, float dt, int n,
                            const float* g, float* gs, float* ga, cudaStream_t st) {
  const int b = 128, gr = (n + b - 1) / b;
  if (n <= 0) return cudaSuccess;
  if (system == SYS_QUAD) APG_LAUNCH(gr, b, 0, st, apg_step_adj_kernel<Quad>)(pc, s, a, dt, n, g, gs, ga);
  else if (system == SYS_WING) APG_LAUNCH(gr, b, 0, st, apg_step_adj_kernel<Wing>)(pc, s, a, dt, n, g, gs, ga);
  else if (system == SYS_CARTPOLE) APG_LAUNCH(gr, b, 0, st, apg_step_adj_kernel<Cartpole>)(pc, s, a, dt, n, g, gs, ga);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

__global__ void apg_features_kernel(const float* __restrict__ s, int n, float* __restrict__ f) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float si[12], fi[15];
#pragma unroll
  for (int j = 0; j < 12; ++j) si[j] = s[(size_t)i * 12 + j];
  Quad<float>::features(si, fi);
#pragma unroll
  for (int j = 0; j < 15; ++j) f[(size_t)i * 15 + j] = fi[j];
}

__global__ void apg_features_adj_kernel(const float* __restrict__ s, const float* __restrict__ gf, int n,
                                        float* __restrict__ gs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float si[12], gi[15], go[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) { si[j] = s[(size_t)i * 12 + j]; go[j] = 0.f; }
#pragma unroll
  for (int j = 0; j < 15; ++j) gi[j] = gf[(size_t)i * 15 + j];
  Quad<float>::features_adj(si, gi, go);
#pragma unroll
  for (int j = 0; j < 12; ++j) gs[(size_t)i * 12 + j] = go[j];
}

cudaError_t launch_features(const float* s, int n, float* f, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  APG_LAUNCH((n + 127) / 128, 128, 0, st, apg_features_kernel)(s, n, f);
  return cudaGetLastError();
}
cudaError_t launch_features_adj(const float* s, const float* gf, int n, float* gs, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  APG_LAUNCH((n + 127) / 128, 128, 0, st, apg_features_adj_kernel)(s, gf, n, gs);
  return cudaGetLastError();
}

}  // namespace apg
