// The path's collective as this package's own kernels over NVLink peer memory (OPTIONAL, APG_P2P_GRAD=1 in the Python
// layer; default: NCCL all-reduce).  Protocol and memory layout: p2p_math.cuh.
//
//   apg_reduce_scatter_p2p_kernel   the gradient reduction that ends the adjoint pass (sum of the per-CTA partials in
//       fixed order) with its result stored straight into slot `rank` of every rank's receive set (coalesced 4-byte
//       stores through the peer mappings); the last CTA to finish raises this rank's flag on every peer.
//   apg_gather_sgd_p2p_kernel       waits for the `world` flags of the local set, sums the slots in rank order and, if
//       asked, applies the SGD-momentum update in the same pass (replaces the all-reduce AND the two optimizer
//       launches).  The wait is bounded (~10 s of SM clocks; ranks that iterate together are microseconds apart):
//       a lost peer poisons the gradient with NaN instead of hanging the GPU.
#include "p2p_math.cuh"
#include "kernels.h"

namespace apg {

namespace {
#ifdef APG_SIM
// CPU model (tests/hostcheck): sequentially consistent atomics stand in for the system-scope release / acquire
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { simte::st_release(p, v); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) { return simte::ld_acquire(p); }
__device__ __forceinline__ float ld_relaxed_sys(const float* p) { return *p; }
__device__ __forceinline__ unsigned ticket_add(unsigned* p) { return simte::atomic_inc(p); }
__device__ __forceinline__ void fence_system() {}
__device__ __forceinline__ long long p2p_clock() { return simte::slow_clock(); }
__device__ __forceinline__ void p2p_griddep_wait() {}
__device__ __forceinline__ void p2p_griddep_launch() {}
#else
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// the slots are written by OTHER GPUs: read them at system scope (no L1, coherent with the peers' stores once the
// acquire on the flag has been observed)
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];\n" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ticket_add(unsigned* p) { return atomicAdd(p, 1u); }
__device__ __forceinline__ void fence_system() { __threadfence_system(); }
__device__ __forceinline__ long long p2p_clock() { return clock64(); }
// programmatic dependent launch (layouts.h APG_LAUNCH_PDL): the kernel before in the stream has completed / the kernel
// after may start
__device__ __forceinline__ void p2p_griddep_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void p2p_griddep_launch() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
#endif
}  // namespace

// 32 parameters x 4 CTA slices per block of 128 threads (the association of apg_reduce4_kernel when `sliced`, of
// apg_reduce_kernel otherwise: fixed order either way -> bitwise reproducible on every rank)
__global__ void __launch_bounds__(128)
    apg_reduce_scatter_p2p_kernel(const float* __restrict__ partials, int ncta, int n, float scale, int pm_off,
                                  int pm_k1, int pm_npos, float* const* __restrict__ slots,
                                  unsigned* const* __restrict__ flags, int rank, int world, unsigned epoch,
                                  unsigned* __restrict__ ticket) {
  __shared__ float s_part[4][32];
  p2p_griddep_wait();                          // the partials of the adjoint kernel before are complete
  p2p_griddep_launch();                        // the gather kernel may start polling its flags
  const int pl = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int p = blockIdx.x * 32 + pl;
  const int per = (ncta + 3) / 4;
  const int c0 = slice * per < ncta ? slice * per : ncta, c1 = (slice + 1) * per < ncta ? (slice + 1) * per : ncta;
  s_part[slice][pl] = p < n ? p2p_reduce_slice(partials, n, p2p_partial_column(p, pm_off, pm_k1, pm_npos), c0, c1) : 0.f;
  __syncthreads();
  if (slice == 0 && p < n) {
    const float v = scale * ((s_part[0][pl] + s_part[1][pl]) + (s_part[2][pl] + s_part[3][pl]));
    for (int q = 0; q < world; ++q) slots[q][(size_t)rank * n + p] = v;
  }
  __syncthreads();                              // the CTA's peer stores happen-before thread 0's fence (cumulative)
  if (threadIdx.x == 0) {
    fence_system();
    const unsigned t = ticket_add(ticket);
    if (t == gridDim.x - 1) {                   // every CTA's stores have been fenced: signal all peers
      *ticket = 0u;                             // ready for the next launch (stream order)
      fence_system();
      for (int q = 0; q < world; ++q) st_release_sys(flags[q] + rank, epoch);
    }
  }
}

__global__ void apg_gather_sgd_p2p_kernel(const float* __restrict__ slots_local, const unsigned* __restrict__ flags_local,
                                          int world, int n, unsigned epoch, float* __restrict__ grad_out,
                                          float* __restrict__ param, float* __restrict__ momentum_buf, float lr,
                                          float momentum) {
  __shared__ int s_ok;
  // Launched with programmatic serialization and WITHOUT a griddepcontrol.wait: what it consumes arrives through the
  // flags (this rank's own contribution included), and nothing earlier in the stream still touches what it writes
  // (parameters, momentum buffer, gradient); the next kernel of the stream is launched normally and waits for it.
  if (threadIdx.x == 0) {
    int ok = 1;
    const long long t0 = p2p_clock();
    for (int q = 0; q < world && ok; ++q) {
      // flags only grow; (int)(flag - epoch) >= 0 also survives the 32-bit wrap of the step counter
      while ((int)(ld_acquire_sys(flags_local + q) - epoch) < 0) {
        if (p2p_clock() - t0 > 20000000000LL) { ok = 0; break; }
      }
    }
    s_ok = ok;
  }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  float g = 0.f;
  if (s_ok) {
    for (int q = 0; q < world; ++q) g += ld_relaxed_sys(slots_local + (size_t)q * n + p);
  } else {
    g = __int_as_float(0x7fc00000);
  }
  if (grad_out) grad_out[p] = g;
  if (param) p2p_sgd_entry(g, lr, momentum, momentum_buf + p, param + p);
}

cudaError_t launch_reduce_scatter_p2p(const float* partials, int ncta, int n, float scale, int pm_off, int pm_k1,
                                      int pm_npos, float* const* slots, unsigned* const* flags, int rank, int world,
                                      unsigned epoch, unsigned* ticket, cudaStream_t st) {
  APG_LAUNCH_PDL((n + 31) / 32, 128, 0, st, apg_reduce_scatter_p2p_kernel)(partials, ncta, n, scale, pm_off, pm_k1, pm_npos,
                                                                 slots, flags, rank, world, epoch, ticket);
  return cudaGetLastError();
}

cudaError_t launch_gather_sgd_p2p(const float* slots_local, const unsigned* flags_local, int world, int n,
                                  unsigned epoch, float* grad_out, float* param, float* momentum_buf, float lr,
                                  float momentum, cudaStream_t st) {
  APG_LAUNCH_PDL((n + 127) / 128, 128, 0, st, apg_gather_sgd_p2p_kernel)(slots_local, flags_local, world, n, epoch, grad_out, param,
                                                             momentum_buf, lr, momentum);
  return cudaGetLastError();
}


}  // namespace apg
