// Test-only harness: the gradient exchange over peer memory (csrc/p2p_math.cuh, p2p_kernels.cu) with all ranks
// emulated in one process.  `mem` is the concatenation of every rank's symmetric buffer; the per-thread code of
// the two kernels is replayed entry by entry (the memory-ordering part of the protocol needs hardware).
#include <stdint.h>
#include <string.h>
#define __host__
#define __device__
#include "p2p_math.cuh"

using namespace apg;

// kernel 1 of `rank`: partials [ncta][n] -> slot `rank` of every rank's set, then the flags
extern "C" void hc_p2p_reduce_scatter(float* mem, int world, int n, int rank, unsigned epoch, const float* partials,
                                      int ncta, float scale, int pm_off, int pm_k1, int pm_npos) {
  const GradCommLayout L{world, n};
  const int set = (int)(epoch & 1u);
  for (int p = 0; p < n; ++p) {
    const float v = p2p_reduce_entry4(partials, ncta, n, p2p_partial_column(p, pm_off, pm_k1, pm_npos), scale);
    for (int q = 0; q < world; ++q) (mem + (size_t)q * L.total_floats() + L.slot_off(set, 0))[(size_t)rank * n + p] = v;
  }
  for (int q = 0; q < world; ++q) {
    unsigned* flags = reinterpret_cast<unsigned*>(mem + (size_t)q * L.total_floats() + L.flag_off(set, 0));
    flags[rank] = epoch;
  }
}

// kernel 2 of `rank`; returns 0 if a flag of the set has not reached `epoch` (the kernel would keep waiting)
extern "C" int hc_p2p_gather_sgd(const float* mem, int world, int n, int rank, unsigned epoch, float* grad_out,
                                 float* param, float* buf, float lr, float momentum) {
  const GradCommLayout L{world, n};
  const int set = (int)(epoch & 1u);
  const float* slots = mem + (size_t)rank * L.total_floats() + L.slot_off(set, 0);
  const unsigned* flags = reinterpret_cast<const unsigned*>(mem + (size_t)rank * L.total_floats() + L.flag_off(set, 0));
  for (int q = 0; q < world; ++q)
    if ((int)(flags[q] - epoch) < 0) return 0;
  for (int p = 0; p < n; ++p) {
    float g = 0.f;
    for (int q = 0; q < world; ++q) g += slots[(size_t)q * n + p];
    if (grad_out) grad_out[p] = g;
    if (param) p2p_sgd_entry(g, lr, momentum, buf + p, param + p);
  }
  return 1;
}

extern "C" size_t hc_p2p_total_floats(int world, int n) { return GradCommLayout{world, n}.total_floats(); }
