// Gradient exchange over NVLink peer memory (the path's one collective is a sum over the ranks of the
// flat weight gradient, 131 KB for the quadrotor MLP): index arithmetic shared by the kernels (p2p_kernels.cu) and the
// CPU check (tests/hostcheck/hostcheck_p2p.cpp).
//
// Every rank owns, in symmetric (peer-mapped) memory, two receive SETS used by alternate steps; a set holds one slot
// of n floats per source rank plus one flag word per source rank:
//     set s:  slots [world][n]   slot q = the reduced gradient of rank q for this step
//             flags [world]      flag q = number of the last step whose slot q is complete
// Step e (1, 2, ...) uses set e & 1.  Kernel 1 (the gradient reduction of the adjoint pass) writes its result into
// slot `rank` of EVERY rank's set and then raises flag `rank` there; kernel 2 waits for all `world` flags of its own
// set to reach e and sums the slots in rank order - the same values in the same order on every rank, so the
// all-reduced gradient is bitwise identical across ranks and independent of arrival order.  A rank can be at most
// one step ahead of a peer (its kernel 2 of step e+1 needs the peer's kernel 1 of step e+1, which runs after the
// peer's kernel 2 of step e), hence two sets are enough.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include "apg_math.cuh"

namespace apg {

struct GradCommLayout {
  int world, n;
  APG_HD size_t set_floats() const { return (size_t)world * n + (size_t)world; }      // slots + flags (as 4-byte words)
  APG_HD size_t total_floats() const { return 2 * set_floats(); }
  APG_HD size_t slot_off(int set, int src) const { return (size_t)set * set_floats() + (size_t)src * n; }
  APG_HD size_t flag_off(int set, int src) const { return (size_t)set * set_floats() + (size_t)world * n + src; }
};

// partial column that holds torch entry p (see apg_reduce_kernel: the fc1 block of the hutter conv nets is kept
// position-major by the kernels)
APG_HD int p2p_partial_column(int p, int pm_off, int pm_k1, int pm_npos) {
  if (pm_npos > 0 && p >= pm_off && p < pm_off + 64 * pm_k1) {
    const int j = (p - pm_off) / pm_k1, k = (p - pm_off) - j * pm_k1;
    if (k >= 64) {
      const int c = (k - 64) / pm_npos, tt = (k - 64) - c * pm_npos;
      return pm_off + j * pm_k1 + 64 + tt * 20 + c;
    }
  }
  return p;
}

// fixed-order sum over the CTA partials, the same association as apg_reduce_kernel
APG_HD float p2p_reduce_entry(const float* partials, int ncta, int n, int q, float scale) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int c = 0;
  for (; c + 3 < ncta; c += 4) {
    s0 += partials[(size_t)(c + 0) * n + q];
    s1 += partials[(size_t)(c + 1) * n + q];
    s2 += partials[(size_t)(c + 2) * n + q];
    s3 += partials[(size_t)(c + 3) * n + q];
  }
  for (; c < ncta; ++c) s0 += partials[(size_t)c * n + q];
  return scale * ((s0 + s1) + (s2 + s3));
}

// the CTAs [c0, c1) of one slice: four interleaved running sums, then (s0+s1)+(s2+s3)
APG_HD float p2p_reduce_slice(const float* partials, int n, int q, int c0, int c1) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int c = c0;
  for (; c + 3 < c1; c += 4) {
    s0 += partials[(size_t)(c + 0) * n + q];
    s1 += partials[(size_t)(c + 1) * n + q];
    s2 += partials[(size_t)(c + 2) * n + q];
    s3 += partials[(size_t)(c + 3) * n + q];
  }
  for (; c < c1; ++c) s0 += partials[(size_t)c * n + q];
  return (s0 + s1) + (s2 + s3);
}
// value of slot entry q as apg_reduce_scatter_p2p_kernel computes it (four slices of CTAs)
APG_HD float p2p_reduce_entry4(const float* partials, int ncta, int n, int q, float scale) {
  const int per = (ncta + 3) / 4;
  float s[4];
  for (int k = 0; k < 4; ++k) {
    const int c0 = k * per < ncta ? k * per : ncta, c1 = (k + 1) * per < ncta ? (k + 1) * per : ncta;
    s[k] = p2p_reduce_slice(partials, n, q, c0, c1);
  }
  return scale * ((s[0] + s[1]) + (s[2] + s[3]));
}

// SGD with momentum as torch.optim.SGD applies it (train_base.py:139-143): buf = momentum * buf + g; p -= lr * buf
APG_HD void p2p_sgd_entry(float g, float lr, float momentum, float* buf, float* param) {
  const float b = momentum * (*buf) + g;
  *buf = b;
  *param = *param - lr * b;
}

}  // namespace apg
