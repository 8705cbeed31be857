// Test-only software model of tcgen05.mma.kind::tf32 (one CTA, M = 128): DECODES the shared-memory and instruction
// descriptors the kernels build (csrc/tc_layout.cuh: kmajor_desc / mnmajor_desc / idesc_tf32) and gathers the
// operands from a byte image of shared memory by the canonical unswizzled forms of cute/atom/mma_traits_sm100.hpp
//   Major-K  : ((8,n),2):((1,SBO),LBO)         [16-byte units]  element (i, k): (i%8)*16 + (i/8)*SBO + (k/4)*LBO + (k%4)*4
//   Major-MN : ((1,n),(8,k)):((X,SBO),(1,LBO))                  element (i, k): (i/4)*SBO + (k%8)*16 + (k/8)*LBO + (i%4)*4
// Operands are truncated to TF32 (top 19 bits) like the tensor core input path; accumulation in double, rounded to
// float per instruction.  Shared by hostcheck_tc.cpp (issuer loops replayed by hand) and tc_sim.h (the kernels
// themselves running on the CPU).
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

namespace emu {

struct Desc { uint32_t start, lbo, sbo; bool ok, sw128; };
static inline Desc decode(uint64_t d) {
  Desc r;
  r.start = (uint32_t)(d & 0x3fff) << 4; r.lbo = (uint32_t)((d >> 16) & 0x3fff) << 4; r.sbo = (uint32_t)((d >> 32) & 0x3fff) << 4;
  r.sw128 = (d >> 61) == 2;
  r.ok = ((d >> 46) & 3) == 1 && ((d >> 61) == 0 || r.sw128);
  return r;
}
static inline float tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }
static inline float smem_f(const unsigned char* sm, size_t sm_size, uint32_t addr) {
  float v;
  if ((size_t)addr + 4 > sm_size) return NAN;
  memcpy(&v, sm + addr, 4);
  return v;
}
// K-major, 128B swizzle (measured with tools/micro/tcgen05_probe.cu on B200: profiles/r2_tcgen05_probe.jsonl):
// linear address = start + (i/8)*SBO + (i%8)*128 + (k/4)*16 + (k%4)*4, then address bits [4,7) ^= bits [7,10)
static inline float elem(const unsigned char* sm, size_t sm_size, const Desc& d, bool mn_major, int i, int k) {
  if (d.sw128) {
    uint32_t a = d.start + (uint32_t)((i / 8) * d.sbo + (i % 8) * 128 + (k / 4) * 16 + (k % 4) * 4);
    a ^= ((a >> 7) & 7u) << 4;
    return smem_f(sm, sm_size, a);
  }
  const uint32_t off = mn_major ? (uint32_t)((i / 4) * d.sbo + (k % 8) * 16 + (k / 8) * d.lbo + (i % 4) * 4)
                                : (uint32_t)((i % 8) * 16 + (i / 8) * d.sbo + (k / 4) * d.lbo + (k % 4) * 4);
  return smem_f(sm, sm_size, d.start + off);
}
// tmem: [128][512] floats.  D[lane][d_col + n] (+)= sum_k A[lane][k] B[n][k], K = 8.  A from TMEM columns
// (a_col >= 0) or from a K-major descriptor.  false: the model rejects the instruction.
static inline bool mma(float* tmem, const unsigned char* sm, size_t sm_size, int d_col, int a_col, uint64_t a_desc,
                       uint64_t b_desc, uint32_t idesc, bool acc) {
  const int M = (int)((idesc >> 24) & 31) * 16, N = (int)((idesc >> 17) & 63) * 8;
  const bool bmn = (idesc >> 16) & 1, amn = (idesc >> 15) & 1;
  if (((idesc >> 4) & 3) != 1 || ((idesc >> 7) & 7) != 2 || ((idesc >> 10) & 7) != 2 || amn || M != 128 || N % 16 ||
      N < 16 || N > 256)
    return false;
#ifndef APG_EMU_ALLOW_MN_MAJOR
  if (bmn) return false;      // measured on B200: kind::tf32 with an MN-major operand yields zeros (tcgen05_probe.cu)
#endif
  const Desc bd0 = decode(b_desc);
  if (bd0.sw128 && (bd0.start & 1023u) >= 128u) return false;   // panel base must be 1024-byte aligned (+ 32 B k-steps)
  if (d_col < 0 || d_col + N > 512 || (a_col >= 0 && a_col + 8 > 512)) return false;
  const Desc bd = decode(b_desc), ad = decode(a_desc);
  if (!bd.ok || (a_col < 0 && !ad.ok)) return false;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = acc ? (double)tmem[m * 512 + d_col + n] : 0.0;
      for (int k = 0; k < 8; ++k) {
        const float a = a_col >= 0 ? tmem[m * 512 + a_col + k] : elem(sm, sm_size, ad, false, m, k);
        s += (double)tf32(a) * (double)tf32(elem(sm, sm_size, bd, bmn, n, k));
      }
      tmem[m * 512 + d_col + n] = (float)s;
    }
  return true;
}

}  // namespace emu
