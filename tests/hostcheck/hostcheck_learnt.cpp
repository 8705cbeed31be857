// Test-only harness: the per-drone learnt-dynamics math (csrc/learnt_math.cuh) on the CPU, with the parameter
// gradient assembled from the per-drone factors exactly like the adjoint kernel does.  Not part of the product.
#include <vector>
#include "learnt_math.cuh"
#include "learnt_wing_math.cuh"

using namespace apg;
using Y = LearntLayout;

template <typename T>
static void run_fwd(const T* P, const float* pc, const T* s, const T* a, T dt, int n, T* out) {
  std::vector<T> h(Y::HD);
  for (int d = 0; d < n; ++d) {
    T at[4];
    LearntQuad<T>::forward(P, pc, s + d * 12, a + d * 4, dt, out + d * 12, at, h.data(), 1);
  }
}

template <typename T>
static void run_adj(const T* P, const float* pc, const T* s, const T* a, T dt, int n, const T* g, T* gs, T* ga,
                    T* gP) {
  std::vector<T> h(Y::HD), dh(Y::HD);
  for (int i = 0; i < Y::NP; ++i) gP[i] = 0;
  for (int d = 0; d < n; ++d) {
    T at[4], out[12], gat[4], dk[3], dj[3];
    const T* sd = s + d * 12;
    const T* ad = a + d * 4;
    const T* gd = g + d * 12;
    LearntQuad<T>::forward(P, pc, sd, ad, dt, out, at, h.data(), 1);
    LearntQuad<T>::adjoint(P, pc, sd, ad, at, h.data(), 1, dt, gd, gs + d * 12, ga + d * 4, dh.data(), gat, dk, dj);
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) gP[Y::O_LAT + r * 4 + c] += gat[r] * ad[c];
    for (int i = 0; i < 3; ++i) { gP[Y::O_K + i] += dk[i]; gP[Y::O_J + i] += dj[i]; }
    for (int j = 0; j < Y::HD; ++j) {
      for (int k = 0; k < 12; ++k) gP[Y::O_W1 + j * Y::XD + k] += dh[j] * sd[k];
      for (int k = 0; k < 4; ++k) gP[Y::O_W1 + j * Y::XD + 12 + k] += dh[j] * at[k];
      gP[Y::O_B1 + j] += dh[j];
    }
    for (int i = 0; i < 12; ++i) {
      for (int j = 0; j < Y::HD; ++j) gP[Y::O_W2 + i * Y::HD + j] += gd[i] * h[j];
      gP[Y::O_B2 + i] += gd[i];
    }
  }
}

extern "C" int hc_learnt_num_params() { return Y::NP; }
extern "C" void hc_learnt_fwd_f64(const double* P, const float* pc, const double* s, const double* a, double dt, int n,
                                  double* out) { run_fwd<double>(P, pc, s, a, dt, n, out); }
extern "C" void hc_learnt_fwd_f32(const float* P, const float* pc, const float* s, const float* a, float dt, int n,
                                  float* out) { run_fwd<float>(P, pc, s, a, dt, n, out); }
extern "C" void hc_learnt_adj_f64(const double* P, const float* pc, const double* s, const double* a, double dt, int n,
                                  const double* g, double* gs, double* ga, double* gP) {
  run_adj<double>(P, pc, s, a, dt, n, g, gs, ga, gP);
}
extern "C" void hc_learnt_adj_f32(const float* P, const float* pc, const float* s, const float* a, float dt, int n,
                                  const float* g, float* gs, float* ga, float* gP) {
  run_adj<float>(P, pc, s, a, dt, n, g, gs, ga, gP);
}

// The adjoint KERNEL's data flow on the host (same for both models): tiles of LT drones, per-drone factors written to
// rows of LP floats, every parameter-gradient entry a dot product of two rows over the tile (LearntRows::entry_rows),
// summed over the tiles.
template <class M>
static void adj_tiled(const float* P, const float* pc, const float* s, const float* a, float dt, int n, const float* g,
                      float* gs, float* ga, float* gP, int LT) {
  using RW = LearntRows<M::NPH>;
  const int LP = LT + 1;
  std::vector<float> F((size_t)RW::R_TOTAL * LP);
  for (int i = 0; i < RW::NP; ++i) gP[i] = 0;
  for (int base = 0; base < n; base += LT) {
    const int valid = n - base < LT ? n - base : LT;
    for (int t = 0; t < valid; ++t) {
      const int d = base + t;
      float x[16], o[12], dph[M::NPH];
      M::fwd(P, pc, s + d * 12, a + d * 4, dt, o, x, F.data() + RW::R_H * LP + t, LP);
      M::adj(P, pc, s + d * 12, a + d * 4, x, F.data() + RW::R_H * LP + t, LP, dt, g + d * 12, gs + d * 12, ga + d * 4,
             F.data() + RW::R_DH * LP + t, dph);
      for (int j = 0; j < 12; ++j) F[(RW::R_G + j) * LP + t] = g[d * 12 + j];
      for (int j = 0; j < 16; ++j) F[(RW::R_X + j) * LP + t] = x[j];
      for (int j = 0; j < M::NPH; ++j) F[(RW::R_DP + j) * LP + t] = dph[j];
      F[RW::R_ONE * LP + t] = 1.f;
    }
    for (int e = 0; e < RW::NP; ++e) {
      int ra, rb;
      RW::entry_rows(e, &ra, &rb);
      float v = 0.f;
      for (int k = 0; k < valid; ++k) v = fmaf(F[ra * LP + k], F[rb * LP + k], v);
      gP[e] += v;
    }
  }
}
extern "C" void hc_learnt_adj_tiled_f32(const float* P, const float* pc, const float* s, const float* a, float dt,
                                        int n, const float* g, float* gs, float* ga, float* gP, int LT) {
  adj_tiled<LearntQuad<float>>(P, pc, s, a, dt, n, g, gs, ga, gP, LT);
}

// ---------------------------------------------------------------------------------------------------------------
// fixed wing (csrc/learnt_wing_math.cuh): parameter gradient assembled from the per-drone factors
// ---------------------------------------------------------------------------------------------------------------
using YW = LearntWingLayout;

template <typename T>
static void wing_fwd(const T* P, const T* s, const T* a, T dt, int n, T* out) {
  std::vector<T> h(YW::HD);
  for (int d = 0; d < n; ++d) LearntWing<T>::forward(P, s + d * 12, a + d * 4, dt, out + d * 12, h.data(), 1);
}
template <typename T>
static void wing_adj(const T* P, const T* s, const T* a, T dt, int n, const T* g, T* gs, T* ga, T* gP) {
  std::vector<T> h(YW::HD), dh(YW::HD);
  for (int i = 0; i < YW::NP; ++i) gP[i] = 0;
  for (int d = 0; d < n; ++d) {
    T out[12], dph[YW::NPH];
    const T* sd = s + d * 12;
    const T* ad = a + d * 4;
    const T* gd = g + d * 12;
    LearntWing<T>::forward(P, sd, ad, dt, out, h.data(), 1);
    LearntWing<T>::adjoint(P, sd, ad, h.data(), 1, dt, gd, gs + d * 12, ga + d * 4, dh.data(), dph);
    for (int i = 0; i < YW::NPH; ++i) gP[i] += dph[i];
    for (int j = 0; j < YW::HD; ++j) {
      for (int k = 0; k < 12; ++k) gP[YW::O_W1 + j * YW::XD + k] += dh[j] * sd[k];
      for (int k = 0; k < 4; ++k) gP[YW::O_W1 + j * YW::XD + 12 + k] += dh[j] * ad[k];
      gP[YW::O_B1 + j] += dh[j];
    }
    for (int i = 0; i < 12; ++i) {
      for (int j = 0; j < YW::HD; ++j) gP[YW::O_W2 + i * YW::HD + j] += gd[i] * h[j];
      gP[YW::O_B2 + i] += gd[i];
    }
  }
}
extern "C" int hc_learnt_wing_num_params() { return YW::NP; }
extern "C" void hc_learnt_wing_fwd_f32(const float* P, const float* s, const float* a, float dt, int n, float* out) {
  wing_fwd<float>(P, s, a, dt, n, out);
}
extern "C" void hc_learnt_wing_fwd_f64(const double* P, const double* s, const double* a, double dt, int n, double* out) {
  wing_fwd<double>(P, s, a, dt, n, out);
}
extern "C" void hc_learnt_wing_adj_f32(const float* P, const float* s, const float* a, float dt, int n, const float* g,
                                       float* gs, float* ga, float* gP) { wing_adj<float>(P, s, a, dt, n, g, gs, ga, gP); }
extern "C" void hc_learnt_wing_adj_f64(const double* P, const double* s, const double* a, double dt, int n,
                                       const double* g, double* gs, double* ga, double* gP) {
  wing_adj<double>(P, s, a, dt, n, g, gs, ga, gP);
}

extern "C" void hc_learnt_wing_adj_tiled_f32(const float* P, const float* s, const float* a, float dt, int n,
                                             const float* g, float* gs, float* ga, float* gP, int LT) {
  adj_tiled<LearntWing<float>>(P, nullptr, s, a, dt, n, g, gs, ga, gP, LT);
}
