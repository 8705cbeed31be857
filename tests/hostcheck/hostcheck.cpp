// Test-only harness: compiles the product's __host__ __device__ math header (csrc/apg_math.cuh) with g++ so
// the per-drone forward/adjoint formulas can be checked on the CPU against autograd of the oracle.
// Not part of the product; nothing in the package loads this library.
#include "apg_math.cuh"

using namespace apg;

template <typename T, template <typename> class Sys>
static void run_step(const T* s, const T* a, T dt, const float* pc, T* out, int n) {
  for (int i = 0; i < n; ++i) Sys<T>::step(s + i * Sys<T>::S, a + i * Sys<T>::A, dt, pc, out + i * Sys<T>::S);
}
template <typename T, template <typename> class Sys>
static void run_adj(const T* s, const T* a, T dt, const float* pc, const T* g, T* gs, T* ga, int n) {
  for (int i = 0; i < n; ++i)
    Sys<T>::step_adj(s + i * Sys<T>::S, a + i * Sys<T>::A, dt, pc, g + i * Sys<T>::S, gs + i * Sys<T>::S,
                     ga + i * Sys<T>::A);
}
// rollout loss + gradients w.r.t. the action sequence for one batch (states recomputed; plain reverse sweep)
template <typename T, template <typename> class Sys>
static T run_rollout(const T* cur, const T* act, const T* ref, T dt, const float* pc, int n, int h, T* gact,
                     T* states_out) {
  constexpr int S = Sys<T>::S, A = Sys<T>::A, R = Sys<T>::REFW;
  T total = 0;
  for (int i = 0; i < n; ++i) {
    T st[65][12];
    for (int j = 0; j < S; ++j) st[0][j] = cur[i * S + j];
    for (int k = 0; k < h; ++k) {
      Sys<T>::step(st[k], act + (i * h + k) * A, dt, pc, st[k + 1]);
      total += Sys<T>::loss(st[k + 1], ref + (i * h + k) * R, act + (i * h + k) * A, st[0], k, h);
      if (states_out) for (int j = 0; j < S; ++j) states_out[(i * h + k) * S + j] = st[k + 1][j];
    }
    T g[12] = {0};
    for (int k = h - 1; k >= 0; --k) {
      T ga[4] = {0}, gs[12], ga2[4];
      Sys<T>::loss_grad(st[k + 1], ref + (i * h + k) * R, act + (i * h + k) * A, st[0], k, h, g, ga);
      Sys<T>::step_adj(st[k], act + (i * h + k) * A, dt, pc, g, gs, ga2);
      for (int j = 0; j < A; ++j) gact[(i * h + k) * A + j] = ga[j] + ga2[j];
      for (int j = 0; j < S; ++j) g[j] = gs[j];
    }
  }
  return total;
}

#define EXPORT_SYS(NAME, SYS)                                                                                     \
  extern "C" void hc_step_##NAME##_f64(const double* s, const double* a, double dt, const float* pc, double* o,   \
                                       int n) { run_step<double, SYS>(s, a, dt, pc, o, n); }                      \
  extern "C" void hc_step_##NAME##_f32(const float* s, const float* a, float dt, const float* pc, float* o,       \
                                       int n) { run_step<float, SYS>(s, a, dt, pc, o, n); }                       \
  extern "C" void hc_adj_##NAME##_f64(const double* s, const double* a, double dt, const float* pc,               \
                                      const double* g, double* gs, double* ga, int n) {                           \
    run_adj<double, SYS>(s, a, dt, pc, g, gs, ga, n); }                                                           \
  extern "C" void hc_adj_##NAME##_f32(const float* s, const float* a, float dt, const float* pc, const float* g,  \
                                      float* gs, float* ga, int n) { run_adj<float, SYS>(s, a, dt, pc, g, gs, ga, n); } \
  extern "C" double hc_rollout_##NAME##_f64(const double* cur, const double* act, const double* ref, double dt,   \
                                            const float* pc, int n, int h, double* gact, double* st) {            \
    return run_rollout<double, SYS>(cur, act, ref, dt, pc, n, h, gact, st); }                                     \
  extern "C" float hc_rollout_##NAME##_f32(const float* cur, const float* act, const float* ref, float dt,        \
                                           const float* pc, int n, int h, float* gact, float* st) {               \
    return run_rollout<float, SYS>(cur, act, ref, dt, pc, n, h, gact, st); }

EXPORT_SYS(quad, Quad)
EXPORT_SYS(wing, Wing)
EXPORT_SYS(cartpole, Cartpole)

extern "C" void hc_features_f64(const double* s, double* f, int n) {
  for (int i = 0; i < n; ++i) Quad<double>::features(s + 12 * i, f + 15 * i);
}
extern "C" void hc_features_adj_f64(const double* s, const double* gf, double* gs, int n) {
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < 12; ++j) gs[12 * i + j] = 0;
    Quad<double>::features_adj(s + 12 * i, gf + 15 * i, gs + 12 * i);
  }
}
extern "C" void hc_features_f32(const float* s, float* f, int n) {
  for (int i = 0; i < n; ++i) Quad<float>::features(s + 12 * i, f + 15 * i);
}
