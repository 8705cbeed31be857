// Per-drone math of the learnt fixed-wing dynamics:
//   next = simulate_fixed_wing(state, action, dt) + linear_state_2(relu(linear_state_1([state, action])))
// where EVERY physical constant is a parameter used live by the simulator: the full 3x3 inertia matrix and the 37
// config scalars.  Forward and hand-written adjoint w.r.t. state, action and all 1914 parameters.
// `__host__ __device__`, templated on the scalar type (kernels: csrc/learnt_kernels.cu; CPU check:
// tests/hostcheck/hostcheck_learnt.cpp).
//
// Reference: neural_control/dynamics/fixed_wing_dynamics.py:270-326 (LearntFixedWingDynamics) over :98-267
// (simulate_fixed_wing).  Kept quirks: the three moments are scaled by the chord; gravity enters through
// `torch.tensor(g_m)` (:197), a detached copy, so `g` receives no gradient and `mass` only the one through 1/mass;
// `I` is a free 3x3 matrix (its inverse and products are general, not the symmetric two-parameter form).
// Flat parameter vector (named_parameters() order): I [3][3] | cfg.<key>, keys SORTED (ParameterDict) |
//   linear_state_1.weight [64][16] | .bias (64) | linear_state_2.weight [12][64] | .bias (12)
#pragma once
#include "apg_math.cuh"

namespace apg {

struct LearntWingLayout {
  static constexpr int XD = 16, HD = 64, SD = 12, AD = 4;
  // sorted config keys (ASCII order: upper case first)
  enum Key { K_CD0 = 0, K_CD_ALPHA, K_CD_DE, K_CD_Q, K_CL0, K_CL_ALPHA, K_CL_DE, K_CL_Q, K_CY0, K_CY_BETA, K_CY_DA,
             K_CY_DR, K_CY_P, K_CY_R, K_CLL0, K_CLL_BETA, K_CLL_DA, K_CLL_DR, K_CLL_P, K_CLL_R, K_CM0, K_CM_ALPHA,
             K_CM_DE, K_CM_Q, K_CN0, K_CN_BETA, K_CN_DA, K_CN_DR, K_CN_P, K_CN_R, K_S, K_B, K_C, K_EPS, K_G, K_MASS,
             K_RHO, NKEYS };
  static constexpr int O_I = 0, O_CFG = 9, NPH = O_CFG + NKEYS;            // 46 physical parameters
  static constexpr int O_W1 = NPH, O_B1 = O_W1 + HD * XD, O_W2 = O_B1 + HD, O_B2 = O_W2 + SD * HD,
                       NP = O_B2 + SD;                                       // 1914
};

template <typename T>
struct LearntWing {
  using Y = LearntWingLayout;

  // simulator step with the constants taken from P; when ADJ also the reverse sweep: gs (12), ga (4) and the
  // cotangents dph (46) of the physical parameters (I row-major, then the sorted config keys)
  template <bool ADJ>
  APG_HD static void sim(const T* P, const T* s, const T* a, T dt, T* o, const T* g, T* gs, T* ga, T* dph) {
    const T* C = P + Y::O_CFG;
    const T PI = T(3.14159265358979323846);
    const T BOUND = T(10.0 / 180.0 * 3.14159265358979323846);
    const T u = s[3], v = s[4], w = s[5], phi = s[6], th = s[7], psi = s[8], p = s[9], q = s[10], r = s[11];
    const T mass = C[Y::K_MASS], cch = C[Y::K_C], bsp = C[Y::K_B];
    const T gm = C[Y::K_G] * mass;                            // detached in the reference: no cotangent below
    const T thr = a[0] * T(7);
    const T de = PI * (a[1] * T(40) - T(20)) / T(180);
    const T da = PI * (a[2] * T(5) - T(2.5)) / T(180);
    const T dr = PI * (a[3] * T(40) - T(20)) / T(180);
    const T V2 = u * u + v * v + w * w;
    const T V = sqrt_(V2);
    const T ra = w / u, rb = v / V;
    const T al0 = atan_(ra), be0 = atan_(rb);
    const T alpha = al0 < -BOUND ? -BOUND : (al0 > BOUND ? BOUND : al0);
    const T beta = be0 < -BOUND ? -BOUND : (be0 > BOUND ? BOUND : be0);
    const T c2v = cch / (T(2) * V), b2v = bsp / (T(2) * V);
    const T CL = C[Y::K_CL0] + C[Y::K_CL_ALPHA] * alpha + C[Y::K_CL_Q] * c2v * q + C[Y::K_CL_DE] * de;
    const T CD = C[Y::K_CD0] + C[Y::K_CD_ALPHA] * alpha + C[Y::K_CD_Q] * c2v * q + C[Y::K_CD_DE] * de;
    const T Cm = C[Y::K_CM0] + C[Y::K_CM_ALPHA] * alpha + C[Y::K_CM_Q] * c2v * q + C[Y::K_CM_DE] * de;
    const T CY = C[Y::K_CY0] + C[Y::K_CY_BETA] * beta + C[Y::K_CY_P] * b2v * p + C[Y::K_CY_R] * b2v * r
               + C[Y::K_CY_DA] * da + C[Y::K_CY_DR] * dr;
    const T Cl = C[Y::K_CLL0] + C[Y::K_CLL_BETA] * beta + C[Y::K_CLL_P] * b2v * p + C[Y::K_CLL_R] * b2v * r
               + C[Y::K_CLL_DA] * da + C[Y::K_CLL_DR] * dr;
    const T Cn = C[Y::K_CN0] + C[Y::K_CN_BETA] * beta + C[Y::K_CN_P] * b2v * p + C[Y::K_CN_R] * b2v * r
               + C[Y::K_CN_DA] * da + C[Y::K_CN_DR] * dr;
    const T rho = C[Y::K_RHO], Sw = C[Y::K_S];
    const T hrs = T(0.5) * rho * Sw;
    const T qS = hrs * V2;
    const T L = qS * CL, D = qS * CD, Yf = qS * CY;
    const T qSc = qS * cch;                                   // all three moments use the chord (reference quirk)
    const T lm = qSc * Cl, mm = qSc * Cm, nm = qSc * Cn;
    T sa, ca, sb, cb, sph, cph, sth, cth, sps, cps, se, ce;
    sincos_(alpha, &sa, &ca); sincos_(beta, &sb, &cb);
    sincos_(phi, &sph, &cph); sincos_(th, &sth, &cth); sincos_(psi, &sps, &cps);
    sincos_(C[Y::K_EPS], &se, &ce);
    const T fx = -ca * cb * D - ca * sb * Yf + sa * L - sth * gm + thr * ce;
    const T fy = -sb * D + cb * Yf + sph * cth * gm;
    const T fz = -sa * cb * D - sa * sb * Yf - ca * L + cph * cth * gm + thr * se;
    const T m01 = -cph * sps + sph * sth * cps, m02 = sph * sps + cph * sth * cps;
    const T m11 = cph * cps + sph * sth * sps,  m12 = -sph * cps + cph * sth * sps;
    const T xd = cth * cps * u + m01 * v + m02 * w;
    const T yd = cth * sps * u + m11 * v + m12 * w;
    const T zd = -sth * u + sph * cth * v + cph * cth * w;
    const T im = T(1) / mass;
    const T ud = im * fx - (q * w - r * v);
    const T vd = im * fy - (r * u - p * w);
    const T wd = im * fz - (p * v - q * u);
    const T tth = sth / cth;
    const T phid = p + sph * tth * q + cph * tth * r;
    const T thd = cph * q - sph * r;
    const T ict = T(1) / cth;
    const T psid = (sph * q + cph * r) * ict;
    // omega_dot = I^-1 (M - omega x (I omega)) with the general 3x3 matrix I (row-major)
    const T* I = P + Y::O_I;
    const T om[3] = {p, q, r};
    T Iw[3];
    for (int i = 0; i < 3; ++i) Iw[i] = I[3 * i] * p + I[3 * i + 1] * q + I[3 * i + 2] * r;
    const T rv[3] = {lm - (q * Iw[2] - r * Iw[1]), mm - (r * Iw[0] - p * Iw[2]), nm - (p * Iw[1] - q * Iw[0])};
    T inv[9];
    {
      const T c00 = I[4] * I[8] - I[5] * I[7], c01 = I[5] * I[6] - I[3] * I[8], c02 = I[3] * I[7] - I[4] * I[6];
      const T idet = T(1) / (I[0] * c00 + I[1] * c01 + I[2] * c02);
      inv[0] = c00 * idet; inv[1] = (I[2] * I[7] - I[1] * I[8]) * idet; inv[2] = (I[1] * I[5] - I[2] * I[4]) * idet;
      inv[3] = c01 * idet; inv[4] = (I[0] * I[8] - I[2] * I[6]) * idet; inv[5] = (I[2] * I[3] - I[0] * I[5]) * idet;
      inv[6] = c02 * idet; inv[7] = (I[1] * I[6] - I[0] * I[7]) * idet; inv[8] = (I[0] * I[4] - I[1] * I[3]) * idet;
    }
    T od[3];
    for (int i = 0; i < 3; ++i) od[i] = inv[3 * i] * rv[0] + inv[3 * i + 1] * rv[1] + inv[3 * i + 2] * rv[2];
    if (!ADJ) {
      o[0] = s[0] + dt * xd; o[1] = s[1] + dt * yd; o[2] = s[2] + dt * zd;
      o[3] = u + dt * ud; o[4] = v + dt * vd; o[5] = w + dt * wd;
      o[6] = phi + dt * phid; o[7] = th + dt * thd; o[8] = psi + dt * psid;
      o[9] = p + dt * od[0]; o[10] = q + dt * od[1]; o[11] = r + dt * od[2];
      return;
    }
    // ------------------------------------------------------------------ reverse sweep
    for (int i = 0; i < Y::NPH; ++i) dph[i] = T(0);
    T* dI = dph + Y::O_I;
    T* dC = dph + Y::O_CFG;
    const T Gx = dt * g[0], Gy = dt * g[1], Gz = dt * g[2], Gu = dt * g[3], Gv = dt * g[4], Gw = dt * g[5];
    const T Gphi = dt * g[6], Gth = dt * g[7], Gpsi = dt * g[8];
    const T God[3] = {dt * g[9], dt * g[10], dt * g[11]};
    T bu = 0, bv = 0, bw = 0, bphi = 0, bth = 0, bpsi = 0;
    T bom[3] = {0, 0, 0};
    // od = inv rv:  b_rv = inv^T God;  d inv -> dI = - b_rv od^T
    T brv[3];
    for (int i = 0; i < 3; ++i) brv[i] = inv[i] * God[0] + inv[3 + i] * God[1] + inv[6 + i] * God[2];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) dI[3 * i + j] -= brv[i] * od[j];
    const T blm = brv[0], bmm = brv[1], bnm = brv[2];
    // rv = M - om x Iw:  b_c = -b_rv;  c = a x b: b_a = b x b_c, b_b = b_c x a
    const T bc[3] = {-brv[0], -brv[1], -brv[2]};
    bom[0] += Iw[1] * bc[2] - Iw[2] * bc[1];
    bom[1] += Iw[2] * bc[0] - Iw[0] * bc[2];
    bom[2] += Iw[0] * bc[1] - Iw[1] * bc[0];
    const T bIw[3] = {bc[1] * om[2] - bc[2] * om[1], bc[2] * om[0] - bc[0] * om[2], bc[0] * om[1] - bc[1] * om[0]};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) { bom[j] += I[3 * i + j] * bIw[i]; dI[3 * i + j] += bIw[i] * om[j]; }
    T bp = bom[0], bq = bom[1], br = bom[2];
    // euler kinematics
    bp += Gphi;
    bq += Gphi * sph * tth + Gth * cph + Gpsi * sph * ict;
    br += Gphi * cph * tth - Gth * sph + Gpsi * cph * ict;
    bphi += Gphi * (cph * tth * q - sph * tth * r) + Gth * (-sph * q - cph * r) + Gpsi * (cph * q - sph * r) * ict;
    bth += Gphi * (sph * q + cph * r) * ict * ict + Gpsi * (sph * q + cph * r) * sth * ict * ict;
    // body accelerations: ud = im fx - ..., im = 1 / mass
    const T bfx = Gu * im, bfy = Gv * im, bfz = Gw * im;
    dC[Y::K_MASS] += -(Gu * fx + Gv * fy + Gw * fz) * im * im;
    bq += -Gu * w + Gw * u;  br += Gu * v - Gv * u;  bp += Gv * w - Gw * v;
    bw += -Gu * q + Gv * p;  bv += Gu * r - Gw * p;  bu += -Gv * r + Gw * q;
    // position kinematics
    bu += Gx * cth * cps + Gy * cth * sps - Gz * sth;
    bv += Gx * m01 + Gy * m11 + Gz * sph * cth;
    bw += Gx * m02 + Gy * m12 + Gz * cph * cth;
    bphi += Gx * ((sph * sps + cph * sth * cps) * v + (cph * sps - sph * sth * cps) * w)
          + Gy * ((-sph * cps + cph * sth * sps) * v + (-cph * cps - sph * sth * sps) * w)
          + Gz * (cph * cth * v - sph * cth * w);
    bth += Gx * (-sth * cps * u + sph * cth * cps * v + cph * cth * cps * w)
         + Gy * (-sth * sps * u + sph * cth * sps * v + cph * cth * sps * w)
         + Gz * (-cth * u - sph * sth * v - cph * sth * w);
    bpsi += Gx * (-cth * sps * u + (-cph * cps - sph * sth * sps) * v + (sph * cps - cph * sth * sps) * w)
          + Gy * (cth * cps * u + (-cph * sps + sph * sth * cps) * v + (sph * sps + cph * sth * cps) * w);
    // forces (gm carries no cotangent to g / mass: detached in the reference)
    const T bD = -bfx * ca * cb - bfy * sb - bfz * sa * cb;
    const T bY = -bfx * ca * sb + bfy * cb - bfz * sa * sb;
    const T bL = bfx * sa - bfz * ca;
    T balpha = bfx * (sa * cb * D + sa * sb * Yf + ca * L) + bfz * (-ca * cb * D - ca * sb * Yf + sa * L);
    T bbeta = bfx * (ca * sb * D - ca * cb * Yf) + bfy * (-cb * D - sb * Yf) + bfz * (sa * sb * D - sa * cb * Yf);
    bth += -bfx * cth * gm - bfy * sph * sth * gm - bfz * cph * sth * gm;
    bphi += bfy * cph * cth * gm - bfz * sph * cth * gm;
    const T bthr = bfx * ce + bfz * se;
    dC[Y::K_EPS] += thr * (-bfx * se + bfz * ce);
    // L = qS CL, ...; moments = qS c C
    const T bqS = bL * CL + bD * CD + bY * CY + cch * (blm * Cl + bmm * Cm + bnm * Cn);
    const T bCL = bL * qS, bCD = bD * qS, bCY = bY * qS, bCl = blm * qSc, bCm = bmm * qSc, bCn = bnm * qSc;
    dC[Y::K_CL0] += bCL; dC[Y::K_CL_ALPHA] += bCL * alpha; dC[Y::K_CL_Q] += bCL * c2v * q; dC[Y::K_CL_DE] += bCL * de;
    dC[Y::K_CD0] += bCD; dC[Y::K_CD_ALPHA] += bCD * alpha; dC[Y::K_CD_Q] += bCD * c2v * q; dC[Y::K_CD_DE] += bCD * de;
    dC[Y::K_CM0] += bCm; dC[Y::K_CM_ALPHA] += bCm * alpha; dC[Y::K_CM_Q] += bCm * c2v * q; dC[Y::K_CM_DE] += bCm * de;
    dC[Y::K_CY0] += bCY; dC[Y::K_CY_BETA] += bCY * beta; dC[Y::K_CY_P] += bCY * b2v * p; dC[Y::K_CY_R] += bCY * b2v * r;
    dC[Y::K_CY_DA] += bCY * da; dC[Y::K_CY_DR] += bCY * dr;
    dC[Y::K_CLL0] += bCl; dC[Y::K_CLL_BETA] += bCl * beta; dC[Y::K_CLL_P] += bCl * b2v * p;
    dC[Y::K_CLL_R] += bCl * b2v * r; dC[Y::K_CLL_DA] += bCl * da; dC[Y::K_CLL_DR] += bCl * dr;
    dC[Y::K_CN0] += bCn; dC[Y::K_CN_BETA] += bCn * beta; dC[Y::K_CN_P] += bCn * b2v * p;
    dC[Y::K_CN_R] += bCn * b2v * r; dC[Y::K_CN_DA] += bCn * da; dC[Y::K_CN_DR] += bCn * dr;
    dC[Y::K_RHO] += bqS * T(0.5) * Sw * V2;
    dC[Y::K_S] += bqS * T(0.5) * rho * V2;
    balpha += bCL * C[Y::K_CL_ALPHA] + bCD * C[Y::K_CD_ALPHA] + bCm * C[Y::K_CM_ALPHA];
    bbeta += bCY * C[Y::K_CY_BETA] + bCl * C[Y::K_CLL_BETA] + bCn * C[Y::K_CN_BETA];
    const T lonq = bCL * C[Y::K_CL_Q] + bCD * C[Y::K_CD_Q] + bCm * C[Y::K_CM_Q];
    const T latp = bCY * C[Y::K_CY_P] + bCl * C[Y::K_CLL_P] + bCn * C[Y::K_CN_P];
    const T latr = bCY * C[Y::K_CY_R] + bCl * C[Y::K_CLL_R] + bCn * C[Y::K_CN_R];
    bq += lonq * c2v;
    bp += latp * b2v;
    br += latr * b2v;
    const T bc2v = lonq * q, bb2v = latp * p + latr * r;
    dC[Y::K_C] += bc2v / (T(2) * V) + qS * (blm * Cl + bmm * Cm + bnm * Cn);
    dC[Y::K_B] += bb2v / (T(2) * V);
    const T bde = bCL * C[Y::K_CL_DE] + bCD * C[Y::K_CD_DE] + bCm * C[Y::K_CM_DE];
    const T bda = bCY * C[Y::K_CY_DA] + bCl * C[Y::K_CLL_DA] + bCn * C[Y::K_CN_DA];
    const T bdr = bCY * C[Y::K_CY_DR] + bCl * C[Y::K_CLL_DR] + bCn * C[Y::K_CN_DR];
    T bV = -(bc2v * c2v + bb2v * b2v) / V;
    T bV2 = bqS * hrs;
    const T bra = (al0 >= -BOUND && al0 <= BOUND) ? balpha / (T(1) + ra * ra) : T(0);
    const T brb = (be0 >= -BOUND && be0 <= BOUND) ? bbeta / (T(1) + rb * rb) : T(0);
    bw += bra / u;  bu += -bra * ra / u;
    bv += brb / V;  bV += -brb * rb / V;
    bV2 += bV * T(0.5) / V;
    bu += bV2 * T(2) * u; bv += bV2 * T(2) * v; bw += bV2 * T(2) * w;
    gs[0] = g[0]; gs[1] = g[1]; gs[2] = g[2];
    gs[3] = g[3] + bu; gs[4] = g[4] + bv; gs[5] = g[5] + bw;
    gs[6] = g[6] + bphi; gs[7] = g[7] + bth; gs[8] = g[8] + bpsi;
    gs[9] = g[9] + bp; gs[10] = g[10] + bq; gs[11] = g[11] + br;
    ga[0] = T(7) * bthr;
    ga[1] = bde * PI * T(40) / T(180);
    ga[2] = bda * PI * T(5) / T(180);
    ga[3] = bdr * PI * T(40) / T(180);
  }

  APG_HD static T hidden(const T* P, const T* s, const T* a, int j) {
    T v = P[Y::O_B1 + j];
    const T* w = P + Y::O_W1 + j * Y::XD;
    for (int k = 0; k < 12; ++k) v += w[k] * s[k];
    for (int k = 0; k < 4; ++k) v += w[12 + k] * a[k];
    return v > T(0) ? v : T(0);
  }
  // out = sim(s, a) + W2 relu(W1 [s, a] + b1) + b2;  h: HD values with stride hs
  APG_HD static void forward(const T* P, const T* s, const T* a, T dt, T* out, T* h, int hs) {
    sim<false>(P, s, a, dt, out, nullptr, nullptr, nullptr, nullptr);
    for (int j = 0; j < Y::HD; ++j) h[j * hs] = hidden(P, s, a, j);
    for (int i = 0; i < Y::SD; ++i) {
      T v = P[Y::O_B2 + i];
      const T* w = P + Y::O_W2 + i * Y::HD;
      for (int j = 0; j < Y::HD; ++j) v += w[j] * h[j * hs];
      out[i] += v;
    }
  }
  // adjoint for one drone: gs (12), ga (4), dh (HD, stride hs: cotangent of the hidden pre-activations),
  // dph (46: cotangents of I and of the config scalars);  dW2 = sum g (x) h, db2 = sum g, dW1 = sum dh (x) [s, a]
  APG_HD static void adjoint(const T* P, const T* s, const T* a, const T* h, int hs, T dt, const T* g, T* gs, T* ga,
                             T* dh, T* dph) {
    sim<true>(P, s, a, dt, nullptr, g, gs, ga, dph);
    for (int j = 0; j < Y::HD; ++j) {
      T v = 0;
      if (h[j * hs] > T(0))
        for (int i = 0; i < Y::SD; ++i) v += P[Y::O_W2 + i * Y::HD + j] * g[i];
      dh[j * hs] = v;
    }
    for (int k = 0; k < Y::XD; ++k) {
      T v = 0;
      for (int j = 0; j < Y::HD; ++j) v += P[Y::O_W1 + j * Y::XD + k] * dh[j * hs];
      if (k < 12) gs[k] += v; else ga[k - 12] += v;
    }
  }

  // ---- the interface the kernels use (same as LearntQuad; the simulator constants are all parameters: pc unused)
  static constexpr int NPH = Y::NPH;
  APG_HD static void fwd(const T* P, const float* /*pc*/, const T* s, const T* a, T dt, T* out, T* x, T* h, int hs) {
    forward(P, s, a, dt, out, h, hs);
    for (int k = 0; k < 12; ++k) x[k] = s[k];
    for (int k = 0; k < 4; ++k) x[12 + k] = a[k];
  }
  APG_HD static void adj(const T* P, const float* /*pc*/, const T* s, const T* a, const T* /*x*/, const T* h, int hs,
                         T dt, const T* g, T* gs, T* ga, T* dh, T* dph) {
    adjoint(P, s, a, h, hs, dt, g, gs, ga, dh, dph);
  }
};

}  // namespace apg
