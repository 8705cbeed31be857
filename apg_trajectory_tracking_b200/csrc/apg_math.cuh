// Per-drone math of the rollout hot path: one dynamics step (forward + hand-written adjoint), the per-step
// tracking loss (+ gradient) and the quadrotor featurizer (+ adjoint) for the three systems of the reference.
//
// Everything here is `__host__ __device__` and templated on the scalar type so that the *same* source is
//   * inlined into the sm_100a kernels (T = float), and
//   * compiled with g++ into a test-only harness (tests/hostcheck) that checks the adjoints against autograd
//     of the CPU oracle in fp64.  The harness is test infrastructure; the product never runs this on the CPU.
//
// Reference behaviour restated here (paths relative to the reference checkout):
//   quad      neural_control/dynamics/quad_dynamics_flightmare.py:128-216, quad_dynamics_base.py:59-127
//   wing      neural_control/dynamics/fixed_wing_dynamics.py:98-267
//   cartpole  neural_control/dynamics/cartpole_dynamics.py:53-119
//   losses    neural_control/drone_loss.py:12-39, 72-82, 136-145; scripts/train_cartpole.py:103-110
//   features  neural_control/dataset.py:207-220
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define APG_HD __host__ __device__ __forceinline__
#else
#define APG_HD inline
#endif

namespace apg {

enum System { SYS_QUAD = 0, SYS_WING = 1, SYS_CARTPOLE = 2 };

// ---------------------------------------------------------------------------------------------------------
// scalar helpers (accurate libm / CUDA math versions; no fast-math intrinsics: parity is fp32-tight)
// ---------------------------------------------------------------------------------------------------------
APG_HD void sincos_(float x, float* s, float* c) {
#if defined(__CUDA_ARCH__)
  sincosf(x, s, c);
#else
  *s = sinf(x); *c = cosf(x);
#endif
}
APG_HD void sincos_(double x, double* s, double* c) { *s = sin(x); *c = cos(x); }
APG_HD float sqrt_(float x) { return sqrtf(x); }
APG_HD double sqrt_(double x) { return sqrt(x); }
APG_HD float atan_(float x) { return atanf(x); }
APG_HD double atan_(double x) { return atan(x); }
APG_HD float atan2_(float y, float x) { return atan2f(y, x); }
APG_HD double atan2_(double y, double x) { return atan2(y, x); }
APG_HD float asin_(float x) { return asinf(x); }
APG_HD double asin_(double x) { return asin(x); }
APG_HD float tanh_(float x) { return tanhf(x); }
APG_HD double tanh_(double x) { return tanh(x); }
APG_HD float exp_(float x) { return expf(x); }
APG_HD double exp_(double x) { return exp(x); }
template <typename T> APG_HD T sigmoid_(T x) { return T(1) / (T(1) + exp_(-x)); }

// ---------------------------------------------------------------------------------------------------------
// physical constants, passed to the kernels by value.  Filled on the host from the reference's config json
// (+ `modified_params` overrides); see apg_trajectory_tracking_b200/params.py.
// ---------------------------------------------------------------------------------------------------------
enum QuadC { Q_MASS = 0, Q_JX, Q_JY, Q_JZ, Q_KX, Q_KY, Q_KZ, Q_GX, Q_GY, Q_GZ, Q_TDX, Q_TDY, Q_TDZ, Q_RDX, Q_RDY,
             Q_RDZ, Q_NCONST };
enum WingC { W_MASS = 0, W_IXX, W_IYY, W_IZZ, W_IXZ, W_RHO, W_S, W_C, W_B, W_G,
             W_CL0, W_CL_ALPHA, W_CL_Q, W_CL_DE, W_CD0, W_CD_ALPHA, W_CD_Q, W_CD_DE,
             W_CY0, W_CY_BETA, W_CY_P, W_CY_R, W_CY_DA, W_CY_DR,
             W_CLL0, W_CLL_BETA, W_CLL_P, W_CLL_R, W_CLL_DA, W_CLL_DR,
             W_CM0, W_CM_ALPHA, W_CM_Q, W_CM_DE,
             W_CN0, W_CN_BETA, W_CN_P, W_CN_R, W_CN_DA, W_CN_DR, W_EPS, W_NCONST };
enum CartC { C_MASSCART = 0, C_MASSPOLE, C_LENGTH, C_MAXFORCE, C_FRICTION, C_NCONST };
constexpr int MAX_PHYS = 48;
struct PhysConsts { float v[MAX_PHYS]; };

// =========================================================================================================
// Quadrotor
// =========================================================================================================
template <typename T>
struct Quad {
  static constexpr int S = 12, A = 4, REFW = 9;
  static constexpr bool SIGMOID_ACTIONS = true;

  // next = f(s, a).  s = [pos(3), roll pitch yaw, vel(3), body rates(3)]
  APG_HD static void step(const T* s, const T* a, T dt, const float* pc, T* o) {
    const T m = pc[Q_MASS];
    T Sr, Cr, Sp, Cp, Sy, Cy;
    sincos_(s[3], &Sr, &Cr); sincos_(s[4], &Sp, &Cp); sincos_(s[5], &Sy, &Cy);
    const T thrust = a[0] * T(15) - T(7.5) + T(9.81);
    const T fom = (T(1) / m) * (m * thrust);
    const T ax = fom * (Cy * Sp * Cr + Sr * Sy) + T(pc[Q_GX]) + T(pc[Q_TDX]);
    const T ay = fom * (Cr * Sy * Sp - Cy * Sr) + T(pc[Q_GY]) + T(pc[Q_TDY]);
    const T az = fom * (Cr * Cp) + T(pc[Q_GZ]) + T(pc[Q_TDZ]);
    const T hdt2 = T(0.5) * dt * dt, hdt = T(0.5) * dt;
    o[0] = s[0] + hdt2 * ax + hdt * s[6];      // the 0.5*dt*vel term is the reference's (flightmare.py:172-174)
    o[1] = s[1] + hdt2 * ay + hdt * s[7];
    o[2] = s[2] + hdt2 * az + hdt * s[8];
    o[6] = s[6] + dt * ax; o[7] = s[7] + dt * ay; o[8] = s[8] + dt * az;
    // w' = w + dt J^-1 (J K (br - w) + c + rot_drag - c); the cross term c cancels algebraically
    const T wx = s[9], wy = s[10], wz = s[11];
    o[9]  = wx + dt * (T(pc[Q_KX]) * ((a[1] - T(0.5)) - wx) + T(pc[Q_RDX]) / T(pc[Q_JX]));
    o[10] = wy + dt * (T(pc[Q_KY]) * ((a[2] - T(0.5)) - wy) + T(pc[Q_RDY]) / T(pc[Q_JY]));
    o[11] = wz + dt * (T(pc[Q_KZ]) * ((a[3] - T(0.5)) - wz) + T(pc[Q_RDZ]) / T(pc[Q_JZ]));
    // attitude integrates the OLD body rates through the Euler-rate matrix
    o[3] = s[3] + dt * (wx - Sp * wz);
    o[4] = s[4] + dt * (Cr * wy + Cp * Sr * wz);
    o[5] = s[5] + dt * (-Sr * wy + Cp * Cr * wz);
  }

  // gs = (d next / d s)^T g ,  ga = (d next / d a)^T g
  APG_HD static void step_adj(const T* s, const T* a, T dt, const float* pc, const T* g, T* gs, T* ga) {
    T Sr, Cr, Sp, Cp, Sy, Cy;
    sincos_(s[3], &Sr, &Cr); sincos_(s[4], &Sp, &Cp); sincos_(s[5], &Sy, &Cy);
    const T thrust = a[0] * T(15) - T(7.5) + T(9.81);
    const T wy = s[10], wz = s[11];
    const T hdt2 = T(0.5) * dt * dt, hdt = T(0.5) * dt;
    gs[0] = g[0]; gs[1] = g[1]; gs[2] = g[2];
    gs[6] = g[6] + hdt * g[0]; gs[7] = g[7] + hdt * g[1]; gs[8] = g[8] + hdt * g[2];
    const T gax = hdt2 * g[0] + dt * g[6], gay = hdt2 * g[1] + dt * g[7], gaz = hdt2 * g[2] + dt * g[8];
    const T r20 = Cy * Sp * Cr + Sr * Sy, r21 = Cr * Sy * Sp - Cy * Sr, r22 = Cr * Cp;
    ga[0] = T(15) * (gax * r20 + gay * r21 + gaz * r22);
    const T hx = thrust * gax, hy = thrust * gay, hz = thrust * gaz;      // cotangent of the body z axis in world
    T groll  = hx * (-Cy * Sp * Sr + Cr * Sy) + hy * (-Sr * Sy * Sp - Cy * Cr) + hz * (-Sr * Cp);
    T gpitch = hx * (Cy * Cp * Cr) + hy * (Cr * Sy * Cp) + hz * (-Cr * Sp);
    T gyaw   = hx * (-Sy * Sp * Cr + Sr * Cy) + hy * (Cr * Cy * Sp + Sy * Sr);
    const T gr = g[3], gp = g[4], gy = g[5];
    groll  += gr + dt * (gp * (-Sr * wy + Cp * Cr * wz) + gy * (-Cr * wy - Cp * Sr * wz));
    gpitch += gp + dt * (gr * (-Cp * wz) + gp * (-Sp * Sr * wz) + gy * (-Sp * Cr * wz));
    gyaw   += gy;
    gs[3] = groll; gs[4] = gpitch; gs[5] = gyaw;
    const T kx = dt * T(pc[Q_KX]), ky = dt * T(pc[Q_KY]), kz = dt * T(pc[Q_KZ]);
    gs[9]  = g[9]  * (T(1) - kx) + dt * gr;
    gs[10] = g[10] * (T(1) - ky) + dt * (gp * Cr - gy * Sr);
    gs[11] = g[11] * (T(1) - kz) + dt * (-gr * Sp + gp * Cp * Sr + gy * Cp * Cr);
    ga[1] = g[9] * kx; ga[2] = g[10] * ky; ga[3] = g[11] * kz;
  }

  // loss contribution of step k: state after the step, reference row (REFW floats), the action of the step
  APG_HD static T loss(const T* sn, const T* ref, const T* a, const T* /*cur0*/, int /*k*/, int /*h*/) {
    T l = 0;
    for (int i = 0; i < 3; ++i) {
      const T dp = sn[i] - ref[i], dv = sn[6 + i] - ref[6 + i], da = a[1 + i] - T(0.5);
      l += T(10) * dp * dp + dv * dv + T(0.1) * sn[9 + i] * sn[9 + i] + T(0.1) * da * da;
    }
    const T d0 = a[0] - T(0.5);
    return l + T(5) * d0 * d0;
  }
  // gs += dl/dsn, ga += dl/da
  APG_HD static void loss_grad(const T* sn, const T* ref, const T* a, const T* /*cur0*/, int /*k*/, int /*h*/,
                               T* gs, T* ga) {
    for (int i = 0; i < 3; ++i) {
      gs[i] += T(20) * (sn[i] - ref[i]);
      gs[6 + i] += T(2) * (sn[6 + i] - ref[6 + i]);
      gs[9 + i] += T(0.2) * sn[9 + i];
      ga[1 + i] += T(0.2) * (a[1 + i] - T(0.5));
    }
    ga[0] += T(10) * (a[0] - T(0.5));
  }

  // ---- featurizer (dataset.py:207-220): (12) -> (15) = [vel, W00 W01 W10 W11 W20 W21, W vel, body rates]
  APG_HD static void features(const T* s, T* f) {
    T Sr, Cr, Sp, Cp, Sy, Cy;
    sincos_(s[3], &Sr, &Cr); sincos_(s[4], &Sp, &Cp); sincos_(s[5], &Sy, &Cy);
    const T w00 = Cy * Cp, w01 = Sy * Cp, w02 = -Sp;
    const T w10 = Cy * Sp * Sr - Cr * Sy, w11 = Cr * Cy + Sr * Sy * Sp, w12 = Cp * Sr;
    const T w20 = Cy * Sp * Cr + Sr * Sy, w21 = Cr * Sy * Sp - Cy * Sr, w22 = Cr * Cp;
    const T vx = s[6], vy = s[7], vz = s[8];
    f[0] = vx; f[1] = vy; f[2] = vz;
    f[3] = w00; f[4] = w01; f[5] = w10; f[6] = w11; f[7] = w20; f[8] = w21;
    f[9]  = w00 * vx + w01 * vy + w02 * vz;
    f[10] = w10 * vx + w11 * vy + w12 * vz;
    f[11] = w20 * vx + w21 * vy + w22 * vz;
    f[12] = s[9]; f[13] = s[10]; f[14] = s[11];
  }
  // gs += (d f / d s)^T gf   (position gets nothing)
  APG_HD static void features_adj(const T* s, const T* gf, T* gs) {
    T Sr, Cr, Sp, Cp, Sy, Cy;
    sincos_(s[3], &Sr, &Cr); sincos_(s[4], &Sp, &Cp); sincos_(s[5], &Sy, &Cy);
    const T w00 = Cy * Cp, w01 = Sy * Cp, w02 = -Sp;
    const T w10 = Cy * Sp * Sr - Cr * Sy, w11 = Cr * Cy + Sr * Sy * Sp, w12 = Cp * Sr;
    const T w20 = Cy * Sp * Cr + Sr * Sy, w21 = Cr * Sy * Sp - Cy * Sr, w22 = Cr * Cp;
    const T vx = s[6], vy = s[7], vz = s[8];
    gs[6] += gf[0] + gf[9] * w00 + gf[10] * w10 + gf[11] * w20;
    gs[7] += gf[1] + gf[9] * w01 + gf[10] * w11 + gf[11] * w21;
    gs[8] += gf[2] + gf[9] * w02 + gf[10] * w12 + gf[11] * w22;
    // cotangents of the matrix entries
    const T g00 = gf[9] * vx + gf[3], g01 = gf[9] * vy + gf[4], g02 = gf[9] * vz;
    const T g10 = gf[10] * vx + gf[5], g11 = gf[10] * vy + gf[6], g12 = gf[10] * vz;
    const T g20 = gf[11] * vx + gf[7], g21 = gf[11] * vy + gf[8], g22 = gf[11] * vz;
    // d/droll: row1 -> row2, row2 -> -row1, row0 -> 0
    gs[3] += g10 * w20 + g11 * w21 + g12 * w22 - (g20 * w10 + g21 * w11 + g22 * w12);
    // d/dpitch
    gs[4] += g00 * (-Cy * Sp) + g01 * (-Sy * Sp) + g02 * (-Cp)
           + g10 * (Cy * Cp * Sr) + g11 * (Sr * Sy * Cp) + g12 * (-Sp * Sr)
           + g20 * (Cy * Cp * Cr) + g21 * (Cr * Sy * Cp) + g22 * (-Cr * Sp);
    // d/dyaw
    gs[5] += g00 * (-Sy * Cp) + g01 * (Cy * Cp)
           + g10 * (-Sy * Sp * Sr - Cr * Cy) + g11 * (-Cr * Sy + Sr * Cy * Sp)
           + g20 * (-Sy * Sp * Cr + Sr * Cy) + g21 * (Cr * Cy * Sp + Sy * Sr);
    gs[9] += gf[12]; gs[10] += gf[13]; gs[11] += gf[14];
  }
};

// =========================================================================================================
// Fixed wing
// =========================================================================================================
template <typename T>
struct Wing {
  static constexpr int S = 12, A = 4, REFW = 3;
  static constexpr bool SIGMOID_ACTIONS = true;

  // Shared forward body.  When ADJ, also back-propagates the cotangent g (12) into gs (12) / ga (4).
  template <bool ADJ>
  APG_HD static void eval(const T* s, const T* a, T dt, const float* pc, T* o, const T* g, T* gs, T* ga) {
    const T PI = T(3.14159265358979323846);
    const T BOUND = T(10.0 / 180.0 * 3.14159265358979323846);
    const T u = s[3], v = s[4], w = s[5], phi = s[6], th = s[7], psi = s[8], p = s[9], q = s[10], r = s[11];
    const T mass = pc[W_MASS], cch = pc[W_C], bsp = pc[W_B];
    const T Ixx = pc[W_IXX], Iyy = pc[W_IYY], Izz = pc[W_IZZ], off = -T(pc[W_IXZ]);
    const T gm = T(pc[W_G]) * mass;
    // controls (fixed_wing_dynamics.py:41-46)
    const T thr = a[0] * T(7);
    const T de = PI * (a[1] * T(40) - T(20)) / T(180);
    const T da = PI * (a[2] * T(5) - T(2.5)) / T(180);
    const T dr = PI * (a[3] * T(40) - T(20)) / T(180);
    // air data
    const T V2 = u * u + v * v + w * w;
    const T V = sqrt_(V2);
    const T ra = w / u, rb = v / V;
    const T al0 = atan_(ra), be0 = atan_(rb);
    const T alpha = al0 < -BOUND ? -BOUND : (al0 > BOUND ? BOUND : al0);
    const T beta = be0 < -BOUND ? -BOUND : (be0 > BOUND ? BOUND : be0);
    const T c2v = cch / (T(2) * V), b2v = bsp / (T(2) * V);
    const T CL = T(pc[W_CL0]) + T(pc[W_CL_ALPHA]) * alpha + T(pc[W_CL_Q]) * c2v * q + T(pc[W_CL_DE]) * de;
    const T CD = T(pc[W_CD0]) + T(pc[W_CD_ALPHA]) * alpha + T(pc[W_CD_Q]) * c2v * q + T(pc[W_CD_DE]) * de;
    const T Cm = T(pc[W_CM0]) + T(pc[W_CM_ALPHA]) * alpha + T(pc[W_CM_Q]) * c2v * q + T(pc[W_CM_DE]) * de;
    const T CY = T(pc[W_CY0]) + T(pc[W_CY_BETA]) * beta + T(pc[W_CY_P]) * b2v * p + T(pc[W_CY_R]) * b2v * r
               + T(pc[W_CY_DA]) * da + T(pc[W_CY_DR]) * dr;
    const T Cl = T(pc[W_CLL0]) + T(pc[W_CLL_BETA]) * beta + T(pc[W_CLL_P]) * b2v * p + T(pc[W_CLL_R]) * b2v * r
               + T(pc[W_CLL_DA]) * da + T(pc[W_CLL_DR]) * dr;
    const T Cn = T(pc[W_CN0]) + T(pc[W_CN_BETA]) * beta + T(pc[W_CN_P]) * b2v * p + T(pc[W_CN_R]) * b2v * r
               + T(pc[W_CN_DA]) * da + T(pc[W_CN_DR]) * dr;
    const T hrs = T(0.5) * T(pc[W_RHO]) * T(pc[W_S]);
    const T qS = hrs * V2;
    const T L = qS * CL, D = qS * CD, Y = qS * CY;
    const T qSc = qS * cch;                                   // all three moments use the chord (reference quirk)
    const T lm = qSc * Cl, mm = qSc * Cm, nm = qSc * Cn;
    T sa, ca, sb, cb, sph, cph, sth, cth, sps, cps;
    sincos_(alpha, &sa, &ca); sincos_(beta, &sb, &cb);
    sincos_(phi, &sph, &cph); sincos_(th, &sth, &cth); sincos_(psi, &sps, &cps);
    const T eps = pc[W_EPS];
    T se, ce; sincos_(eps, &se, &ce);
    const T fx = -ca * cb * D - ca * sb * Y + sa * L - sth * gm + thr * ce;
    const T fy = -sb * D + cb * Y + sph * cth * gm;
    const T fz = -sa * cb * D - sa * sb * Y - ca * L + cph * cth * gm + thr * se;
    // rotation body -> inertial (rows used for pos_dot)
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
    const T Iwx = Ixx * p + off * r, Iwy = Iyy * q, Iwz = off * p + Izz * r;
    const T rx = lm - (q * Iwz - r * Iwy);
    const T ry = mm - (r * Iwx - p * Iwz);
    const T rz = nm - (p * Iwy - q * Iwx);
    const T idet = T(1) / (Ixx * Izz - off * off);
    const T pd = (Izz * rx - off * rz) * idet;
    const T qd = ry / Iyy;
    const T rd = (-off * rx + Ixx * rz) * idet;
    if (!ADJ) {
      o[0] = s[0] + dt * xd; o[1] = s[1] + dt * yd; o[2] = s[2] + dt * zd;
      o[3] = u + dt * ud; o[4] = v + dt * vd; o[5] = w + dt * wd;
      o[6] = phi + dt * phid; o[7] = th + dt * thd; o[8] = psi + dt * psid;
      o[9] = p + dt * pd; o[10] = q + dt * qd; o[11] = r + dt * rd;
      return;
    }
    // ------------------------------------------------------------------ reverse sweep
    const T Gx = dt * g[0], Gy = dt * g[1], Gz = dt * g[2], Gu = dt * g[3], Gv = dt * g[4], Gw = dt * g[5];
    const T Gphi = dt * g[6], Gth = dt * g[7], Gpsi = dt * g[8], Gp = dt * g[9], Gq = dt * g[10], Gr = dt * g[11];
    T bu = 0, bv = 0, bw = 0, bphi = 0, bth = 0, bpsi = 0, bp = 0, bq = 0, br = 0;
    // omega_dot
    const T brx = (Izz * Gp - off * Gr) * idet;
    const T bry = Gq / Iyy;
    const T brz = (-off * Gp + Ixx * Gr) * idet;
    T blm = brx, bmm = bry, bnm = brz;
    // rx = lm - (q Iwz - r Iwy); ry = mm - (r Iwx - p Iwz); rz = nm - (p Iwy - q Iwx)
    T bIwx = -bry * r + brz * q;
    T bIwy = brx * r - brz * p;
    T bIwz = -brx * q + bry * p;
    bq += -brx * Iwz + brz * Iwx;
    br += brx * Iwy - bry * Iwx;
    bp += bry * Iwz - brz * Iwy;
    bp += bIwx * Ixx + bIwz * off;
    bq += bIwy * Iyy;
    br += bIwx * off + bIwz * Izz;
    // euler kinematics
    bp += Gphi;
    bq += Gphi * sph * tth + Gth * cph + Gpsi * sph * ict;
    br += Gphi * cph * tth - Gth * sph + Gpsi * cph * ict;
    bphi += Gphi * (cph * tth * q - sph * tth * r) + Gth * (-sph * q - cph * r) + Gpsi * (cph * q - sph * r) * ict;
    // d tan/dth = 1/cth^2 ; d (1/cth)/dth = sth/cth^2
    bth += Gphi * (sph * q + cph * r) * ict * ict + Gpsi * (sph * q + cph * r) * sth * ict * ict;
    // body accelerations
    const T bfx = Gu * im, bfy = Gv * im, bfz = Gw * im;
    bq += -Gu * w + Gw * u;  br += Gu * v - Gv * u;  bp += Gv * w - Gw * v;
    bw += -Gu * q + Gv * p;  bv += Gu * r - Gw * p;  bu += -Gv * r + Gw * q;
    // position kinematics
    bu += Gx * cth * cps + Gy * cth * sps - Gz * sth;
    bv += Gx * m01 + Gy * m11 + Gz * sph * cth;
    bw += Gx * m02 + Gy * m12 + Gz * cph * cth;
    // d/dphi of the rotation entries
    bphi += Gx * ((sph * sps + cph * sth * cps) * v + (cph * sps - sph * sth * cps) * w)
          + Gy * ((-sph * cps + cph * sth * sps) * v + (-cph * cps - sph * sth * sps) * w)
          + Gz * (cph * cth * v - sph * cth * w);
    bth += Gx * (-sth * cps * u + sph * cth * cps * v + cph * cth * cps * w)
         + Gy * (-sth * sps * u + sph * cth * sps * v + cph * cth * sps * w)
         + Gz * (-cth * u - sph * sth * v - cph * sth * w);
    bpsi += Gx * (-cth * sps * u + (-cph * cps - sph * sth * sps) * v + (sph * cps - cph * sth * sps) * w)
          + Gy * (cth * cps * u + (-cph * sps + sph * sth * cps) * v + (sph * sps + cph * sth * cps) * w);
    // forces
    const T bD = -bfx * ca * cb - bfy * sb - bfz * sa * cb;
    const T bY = -bfx * ca * sb + bfy * cb - bfz * sa * sb;
    const T bL = bfx * sa - bfz * ca;
    T balpha = bfx * (sa * cb * D + sa * sb * Y + ca * L) + bfz * (-ca * cb * D - ca * sb * Y + sa * L);
    T bbeta = bfx * (ca * sb * D - ca * cb * Y) + bfy * (-cb * D - sb * Y) + bfz * (sa * sb * D - sa * cb * Y);
    bth += -bfx * cth * gm - bfy * sph * sth * gm - bfz * cph * sth * gm;
    bphi += bfy * cph * cth * gm - bfz * sph * cth * gm;
    const T bthr = bfx * ce + bfz * se;
    // L = qS CL, ...
    T bqS = bL * CL + bD * CD + bY * CY + cch * (blm * Cl + bmm * Cm + bnm * Cn);
    const T bCL = bL * qS, bCD = bD * qS, bCY = bY * qS, bCl = blm * qSc, bCm = bmm * qSc, bCn = bnm * qSc;
    balpha += bCL * T(pc[W_CL_ALPHA]) + bCD * T(pc[W_CD_ALPHA]) + bCm * T(pc[W_CM_ALPHA]);
    bbeta += bCY * T(pc[W_CY_BETA]) + bCl * T(pc[W_CLL_BETA]) + bCn * T(pc[W_CN_BETA]);
    const T lonq = bCL * T(pc[W_CL_Q]) + bCD * T(pc[W_CD_Q]) + bCm * T(pc[W_CM_Q]);
    const T latp = bCY * T(pc[W_CY_P]) + bCl * T(pc[W_CLL_P]) + bCn * T(pc[W_CN_P]);
    const T latr = bCY * T(pc[W_CY_R]) + bCl * T(pc[W_CLL_R]) + bCn * T(pc[W_CN_R]);
    bq += lonq * c2v;
    bp += latp * b2v;
    br += latr * b2v;
    const T bc2v = lonq * q, bb2v = latp * p + latr * r;
    const T bde = bCL * T(pc[W_CL_DE]) + bCD * T(pc[W_CD_DE]) + bCm * T(pc[W_CM_DE]);
    const T bda = bCY * T(pc[W_CY_DA]) + bCl * T(pc[W_CLL_DA]) + bCn * T(pc[W_CN_DA]);
    const T bdr = bCY * T(pc[W_CY_DR]) + bCl * T(pc[W_CLL_DR]) + bCn * T(pc[W_CN_DR]);
    // c2v = c/(2V), b2v = b/(2V)  ->  d/dV = -x/V
    T bV = -(bc2v * c2v + bb2v * b2v) / V;
    T bV2 = bqS * hrs;
    // clamps pass the gradient inside and AT the bounds (torch.clamp), atan' = 1/(1+x^2)
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

  APG_HD static void step(const T* s, const T* a, T dt, const float* pc, T* o) {
    eval<false>(s, a, dt, pc, o, nullptr, nullptr, nullptr);
  }
  APG_HD static void step_adj(const T* s, const T* a, T dt, const float* pc, const T* g, T* gs, T* ga) {
    eval<true>(s, a, dt, pc, nullptr, g, gs, ga);
  }
  APG_HD static T loss(const T* sn, const T* ref, const T* a, const T* /*cur0*/, int /*k*/, int /*h*/) {
    T l = 0;
    for (int i = 0; i < 3; ++i) {
      const T dp = sn[i] - ref[i], da = a[1 + i] - T(0.5);
      l += T(10) * dp * dp + T(0.1) * da * da;
    }
    return l;
  }
  APG_HD static void loss_grad(const T* sn, const T* ref, const T* a, const T* /*cur0*/, int /*k*/, int /*h*/,
                               T* gs, T* ga) {
    for (int i = 0; i < 3; ++i) {
      gs[i] += T(20) * (sn[i] - ref[i]);
      ga[1 + i] += T(0.2) * (a[1 + i] - T(0.5));
    }
  }
};

// =========================================================================================================
// Cartpole
// =========================================================================================================
template <typename T>
struct Cartpole {
  static constexpr int S = 4, A = 1, REFW = 0;      // the reference is made from the start state (make_reference)
  static constexpr bool SIGMOID_ACTIONS = false;     // the net ends in tanh; no sigmoid afterwards

  template <bool ADJ>
  APG_HD static void eval(const T* s, const T* a, T dt, const float* pc, T* o, const T* g, T* gs, T* ga) {
    const T GRAV = T(9.81);
    const T mp = pc[C_MASSPOLE], len = pc[C_LENGTH], fr = pc[C_FRICTION];
    const T M = mp + T(pc[C_MASSCART]), pml = mp * len;
    const T fscale = T(pc[C_MAXFORCE]) * T(0.5);
    const T xd = s[1], th = s[2], thd = s[3];
    const T F = a[0] * fscale;
    T sn, cs; sincos_(th, &sn, &cs);
    const T N1 = T(-2) * pml * thd * thd * sn + T(3) * mp * GRAV * sn * cs + T(4) * F - T(4) * fr * xd;
    const T D1 = T(4) * M - T(3) * mp * cs * cs;
    const T N2 = T(-3) * pml * thd * thd * sn * cs + T(6) * M * GRAV * sn + T(6) * (F - fr * xd) * cs;
    const T D2 = T(4) * len * M - T(3) * pml * cs * cs;
    const T xacc = N1 / D1, thacc = N2 / D2;
    if (!ADJ) {
      T sd, cd; sincos_(thd * dt, &sd, &cd);
      o[0] = s[0] + xd * dt;
      o[1] = xd + xacc * dt;
      o[2] = atan2_(sn * cd + cs * sd, cs * cd - sn * sd);
      o[3] = thd + thacc * dt;
      return;
    }
    const T c2 = cs * cs - sn * sn;
    const T dN1_th = T(-2) * pml * thd * thd * cs + T(3) * mp * GRAV * c2;
    const T dD1_th = T(6) * mp * cs * sn;
    const T dN2_th = T(-3) * pml * thd * thd * c2 + T(6) * M * GRAV * cs - T(6) * (F - fr * xd) * sn;
    const T dD2_th = T(6) * pml * cs * sn;
    const T xa_th = (dN1_th - xacc * dD1_th) / D1, ta_th = (dN2_th - thacc * dD2_th) / D2;
    const T xa_xd = T(-4) * fr / D1, ta_xd = T(-6) * fr * cs / D2;
    const T xa_td = T(-4) * pml * thd * sn / D1, ta_td = T(-6) * pml * thd * sn * cs / D2;
    const T xa_F = T(4) / D1, ta_F = T(6) * cs / D2;
    const T gx1 = dt * g[1], gt3 = dt * g[3];
    gs[0] = g[0];
    gs[1] = g[1] + dt * g[0] + gx1 * xa_xd + gt3 * ta_xd;
    gs[2] = g[2] + gx1 * xa_th + gt3 * ta_th;          // theta' = wrap(theta + thd*dt): unit derivative
    gs[3] = g[3] + dt * g[2] + gx1 * xa_td + gt3 * ta_td;
    ga[0] = fscale * (gx1 * xa_F + gt3 * ta_F);
  }
  APG_HD static void step(const T* s, const T* a, T dt, const float* pc, T* o) {
    eval<false>(s, a, dt, pc, o, nullptr, nullptr, nullptr);
  }
  APG_HD static void step_adj(const T* s, const T* a, T dt, const float* pc, const T* g, T* gs, T* ga) {
    eval<true>(s, a, dt, pc, nullptr, g, gs, ga);
  }
  // ref[k] = cur0 * (1 - k/(h-1)) for k < h-1, 0 for the last step (train_cartpole.py:103-110); weights [0,3,10,1]
  APG_HD static T ref_scale(int k, int h) { return k < h - 1 ? T(1) - T(1) / T(h - 1) * T(k) : T(0); }
  APG_HD static T loss(const T* sn, const T* /*ref*/, const T* a, const T* cur0, int k, int h) {
    const T sc = ref_scale(k, h);
    const T d1 = sn[1] - cur0[1] * sc, d2 = sn[2] - cur0[2] * sc, d3 = sn[3] - cur0[3] * sc;
    return T(3) * d1 * d1 + T(10) * d2 * d2 + d3 * d3 + T(0.01) * a[0] * a[0];
  }
  APG_HD static void loss_grad(const T* sn, const T* /*ref*/, const T* a, const T* cur0, int k, int h, T* gs,
                               T* ga) {
    const T sc = ref_scale(k, h);
    gs[1] += T(6) * (sn[1] - cur0[1] * sc);
    gs[2] += T(20) * (sn[2] - cur0[2] * sc);
    gs[3] += T(2) * (sn[3] - cur0[3] * sc);
    ga[0] += T(0.02) * a[0];
  }
};

}  // namespace apg
