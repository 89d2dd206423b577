// Device dynamics functors for the fused AL-MPC kernels (sm_100a).
//
// Every model is the continuous ODE xdot = f(x,u) that the reference discretises with one explicit
// RK4 step and zero-order-hold control:
//   RexQuad        deqmpc/rex_quadrotor.py:52-127      (rk4 :98-107)
//   FlyingCartpole deqmpc/flying_cartpole2d.py:50-133  (rk4 :81-91)
//   Pendulum       deqmpc/my_envs/pendulum1l/src/dynamics_gpu.cu:13-86   (CasADi straight-line code)
//   Cartpole1L     deqmpc/my_envs/cartpole1l/src/dynamics_gpu.cu:13-2481
//   Cartpole2L     deqmpc/my_envs/cartpole2l/src/dynamics_gpu.cu:13-11079
// They are written from the physics, templated on the value type V so the same
// code runs on plain scalars (line search, residuals) and on dual numbers (forward-mode Jacobian
// columns, one tangent per thread).
#pragma once
#ifndef ALMPC_MODELS_NO_CUDA
#include <cuda_runtime.h>
#endif

namespace almpc {

// ------------------------------------------------------------------------------------------------
// dual numbers: value + one tangent
// ------------------------------------------------------------------------------------------------
template <typename S>
struct Dual {
  S v, d;
  __device__ __forceinline__ Dual() {}
  __device__ __forceinline__ Dual(S v_) : v(v_), d(S(0)) {}
  __device__ __forceinline__ Dual(S v_, S d_) : v(v_), d(d_) {}
};
template <typename S> __device__ __forceinline__ Dual<S> operator+(Dual<S> a, Dual<S> b) { return {a.v + b.v, a.d + b.d}; }
template <typename S> __device__ __forceinline__ Dual<S> operator-(Dual<S> a, Dual<S> b) { return {a.v - b.v, a.d - b.d}; }
template <typename S> __device__ __forceinline__ Dual<S> operator-(Dual<S> a) { return {-a.v, -a.d}; }
template <typename S> __device__ __forceinline__ Dual<S> operator*(Dual<S> a, Dual<S> b) { return {a.v * b.v, a.v * b.d + a.d * b.v}; }
template <typename S> __device__ __forceinline__ Dual<S> operator+(Dual<S> a, S b) { return {a.v + b, a.d}; }
template <typename S> __device__ __forceinline__ Dual<S> operator+(S b, Dual<S> a) { return {a.v + b, a.d}; }
template <typename S> __device__ __forceinline__ Dual<S> operator-(Dual<S> a, S b) { return {a.v - b, a.d}; }
template <typename S> __device__ __forceinline__ Dual<S> operator-(S b, Dual<S> a) { return {b - a.v, -a.d}; }
template <typename S> __device__ __forceinline__ Dual<S> operator*(Dual<S> a, S b) { return {a.v * b, a.d * b}; }
template <typename S> __device__ __forceinline__ Dual<S> operator*(S b, Dual<S> a) { return {a.v * b, a.d * b}; }

template <typename S> __device__ __forceinline__ S recip(S a) { return S(1) / a; }
template <typename S> __device__ __forceinline__ Dual<S> recip(Dual<S> a) {
  S r = S(1) / a.v;
  return {r, -a.d * r * r};
}
__device__ __forceinline__ void sincos_(double a, double& s, double& c) { sincos(a, &s, &c); }
__device__ __forceinline__ void sincos_(float a, float& s, float& c) { sincosf(a, &s, &c); }
template <typename S> __device__ __forceinline__ void sincos_(Dual<S> a, Dual<S>& s, Dual<S>& c) {
  S sv, cv;
  sincos_(a.v, sv, cv);
  s = {sv, cv * a.d};
  c = {cv, -sv * a.d};
}
// v + 2 t with one fused multiply-add (2 t is exact, so this is bit-identical to v + (t + t))
template <typename T, typename W> __device__ __forceinline__ T twice_plus(T t, W v) { return v + (t + t); }
#ifndef ALMPC_MODELS_NO_CUDA
__device__ __forceinline__ double twice_plus(double t, double v) { return fma(2.0, t, v); }
__device__ __forceinline__ float twice_plus(float t, float v) { return fmaf(2.0f, t, v); }
template <typename S> __device__ __forceinline__ Dual<S> twice_plus(Dual<S> t, Dual<S> v) {
  return {twice_plus(t.v, v.v), twice_plus(t.d, v.d)};
}
template <typename S> __device__ __forceinline__ Dual<S> twice_plus(Dual<S> t, S v) { return {twice_plus(t.v, v), t.d + t.d}; }
#endif
template <typename S> __device__ __forceinline__ S value_of(S a) { return a; }
template <typename S> __device__ __forceinline__ S value_of(Dual<S> a) { return a.v; }

// ------------------------------------------------------------------------------------------------
// parameter block layouts (host side: deq_mpc_corl_b200/dynamics.py)
// ------------------------------------------------------------------------------------------------
// common: p[0] = dt
// quadrotor family:
//   1 act_scale  2 u_offset(u_hover)  3 kf(force)  4 kf(moment)  5 bf  6 km  7 1/mass
//   8..10 mass*g (body-frame gravity source, world vector)   11..13 Bf (constant body force)
//   14..22 J row-major   23..31 Jinv row-major   32..39 arm_i = L*ss_i (x,y) for the 4 motors
//   40 pendulum g_z      41 1/L_pend
enum { P_DT = 0, P_ACT = 1, P_UOFF = 2, P_KFF = 3, P_KFM = 4, P_BF = 5, P_KM = 6, P_INVM = 7, P_MG = 8,
       P_BFV = 11, P_J = 14, P_JINV = 23, P_ARM = 32, P_PGZ = 40, P_PINVL = 41 };

template <typename V>
__device__ __forceinline__ void cross3(const V* a, const V* b, V* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

// v' = v + 2 w (r x v) + 2 r x (r x v)   (scalar-first unit quaternion [w, r])
template <typename V, typename W>
__device__ __forceinline__ void quatrot(V w, const V* r, const W* v, V* o) {
  V rv[3], rrv[3];
  rv[0] = r[1] * v[2] - r[2] * v[1];
  rv[1] = r[2] * v[0] - r[0] * v[2];
  rv[2] = r[0] * v[1] - r[1] * v[0];
  cross3(r, rv, rrv);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    V t = w * rv[i] + rrv[i];
    o[i] = twice_plus(t, v[i]);
  }
}

// shared rigid-body part of the quadrotor / flying cartpole ODE.
//   m: MRP, v: body velocity, w: body rates, u: raw controls -> pdot, mdot, vdot, wdot, and q=[q0,qv]
template <typename S, typename V>
__device__ __forceinline__ void quad_body(const V* m, const V* v, const V* w, const V* u, const S* p,
                                          V* pdot, V* mdot, V* vdot, V* wdot, V& q0, V* qv) {
  V n2 = m[0] * m[0] + m[1] * m[1] + m[2] * m[2];
  V den = recip(n2 + S(1));
  q0 = (S(1) - n2) * den;
  V den2 = den + den;  // exact doubling: (2 m) den == m (2 den)
#pragma unroll
  for (int i = 0; i < 3; ++i) qv[i] = m[i] * den2;
  // controls: u <- act_scale * (u + u_hover)
  V us[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) us[i] = (u[i] + p[P_UOFF]) * p[P_ACT];
  // forces (body frame): thrust + gravity rotated by the inverse attitude + constant bias
  V F[3];
  {
    V nqv[3] = {-qv[0], -qv[1], -qv[2]};
    const S mg[3] = {p[P_MG], p[P_MG + 1], p[P_MG + 2]};
    quatrot(q0, nqv, mg, F);
    V Fz = ((us[0] * p[P_KFF] + us[1] * p[P_KFF]) + us[2] * p[P_KFF]) + us[3] * p[P_KFF];
    F[0] = F[0] + p[P_BFV];
    F[1] = F[1] + p[P_BFV + 1];
    F[2] = (F[2] + Fz) + p[P_BFV + 2];
  }
  // moments: yaw drag + sum_i arm_i x [0,0,kf u_i + bf]
  V tau[3];
  {
    V f0 = us[0] * p[P_KFM] + p[P_BF], f1 = us[1] * p[P_KFM] + p[P_BF];
    V f2 = us[2] * p[P_KFM] + p[P_BF], f3 = us[3] * p[P_KFM] + p[P_BF];
    tau[0] = ((f0 * p[P_ARM + 1] + f1 * p[P_ARM + 3]) + f2 * p[P_ARM + 5]) + f3 * p[P_ARM + 7];
    tau[1] = -(((f0 * p[P_ARM + 0] + f1 * p[P_ARM + 2]) + f2 * p[P_ARM + 4]) + f3 * p[P_ARM + 6]);
    tau[2] = (((us[0] - us[1]) + us[2]) - us[3]) * p[P_KM];
  }
  // MRP kinematics: 1/4 ((1-|m|^2) w + 2 m x w + 2 (m.w) m)
  {
    V mw = m[0] * w[0] + m[1] * w[1] + m[2] * w[2];
    V mxw[3];
    cross3(m, w, mxw);
    // 1/4 (a w + 2 (m x w + (m.w) m)) with the exact power-of-two scalings folded in: a/4 w + (m x w + (m.w) m)/2
    V a4 = (S(1) - n2) * S(0.25);
#pragma unroll
    for (int i = 0; i < 3; ++i) mdot[i] = a4 * w[i] + (mxw[i] + mw * m[i]) * S(0.5);
  }
  quatrot(q0, qv, v, pdot);
  {
    V wxv[3];
    cross3(w, v, wxv);
#pragma unroll
    for (int i = 0; i < 3; ++i) vdot[i] = F[i] * p[P_INVM] - wxv[i];
  }
  {
    V Jw[3], wJw[3], r[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) Jw[i] = (w[0] * p[P_J + 3 * i] + w[1] * p[P_J + 3 * i + 1]) + w[2] * p[P_J + 3 * i + 2];
    cross3(w, Jw, wJw);
#pragma unroll
    for (int i = 0; i < 3; ++i) r[i] = tau[i] - wJw[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) wdot[i] = (r[0] * p[P_JINV + 3 * i] + r[1] * p[P_JINV + 3 * i + 1]) + r[2] * p[P_JINV + 3 * i + 2];
  }
}

struct RexQuad {
  static constexpr int NX = 12, NU = 4;
  // the ODE does not depend on the position r = x[0..2]: d x+ / d r = [I; 0], no tangent pass needed
  static constexpr int NTRIV = 3;
  template <typename S, typename V>
  __device__ __forceinline__ static void f(const V* x, const V* u, V* xd, const S* p) {
    V q0, qv[3];
    quad_body<S, V>(x + 3, x + 6, x + 9, u, p, xd, xd + 3, xd + 6, xd + 9, q0, qv);
  }
};

struct FlyingCartpole {
  static constexpr int NX = 14, NU = 4;
  static constexpr int NTRIV = 3;  // position columns are identity (as for RexQuad)
  // state [r(3), m(3), theta, v(3), w(3), theta_dot]
  template <typename S, typename V>
  __device__ __forceinline__ static void f(const V* x, const V* u, V* xd, const S* p) {
    V q0, qv[3];
    quad_body<S, V>(x + 3, x + 7, x + 10, u, p, xd, xd + 3, xd + 7, xd + 10, q0, qv);
    V acc[3];
    quatrot(q0, qv, xd + 7, acc);  // world-frame acceleration; only x is used
    V s, c;
    sincos_(x[6], s, c);
    xd[6] = x[13];
    xd[13] = (s * p[P_PGZ] + acc[0] * c) * p[P_PINVL];
  }
};

struct Pendulum {
  static constexpr int NX = 2, NU = 1;
  static constexpr int NTRIV = 0;
  // qdd = 4 u - 2 * (9.81 sin q)
  template <typename S, typename V>
  __device__ __forceinline__ static void f(const V* x, const V* u, V* xd, const S* p) {
    V s, c;
    sincos_(x[0], s, c);
    xd[0] = x[1];
    xd[1] = u[0] * S(4) - (s * S(9.81)) * S(2);
  }
};

struct Cartpole1L {
  static constexpr int NX = 4, NU = 1;
  static constexpr int NTRIV = 1;  // cart position does not enter the ODE
  // M = [[11,-c],[-c,2]], b = [s thd^2, -9.81 s], M qdd + b = [u, 0]
  template <typename S, typename V>
  __device__ __forceinline__ static void f(const V* x, const V* u, V* xd, const S* p) {
    V s, c;
    sincos_(x[1], s, c);
    V r0 = u[0] - s * x[3] * x[3];
    V r1 = s * S(9.81);
    V idet = recip(S(22) - c * c);
    xd[0] = x[2];
    xd[1] = x[3];
    xd[2] = (r0 * S(2) + c * r1) * idet;
    xd[3] = (c * r0 + r1 * S(11)) * idet;
  }
};

struct Cartpole2L {
  static constexpr int NX = 6, NU = 1;
  static constexpr int NTRIV = 1;  // cart position does not enter the ODE
  // cart (10 kg) + two unit links, relative angles; M(q) qdd = tau - b(q,qd), solved by LDL^T of the 3x3
  template <typename S, typename V>
  __device__ __forceinline__ static void f(const V* x, const V* u, V* xd, const S* p) {
    V s1, c1, s2, c2, sa, ca;
    sincos_(x[1], s1, c1);
    sincos_(x[2], s2, c2);
    sincos_(x[1] + x[2], sa, ca);
    V thd1 = x[4], thd2 = x[5], ad = thd1 + thd2;
    V M01 = -(c1 * S(2)) - ca, M02 = -ca, M11 = c2 * S(2) + S(5), M12 = c2 + S(2);
    V b0 = (s1 * S(2)) * thd1 * thd1 + sa * ad * ad;
    V b1 = -(s2 * ((thd1 * S(2)) * thd2 + thd2 * thd2)) - (s1 * S(2) + sa) * S(9.81);
    V b2 = s2 * thd1 * thd1 - sa * S(9.81);
    V r0 = u[0] - b0, r1 = -b1, r2 = -b2;
    // LDL^T with M00 = 12, M22 = 2
    V l10 = M01 * S(1.0 / 12.0), l20 = M02 * S(1.0 / 12.0);
    V d1 = M11 - l10 * M01;
    V id1 = recip(d1);
    V t21 = M12 - l20 * M01;
    V l21 = t21 * id1;
    V d2 = (S(2) - l20 * M02) - l21 * t21;
    V y1 = r1 - l10 * r0;
    V y2 = (r2 - l20 * r0) - l21 * y1;
    V z2 = y2 * recip(d2);
    V z1 = y1 * id1 - l21 * z2;
    V z0 = (r0 * S(1.0 / 12.0) - l10 * z1) - l20 * z2;
    xd[0] = x[3];
    xd[1] = x[4];
    xd[2] = x[5];
    xd[3] = z0;
    xd[4] = z1;
    xd[5] = z2;
  }
};

// ------------------------------------------------------------------------------------------------
// one explicit RK4 step: y0 + dt/6 (k1 + 2 k2 + 2 k3 + k4)      (rex_quadrotor.py:98-107)
// ------------------------------------------------------------------------------------------------
template <class M, typename S, typename V>
__device__ __forceinline__ void rk4_step(const V* x, const V* u, V* xn, const S* p) {
  constexpr int NX = M::NX;
  const S dt = p[P_DT];
  const S dt2 = dt * S(0.5);
  V k[NX], acc[NX], xt[NX];
  M::template f<S, V>(x, u, k, p);
#pragma unroll
  for (int i = 0; i < NX; ++i) { acc[i] = k[i]; xt[i] = x[i] + k[i] * dt2; }
  M::template f<S, V>(xt, u, k, p);
#pragma unroll
  for (int i = 0; i < NX; ++i) { acc[i] = acc[i] + k[i] * S(2); xt[i] = x[i] + k[i] * dt2; }
  M::template f<S, V>(xt, u, k, p);
#pragma unroll
  for (int i = 0; i < NX; ++i) { acc[i] = acc[i] + k[i] * S(2); xt[i] = x[i] + k[i] * dt; }
  M::template f<S, V>(xt, u, k, p);
  const S dt6 = dt / S(6);
#pragma unroll
  for (int i = 0; i < NX; ++i) xn[i] = x[i] + (acc[i] + k[i]) * dt6;
}

}  // namespace almpc
