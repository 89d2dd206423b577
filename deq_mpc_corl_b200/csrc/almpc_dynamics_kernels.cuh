// almpc_dynamics_kernels.cuh — Jacobian / stand-alone dynamics / tracking-cost / FMA-peak kernels: part of libalmpc (see almpc_kernels.cu for the overview).
#pragma once
#include "almpc_common.cuh"

namespace almpc {

// ------------------------------------------------------------------------------------------------
// kernel 1/3: dynamics + Jacobian columns.  1 thread = 1 (problem, stage t, column j of G_t = [A_t B_t]).
//   G_g[prob][t][i][j] = d f_i / d z_j at (x_t,u_t) ; c_g[prob][t*nx+i] = x_{t+1,i} - f_i ; last block: x_0 - x0
// ------------------------------------------------------------------------------------------------
template <class M, typename S, int V>
__global__ void __launch_bounds__(128, ALMPC_JAC_MINB)
k_jacobian(const KParams<S> P, const S* __restrict__ x0_g, const S* __restrict__ xu_g, S* __restrict__ G_g,
           S* __restrict__ c_g, const Ctl* __restrict__ ctl) {
  constexpr int NX = M::NX, NU = M::NU, NN = NX + NU;
  constexpr int NTRIV = M::NTRIV;    // leading state columns the ODE does not depend on: identity columns
  constexpr int NC = NN - NTRIV;     // tangent columns actually propagated per stage
  grid_dep_wait();
  if (ctl->mode == MODE_DONE) return;
  const int T = P.T;
  const int IT = (T - 1) * NC;
  const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long)P.B * IT) return;
  const int prob = (int)(gid / IT);
  const int it = (int)(gid - (long)prob * IT);
  const int t = it / NC, jj = it - t * NC, j = jj + NTRIV;
  const S* xz = xu_g + (size_t)prob * T * NN + (size_t)t * NN;
  constexpr bool SE = (V == 2);
  const int neq = SE ? (T - 1) * NX : T * NX;
  if (SE && j >= NX) {  // controls are data: their Jacobian columns are zero (al_utils_se.py:153)
    S* Gz = G_g + ((size_t)prob * (T - 1) + t) * NX * NN;
#pragma unroll
    for (int i = 0; i < NX; ++i) Gz[i * NN + j] = S(0);
    return;
  }
  Dual<S> xd[NX], ud[NU], xn[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) xd[i] = Dual<S>(xz[i], (i == j) ? S(1) : S(0));
#pragma unroll
  for (int i = 0; i < NU; ++i) ud[i] = Dual<S>(xz[NX + i], (NX + i == j) ? S(1) : S(0));
  rk4_step<M, S, Dual<S>>(xd, ud, xn, P.dyn);
  S* Gp = G_g + ((size_t)prob * (T - 1) + t) * NX * NN;
#pragma unroll
  for (int i = 0; i < NX; ++i) Gp[i * NN + j] = xn[i].d;
  S* cp = c_g + (size_t)prob * neq;
  if (jj == 0) {
#pragma unroll
    for (int i = 0; i < NX; ++i) cp[t * NX + i] = xz[NN + i] - xn[i].v;
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
      for (int c = 0; c < NTRIV; ++c) Gp[i * NN + c] = (i == c) ? S(1) : S(0);
  }
  if (!SE && it == IT - 1) {  // the x_0 - x0 block (al_utils.py:219)
    const S* z0 = xu_g + (size_t)prob * T * NN;
#pragma unroll
    for (int i = 0; i < NX; ++i) cp[(T - 1) * NX + i] = z0[i] - x0_g[(size_t)prob * NX + i];
  }
}

// ------------------------------------------------------------------------------------------------
// stand-alone dynamics (+ Jacobian): one thread per (row, tangent)
// ------------------------------------------------------------------------------------------------
template <class M, typename S>
__global__ void k_dynamics(const KParams<S> P, int64_t R, const S* __restrict__ x, const S* __restrict__ u,
                           S* __restrict__ xn_g, S* __restrict__ Jx, S* __restrict__ Ju) {
  constexpr int NX = M::NX, NU = M::NU, NN = NX + NU;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (Jx == nullptr) {
    if (gid >= R) return;
    S xs[NX], us[NU], xn[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) xs[i] = x[gid * NX + i];
#pragma unroll
    for (int i = 0; i < NU; ++i) us[i] = u[gid * NU + i];
    rk4_step<M, S, S>(xs, us, xn, P.dyn);
#pragma unroll
    for (int i = 0; i < NX; ++i) xn_g[gid * NX + i] = xn[i];
    return;
  }
  if (gid >= R * NN) return;
  const int64_t r = gid / NN;
  const int j = (int)(gid - r * NN);
  Dual<S> xd[NX], ud[NU], xn[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) xd[i] = Dual<S>(x[r * NX + i], (i == j) ? S(1) : S(0));
#pragma unroll
  for (int i = 0; i < NU; ++i) ud[i] = Dual<S>(u[r * NU + i], (NX + i == j) ? S(1) : S(0));
  rk4_step<M, S, Dual<S>>(xd, ud, xn, P.dyn);
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    if (j < NX) Jx[(r * NX + i) * NX + j] = xn[i].d;
    else Ju[(r * NX + i) * NU + (j - NX)] = xn[i].d;
  }
  if (j == 0 && xn_g) {
#pragma unroll
    for (int i = 0; i < NX; ++i) xn_g[r * NX + i] = xn[i].v;
  }
}

// ------------------------------------------------------------------------------------------------
// Tracking cost build / its backward (deqmpc/policies.py:1285-1297 compute_pf, :1245-1248 q_scaling):
//   Qd[b,t,j] = Q0[j] (1 + s[b,t]) ;  q[b,t,j] = -Qd[b,t,j] ref[b,t,j] ,  ref = [x_ref | u_ref] (fp32 network output)
// fuses the cat / cast / diag-embed / multiply chain of the caller into one pass over the batch.
// ------------------------------------------------------------------------------------------------
template <typename S>
__global__ void k_tracking_cost(int B, int T, int nx, int nu, const S* __restrict__ Q0, const float* __restrict__ x_ref,
                                const float* __restrict__ u_ref, const float* __restrict__ scale,
                                S* __restrict__ Qd, S* __restrict__ q) {
  const int n = nx + nu;
  const long total = (long)B * T * n;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long bt = i / n;
    const int j = (int)(i - bt * n);
    const S r = (S)(j < nx ? x_ref[bt * nx + j] : u_ref[bt * nu + (j - nx)]);
    // (1 + s) is an fp32 sum in the reference (policies.py:1247: q_scaling + ones_like(q_scaling))
    const S Qv = scale ? Q0[j] * (S)(1.0f + scale[bt]) : Q0[j];
    Qd[i] = Qv;
    q[i] = -(Qv * r);
  }
}

// grad_xref / grad_uref = -Qd * grad_q ; grad_scale[b,t] = sum_j Q0[j] (grad_Q - grad_q * ref)   (one thread per (b,t))
template <typename S>
__global__ void k_tracking_cost_bwd(int B, int T, int nx, int nu, const S* __restrict__ Q0,
                                    const float* __restrict__ x_ref, const float* __restrict__ u_ref,
                                    const float* __restrict__ scale, const S* __restrict__ gQ,
                                    const S* __restrict__ gq, float* __restrict__ gx, float* __restrict__ gu,
                                    float* __restrict__ gs) {
  const int n = nx + nu;
  const long bt = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (bt >= (long)B * T) return;
  const S sc = scale ? (S)(1.0f + scale[bt]) : S(1);
  S acc = S(0);
  for (int j = 0; j < n; ++j) {
    const S r = (S)(j < nx ? x_ref[bt * nx + j] : u_ref[bt * nu + (j - nx)]);
    const S g = gq[bt * n + j];
    const S v = -(Q0[j] * sc) * g;
    if (j < nx) { if (gx) gx[bt * nx + j] = (float)v; }
    else if (gu) gu[bt * nu + (j - nx)] = (float)v;
    acc += Q0[j] * ((gQ ? gQ[bt * n + j] : S(0)) - g * r);
  }
  if (gs) gs[bt] = (float)acc;
}

// ------------------------------------------------------------------------------------------------
// FMA-pipe peak micro-benchmark (roofline denominator for bench.py)
// ------------------------------------------------------------------------------------------------
template <typename S>
__global__ void k_fma_peak(int iters, S* sink) {
  S a0 = S(1.0) + threadIdx.x * S(1e-6), a1 = a0 + S(0.1), a2 = a0 + S(0.2), a3 = a0 + S(0.3);
  S a4 = a0 + S(0.4), a5 = a0 + S(0.5), a6 = a0 + S(0.6), a7 = a0 + S(0.7);
  const S m = S(0.999999), c = S(1e-7);
  for (int i = 0; i < iters; ++i) {
    a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
    a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
  }
  S s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == S(-1)) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


}  // namespace almpc
