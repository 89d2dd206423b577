// Counts the floating-point operations of ONE evaluation of each model's continuous ODE f(x,u) and of
// one RK4 step, by running the device functors' source on an op-counting scalar type on the host.
// Convention: add/sub/mul = 1, fma counted as its mul + add = 2, reciprocal/division = 1,
// sincos = 2 (one sin + one cos).  Build: g++ -std=c++17 -I deq_mpc_corl_b200/csrc -DALMPC_HOST_COUNT tools/count_flops.cpp
#include <cstdio>
#include <cmath>
#define __device__
#define __forceinline__ inline
#define __host__
struct Cnt { double v; };
static long g_ops = 0;
inline Cnt operator+(Cnt a, Cnt b) { ++g_ops; return {a.v + b.v}; }
inline Cnt operator-(Cnt a, Cnt b) { ++g_ops; return {a.v - b.v}; }
inline Cnt operator*(Cnt a, Cnt b) { ++g_ops; return {a.v * b.v}; }
inline Cnt operator/(Cnt a, Cnt b) { ++g_ops; return {a.v / b.v}; }
inline Cnt operator-(Cnt a) { return {-a.v}; }
inline Cnt operator+(Cnt a, double b) { ++g_ops; return {a.v + b}; }
inline Cnt operator+(double b, Cnt a) { ++g_ops; return {a.v + b}; }
inline Cnt operator-(Cnt a, double b) { ++g_ops; return {a.v - b}; }
inline Cnt operator-(double b, Cnt a) { ++g_ops; return {b - a.v}; }
inline Cnt operator*(Cnt a, double b) { ++g_ops; return {a.v * b}; }
inline Cnt operator*(double b, Cnt a) { ++g_ops; return {a.v * b}; }
#define ALMPC_MODELS_NO_CUDA
namespace almpc {
inline Cnt recip(Cnt a) { ++g_ops; return {1.0 / a.v}; }
inline void sincos_(Cnt a, Cnt& s, Cnt& c) { g_ops += 2; s = {std::sin(a.v)}; c = {std::cos(a.v)}; }
}
#include "almpc_models.cuh"
using namespace almpc;

template <class M> void count(const char* name) {
  Cnt x[M::NX], u[M::NU], xd[M::NX], xn[M::NX];
  for (int i = 0; i < M::NX; ++i) x[i] = {0.1 * (i + 1)};
  for (int i = 0; i < M::NU; ++i) u[i] = {0.3};
  double p[64];
  for (int i = 0; i < 64; ++i) p[i] = 0.5 + 0.01 * i;
  g_ops = 0;
  M::template f<double, Cnt>(x, u, xd, p);
  long f_ops = g_ops;
  g_ops = 0;
  rk4_step<M, double, Cnt>(x, u, xn, p);
  printf("%-16s nx=%2d nu=%d  F_ode=%5ld  F_dyn(rk4)=%6ld\n", name, M::NX, M::NU, f_ops, g_ops);
}
int main() {
  count<Pendulum>("pendulum");
  count<Cartpole1L>("cartpole1l");
  count<Cartpole2L>("cartpole2l");
  count<RexQuad>("rex_quadrotor");
  count<FlyingCartpole>("flying_cartpole");
  return 0;
}
