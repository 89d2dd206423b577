// A host that is neither Python nor torch: plain C++ + the CUDA runtime, linked against libalmpc.so through
// include/almpc.h.  Solves a batch of pendulum tracking problems (T = 5) and differentiates through the solution.
//
//   nvcc -std=c++17 -I include examples/c_host.cpp -o examples/c_host -L deq_mpc_corl_b200 -lalmpc \
//        -Xlinker -rpath -Xlinker '$ORIGIN/../deq_mpc_corl_b200'
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "almpc.h"

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } \
  } while (0)

template <class T>
static T* to_device(const std::vector<T>& h) {
  T* d = nullptr;
  if (cudaMalloc(&d, h.size() * sizeof(T)) != cudaSuccess) return nullptr;
  cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
  return d;
}

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 1024, T = 5;
  int32_t nx = 0, nu = 0;
  if (almpc_model_dims(ALMPC_DYN_PENDULUM, &nx, &nu)) { printf("%s\n", almpc_last_error()); return 1; }
  const int n = nx + nu, m = T * nx + 2 * T * nu;

  almpc_desc d = {};
  d.B = B; d.T = T; d.dyn_id = ALMPC_DYN_PENDULUM; d.al_iter = 2; d.max_newton = 4;
  d.rho_mult = 10.0; d.rho_max = 1e8;
  d.dyn_params[0] = 0.05;  // dt
  d.u_lower[0] = -2.0; d.u_upper[0] = 2.0;

  // tracking cost 1/2 (z - z_ref)^T diag(Q) (z - z_ref): Qd, q = -Qd * z_ref; warm start = reference
  std::vector<double> x0(B * nx), Qd(B * T * n), q(B * T * n), xu(B * T * n), lam(B * m, 0.0), rho(B, 1.0), grad(B * T * n);
  unsigned s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / (1 << 24) - 0.5; };
  for (int b = 0; b < B; ++b) {
    x0[b * nx + 0] = 3.0 * rnd(); x0[b * nx + 1] = rnd();
    for (int t = 0; t < T; ++t)
      for (int i = 0; i < n; ++i) {
        const int e = (b * T + t) * n + i;
        const double ref = i < nx ? x0[b * nx + i] * (1.0 - 0.1 * t) : 0.0;
        Qd[e] = i < nx ? 1.0 : 1e-3;
        q[e] = -Qd[e] * ref;
        xu[e] = ref;
        grad[e] = rnd();
      }
  }
  double *x0_d = to_device(x0), *Qd_d = to_device(Qd), *q_d = to_device(q), *xu_d = to_device(xu), *lam_d = to_device(lam),
         *rho_d = to_device(rho), *grad_d = to_device(grad);
  double *factor_d, *dyn_res_d, *gQ_d, *gq_d;
  float *x_d, *u_d;
  int32_t *status_d, *info_d;
  void* ws_d;
  CK(cudaMalloc(&factor_d, (size_t)B * almpc_factor_elems(&d) * sizeof(double)));
  CK(cudaMalloc(&ws_d, almpc_workspace_bytes(&d)));
  CK(cudaMalloc(&x_d, (size_t)B * T * nx * sizeof(float)));
  CK(cudaMalloc(&u_d, (size_t)B * T * nu * sizeof(float)));
  CK(cudaMalloc(&dyn_res_d, B * sizeof(double)));
  CK(cudaMalloc(&status_d, B * sizeof(int32_t)));
  CK(cudaMalloc(&info_d, ALMPC_INFO_INTS * sizeof(int32_t)));
  CK(cudaMalloc(&gQ_d, (size_t)B * T * n * sizeof(double)));
  CK(cudaMalloc(&gq_d, (size_t)B * T * n * sizeof(double)));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));

  if (almpc_forward_f64(&d, x0_d, Qd_d, q_d, xu_d, lam_d, rho_d, factor_d, x_d, u_d, dyn_res_d, status_d, info_d, nullptr,
                        ws_d, st)) { printf("forward: %s\n", almpc_last_error()); return 1; }
  if (almpc_backward_f64(&d, factor_d, xu_d, grad_d, status_d, gQ_d, gq_d, st)) { printf("backward: %s\n", almpc_last_error()); return 1; }
  CK(cudaStreamSynchronize(st));

  std::vector<int32_t> info(ALMPC_INFO_INTS), status(B);
  std::vector<double> res(B), gq(B * T * n);
  CK(cudaMemcpy(info.data(), info_d, info.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(status.data(), status_d, B * sizeof(int32_t), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(res.data(), dyn_res_d, B * sizeof(double), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(gq.data(), gq_d, gq.size() * sizeof(double), cudaMemcpyDeviceToHost));
  double worst = 0, gsum = 0;
  int ok = 0;
  for (int b = 0; b < B; ++b) { worst = fmax(worst, res[b]); ok += status[b] == (ALMPC_STATUS_FINITE | ALMPC_STATUS_CHOL_OK); }
  for (double v : gq) gsum += v * v;
  printf("%s: B=%d AL iterations %d, newton steps [%d, %d], max ||c+|| %.3e, status ok %d/%d, ||grad_q|| %.6e\n",
         almpc_version(), B, info[ALMPC_INFO_AL_DONE], info[ALMPC_INFO_NEWTON0], info[ALMPC_INFO_NEWTON0 + 1], worst, ok, B,
         sqrt(gsum));
  return (ok == B && std::isfinite(gsum)) ? 0 : 2;
}
