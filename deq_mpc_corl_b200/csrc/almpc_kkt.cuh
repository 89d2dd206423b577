// almpc_kkt.cuh — k_kkt_solve: the KKT solve of the interior-point path (part of libalmpc, see almpc_kernels.cu).
//
// What it replaces: `solve_kkt` of the reference's primal-dual interior-point QP solver
// (qpth/solvers/pdipm/batch_LU.py:212-244, called from its `forward` :70, :155, :175 and from
// `DenseQPFunction.backward`, qpth/qp.py:249-253).  The reference assembles the dense
// (nz + 2 nineq + neq)^2 matrix
//
//        [ Q    0    G^T  A^T ]   x        Ktilde = K + eps diag(+1 on the x and s rows, -1 on the z and y rows)
//    K = [ 0   D(z)  D(s)  0  ]   s        (eps = 1e-7; batch_LU.py:42,58-60,113-116)
//        [ G    I    0     0  ]   z
//        [ A    0    0     0  ]   y
//
// LU-factors Ktilde with partial pivoting (torch.linalg.lu_factor), solves, and takes ONE step of iterative refinement
// against the unregularised K.  Here the two diagonal blocks are eliminated in closed form,
//
//     s = (r2 - S z) / (Z + eps),   z = w (G x + r2 / (Z + eps) - r3),   w = 1 / (S / (Z + eps) + eps)
//
// which leaves the quasi-definite system  [ Q + eps I + G^T W G   A^T ] [x]   [ r1 - G^T w (r2/(Z+eps) - r3) ]
//                                         [ A                  -eps I ] [y] = [ r4                           ]
// of size nz + neq (140 instead of 220 for the quadrotor horizon-5 QP); it is factored as L D L^T without pivoting
// (stable for quasi-definite matrices) in shared memory by one CTA per problem, and the same refinement step against K
// follows (n_refine = 1).  The backward pass of the reference solves with K itself (qp.py:249); eliminating the
// slack block of the UNREGULARISED K puts z_i / s_i ~ 1e12 of the active rows into the reduced matrix, so the backward
// factors the regularised matrix as well and takes three refinement steps against K instead (each step contracts the
// error by ~eps |Ktilde^-1|).  Run-time sizes; the reduced matrix must fit the opt-in shared memory ((nz + neq)^2
// scalars).
#pragma once
#include "almpc_common.cuh"

namespace almpc {

constexpr int kKktThreads = 256;

// shared-memory scalars of k_kkt_solve
__host__ __device__ inline size_t kkt_smem_elems(int nz, int nineq, int neq) {
  const size_t nr = (size_t)nz + neq, ld = nr | 1;
  return nr * ld + 3 * nr + 4 * (size_t)nineq + 8;
}

template <typename S>
__global__ void __launch_bounds__(kKktThreads)
k_kkt_solve(int B, int nz, int nineq, int neq, const S* __restrict__ Q_g, const S* __restrict__ G_g,
            const S* __restrict__ A_g, const S* __restrict__ s_g, const S* __restrict__ z_g, S eps, int n_refine,
            const S* __restrict__ rx_g, const S* __restrict__ rs_g, const S* __restrict__ rz_g,
            const S* __restrict__ ry_g, S* __restrict__ dx_g, S* __restrict__ ds_g, S* __restrict__ dz_g,
            S* __restrict__ dy_g, int* __restrict__ status_g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int prob = blockIdx.x;
  if (prob >= B) return;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int nr = nz + neq, ld = nr | 1;
  S* M = reinterpret_cast<S*>(smem_raw);  // reduced matrix, lower part; after the factorisation: unscaled columns of L D
  S* bv = M + (size_t)nr * ld;            // right-hand side / solution of the reduced system
  S* lx = bv + nr;                        // accumulated solution [x; y]
  S* dinv = lx + nr;                      // 1 / D_kk
  S* wv = dinv + nr;                      // w
  S* tv = wv + nineq;                     // t, then z of the current solve
  S* lz = tv + nineq;                     // accumulated z
  S* ls = lz + nineq;                     // accumulated s
  const S* Q = Q_g + (size_t)prob * nz * nz;
  const S* G = G_g + (size_t)prob * nineq * nz;
  const S* A = A_g + (size_t)prob * neq * nz;
  const S* sv = s_g + (size_t)prob * nineq;
  const S* zv = z_g + (size_t)prob * nineq;

  for (int k = tid; k < nineq; k += nthr) wv[k] = S(1) / (sv[k] / (zv[k] + eps) + eps);
  __syncthreads();
  // ---- reduced matrix (lower part): H = Q + eps I + G^T W G, then the rows of A and -eps I
  for (int idx = tid; idx < nz * nz; idx += nthr) {
    const int i = idx / nz, j = idx - i * nz;
    if (j > i) continue;
    S v = Q[idx] + ((i == j) ? eps : S(0));
    S a0 = S(0), a1 = S(0);
    int k = 0;
    for (; k + 1 < nineq; k += 2) {
      a0 += G[k * nz + i] * wv[k] * G[k * nz + j];
      a1 += G[(k + 1) * nz + i] * wv[k + 1] * G[(k + 1) * nz + j];
    }
    if (k < nineq) a0 += G[k * nz + i] * wv[k] * G[k * nz + j];
    M[i * ld + j] = v + (a0 + a1);
  }
  for (int idx = tid; idx < neq * nz; idx += nthr) {
    const int a = idx / nz, j = idx - a * nz;
    M[(nz + a) * ld + j] = A[idx];
  }
  for (int idx = tid; idx < neq * neq; idx += nthr) {
    const int a = idx / neq, c = idx - a * neq;
    if (c <= a) M[(nz + a) * ld + nz + c] = (a == c) ? -eps : S(0);
  }
  // ---- L D L^T, right-looking, one barrier per column: column k is final once the updates of columns < k are in;
  // it stays UNSCALED in M (L_ik = M[i][k] / D_kk), so no second pass over it is needed
  const int ti = tid >> 4, tj = tid & 15;  // 16 x 16 thread tile over the trailing submatrix
  bool ok = true;
  for (int k = 0; k < nr; ++k) {
    __syncthreads();
    const S d = M[k * ld + k];
    ok = ok && (d != S(0)) && finite_(d);
    const S inv = S(1) / d;
    if (tid == 0) dinv[k] = inv;
    for (int i = k + 1 + ti; i < nr; i += 16) {
      const S ci = M[i * ld + k] * inv;
      for (int j = k + 1 + tj; j <= i; j += 16) M[i * ld + j] -= ci * M[j * ld + k];
    }
  }
  __syncthreads();

  // one solve of the regularised system for the right-hand side (r1, r2, r3, r4); adds the result to (lx, ls, lz)
  auto solve_add = [&](auto r1, auto r2, auto r3, auto r4) {
    // t = r2 / (Z + eps) - r3 ;  b = [r1 - G^T (w t) ; r4]
    for (int k = tid; k < nineq; k += nthr) tv[k] = r2(k) / (zv[k] + eps) - r3(k);
    __syncthreads();
    for (int i = tid; i < nz; i += nthr) {
      S acc = S(0);
      for (int k = 0; k < nineq; ++k) acc += G[k * nz + i] * (wv[k] * tv[k]);
      bv[i] = r1(i) - acc;
    }
    for (int a = tid; a < neq; a += nthr) bv[nz + a] = r4(a);
    // forward substitution with the unit lower factor, column oriented
    for (int k = 0; k < nr; ++k) {
      __syncthreads();
      const S yk = bv[k] * dinv[k];  // = y_k / D_kk; L_ik y_k = M[i][k] / D_kk * y_k
      for (int i = k + 1 + tid; i < nr; i += nthr) bv[i] -= M[i * ld + k] * yk;
    }
    __syncthreads();
    for (int i = tid; i < nr; i += nthr) bv[i] *= dinv[i];
    // backward substitution with L^T: row k of L^T is column k of L = M[.][k] / D_kk
    for (int k = nr - 1; k >= 0; --k) {
      __syncthreads();
      const S xk = bv[k];
      // x_i -= L_ki x_k for i < k, L_ki = M[k][i] / D_ii
      for (int i = tid; i < k; i += nthr) bv[i] -= M[k * ld + i] * dinv[i] * xk;
    }
    __syncthreads();
    // z = w (G x + t) ; s = (r2 - S z) / (Z + eps)
    for (int k = tid; k < nineq; k += nthr) {
      S acc = S(0);
      for (int j = 0; j < nz; ++j) acc += G[k * nz + j] * bv[j];
      const S zk = wv[k] * (acc + tv[k]);
      lz[k] += zk;
      ls[k] += (r2(k) - sv[k] * zk) / (zv[k] + eps);
    }
    for (int i = tid; i < nr; i += nthr) lx[i] += bv[i];
    __syncthreads();
  };

  for (int i = tid; i < nr; i += nthr) lx[i] = S(0);
  for (int k = tid; k < nineq; k += nthr) { lz[k] = S(0); ls[k] = S(0); }
  __syncthreads();
  // the reference negates the residuals: r = -(rx, rs, rz, ry)   (batch_LU.py:219)
  const S* rx = rx_g + (size_t)prob * nz;
  const S* rs = rs_g + (size_t)prob * nineq;
  const S* rz = rz_g + (size_t)prob * nineq;
  const S* ry = ry_g + (size_t)prob * neq;
  solve_add([&](int i) { return -rx[i]; }, [&](int k) { return -rs[k]; }, [&](int k) { return -rz[k]; },
            [&](int a) { return -ry[a]; });
  // ---- one step of iterative refinement against the UNREGULARISED K (batch_LU.py:229-236): res = r - K l, solve,
  // add.  The residual is materialised in this problem's slice of the OUTPUT buffers (they receive the solution
  // only at the very end); every element is written and later read by the same thread.
  auto q_row = [&](int i) {  // (Q x + G^T z + A^T y)_i
    S acc = S(0);
    for (int j = 0; j < nz; ++j) acc += Q[i * nz + j] * lx[j];
    for (int k = 0; k < nineq; ++k) acc += G[k * nz + i] * lz[k];
    for (int a = 0; a < neq; ++a) acc += A[a * nz + i] * lx[nz + a];
    return acc;
  };
  S* o_x = dx_g + (size_t)prob * nz;
  S* o_s = ds_g + (size_t)prob * nineq;
  S* o_z = dz_g + (size_t)prob * nineq;
  S* o_y = dy_g + (size_t)prob * neq;
  for (int it = 0; it < n_refine; ++it) {
  for (int i = tid; i < nz; i += nthr) o_x[i] = -rx[i] - q_row(i);
  for (int k = tid; k < nineq; k += nthr) {
    S gx = S(0);
    for (int j = 0; j < nz; ++j) gx += G[k * nz + j] * lx[j];
    o_s[k] = -rs[k] - (zv[k] * ls[k] + sv[k] * lz[k]);
    o_z[k] = -rz[k] - (gx + ls[k]);
  }
  for (int a = tid; a < neq; a += nthr) {
    S ax = S(0);
    for (int j = 0; j < nz; ++j) ax += A[a * nz + j] * lx[j];
    o_y[a] = -ry[a] - ax;
  }
  __syncthreads();
  solve_add([&](int i) { return o_x[i]; }, [&](int k) { return o_s[k]; }, [&](int k) { return o_z[k]; },
            [&](int a) { return o_y[a]; });
  }
  // ---- outputs
  ok = __syncthreads_and(ok);
  for (int i = tid; i < nz; i += nthr) o_x[i] = lx[i];
  for (int k = tid; k < nineq; k += nthr) { o_s[k] = ls[k]; o_z[k] = lz[k]; }
  for (int a = tid; a < neq; a += nthr) o_y[a] = lx[nz + a];
  if (tid == 0 && status_g) status_g[prob] = ok ? 1 : 0;
}

}  // namespace almpc
