// libalmpc: batched augmented-Lagrangian MPC solver for sm_100a (B200).
//
// One Gauss-Newton step of the reference's NewtonAL.forward (qpth/al_utils.py:451-576) for every problem
// of the batch is THREE stream-ordered launches, each with the thread mapping / register budget / occupancy
// its phase wants (the first version fused them into one warp-per-problem kernel and was latency-bound at
// 16 warps/SM because the register-hungry dynamics and the pivot-chain-bound factorisation shared one
// register allocation; profiles/r01_v1_*):
//
//   k_jacobian      1 thread = 1 (problem, stage t, tangent column j): dynamics + Jacobians G_t=[A_t B_t]
//                   by forward-mode duals, residuals c_t                                       al_utils.py:233-265
//   k_factor_solve  1 warp = 1 problem (~28 warps/SM): residual statistics, (optional) dual/penalty update of
//                   the previous AL iteration (AL_mpc.py:315-325), merit (al_utils.py:52-77), gradient
//                   (:113-117), block-tridiagonal Cholesky of H = diag(Qd) + rho Jc^T Jc and both triangular
//                   sweeps (:119-120, 510-515; the reference factors the dense matrix, H is exactly block
//                   tridiagonal) -> Newton direction
//   k_line_search   1 thread = 1 (problem, candidate 2^-k, stage t): 20-point line search, first-argmin,
//                   strict-decrease accept (al_utils.py:618-642), iterate update, and the batch-global stop
//                   statistics -> device-side control block for the next launches           al_utils.py:551-564
//
// The batch-global decisions of the reference (stop rule on ||c+||_F over the WHOLE batch, the rho_max
// exit of al_solve_stream) are taken on the device by the last CTA of k_line_search (deterministic two-stage
// reduction), so a solve is a fixed sequence of launches without any host synchronisation; launches that
// find the solve finished return immediately.
//
// The block factor streams to global memory as it is produced (it is the cached factor the implicit
// backward needs, al_utils.py:572,590) and is read back (L2) for the backward sweep.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "almpc.h"
#include "almpc_models.cuh"

// The library is compiled as four translation units built in parallel: (fp64 | fp32) x (entry points with the
// plain / obstacle variants | the state-estimator variant, ALMPC_BUILD_SE=1); the common entry points live in
// the fp64 unit.
#ifndef ALMPC_BUILD_F64
#define ALMPC_BUILD_F64 1
#endif
#ifndef ALMPC_BUILD_F32
#define ALMPC_BUILD_F32 1
#endif
#ifndef ALMPC_BUILD_SE
#define ALMPC_BUILD_SE 0
#endif
#ifndef ALMPC_NO_VEC2
#define ALMPC_NO_VEC2 0
#endif
#ifndef ALMPC_NO_CP16
#define ALMPC_NO_CP16 0
#endif
#ifndef ALMPC_LS_MINB
#define ALMPC_LS_MINB 4  // CTAs/SM the warp line search is compiled for
#endif
#ifndef ALMPC_FAC_MINB
#define ALMPC_FAC_MINB 6  // CTAs/SM the factor kernel is compiled for (register cap 80; 5..7 measured within 2%)
#endif

namespace almpc {

constexpr int NLS = ALMPC_N_LS;
constexpr unsigned FULL = 0xffffffffu;

enum Mode { MODE_FIRST = 0, MODE_CONT = 1, MODE_DUAL_NEWTON = 2, MODE_DONE = 3 };

struct Ctl {
  int mode, a, k, total_newton;
  unsigned ticket;
  int pad;
  double res_old;
  double rho_final_max;
};

template <typename S>
struct KParams {
  S dyn[ALMPC_MAX_PARAMS];
  S ulo[ALMPC_MAX_NU], uhi[ALMPC_MAX_NU];
  double rho_mult, rho_max;
  int B, T, al_iter, max_newton, stream_mode, rho_f32, flags;
  int part_off, ticket_total;  // k_line_search: slot offset of this part's CTA partials / CTAs of all parts
  int n_obs;                   // obstacle rows per stage (Obstacle_MPC), 0 = none
  S obs_r2;                    // obstacle radius squared
  const S* obs;                // [B][T][n_obs][3] obstacle centres (device)
};

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
// 1/sqrt(x): fp32 MUFU seed (2^-22) + one cubically convergent correction in fp64 (~1 ulp):
//   r = 1 - x y^2 ;  x^-1/2 = y (1 + r/2 + 3 r^2/8 + O(r^3)).  Pivots outside the fp32 range are flagged by
// the caller (d < 1e30) and never occur for rho <= 1e12.
__device__ __forceinline__ double rsqrt_(double x) {
  float y32;  // raw MUFU.RSQ: rsqrtf() wraps it in denormal / range handling that the pivots never need
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y32) : "f"((float)x));
  const double y = (double)y32;
  const double r = fma(-x * y, y, 1.0);
  const double p = fma(0.375, r, 0.5);
  return fma(y * r, p, y);
}
__device__ __forceinline__ float rsqrt_(float x) { return rsqrtf(x); }
template <typename S> __device__ __forceinline__ S fmax0(S x) { return x > S(0) ? x : S(0); }
template <typename S> __device__ __forceinline__ bool finite_(S x) { return (x - x) == S(0); }

// Programmatic dependent launch: the step kernels are launched with stream serialization relaxed, so a kernel's
// CTAs are scheduled while the previous kernel drains; nothing of the predecessor may be read before this wait.
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename S> struct alignas(2 * sizeof(S)) Vec2 { S x, y; };

// Ampere-style asynchronous global->shared copies (LDGSTS): the factor kernel prefetches the next stage's
// blocks while the pivot chain of the current stage runs.
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// copy COUNT scalars with copy width W bytes (W = widest size dividing EVERY block size of the problem, so
// that all block offsets are W-aligned)
template <int COUNT, int W, typename S>
__device__ __forceinline__ void warp_copy_async(S* dst, const S* src, int lane) {
  constexpr int BYTES = COUNT * (int)sizeof(S);
  static_assert(BYTES % W == 0, "copy width must divide the block");
  for (int o = lane * W; o < BYTES; o += 32 * W) {
    if constexpr (W == 16) cp_async16((char*)dst + o, (const char*)src + o);
    else if constexpr (W == 8) cp_async8((char*)dst + o, (const char*)src + o);
    else cp_async4((char*)dst + o, (const char*)src + o);
  }
}
// dense global block [ROWS][NN] -> shared block with odd row stride LB (element-wise copies)
template <int ROWS, int NN, int LB, typename S>
__device__ __forceinline__ void warp_copy_block_async(S* dst, const S* src, int lane) {
  for (int e = lane; e < ROWS * NN; e += 32) {
    const int r = e / NN, c = e - r * NN;
    if constexpr (sizeof(S) == 8) cp_async8(dst + r * LB + c, src + e);
    else cp_async4(dst + r * LB + c, src + e);
  }
}
template <int NX, int NU, typename S>
struct CopyWidth {
  static constexpr int NN = NX + NU;
  static constexpr int b1 = NN * NN * (int)sizeof(S), b2 = NX * NN * (int)sizeof(S);
  static constexpr int W = (b1 % 16 == 0 && b2 % 16 == 0 && !ALMPC_NO_CP16) ? 16 : ((b1 % 8 == 0 && b2 % 8 == 0) ? 8 : 4);
};

template <typename S>
__device__ __forceinline__ S warp_sum(S v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

template <typename S>
__device__ __forceinline__ S next_rho(S rho, double mult, int f32) {
  double r = (double)rho * mult;
  if (f32) r = (double)(float)r;
  return (S)r;
}

// obstacle constraint k of stage t: c = r^2 - |p - c_k|^2 and its Jacobian row J = -2 (p - c_k)^T on the position.
// The reference builds J in an fp32 buffer (al_utils.py:378-379: torch.zeros without dtype), so the entries
// are rounded to fp32 before use.
template <typename S>
__device__ __forceinline__ S obstacle_res(const S* p, const S* ck, S r2, S* J) {
  const S d0 = p[0] - ck[0], d1 = p[1] - ck[1], d2 = p[2] - ck[2];
  if (J) {
    J[0] = (S)(float)(S(-2) * d0);
    J[1] = (S)(float)(S(-2) * d1);
    J[2] = (S)(float)(S(-2) * d2);
  }
  return r2 - ((d0 * d0 + d1 * d1) + d2 * d2);
}

// residual of inequality row `sidx` (0..SI-1) of a stage: nu upper bounds, nu lower bounds, obstacle rows
template <int NX, int NU, typename S>
__device__ __forceinline__ S ineq_res(const S* z_t, int sidx, const S* ulo, const S* uhi, const S* obs_t, S r2) {
  if (sidx < NU) return z_t[NX + sidx] - uhi[sidx];
  if (sidx < 2 * NU) return ulo[sidx - NU] - z_t[NX + sidx - NU];
  return obstacle_res<S>(z_t, obs_t + (sidx - 2 * NU) * 3, r2, nullptr);
}

// per-warp shared-memory layout of k_factor_solve, in units of S (every region 16-byte aligned)
template <int NX, int NU>
struct Layout {
  static constexpr int NN = NX + NU;
  int N, neq, nin, m, SI;
  int xu, dv, w, Gt, Ep, scr, total;
  __host__ __device__ Layout(int T, int nobs = 0, bool se = false) {
    // inequality rows per stage: control bounds, then obstacle rows.  State-estimator variant
    // (al_utils_se.py): the T-1 dynamics blocks only
    SI = se ? 0 : 2 * NU + nobs;
    N = T * NN; neq = se ? (T - 1) * NX : T * NX; nin = T * SI; m = neq + nin;
    int o = 0;
    auto take = [&](int cnt) { int at = o; o += (cnt + 3) & ~3; return at; };  // regions 16-byte aligned for float and double
    xu = take(N);
    dv = take(N);
    w = take(neq);
    // n = 16 (row stride 17): the nx staged rows of G / E sit on nx consecutive 8-byte bank pairs; place both
    // buffers so that the right-hand-side row (dv + t*n), moved by the same instructions, lands on the next one
    constexpr bool kAlign = (NN % 16 == 0) && (NX % 4 == 0) && (NX < 16);
    if (kAlign) while ((o + NX) % 16 != dv % 16) o += 4;
    Gt = take(NX * (NN | 1));
    if (kAlign) while (o % 16 != Gt % 16) o += 4;
    Ep = take(NX * (NN | 1));
    scr = take(64 + NN * (NN | 1));
    total = o;
  }
};

// per-problem shared-memory layout of k_line_search, in units of S
template <int NX, int NU>
struct LsLayout {
  static constexpr int NN = NX + NU;
  static constexpr int LS = NN | 1;  // odd stage stride: lanes working on different stages hit different banks
  int N, neq, nin, m, SI;
  int xu, d, lam, Qd, q, x0, obs, part, total;
  __host__ __device__ LsLayout(int T, int nobs = 0, bool se = false) {
    SI = se ? 0 : 2 * NU + nobs;
    N = T * NN; neq = se ? (T - 1) * NX : T * NX; nin = T * SI; m = neq + nin;
    int o = 0;
    auto take = [&](int cnt) { int at = o; o += (cnt + 3) & ~3; return at; };  // regions 16-byte aligned for float and double
    xu = take(T * LS);
    d = take(T * LS);
    lam = take(m);
    Qd = take(T * LS);
    q = take(T * LS);
    x0 = take(NX);
    obs = take(T * nobs * 3);
    part = take(NLS * (T - 1) * 3);
    total = o;
  }
};

// ------------------------------------------------------------------------------------------------
// right-looking elimination on "one row per lane".
//   lanes 0..NN-1 : rows of the SPD block (lower part in a[0..lane])         -> rows of L (diag = 1/L_ii)
//   lanes >= NN   : extra rows b (rhs, sub-diagonal block rows)               -> b L^{-T}
// FUSED=true  : factor + solve in one pass.
// FUSED=false : lanes 0..NN-1 already hold L (read-only); only lanes >= NN are transformed.
// ------------------------------------------------------------------------------------------------
// The pivot column travels through shared memory (one STS per lane, then 16-byte broadcast loads) instead
// of one shuffle pair per entry: the first version was bound by the SM's shuffle throughput.
//   colbuf : FUSED: 2 x 32 scalars (double-buffered pivot column).  !FUSED: the L rows, [NN][LDL].
template <int NN, bool FUSED, typename S>
__device__ __forceinline__ bool eliminate(S (&a)[NN], int lane, S* colbuf, int LDL = 0) {
  bool ok = true;
  if (FUSED) {
    // software-pipelined pivots: as soon as column k+1 has its update, the next pivot is broadcast and its
    // reciprocal square root (the longest dependent chain of the step) runs under the remaining updates
    S d = __shfl_sync(FULL, a[0], 0);
    ok = (d > S(0));
    S inv = rsqrt_(d);
#pragma unroll
    for (int k = 0; k < NN; ++k) {
      const S lk = a[k] * inv;
      a[k] = (lane == k) ? inv : lk;
      if (k + 1 < NN) {
        S* col = colbuf + (k & 1) * 32;
        if (lane < NN) col[lane] = lk;  // only rows k+1..NN-1 of the pivot column are read
        __syncwarp();
        a[k + 1] -= lk * col[k + 1];
        d = __shfl_sync(FULL, a[k + 1], k + 1);
        ok = ok && (d > S(0));
        inv = rsqrt_(d);
        // rest of the pivot column: one scalar head (odd start), 16-byte pairs, one scalar tail (odd NN)
        constexpr bool kVec = !ALMPC_NO_VEC2;
        const int j0 = k + 2;
        const int js = kVec ? ((j0 + 1) & ~1) : NN;  // first even index >= j0
        if (kVec && (j0 & 1) && j0 < NN) a[j0] -= lk * col[j0];
#pragma unroll
        for (int j = 0; j < NN; j += 2) {
          if (kVec && j >= js && j + 1 < NN) {
            const Vec2<S> v = *reinterpret_cast<const Vec2<S>*>(col + j);
            a[j] -= lk * v.x;
            a[j + 1] -= lk * v.y;
          }
        }
        if (kVec && (NN & 1) && (NN - 1) >= js) a[NN - 1] -= lk * col[NN - 1];
        if (!kVec) {
#pragma unroll
          for (int j = 0; j < NN; ++j)
            if (j >= j0) a[j] -= lk * col[j];
        }
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < NN; ++k) {
      const S inv = colbuf[k * LDL + k];
      const S lk = a[k] * inv;
      if (lane >= NN) a[k] = lk;
#pragma unroll
      for (int j = k + 1; j < NN; ++j) {
        if (lane >= NN) a[j] -= lk * colbuf[j * LDL + k];
      }
    }
  }
  return ok;
}

// factor storage per problem: T blocks [NN][NN] then (T-1) blocks [NX][NN]
template <int NX, int NU>
__host__ __device__ inline size_t factor_elems(int T) {
  constexpr int NN = NX + NU;
  return (size_t)T * NN * NN + (size_t)(T - 1) * NX * NN;
}

// D(8x8) += A(8x4) B(4x8), FP64 tensor-core MMA (legacy mma.sync path; on B200 it runs at the DFMA rate
// but replaces 8 DFMA + their shared-memory operand loads per lane by one instruction).
// Fragment layout (PTX ISA, m8n8k4 .f64): g = lane>>2, q = lane&3:  a = A[g][q], b = B[q][g],
// c0 = C[g][2q], c1 = C[g][2q+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int NN>
struct ScrLd {
  static constexpr int LD = NN | 1;  // odd row stride: one row per lane is bank-conflict free
};

// D_t = diag(Qd_t) + rho (G_t^T G_t + I_x + active bounds) - E_{t-1} E_{t-1}^T   (lower tiles) -> scr
template <int NX, int NU, typename S>
__device__ __forceinline__ void build_block(S* scr, const S* xu_t, const S* Gt, const S* Ep, bool has_G, bool has_E,
                                            int lane, S rho, const S* __restrict__ Qd_t, const S* ulo, const S* uhi,
                                            bool add_ix = true, bool bounds = true) {
  constexpr int NN = NX + NU;
  constexpr int LD = ScrLd<NN>::LD;
  constexpr int LB = NN | 1;  // row stride of the staged G / E blocks
  if constexpr (sizeof(S) == 8) {
    constexpr int NT = (NN + 7) / 8, KG = (NX + 3) / 4, NTX = (NX + 7) / 8, KE = (NN + 3) / 4;
    const int g = lane >> 2, q4 = lane & 3;
    // diagonal cost entries this lane will add in the epilogue: issued now, their latency hides under the MMAs
    double qd[NT];
#pragma unroll
    for (int rt = 0; rt < NT; ++rt) qd[rt] = (rt * 8 + g < NN) ? (double)Qd_t[rt * 8 + g] : 0.0;
    double acc[NT][NT][2];
#pragma unroll
    for (int rt = 0; rt < NT; ++rt)
#pragma unroll
      for (int ct = 0; ct < NT; ++ct) acc[rt][ct][0] = acc[rt][ct][1] = 0.0;
    if (has_G) {
      double f[NT][KG];
#pragma unroll
      for (int tl = 0; tl < NT; ++tl)
#pragma unroll
        for (int ch = 0; ch < KG; ++ch) {
          // the contraction index may be visited in any order as long as both operands agree: row = KG*q4 + ch
          // (instead of 4*ch + q4) halves the bank conflicts of these fragment loads for the odd stride LB
          const int col = tl * 8 + g, row = KG * q4 + ch;
          f[tl][ch] = (col < NN && row < NX) ? (double)Gt[row * LB + col] : 0.0;
        }
#pragma unroll
      for (int rt = 0; rt < NT; ++rt)
#pragma unroll
        for (int ct = 0; ct <= rt; ++ct)
#pragma unroll
          for (int ch = 0; ch < KG; ++ch) dmma884(acc[rt][ct][0], acc[rt][ct][1], f[rt][ch], f[ct][ch]);
    }
    double sy[NTX][NTX][2];
#pragma unroll
    for (int rt = 0; rt < NTX; ++rt)
#pragma unroll
      for (int ct = 0; ct < NTX; ++ct) sy[rt][ct][0] = sy[rt][ct][1] = 0.0;
    if (has_E) {
      double f[NTX][KE];
#pragma unroll
      for (int tl = 0; tl < NTX; ++tl)
#pragma unroll
        for (int ch = 0; ch < KE; ++ch) {
          const int row = tl * 8 + g, k = KE * q4 + ch;  // permuted contraction index: conflict-free for LB = 17
          f[tl][ch] = (row < NX && k < NN) ? (double)Ep[row * LB + k] : 0.0;
        }
#pragma unroll
      for (int rt = 0; rt < NTX; ++rt)
#pragma unroll
        for (int ct = 0; ct <= rt; ++ct)
#pragma unroll
          for (int ch = 0; ch < KE; ++ch) dmma884(sy[rt][ct][0], sy[rt][ct][1], f[rt][ch], f[ct][ch]);
    }
#pragma unroll
    for (int rt = 0; rt < NT; ++rt) {
      const int r = rt * 8 + g;
#pragma unroll
      for (int ct = 0; ct <= rt; ++ct) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = ct * 8 + 2 * q4 + e;
          double v = acc[rt][ct][e];
          if (r == c) {
            if (r < NX) {
              if (add_ix) v += 1.0;
            } else if (r < NN && bounds) {
              const double u = xu_t[r];
              v += ((u - uhi[r - NX]) >= 0.0 ? 1.0 : 0.0) + ((ulo[r - NX] - u) >= 0.0 ? 1.0 : 0.0);
            }
          }
          v *= rho;
          if (r == c && r < NN) v += qd[rt];
          if (rt < NTX && ct < NTX) v -= sy[rt][ct][e];
          if (r < NN && c < NN) scr[r * LD + c] = v;
        }
      }
    }
  } else {
    constexpr int NTRI = NN * (NN + 1) / 2;
    for (int idx = lane; idx < NTRI; idx += 32) {
      int r = (int)((sqrtf(8.0f * idx + 1.0f) - 1.0f) * 0.5f);
      while (r * (r + 1) / 2 > idx) --r;
      while ((r + 1) * (r + 2) / 2 <= idx) ++r;
      const int c = idx - r * (r + 1) / 2;
      S v0 = S(0), v1 = S(0);
      if (has_G) {
#pragma unroll
        for (int i = 0; i + 1 < NX; i += 2) {
          v0 += Gt[i * LB + r] * Gt[i * LB + c];
          v1 += Gt[(i + 1) * LB + r] * Gt[(i + 1) * LB + c];
        }
        if (NX & 1) v0 += Gt[(NX - 1) * LB + r] * Gt[(NX - 1) * LB + c];
      }
      S v = v0 + v1;
      if (r == c) {
        if (r < NX) {
          if (add_ix) v += S(1);
        } else if (bounds) {
          const S u = xu_t[r];
          v += ((u - uhi[r - NX]) >= S(0) ? S(1) : S(0)) + ((ulo[r - NX] - u) >= S(0) ? S(1) : S(0));
        }
      }
      v *= rho;
      if (r == c) v += Qd_t[r];
      if (has_E && r < NX) {
        S s0 = S(0), s1 = S(0);
#pragma unroll
        for (int k = 0; k + 1 < NN; k += 2) {
          s0 += Ep[r * LB + k] * Ep[c * LB + k];
          s1 += Ep[r * LB + k + 1] * Ep[c * LB + k + 1];
        }
        if (NN & 1) s0 += Ep[r * LB + NN - 1] * Ep[c * LB + NN - 1];
        v -= s0 + s1;
      }
      scr[r * LD + c] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// kernel 1/3: dynamics + Jacobian columns.  1 thread = 1 (problem, stage t, column j of G_t = [A_t B_t]).
//   G_g[prob][t][i][j] = d f_i / d z_j at (x_t,u_t) ; c_g[prob][t*nx+i] = x_{t+1,i} - f_i ; last block: x_0 - x0
// ------------------------------------------------------------------------------------------------
template <class M, typename S, int V>
__global__ void __launch_bounds__(128)
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
// kernel 2/3: residual statistics, dual update, merit, gradient, block-tridiagonal factor + solve.
// 1 warp = 1 problem.
// ------------------------------------------------------------------------------------------------
template <int NX, int NU, typename S, int WPB, int V>
__global__ void __launch_bounds__(WPB * 32, ALMPC_FAC_MINB)
k_factor_solve(const KParams<S> P, const S* __restrict__ Qd_g, const S* __restrict__ q_g, const S* __restrict__ xu_g,
               S* __restrict__ lam_g, S* __restrict__ rho_g, S* fac_g, const S* __restrict__ G_g,
               const S* __restrict__ c_g, S* __restrict__ d_g, S* __restrict__ merit_g, double* __restrict__ rentry_g,
               int* __restrict__ status_g, S* __restrict__ debug_g, const Ctl* __restrict__ ctl) {
  constexpr int NN = NX + NU;
  constexpr int LD = ScrLd<NN>::LD;
  constexpr int EX0 = 32 - NN - 1;  // sub-diagonal rows that fit next to the L rows and the rhs row
  constexpr int CW = CopyWidth<NX, NU, S>::W;
  constexpr int LB = NN | 1;  // odd row stride of the staged G / E blocks (bank-conflict free rows)
  constexpr bool OBS = (V == 1), SE = (V == 2);  // variant: 0 plain, 1 obstacle rows, 2 state estimator
  extern __shared__ __align__(16) unsigned char smem_raw[];
  grid_dep_wait();
  const int mode = ctl->mode;
  if (mode == MODE_DONE) return;
  const int T = P.T;
  const int n_obs = OBS ? P.n_obs : 0;
  const Layout<NX, NU> L(T, n_obs, SE);
  const int SI = OBS ? L.SI : 2 * NU;  // compile-time constant without obstacle rows
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int prob = blockIdx.x * WPB + warp;
  if (prob >= P.B) return;
  S* sm = reinterpret_cast<S*>(smem_raw) + (size_t)warp * L.total;
  S* xu = sm + L.xu;
  S* dv = sm + L.dv;
  S* wv = sm + L.w;
  S* Gt = sm + L.Gt;
  S* Ep = sm + L.Ep;
  S* scr = sm + L.scr;
  const size_t pN = (size_t)prob * L.N;
  const S* Qd = Qd_g + pN;
  const S* qv = q_g + pN;
  S* lam = lam_g + (size_t)prob * L.m;
  const S* cres = c_g + (size_t)prob * L.neq;
  const S* Gp = G_g + (size_t)prob * (T - 1) * NX * NN;
  S* fac = fac_g + (size_t)prob * factor_elems<NX, NU>(T);
  S* dbg = debug_g ? debug_g + (size_t)prob * (2 * L.N + 24) : nullptr;
  const S* obs = OBS ? P.obs + (size_t)prob * T * n_obs * 3 : nullptr;

  if (T > 1) warp_copy_block_async<NX, NN, LB>(Gt, Gp, lane);  // G_0, consumed after the residual statistics
  cp_async_commit();
  S rho = rho_g[prob];
  for (int base = 0; base < L.N; base += 128) {
    S xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + lane + 32 * u;
      xv[u] = (i < L.N) ? xu_g[pN + i] : S(0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + lane + 32 * u;
      if (i < L.N) xu[i] = xv[u];
    }
  }
  __syncwarp();

  // ---- dual / penalty update of the previous AL iteration (AL_mpc.py:316-317,325), entry residual
  // ||c+||^2, merit at the entry point, w = lam_eq + rho c, cost gradient.  Global loads are issued four
  // at a time per lane so that their latencies overlap.
  const bool dual = (mode == MODE_DUAL_NEWTON);
  const S rho_old = rho;
  if (dual) rho = next_rho(rho, P.rho_mult, P.rho_f32);
  {
    S s2 = S(0), sl = S(0), cost = S(0);
    for (int base = 0; base < L.neq; base += 128) {
      S cv[4], lv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + lane + 32 * u;
        cv[u] = (i < L.neq) ? cres[i] : S(0);
        lv[u] = (i < L.neq) ? lam[i] : S(0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + lane + 32 * u;
        if (i < L.neq) {
          S li = lv[u];
          if (dual) { li += rho_old * cv[u]; lam[i] = li; }
          s2 += cv[u] * cv[u];
          sl += li * cv[u];
          wv[i] = li + rho * cv[u];
        }
      }
    }
    for (int base = 0; base < L.nin; base += 128) {
      S lv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + lane + 32 * u;
        lv[u] = (i < L.nin) ? lam[L.neq + i] : S(0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + lane + 32 * u;
        if (i < L.nin) {
          const int t = i / SI, sidx = i - t * SI;
          const S ci = ineq_res<NX, NU, S>(xu + t * NN, sidx, P.ulo, P.uhi, OBS ? obs + (size_t)t * n_obs * 3 : nullptr, P.obs_r2);
          S li = lv[u];
          if (dual) { li = fmax0(li + rho_old * ci); lam[L.neq + i] = li; }
          const S cp = fmax0(ci);
          s2 += cp * cp;
          sl += li * ci;
        }
      }
    }
    for (int base = 0; base < L.N; base += 128) {
      S Qv[4], qq[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + lane + 32 * u;
        Qv[u] = (i < L.N) ? Qd[i] : S(0);
        qq[u] = (i < L.N) ? qv[i] : S(0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + lane + 32 * u;
        if (i < L.N) {
          const S z = xu[i];
          // state estimator: cost and cost gradient see the states only (al_utils_se.py:272-305)
          const bool in_cost = !SE || (i % NN) < NX;
          if (in_cost) cost += S(0.5) * (z * Qv[u] * z) + qq[u] * z;
          dv[i] = in_cost ? Qv[u] * z + qq[u] : S(0);  // cost gradient (al_utils.py:438-446), completed per stage below
        }
      }
    }
    s2 = warp_sum(s2);
    sl = warp_sum(sl);
    cost = warp_sum(cost);
    if (lane == 0) {
      if (dual) rho_g[prob] = rho;
      rentry_g[prob] = (double)s2;
      if (mode != MODE_CONT) merit_g[prob] = cost + S(0.5) * rho * s2 + sl;
    }
  }
  __syncwarp();

  // ---- forward sweep over the stages: gradient, Hessian block, fused elimination.
  // Two block buffers ping-pong: X holds G_t (prefetched) and is overwritten by E_t; Y holds E_{t-1} and,
  // once its last reader is done, receives the prefetch of G_{t+1}.
  bool ok = true;
  S* X = Gt;
  S* Y = Ep;
  for (int t = 0; t < T; ++t) {
    S* Lg = fac + (size_t)t * NN * NN;
    S* Eg = fac + (size_t)T * NN * NN + (size_t)t * NX * NN;
    const bool has_G = t < T - 1, has_E = t > 0;
    cp_async_wait_all();  // G_t (issued one stage ago / in the prologue)
    __syncwarp();
    // obstacle rows of stage t: (lam_k + rho c_k+) and the active-masked Jacobian rows, staged in scr[0..]
    if constexpr (OBS) {
      if (lane < n_obs) {
        S J[3];
        const S ck = obstacle_res<S>(xu + t * NN, obs + ((size_t)t * n_obs + lane) * 3, P.obs_r2, J);
        const S lk = lam[L.neq + t * SI + 2 * NU + lane];
        scr[lane * 8 + 0] = J[0];
        scr[lane * 8 + 1] = J[1];
        scr[lane * 8 + 2] = J[2];
        scr[lane * 8 + 3] = lk + rho * fmax0(ck);        // multiplies J in the gradient (al_utils.py:113-117)
        scr[lane * 8 + 4] = (ck >= S(0)) ? rho : S(0);   // multiplies J^T J in the Hessian (active rows only)
      }
      __syncwarp();
    }
    // gradient of stage t (al_utils.py:113-117) minus the coupling E_{t-1} y_{t-1}
    if (lane < NN) {
      const int j = lane, idx = t * NN + j;
      S g = dv[idx];
      if (has_G) {
        S s0 = S(0), s1 = S(0);
#pragma unroll
        for (int i = 0; i + 1 < NX; i += 2) {
          s0 += X[i * LB + j] * wv[t * NX + i];
          s1 += X[(i + 1) * LB + j] * wv[t * NX + i + 1];
        }
        if (NX & 1) s0 += X[(NX - 1) * LB + j] * wv[t * NX + NX - 1];
        g -= s0 + s1;
      }
      if (j < NX) {
        if (has_E) g += wv[(t - 1) * NX + j];
        else if (!SE) g += wv[(T - 1) * NX + j];  // the x_0 - x0 block
      } else if (!SE) {
        const int sidx = j - NX;
        const S u = xu[idx];
        const S cu = u - P.uhi[sidx], cl = P.ulo[sidx] - u;
        const S lu = lam[L.neq + t * SI + sidx], ll = lam[L.neq + t * SI + NU + sidx];
        g += (lu - ll) + rho * (fmax0(cu) - fmax0(cl));
      }
      if constexpr (OBS) {
        if (j < 3)
          for (int k = 0; k < n_obs; ++k) g += scr[k * 8 + 3] * scr[k * 8 + j];
      }
      if (dbg) dbg[idx] = g;
      if (has_E && j < NX) {
        S s0 = S(0), s1 = S(0);
#pragma unroll
        for (int k = 0; k + 1 < NN; k += 2) {
          s0 += Y[j * LB + k] * dv[(t - 1) * NN + k];
          s1 += Y[j * LB + k + 1] * dv[(t - 1) * NN + k + 1];
        }
        if (NN & 1) s0 += Y[j * LB + NN - 1] * dv[(t - 1) * NN + NN - 1];
        g -= s0 + s1;
      }
      dv[idx] = g;
    }
    build_block<NX, NU, S>(scr + 64, xu + t * NN, X, Y, has_G, has_E, lane, rho, Qd + t * NN, P.ulo, P.uhi,
                           !SE || has_E, !SE);
    __syncwarp();
    if constexpr (OBS) {  // + rho sum_k active J_k^T J_k on the position block (lower part)
      if (lane < 6) {
        const int r = (lane < 1) ? 0 : (lane < 3 ? 1 : 2), c = lane - (r * (r + 1)) / 2;
        S acc = S(0);
        for (int k = 0; k < n_obs; ++k) acc += scr[k * 8 + 4] * scr[k * 8 + r] * scr[k * 8 + c];
        scr[64 + r * LD + c] += acc;
      }
      __syncwarp();
    }
    // rows into registers: L rows | rhs | -rho G_t rows.  One uniform load loop with a per-lane row pointer
    // (three divergent loops cost three times the LSU instructions).
    S a[NN];
    {
      const int ei = lane - NN - 1;
      const bool e_act = (ei >= 0) && (ei < NX) && has_G;
      const S* rowp = (lane < NN) ? (scr + 64 + lane * LD) : ((lane == NN) ? (dv + t * NN) : (X + (e_act ? ei : 0) * LB));
      const S scale = (lane <= NN) ? S(1) : (e_act ? -rho : S(0));
      // entries above the diagonal of an L row are never read by another lane (the pivot column is consumed
      // from row k+1 on and the write-out masks them), so they are loaded as they are
#pragma unroll
      for (int c = 0; c < NN; ++c) a[c] = rowp[c] * scale;
    }
    __syncwarp();  // Y (E_{t-1}) has no reader left: prefetch G_{t+1} into it behind the pivot chain
    if (t + 1 < T - 1) warp_copy_block_async<NX, NN, LB>(Y, Gp + (size_t)(t + 1) * NX * NN, lane);
    cp_async_commit();
    ok = eliminate<NN, true>(a, lane, scr) && ok;
    __syncwarp();
    {
      // write-outs with per-lane destinations: L rows -> global factor; rhs -> dv (y_t); E rows -> X + global factor
      const int ei = lane - NN - 1;
      const bool e_wr = (ei >= 0) && (ei < NX) && has_G && (NX <= EX0);
      const int last = (lane < NN) ? lane : NN - 1;
      S* gdst = (lane < NN) ? (Lg + lane * NN) : (e_wr ? (Eg + ei * NN) : nullptr);
      S* sdst = (lane == NN) ? (dv + t * NN) : (e_wr ? (X + ei * LB) : nullptr);
      if (gdst) {
        if constexpr (NN % 2 == 0) {  // rows are 16-byte aligned: 128-bit stores
#pragma unroll
          for (int c = 0; c < NN; c += 2) {
            Vec2<S> v;
            v.x = (c <= last) ? a[c] : S(0);
            v.y = (c + 1 <= last) ? a[c + 1] : S(0);
            *reinterpret_cast<Vec2<S>*>(gdst + c) = v;
          }
        } else {
#pragma unroll
          for (int c = 0; c < NN; ++c) gdst[c] = (c <= last) ? a[c] : S(0);
        }
      }
      if (sdst) {
#pragma unroll
        for (int c = 0; c < NN; ++c) sdst[c] = a[c];
      }
      if (lane < NN && NX > EX0 && has_G) {  // publish L for the second elimination pass
#pragma unroll
        for (int c = 0; c < NN; ++c) scr[64 + lane * LD + c] = a[c];
      }
    }
    if (NX > EX0 && has_G) {  // sub-diagonal rows that did not fit (flying cartpole: nx=14, n=18)
      // rows of the first pass are parked in registers until every row of -rho G_t has been read from X
      S keep[NN];
#pragma unroll
      for (int c = 0; c < NN; ++c) keep[c] = a[c];
      const int i0 = lane - NN - 1;
      constexpr int PER = 32 - NN;
      for (int base = EX0; base < NX; base += PER) {
        const int i = base + (lane - NN);
        const bool act = (lane >= NN) && (i < NX);
        if (lane >= NN) {
#pragma unroll
          for (int c = 0; c < NN; ++c) a[c] = act ? -rho * X[i * LB + c] : S(0);
        }
        __syncwarp();
        eliminate<NN, false>(a, lane, scr + 64, LD);
        __syncwarp();
        if (act) {
#pragma unroll
          for (int c = 0; c < NN; ++c) {
            X[i * LB + c] = a[c];
            Eg[i * NN + c] = a[c];
          }
        }
      }
      if (lane > NN && i0 < NX) {
#pragma unroll
        for (int c = 0; c < NN; ++c) {
          X[i0 * LB + c] = keep[c];
          Eg[i0 * NN + c] = keep[c];
        }
      }
    }
    __syncwarp();
    S* tmp = X; X = Y; Y = tmp;  // Y := E_t, X := G_{t+1}
  }
  ok = __all_sync(FULL, ok);
  // ---- backward sweep, d = -H^{-1} g.  L_tt / E_t come back from the streamed factor (L2) through
  // cp.async: the blocks of stage t-1 are in flight while the substitution of stage t runs.
  cp_async_wait_all();
  __syncwarp();
  S* Lsm = scr + 64;                    // [NN][NN] row-major copy of L_tt
  S* Ecur = X;                          // E_t
  S* Enext = Y;
  warp_copy_async<NN * NN, CW>(Lsm, fac + (size_t)(T - 1) * NN * NN, lane);
  cp_async_commit();
  for (int t = T - 1; t >= 0; --t) {
    cp_async_wait_all();
    __syncwarp();
    S l[NN];
    if (lane < NN) {
#pragma unroll
      for (int r = 0; r < NN; ++r) l[r] = Lsm[r * NN + lane];  // stored factor: zeros above the diagonal
    }
    S b = (lane < NN) ? dv[t * NN + lane] : S(0);
    if (t < T - 1 && lane < NN) {
      S s0 = S(0), s1 = S(0);
#pragma unroll
      for (int i = 0; i + 1 < NX; i += 2) {
        s0 += Ecur[i * LB + lane] * dv[(t + 1) * NN + i];
        s1 += Ecur[(i + 1) * LB + lane] * dv[(t + 1) * NN + i + 1];
      }
      if (NX & 1) s0 += Ecur[(NX - 1) * LB + lane] * dv[(t + 1) * NN + NX - 1];
      b -= s0 + s1;
    }
    __syncwarp();
    if (t > 0) {
      warp_copy_async<NN * NN, CW>(Lsm, fac + (size_t)(t - 1) * NN * NN, lane);
      warp_copy_block_async<NX, NN, LB>(Enext, fac + (size_t)T * NN * NN + (size_t)(t - 1) * NX * NN, lane);
    }
    cp_async_commit();
#pragma unroll
    for (int r = NN - 1; r >= 0; --r) {
      const S dr = __shfl_sync(FULL, b * l[r], r);  // lane r: l[r] = 1/L_rr
      b = (lane == r) ? dr : (b - l[r] * dr);       // l[r] = 0 on lanes > r
    }
    if (lane < NN) dv[t * NN + lane] = b;
    S* tmp = Ecur; Ecur = Enext; Enext = tmp;
  }
  __syncwarp();
  // A non-positive pivot (never seen in fp64; routine in fp32 once rho/R exceeds ~1e10) leaves no
  // usable direction: the step is skipped (d = 0 keeps the iterate) and the problem is flagged, instead of the
  // backend-specific garbage direction the reference would push through its line search.
  for (int i = lane; i < L.N; i += 32) {
    const S d = ok ? -dv[i] : S(0);
    d_g[pN + i] = d;
    if (dbg) dbg[L.N + i] = d;
  }
  if (lane == 0) {
    int status = (mode == MODE_FIRST) ? (ALMPC_STATUS_FINITE | ALMPC_STATUS_CHOL_OK) : status_g[prob];
    if (!ok) status &= ~ALMPC_STATUS_CHOL_OK;
    status_g[prob] = status;
  }
}

// ------------------------------------------------------------------------------------------------
// kernel 3/3: 20-point line search + iterate update + batch-global stop statistics / control block.
// 1 thread = 1 (problem, candidate k, stage t); `pls` problems per CTA.
// ------------------------------------------------------------------------------------------------
// Executed by the LAST CTA of the line-search launch: deterministic reduction of the per-CTA statistics and the
// batch-global decisions of the reference (stop rule al_utils.py:560-564, rho_max exit AL_mpc.py:412).
template <typename S>
__device__ __forceinline__ void steer_solve(const KParams<S>& P, Ctl* ctl, int* info_g, const double* partials,
                                            double (*fin_red)[3], int tid, int nthr, int al_a, int nk) {
  {
    double s0 = 0, s1 = 0, s2 = 0;
    const volatile double* pv = partials;
    for (int b = tid; b < P.ticket_total; b += nthr) {
      s0 += pv[b * 4 + 0];
      s1 += pv[b * 4 + 1];
      s2 = fmax(s2, pv[b * 4 + 2]);
    }
    fin_red[tid][0] = s0;
    fin_red[tid][1] = s1;
    fin_red[tid][2] = s2;
  }
  __syncthreads();
  if (tid == 0) {
    double s0 = 0, s1 = 0, s2 = 0;
    for (int i = 0; i < nthr; ++i) { s0 += fin_red[i][0]; s1 += fin_red[i][1]; s2 = fmax(s2, fin_red[i][2]); }
    const double R_entry = sqrt(s0), R_exit = sqrt(s1), rho_max_used = s2;
    const double res_old = (nk == 0) ? R_entry : ctl->res_old;
    // al_utils.py:560-564 — evaluated exactly as written there (NaN compares false -> no break)
    bool stop = (fabs(res_old - R_exit) / R_exit < 1e-3) || (R_exit < 1e-3);
    const int steps_done = nk + 1;
    if (steps_done >= P.max_newton) stop = true;
    ctl->ticket = 0;
    ctl->total_newton += 1;
    info_g[ALMPC_INFO_TOTAL_NEWTON] = ctl->total_newton;
    if (!stop) {
      ctl->k = steps_done;
      ctl->mode = MODE_CONT;
      ctl->res_old = R_exit;
    } else {
      info_g[ALMPC_INFO_NEWTON0 + al_a] = steps_done;
      info_g[ALMPC_INFO_AL_DONE] = al_a + 1;
      double rho_next = rho_max_used * P.rho_mult;
      if (P.rho_f32) rho_next = (double)(float)rho_next;
      ctl->rho_final_max = rho_next;
      const bool exceed = P.stream_mode && (rho_next > P.rho_max);
      if (al_a + 1 >= P.al_iter || exceed) {
        ctl->mode = MODE_DONE;
      } else {
        ctl->a = al_a + 1;
        ctl->k = 0;
        ctl->mode = MODE_DUAL_NEWTON;
      }
    }
  }
}

// One line-search item: candidate alpha = 2^-kk, stage t.  Evaluates the stage's share of the merit
// (al_utils.py:52-77): cost, sum c+^2 and lam.c of the dynamics residual x_{t+1} - f(x_t,u_t), the control
// bounds of stage t (and of the last stage for t = T-2) and the initial-state residual (t = 0).
template <class M, typename S, int V>
__device__ __forceinline__ void ls_item(const S* sp, const LsLayout<M::NX, M::NU>& L, const KParams<S>& P, int T,
                                        int kk, int t, S& cost, S& s2, S& sl) {
  constexpr int NX = M::NX, NU = M::NU, NN = NX + NU, LS = NN | 1;
  const S* xu = sp + L.xu;
  const S* dv = sp + L.d;
  const S* lam = sp + L.lam;
  const S* Qd = sp + L.Qd;
  const S* qv = sp + L.q;
  const S* x0 = sp + L.x0;
  constexpr bool OBS = (V == 1), SE = (V == 2);
  const int n_obs = OBS ? P.n_obs : 0;
  const int SI = OBS ? L.SI : 2 * NU;
    const S alpha = (S)ldexpf(1.0f, -kk);
    S xs[NX], us[NU], xn[NX];
    cost = S(0); s2 = S(0); sl = S(0);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      const S z = xu[t * LS + i] + alpha * dv[t * LS + i];
      xs[i] = z;
      cost += S(0.5) * (z * Qd[t * LS + i] * z) + qv[t * LS + i] * z;
    }
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      const S z = xu[t * LS + NX + i] + alpha * dv[t * LS + NX + i];
      us[i] = z;
      if (!SE) {
        cost += S(0.5) * (z * Qd[t * LS + NX + i] * z) + qv[t * LS + NX + i] * z;
        const S cu = z - P.uhi[i], cl = P.ulo[i] - z;
        sl += lam[L.neq + t * SI + i] * cu + lam[L.neq + t * SI + NU + i] * cl;
        const S pu = fmax0(cu), pl = fmax0(cl);
        s2 += pu * pu + pl * pl;
      }
    }
    if (!SE && t == 0) {
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        const S ci = xs[i] - x0[i];
        s2 += ci * ci;
        sl += lam[(T - 1) * NX + i] * ci;
      }
    }
    if constexpr (OBS) {  // obstacle rows of stage t at the candidate position
      const S* ob = sp + L.obs + t * n_obs * 3;
      for (int k = 0; k < n_obs; ++k) {
        const S ck = obstacle_res<S>(xs, ob + k * 3, P.obs_r2, nullptr);
        sl += lam[L.neq + t * SI + 2 * NU + k] * ck;
        const S cp = fmax0(ck);
        s2 += cp * cp;
      }
    }
    S zlast[3];
    rk4_step<M, S, S>(xs, us, xn, P.dyn);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      const S z1 = xu[(t + 1) * LS + i] + alpha * dv[(t + 1) * LS + i];
      const S ci = z1 - xn[i];
      s2 += ci * ci;
      sl += lam[t * NX + i] * ci;
      if (OBS && i < 3) zlast[i] = z1;
      if (t == T - 2) cost += S(0.5) * (z1 * Qd[(t + 1) * LS + i] * z1) + qv[(t + 1) * LS + i] * z1;
    }
    if (!SE && t == T - 2) {
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        const int e = (T - 1) * LS + NX + i;
        const S z = xu[e] + alpha * dv[e];
        cost += S(0.5) * (z * Qd[e] * z) + qv[e] * z;
        const S cu = z - P.uhi[i], cl = P.ulo[i] - z;
        sl += lam[L.neq + (T - 1) * SI + i] * cu + lam[L.neq + (T - 1) * SI + NU + i] * cl;
        const S pu = fmax0(cu), pl = fmax0(cl);
        s2 += pu * pu + pl * pl;
      }
      if constexpr (OBS) {  // obstacle rows of the last stage
        const S* ob = sp + L.obs + (T - 1) * n_obs * 3;
        for (int k = 0; k < n_obs; ++k) {
          const S ck = obstacle_res<S>(zlast, ob + k * 3, P.obs_r2, nullptr);
          sl += lam[L.neq + (T - 1) * SI + 2 * NU + k] * ck;
          const S cp = fmax0(ck);
          s2 += cp * cp;
        }
      }
    }
}

// asynchronous staging of one problem's line-search inputs into shared memory (all threads of the CTA)
template <int NX, int NU, typename S>
__device__ __forceinline__ void ls_stage_async(S* sp, const LsLayout<NX, NU>& L, int prob, int tid, int nthr,
                                               const S* x0_g, const S* Qd_g, const S* q_g, const S* xu_g,
                                               const S* lam_g, const S* d_g, const S* obs_g = nullptr) {
  const size_t pN = (size_t)prob * L.N;
  auto copy = [&](S* dst, const S* src, int count) {
    if ((((size_t)count * sizeof(S)) & 15) == 0 && ((reinterpret_cast<size_t>(src) & 15) == 0)) {
      const int bytes = count * (int)sizeof(S);
      for (int o = tid * 16; o < bytes; o += nthr * 16) cp_async16((char*)dst + o, (const char*)src + o);
    } else {
      for (int e = tid; e < count; e += nthr) {
        if constexpr (sizeof(S) == 8) cp_async8(dst + e, src + e);
        else cp_async4(dst + e, src + e);
      }
    }
  };
  constexpr int NN = NX + NU, LS = NN | 1;
  for (int e = tid; e < L.N; e += nthr) {
    const int t = e / NN, c = e - t * NN;
    const int o = t * LS + c;
    if constexpr (sizeof(S) == 8) {
      cp_async8(sp + L.xu + o, xu_g + pN + e);
      cp_async8(sp + L.d + o, d_g + pN + e);
      cp_async8(sp + L.Qd + o, Qd_g + pN + e);
      cp_async8(sp + L.q + o, q_g + pN + e);
    } else {
      cp_async4(sp + L.xu + o, xu_g + pN + e);
      cp_async4(sp + L.d + o, d_g + pN + e);
      cp_async4(sp + L.Qd + o, Qd_g + pN + e);
      cp_async4(sp + L.q + o, q_g + pN + e);
    }
  }
  copy(sp + L.lam, lam_g + (size_t)prob * L.m, L.m);
  copy(sp + L.x0, x0_g + (size_t)prob * NX, NX);
  if (obs_g) copy(sp + L.obs, obs_g + (size_t)prob * (L.nin / L.SI) * (L.SI - 2 * NU) * 3, (L.nin / L.SI) * (L.SI - 2 * NU) * 3);
}

template <class M, typename S, int V>
__global__ void __launch_bounds__(256, 2)
k_line_search(const KParams<S> P, const int pls, const S* __restrict__ x0_g, const S* __restrict__ Qd_g,
              const S* __restrict__ q_g, S* xu_g, const S* __restrict__ lam_g,
              const S* __restrict__ rho_g, const S* __restrict__ d_g, S* __restrict__ merit_g,
              const double* __restrict__ rentry_g, int* __restrict__ status_g, int* __restrict__ info_g,
              S* __restrict__ debug_g, Ctl* ctl, double* __restrict__ partials) {
  constexpr int NX = M::NX, NU = M::NU, NN = NX + NU, LS = NN | 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[8][3];
  __shared__ bool is_last;
  grid_dep_wait();
  const int mode = ctl->mode;
  if (mode == MODE_DONE) return;
  const int al_a = ctl->a, nk = ctl->k;
  const int T = P.T;
  constexpr bool OBS = (V == 1), SE = (V == 2);
  const LsLayout<NX, NU> L(T, OBS ? P.n_obs : 0, SE);
  const int IT = NLS * (T - 1);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31;
  S* smem = reinterpret_cast<S*>(smem_raw);
  const int ngroups = (P.B + pls - 1) / pls;
  const size_t buf_elems = (size_t)pls * L.total;
  double blk0 = 0.0, blk1 = 0.0, blk2 = 0.0;  // thread 0: this CTA's running statistics

  // persistent CTA: groups of `pls` problems, double-buffered; the next group's inputs stream in (cp.async)
  // while the current group's 20 x (T-1) dynamics evaluations run
  int cur = 0;
  if ((int)blockIdx.x < ngroups) {
    for (int p = 0; p < pls; ++p) {
      const int prob = blockIdx.x * pls + p;
      if (prob < P.B) ls_stage_async<NX, NU, S>(smem + (size_t)p * L.total, L, prob, tid, nthr, x0_g, Qd_g, q_g, xu_g, lam_g, d_g, OBS ? P.obs : nullptr);
    }
  }
  cp_async_commit();
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int prob0 = grp * pls;
    const int nxt = grp + gridDim.x;
    if (nxt < ngroups) {
      S* nb = smem + (size_t)(cur ^ 1) * buf_elems;
      for (int p = 0; p < pls; ++p) {
        const int prob = nxt * pls + p;
        if (prob < P.B) ls_stage_async<NX, NU, S>(nb + (size_t)p * L.total, L, prob, tid, nthr, x0_g, Qd_g, q_g, xu_g, lam_g, d_g, OBS ? P.obs : nullptr);
      }
    }
    cp_async_commit();
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    S* base = smem + (size_t)cur * buf_elems;

  for (int wi = tid; wi < pls * IT; wi += nthr) {
    const int p = wi / IT, it = wi - p * IT;
    if (prob0 + p >= P.B) continue;
    const S* sp = base + (size_t)p * L.total;
    S* part = const_cast<S*>(sp) + L.part;
    const int kk = it / (T - 1), t = it - kk * (T - 1);
    S cost, s2, sl;
    ls_item<M, S, V>(sp, L, P, T, kk, t, cost, s2, sl);
    part[it * 3 + 0] = cost;
    part[it * 3 + 1] = s2;
    part[it * 3 + 2] = sl;
  }
  __syncthreads();

  // one warp per problem: candidate sums, first-argmin, accept, iterate update
  for (int p = warp; p < pls; p += (nthr >> 5)) {
    const int prob = prob0 + p;
    double r_entry = 0.0, r_exit = 0.0, rho_used = 0.0;
    if (prob < P.B) {
      const S* sp = base + (size_t)p * L.total;
      const S* xu = sp + L.xu;
      const S* dv = sp + L.d;
      const S* part = sp + L.part;
      const S rho = rho_g[prob];
      const S merit_cur = merit_g[prob];
      S* dbg = debug_g ? debug_g + (size_t)prob * (2 * L.N + 24) + 2 * L.N : nullptr;
      r_entry = rentry_g[prob];
      rho_used = (double)rho;
      S mk = S(0), r2k = S(0);
      if (lane < NLS) {
        S cost = S(0), s2 = S(0), sl = S(0);
        for (int t = 0; t < T - 1; ++t) {
          cost += part[(lane * (T - 1) + t) * 3 + 0];
          s2 += part[(lane * (T - 1) + t) * 3 + 1];
          sl += part[(lane * (T - 1) + t) * 3 + 2];
        }
        mk = cost + S(0.5) * rho * s2 + sl;
        r2k = s2;
        if (dbg) dbg[lane] = mk;
      }
      // first-argmin with torch.min NaN semantics (any NaN -> NaN, step rejected)
      const bool isn = (lane < NLS) && (mk != mk);
      const bool any_nan = __any_sync(FULL, isn);
      S bv = (lane < NLS) ? mk : S(INFINITY);
      int bi = (lane < NLS) ? lane : NLS;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const S ov = __shfl_xor_sync(FULL, bv, o);
        const int oi = __shfl_xor_sync(FULL, bi, o);
        if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (bi >= NLS) bi = 0;  // all +inf
      const S r2_best = __shfl_sync(FULL, r2k, bi);
      if (any_nan) bv = S(NAN);
      const bool accept = bv < merit_cur;
      const S alpha = (S)ldexpf(1.0f, -bi);
      bool fin = true;
      const size_t pN = (size_t)prob * L.N;
      for (int i = lane; i < L.N; i += 32) {
        const int ts = i / NN, si = ts * LS + (i - ts * NN);
        const S xo = xu[si];
        const S xnew = xo + alpha * dv[si];
        // status*x_next + (1-status)*x_est (al_utils.py:640-641): a non-finite rejected candidate poisons the element
        const S xr = accept ? xnew : (finite_(xnew) ? xo : S(NAN));
        fin = fin && finite_(xr);
        if (accept || !finite_(xnew)) xu_g[pN + i] = xr;
      }
      fin = __all_sync(FULL, fin);
      r_exit = accept ? (double)r2_best : (fin ? r_entry : (double)NAN);
      if (lane == 0) {
        merit_g[prob] = bv;  // merit <- new_merit even when the step was rejected (al_utils.py:569)
        if (!fin) status_g[prob] &= ~ALMPC_STATUS_FINITE;
        if (dbg) {
          dbg[NLS + 0] = (S)bi;
          dbg[NLS + 1] = accept ? S(1) : S(0);
          dbg[NLS + 2] = merit_cur;
        }
      }
    }
    if (lane == 0) {
      red[p][0] = r_entry;
      red[p][1] = r_exit;
      red[p][2] = rho_used;
    }
  }
  __syncthreads();
    // fold this group's statistics into the CTA totals in a fixed order
    if (tid == 0) {
      for (int p = 0; p < pls; ++p) { blk0 += red[p][0]; blk1 += red[p][1]; blk2 = fmax(blk2, red[p][2]); }
    }
    __syncthreads();  // buffer `cur` may be overwritten by the prefetch of the iteration after next
    cur ^= 1;
  }
  cp_async_wait_all();
  // ---- batch-global statistics: per-CTA partials, last CTA reduces in a fixed order and steers the solve
  if (tid == 0) {
    const int slot = P.part_off + blockIdx.x;
    partials[slot * 4 + 0] = blk0;
    partials[slot * 4 + 1] = blk1;
    partials[slot * 4 + 2] = blk2;
    __threadfence();
    const unsigned tk = atomicAdd(&ctl->ticket, 1u);
    is_last = (tk == (unsigned)P.ticket_total - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  steer_solve<S>(P, ctl, info_g, partials, reinterpret_cast<double (*)[3]>(smem_raw), tid, nthr, al_a, nk);
}

// ------------------------------------------------------------------------------------------------
// kernel 3/3, warp-synchronous variant for horizons with (T-1) | 32  (T = 2, 3, 5, 9, 17, 33).
// One warp owns PW = max(1, 8/(T-1)) problems at a time (PW * 20 * (T-1) items = a whole number of 32-lane
// rounds, no idle lanes); the T-1 stage items of a candidate sit in adjacent lanes, so candidate totals are
// segmented shuffle reductions and there is no CTA barrier and no partial-sum traffic in the loop.
// ------------------------------------------------------------------------------------------------
template <class M, typename S, int TM1, int WPB, int V>
__global__ void __launch_bounds__(WPB * 32, ALMPC_LS_MINB)
k_line_search_warp(const KParams<S> P, const S* __restrict__ x0_g, const S* __restrict__ Qd_g,
                   const S* __restrict__ q_g, S* xu_g, const S* __restrict__ lam_g, const S* __restrict__ rho_g,
                   const S* __restrict__ d_g, S* __restrict__ merit_g, const double* __restrict__ rentry_g,
                   int* __restrict__ status_g, int* __restrict__ info_g, S* __restrict__ debug_g, Ctl* ctl,
                   double* __restrict__ partials) {
  constexpr int NX = M::NX, NU = M::NU, NN = NX + NU, LS = NN | 1;
  constexpr int PW = (8 / TM1) > 0 ? (8 / TM1) : 1;
  constexpr int IT = NLS * TM1;          // items per problem
  constexpr int ROUNDS = PW * IT / 32;
  static_assert((PW * IT) % 32 == 0 && 32 % TM1 == 0, "warp variant needs (T-1) | 32");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double wred[WPB][3];
  __shared__ bool is_last;
  grid_dep_wait();
  const int mode = ctl->mode;
  if (mode == MODE_DONE) return;
  const int al_a = ctl->a, nk = ctl->k;
  constexpr int T = TM1 + 1;
  constexpr bool OBS = (V == 1), SE = (V == 2);
  const LsLayout<NX, NU> L(T, OBS ? P.n_obs : 0, SE);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int stride = L.part;  // staged scalars per problem (the partial-sum area of the CTA variant is unused)
  S* wsm = reinterpret_cast<S*>(smem_raw) + (size_t)warp * PW * stride;
  const int ngroups = (P.B + PW - 1) / PW;
  const int gwarp = blockIdx.x * WPB + warp, nwarps = gridDim.x * WPB;
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;  // this warp's running statistics (identical on all lanes)

  for (int grp = gwarp; grp < ngroups; grp += nwarps) {
    const int prob0 = grp * PW;
    __syncwarp();
#pragma unroll
    for (int p = 0; p < PW; ++p) {
      if (prob0 + p < P.B)
        ls_stage_async<NX, NU, S>(wsm + (size_t)p * stride, L, prob0 + p, lane, 32, x0_g, Qd_g, q_g, xu_g, lam_g, d_g, OBS ? P.obs : nullptr);
    }
    cp_async_commit();
    // per-problem scalars travel in registers while the copies are in flight
    S rho_p[PW], mcur_p[PW];
    double rent_p[PW];
#pragma unroll
    for (int p = 0; p < PW; ++p) {
      const bool v = prob0 + p < P.B;
      rho_p[p] = v ? rho_g[prob0 + p] : S(0);
      mcur_p[p] = v ? merit_g[prob0 + p] : S(0);
      rent_p[p] = v ? rentry_g[prob0 + p] : 0.0;
    }
    cp_async_wait_all();
    __syncwarp();

    S best_v[PW], best_r2[PW];
    int best_i[PW];
    bool nan_p[PW];
#pragma unroll
    for (int p = 0; p < PW; ++p) { best_v[p] = S(INFINITY); best_r2[p] = S(0); best_i[p] = NLS; nan_p[p] = false; }
#pragma unroll 1
    for (int r = 0; r < ROUNDS; ++r) {
      const int w = r * 32 + lane;
      const int p = w / IT, it = w - p * IT;
      const int kk = it / TM1, t = it - kk * TM1;
      const bool valid = prob0 + p < P.B;
      S cost = S(0), s2 = S(0), sl = S(0);
      if (valid) ls_item<M, S, V>(wsm + (size_t)p * stride, L, P, T, kk, t, cost, s2, sl);
      // candidate totals: the TM1 stage items of a candidate are adjacent lanes; fixed summation order t=0..TM1-1
      S c_tot = S(0), s_tot = S(0), l_tot = S(0);
#pragma unroll
      for (int tt = 0; tt < TM1; ++tt) {
        const int src = (lane & ~(TM1 - 1)) + tt;
        c_tot += __shfl_sync(FULL, cost, src);
        s_tot += __shfl_sync(FULL, s2, src);
        l_tot += __shfl_sync(FULL, sl, src);
      }
#pragma unroll
      for (int q = 0; q < PW; ++q) {
        if (valid && p == q && t == 0) {
          const S mk = c_tot + S(0.5) * rho_p[q] * s_tot + l_tot;
          if (debug_g) debug_g[(size_t)(prob0 + q) * (2 * L.N + 24) + 2 * L.N + kk] = mk;
          if (mk != mk) nan_p[q] = true;
          if (mk < best_v[q] || (mk == best_v[q] && kk < best_i[q])) { best_v[q] = mk; best_i[q] = kk; best_r2[q] = s_tot; }
        }
      }
    }
    // per problem: first-argmin over the lanes (torch.min NaN semantics), accept, iterate update
#pragma unroll
    for (int q = 0; q < PW; ++q) {
      const int prob = prob0 + q;
      if (prob >= P.B) continue;  // uniform across the warp
      S bv = best_v[q], br2 = best_r2[q];
      int bi = best_i[q];
      const bool any_nan = __any_sync(FULL, nan_p[q]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const S ov = __shfl_xor_sync(FULL, bv, o);
        const S or2 = __shfl_xor_sync(FULL, br2, o);
        const int oi = __shfl_xor_sync(FULL, bi, o);
        if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; br2 = or2; }
      }
      if (bi >= NLS) { bi = 0; br2 = S(0); }  // all +inf
      if (any_nan) bv = S(NAN);
      const bool accept = bv < mcur_p[q];
      const S alpha = (S)ldexpf(1.0f, -bi);
      const S* xu = wsm + (size_t)q * stride + L.xu;
      const S* dv = wsm + (size_t)q * stride + L.d;
      bool fin = true;
      const size_t pN = (size_t)prob * L.N;
      for (int i = lane; i < L.N; i += 32) {
        const int ts = i / NN, si = ts * LS + (i - ts * NN);
        const S xo = xu[si];
        const S xnew = xo + alpha * dv[si];
        const S xr = accept ? xnew : (finite_(xnew) ? xo : S(NAN));  // al_utils.py:640-641
        fin = fin && finite_(xr);
        if (accept || !finite_(xnew)) xu_g[pN + i] = xr;
      }
      fin = __all_sync(FULL, fin);
      const double r_exit = accept ? (double)br2 : (fin ? rent_p[q] : (double)NAN);
      acc0 += rent_p[q];
      acc1 += r_exit;
      acc2 = fmax(acc2, (double)rho_p[q]);
      if (lane == 0) {
        merit_g[prob] = bv;  // merit <- new_merit even when the step was rejected (al_utils.py:569)
        if (!fin) status_g[prob] &= ~ALMPC_STATUS_FINITE;
        if (debug_g) {
          S* dbg = debug_g + (size_t)prob * (2 * L.N + 24) + 2 * L.N + NLS;
          dbg[0] = (S)bi;
          dbg[1] = accept ? S(1) : S(0);
          dbg[2] = mcur_p[q];
        }
      }
    }
  }
  if (lane == 0) { wred[warp][0] = acc0; wred[warp][1] = acc1; wred[warp][2] = acc2; }
  __syncthreads();
  if (tid == 0) {
    double b0 = 0, b1 = 0, b2 = 0;
    for (int w = 0; w < WPB; ++w) { b0 += wred[w][0]; b1 += wred[w][1]; b2 = fmax(b2, wred[w][2]); }
    const int slot = P.part_off + blockIdx.x;
    partials[slot * 4 + 0] = b0;
    partials[slot * 4 + 1] = b1;
    partials[slot * 4 + 2] = b2;
    __threadfence();
    const unsigned tk = atomicAdd(&ctl->ticket, 1u);
    is_last = (tk == (unsigned)P.ticket_total - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  steer_solve<S>(P, ctl, info_g, partials, reinterpret_cast<double (*)[3]>(smem_raw), tid, nthr, al_a, nk);
}

// ------------------------------------------------------------------------------------------------
// final dual update + outputs (always runs once at the end of a forward solve)
//   lam <- lam + rho c(xu), clamp ; rho <- 10 rho ; dyn_res = ||c+|| ; x,u -> fp32     AL_mpc.py:315-338
// ------------------------------------------------------------------------------------------------
template <class M, typename S, int WPB, int V>
__global__ void __launch_bounds__(WPB * 32)
k_final_dual(const KParams<S> P, const S* __restrict__ x0_g, const S* __restrict__ xu_g, S* __restrict__ lam_g,
             S* __restrict__ rho_g, float* __restrict__ x_out, float* __restrict__ u_out, S* __restrict__ dyn_res_g,
             int* __restrict__ info_g, Ctl* ctl) {
  constexpr int NX = M::NX, NU = M::NU, NN = NX + NU;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int T = P.T;
  constexpr bool OBS = (V == 1), SE = (V == 2);
  const int n_obs = OBS ? P.n_obs : 0;
  const Layout<NX, NU> L(T, n_obs, SE);
  const int SI = OBS ? L.SI : 2 * NU;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int prob = blockIdx.x * WPB + warp;
  grid_dep_wait();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    info_g[ALMPC_INFO_RHO_EXCEEDED] = (ctl->rho_final_max > P.rho_max) ? 1 : 0;
  }
  if (prob >= P.B) return;
  S* sm = reinterpret_cast<S*>(smem_raw) + (size_t)warp * (L.N + L.neq);
  S* xu = sm;
  S* cres = sm + L.N;
  const size_t pN = (size_t)prob * L.N;
  for (int i = lane; i < L.N; i += 32) xu[i] = xu_g[pN + i];
  __syncwarp();
  for (int t = lane; t < T - 1; t += 32) {
    S xs[NX], us[NU], xn[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) xs[i] = xu[t * NN + i];
#pragma unroll
    for (int i = 0; i < NU; ++i) us[i] = xu[t * NN + NX + i];
    rk4_step<M, S, S>(xs, us, xn, P.dyn);
#pragma unroll
    for (int i = 0; i < NX; ++i) cres[t * NX + i] = xu[(t + 1) * NN + i] - xn[i];
  }
  if (!SE)
    for (int i = lane; i < NX; i += 32) cres[(T - 1) * NX + i] = xu[i] - x0_g[(size_t)prob * NX + i];
  __syncwarp();
  const S rho = rho_g[prob];
  S* lamp = lam_g + (size_t)prob * L.m;
  const bool upd = !(P.flags & ALMPC_FLAG_SKIP_FINAL_DUAL);
  S s2 = S(0);
  for (int i = lane; i < L.neq; i += 32) {
    S ci = cres[i];
    if (upd) lamp[i] += rho * ci;
    s2 += ci * ci;
  }
  const S* obs = OBS ? P.obs + (size_t)prob * T * n_obs * 3 : nullptr;
  for (int i = lane; i < L.nin; i += 32) {
    const int t = i / SI, s = i - t * SI;
    const S ci = ineq_res<NX, NU, S>(xu + t * NN, s, P.ulo, P.uhi, OBS ? obs + (size_t)t * n_obs * 3 : nullptr, P.obs_r2);
    if (upd) lamp[L.neq + i] = fmax0(lamp[L.neq + i] + rho * ci);
    S cp = fmax0(ci);
    s2 += cp * cp;
  }
  s2 = warp_sum(s2);
  if (lane == 0) {
    if (upd) rho_g[prob] = next_rho(rho, P.rho_mult, P.rho_f32);
    if (dyn_res_g) dyn_res_g[prob] = sqrt(s2);
  }
  if (x_out) {
    for (int i = lane; i < T * NX; i += 32) {
      int t = i / NX, j = i - t * NX;
      x_out[(size_t)prob * T * NX + i] = (float)xu[t * NN + j];
    }
  }
  if (u_out) {
    for (int i = lane; i < T * NU; i += 32) {
      int t = i / NU, j = i - t * NU;
      u_out[(size_t)prob * T * NU + i] = (float)xu[t * NN + NX + j];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// implicit backward: gbar = -H^{-1} grad ; grad_q = gbar ; grad_Q = gbar * xu     al_utils.py:578-615
// ------------------------------------------------------------------------------------------------
template <int NX, int NU, typename S, int WPB>
__global__ void __launch_bounds__(WPB * 32)
k_backward(int B, int T, const S* __restrict__ fac_g, const S* __restrict__ xu_g, const S* __restrict__ grad_g,
           S* __restrict__ gQ_g, S* __restrict__ gq_g) {
  constexpr int NN = NX + NU;
  constexpr int LB = NN | 1;  // odd row stride of the staged blocks: rows AND columns per lane are conflict free
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int prob = blockIdx.x * WPB + warp;
  if (prob >= B) return;
  const int N = T * NN;
  const int per_warp = ((N + 3) & ~3) + 2 * NN * LB + 2 * NX * LB;
  S* sm = reinterpret_cast<S*>(smem_raw) + (size_t)warp * per_warp;
  S* dv = sm;
  S* Lb[2] = {sm + ((N + 3) & ~3), sm + ((N + 3) & ~3) + NN * LB};
  S* Eb[2] = {Lb[1] + NN * LB, Lb[1] + NN * LB + NX * LB};
  const size_t pN = (size_t)prob * N;
  const S* fac = fac_g + (size_t)prob * factor_elems<NX, NU>(T);
  const S* facE = fac + (size_t)T * NN * NN;
  // both sweeps stream the factor through shared memory with cp.async, one stage ahead of the substitution
  warp_copy_block_async<NN, NN, LB>(Lb[0], fac, lane);
  cp_async_commit();
  for (int i = lane; i < N; i += 32) dv[i] = grad_g[pN + i];
  // ---- forward: y = L^{-1} grad   (row per lane; stage t needs L_tt and E_{t-1})
  for (int t = 0; t < T; ++t) {
    const int cur = t & 1, nxt = cur ^ 1;
    cp_async_wait_all();
    __syncwarp();
    if (t + 1 < T) {
      warp_copy_block_async<NN, NN, LB>(Lb[nxt], fac + (size_t)(t + 1) * NN * NN, lane);
      warp_copy_block_async<NX, NN, LB>(Eb[nxt], facE + (size_t)t * NX * NN, lane);  // E_t, used by stage t+1
    }
    cp_async_commit();
    if (t > 0 && lane < NX) {
      const S* E = Eb[cur];
      S s0 = S(0), s1 = S(0);
#pragma unroll
      for (int k = 0; k + 1 < NN; k += 2) {
        s0 += E[lane * LB + k] * dv[(t - 1) * NN + k];
        s1 += E[lane * LB + k + 1] * dv[(t - 1) * NN + k + 1];
      }
      if (NN & 1) s0 += E[lane * LB + NN - 1] * dv[(t - 1) * NN + NN - 1];
      dv[t * NN + lane] -= s0 + s1;
    }
    __syncwarp();
    S a[NN];
    if (lane < NN) {
#pragma unroll
      for (int c = 0; c < NN; ++c) a[c] = Lb[cur][lane * LB + c];
    }
    S b = (lane < NN) ? dv[t * NN + lane] : S(0);
#pragma unroll
    for (int k = 0; k < NN; ++k) {
      const S yk = __shfl_sync(FULL, b * a[k], k);  // lane k: a[k] = 1/L_kk
      b = (lane == k) ? yk : ((lane > k) ? (b - a[k] * yk) : b);
    }
    if (lane < NN) dv[t * NN + lane] = b;
  }
  // ---- backward: gbar = -L^{-T} y   (column per lane; stage t needs L_tt and E_t)
  // after the forward loop Lb[(T-1)&1] still holds L_{T-1}
  int cur = (T - 1) & 1;
  cp_async_wait_all();
  __syncwarp();
  for (int t = T - 1; t >= 0; --t) {
    const int nxt = cur ^ 1;
    if (t > 0) {
      warp_copy_block_async<NN, NN, LB>(Lb[nxt], fac + (size_t)(t - 1) * NN * NN, lane);
      warp_copy_block_async<NX, NN, LB>(Eb[nxt], facE + (size_t)(t - 1) * NX * NN, lane);  // E_{t-1}, used by stage t-1
    }
    cp_async_commit();
    S l[NN];
    if (lane < NN) {
#pragma unroll
      for (int r = 0; r < NN; ++r) l[r] = (r >= lane) ? Lb[cur][r * LB + lane] : S(0);
    }
    S b = (lane < NN) ? dv[t * NN + lane] : S(0);
    if (t < T - 1 && lane < NN) {
      const S* E = Eb[cur];  // E_t (staged one iteration ago)
      S s0 = S(0), s1 = S(0);
#pragma unroll
      for (int i = 0; i + 1 < NX; i += 2) {
        s0 += E[i * LB + lane] * dv[(t + 1) * NN + i];
        s1 += E[(i + 1) * LB + lane] * dv[(t + 1) * NN + i + 1];
      }
      if (NX & 1) s0 += E[(NX - 1) * LB + lane] * dv[(t + 1) * NN + NX - 1];
      b -= s0 + s1;
    }
#pragma unroll
    for (int r = NN - 1; r >= 0; --r) {
      const S dr = __shfl_sync(FULL, b * l[r], r);
      b = (lane == r) ? dr : (b - l[r] * dr);
    }
    if (lane < NN) dv[t * NN + lane] = b;
    cp_async_wait_all();
    __syncwarp();
    cur = nxt;
  }
  // a factor with non-positive pivots (flagged in the forward status) yields no gradient instead of NaNs
  bool fin = true;
  for (int i = lane; i < N; i += 32) fin = fin && finite_(dv[i]);
  fin = __all_sync(FULL, fin);
  for (int i = lane; i < N; i += 32) {
    const S g = fin ? -dv[i] : S(0);
    gq_g[pN + i] = g;
    if (gQ_g) gQ_g[pN + i] = g * xu_g[pN + i];
  }
}


// ------------------------------------------------------------------------------------------------
// generic-callable dynamics mode: block-tridiagonal factor + solve with RUN-TIME dimensions.
// When the dynamics are a user's Python callable, the host loop (deq_mpc_corl_b200/generic.py) evaluates
// it with torch and hands this kernel the dynamics Jacobians; it is the one piece of the Newton step the
// reference does densely (H[N,N], cholesky_ex + cholesky_solve, al_utils.py:510-513,586-590).
//   H = diag(dD) + rho * sum_t [ -G_t | I ]^T [ -G_t | I ]   restricted to the G^T G and -rho G_t blocks
//       (every diagonal contribution, incl. the rho I of the identity blocks, is in dD)
//   sol = -H^{-1} rhs ;  factor layout = the fused path's (T blocks L_tt with 1/L_ii on the diagonal, then E_t)
// 1 warp = 1 problem; row per lane, blocks in shared memory (n = nx+nu <= 32).
// ------------------------------------------------------------------------------------------------
template <typename S, int WPB>
__global__ void __launch_bounds__(WPB * 32)
k_blocktri(int B, int T, int nx, int nu, int do_factor, const S* __restrict__ G_g, const S* __restrict__ dD_g,
           const S* __restrict__ rho_g, S* fac_g, const S* __restrict__ rhs_g, S* __restrict__ sol_g,
           int* __restrict__ status_g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = nx + nu, LD = n | 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int prob = blockIdx.x * WPB + warp;
  if (prob >= B) return;
  const int per_warp = n * LD + 2 * nx * LD + T * n;
  S* Dm = reinterpret_cast<S*>(smem_raw) + (size_t)warp * per_warp;  // D_t, then L_tt (diag = 1/L_ii)
  S* Gm = Dm + n * LD;                                               // G_t
  S* Em = Gm + nx * LD;                                              // E_{t-1}, then E_t
  S* yv = Em + nx * LD;                                              // rhs -> y -> solution
  const size_t FE = (size_t)T * n * n + (size_t)(T - 1) * nx * n;
  S* fac = fac_g + (size_t)prob * FE;
  S* facE = fac + (size_t)T * n * n;
  const S* Gp = G_g ? G_g + (size_t)prob * (T - 1) * nx * n : nullptr;
  const S rho = do_factor ? rho_g[prob] : S(0);
  for (int i = lane; i < T * n; i += 32) yv[i] = rhs_g[(size_t)prob * T * n + i];
  bool ok = true;
  for (int t = 0; t < T; ++t) {
    const bool has_G = t < T - 1, has_E = t > 0;
    S* Lg = fac + (size_t)t * n * n;
    __syncwarp();
    if (do_factor) {
      if (has_G)
        for (int e = lane; e < nx * n; e += 32) Gm[(e / n) * LD + (e % n)] = Gp[(size_t)t * nx * n + e];
      __syncwarp();
      for (int idx = lane; idx < n * n; idx += 32) {
        const int r = idx / n, c = idx - r * n;
        if (c > r) continue;
        S v = (r == c) ? dD_g[((size_t)prob * T + t) * n + r] : S(0);
        if (has_G) {
          S a = S(0);
          for (int i = 0; i < nx; ++i) a += Gm[i * LD + r] * Gm[i * LD + c];
          v += rho * a;
        }
        if (has_E && r < nx) {
          S a = S(0);
          for (int k = 0; k < n; ++k) a += Em[r * LD + k] * Em[c * LD + k];
          v -= a;
        }
        Dm[r * LD + c] = v;
      }
      __syncwarp();
      for (int k = 0; k < n; ++k) {  // right-looking Cholesky, row per lane
        const S dkk = Dm[k * LD + k];
        ok = ok && (dkk > S(0));
        const S inv = S(1) / sqrt(dkk);
        __syncwarp();
        if (lane == k) Dm[k * LD + k] = inv;
        else if (lane > k && lane < n) Dm[lane * LD + k] *= inv;
        __syncwarp();
        if (lane > k && lane < n) {
          const S lrk = Dm[lane * LD + k];
          for (int c = k + 1; c <= lane; ++c) Dm[lane * LD + c] -= lrk * Dm[c * LD + k];
        }
        __syncwarp();
      }
      for (int idx = lane; idx < n * n; idx += 32) {
        const int r = idx / n, c = idx - r * n;
        Lg[idx] = (c <= r) ? Dm[r * LD + c] : S(0);
      }
    } else {
      for (int idx = lane; idx < n * n; idx += 32) Dm[(idx / n) * LD + (idx % n)] = Lg[idx];
      if (has_E)
        for (int e = lane; e < nx * n; e += 32) Em[(e / n) * LD + (e % n)] = facE[(size_t)(t - 1) * nx * n + e];
      __syncwarp();
    }
    // y_t = L^{-1} (rhs_t - E_{t-1} y_{t-1})
    S b = (lane < n) ? yv[t * n + lane] : S(0);
    if (has_E && lane < nx) {
      S a = S(0);
      for (int k = 0; k < n; ++k) a += Em[lane * LD + k] * yv[(t - 1) * n + k];
      b -= a;
    }
    for (int k = 0; k < n; ++k) {
      const S yk = __shfl_sync(FULL, b * Dm[k * LD + k], k);
      if (lane == k) b = yk;
      else if (lane > k && lane < n) b -= Dm[lane * LD + k] * yk;
    }
    __syncwarp();
    if (lane < n) yv[t * n + lane] = b;
    // E_t = -rho G_t L^{-T}  (row per lane), kept for stage t+1 and stored for the backward sweep
    if (do_factor && has_G) {
      __syncwarp();
      if (lane < nx) {
        for (int k = 0; k < n; ++k) {
          S a = -rho * Gm[lane * LD + k];
          for (int j = 0; j < k; ++j) a -= Em[lane * LD + j] * Dm[k * LD + j];
          Em[lane * LD + k] = a * Dm[k * LD + k];
        }
      }
      __syncwarp();
      for (int e = lane; e < nx * n; e += 32) facE[(size_t)t * nx * n + e] = Em[(e / n) * LD + (e % n)];
    }
  }
  ok = __all_sync(FULL, ok);
  __syncwarp();
  // backward sweep: d_t = L^{-T} (y_t - E_t^T d_{t+1}), factor read back from global memory
  for (int t = T - 1; t >= 0; --t) {
    const S* Lg = fac + (size_t)t * n * n;
    __syncwarp();
    for (int idx = lane; idx < n * n; idx += 32) Dm[(idx / n) * LD + (idx % n)] = Lg[idx];
    if (t < T - 1)
      for (int e = lane; e < nx * n; e += 32) Em[(e / n) * LD + (e % n)] = facE[(size_t)t * nx * n + e];
    __syncwarp();
    S b = (lane < n) ? yv[t * n + lane] : S(0);
    if (t < T - 1 && lane < n) {
      S a = S(0);
      for (int i = 0; i < nx; ++i) a += Em[i * LD + lane] * yv[(t + 1) * n + i];
      b -= a;
    }
    for (int r = n - 1; r >= 0; --r) {
      const S dr = __shfl_sync(FULL, b * Dm[r * LD + r], r);
      if (lane == r) b = dr;
      else if (lane < r) b -= Dm[r * LD + lane] * dr;
    }
    __syncwarp();
    if (lane < n) yv[t * n + lane] = b;
  }
  __syncwarp();
  bool fin = true;
  for (int i = lane; i < T * n; i += 32) fin = fin && finite_(yv[i]);
  fin = __all_sync(FULL, fin) && ok;  // failed pivot / non-finite: no direction (same policy as k_factor_solve)
  for (int i = lane; i < T * n; i += 32) sol_g[(size_t)prob * T * n + i] = fin ? -yv[i] : S(0);
  if (status_g && lane == 0 && do_factor) {
    int st = status_g[prob];
    if (!ok) st &= ~ALMPC_STATUS_CHOL_OK;
    status_g[prob] = st;
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

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
#if ALMPC_BUILD_F64 && !ALMPC_BUILD_SE
thread_local char g_err[512] = "";
#else
extern thread_local char g_err[512];
#endif
static int set_err(int code, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s", what);
  return code;
}
static int cuda_err(cudaError_t e, const char* where) {
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return (int)e;
}

template <typename S>
static KParams<S> make_params(const almpc_desc* d) {
  KParams<S> P;
  for (int i = 0; i < ALMPC_MAX_PARAMS; ++i) P.dyn[i] = (S)d->dyn_params[i];
  for (int i = 0; i < ALMPC_MAX_NU; ++i) { P.ulo[i] = (S)d->u_lower[i]; P.uhi[i] = (S)d->u_upper[i]; }
  P.rho_mult = d->rho_mult;
  P.rho_max = d->rho_max;
  P.B = d->B; P.T = d->T; P.al_iter = d->al_iter; P.max_newton = d->max_newton;
  P.stream_mode = d->stream_mode; P.rho_f32 = d->rho_f32; P.flags = d->flags;
  P.part_off = 0; P.ticket_total = 0;
  P.n_obs = d->obstacles ? d->n_obs : 0;
  P.obs_r2 = (S)(d->obs_radius * d->obs_radius);
  P.obs = (const S*)d->obstacles;
  return P;
}

static int check_desc(const almpc_desc* d) {
  if (!d) return set_err(-1, "null descriptor");
  if (d->dyn_id < 0 || d->dyn_id >= ALMPC_DYN_COUNT) return set_err(-2, "unknown dyn_id");
  if (d->B < 0) return set_err(-3, "negative batch");
  if (d->T < 2 || d->T > 64) return set_err(-4, "horizon T must be in [2,64]");
  if (d->al_iter < 1 || d->al_iter > ALMPC_MAX_AL_ITER) return set_err(-5, "al_iter out of range");
  if (d->max_newton < 1) return set_err(-6, "max_newton must be >= 1");
  if (d->n_obs < 0 || d->n_obs > 8) return set_err(-7, "n_obs must be in [0,8]");
  return 0;
}

constexpr int kWPB = 4;
constexpr int kMaxParts = 4;

// optional per-kernel timing (bench.py): events around every launch of one forward call
// (shared by the two translation units: defined in the fp64 unit)
#if ALMPC_BUILD_F64 && !ALMPC_BUILD_SE
int g_profile_on = 0;
double g_profile[8] = {0};  // ms: [0] jacobian [1] factor_solve [2] line_search [3] final_dual, [4..7] launch counts
#else
extern int g_profile_on;
extern double g_profile[8];
#endif
struct ProfEv { cudaEvent_t a, b; int kind; };

static int n_blocks(int B) { return (B + kWPB - 1) / kWPB; }

// line-search launch geometry: `pls` problems per CTA, threads = one per (candidate, stage) item when it fits
struct LsGeom { int pls, threads, blocks; };
static LsGeom ls_geom(int B, int T) {
  const int IT = NLS * (T - 1);
  LsGeom g;
  g.pls = 160 / IT;
  if (g.pls < 1) g.pls = 1;
  if (g.pls > 8) g.pls = 8;
  int thr = g.pls * IT;
  if (thr > 256) thr = 256;
  g.threads = (thr + 31) & ~31;
  g.blocks = (B + g.pls - 1) / g.pls;
  return g;
}

// device workspace carved out of the caller's buffer
struct WsPlan {
  size_t ctl, partials, merit, rentry, status, c, d, G, total;
};
static WsPlan ws_plan(int B, int T, int nx, int nu, size_t es) {
  const size_t nn = (size_t)nx + nu;
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  WsPlan w;
  size_t o = 0;
  w.ctl = o; o += 256;
  w.partials = o; o = up(o + ((size_t)ls_geom(B, T).blocks + kMaxParts) * 4 * sizeof(double));
  w.merit = o; o = up(o + (size_t)B * es);
  w.rentry = o; o = up(o + (size_t)B * sizeof(double));
  w.status = o; o = up(o + (size_t)B * sizeof(int32_t));
  w.c = o; o = up(o + (size_t)B * T * nx * es);
  w.d = o; o = up(o + (size_t)B * T * nn * es);
  w.G = o; o = up(o + (size_t)B * (T - 1) * nx * nn * es);
  w.total = o;
  return w;
}

// launch with (optional) programmatic stream serialization
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*k)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t s, bool pdl,
                            Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, k, KArgs(args)...);
}

constexpr int kWLS = 4;  // warps per CTA of the warp-synchronous line search

struct LsWarpCfg { const void* fn; int grid_cap; size_t smem; int pw; };

template <class M, typename S, int TM1, int V>
static cudaError_t ls_warp_cfg(LsWarpCfg& c, int n_sm, int n_obs) {
  constexpr int PW = (8 / TM1) > 0 ? (8 / TM1) : 1;
  const LsLayout<M::NX, M::NU> LL(TM1 + 1, n_obs, V == 2);
  auto k = k_line_search_warp<M, S, TM1, kWLS, V>;
  size_t smem = (size_t)kWLS * PW * LL.part * sizeof(S);
  if (smem < (size_t)kWLS * 32 * 3 * sizeof(double)) smem = (size_t)kWLS * 32 * 3 * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kWLS * 32, smem);
  if (e != cudaSuccess) return e;
  c.fn = (const void*)k;
  c.grid_cap = n_sm * (per_sm > 0 ? per_sm : 1);
  c.smem = smem;
  c.pw = PW;
  return cudaSuccess;
}

template <class M, typename S, int V>
static int forward_impl(const almpc_desc* d, const S* x0, const S* Qd, const S* q, S* xu, S* lam, S* rho, S* factor,
                        float* x_out, float* u_out, S* dyn_res_out, int32_t* status_out, int32_t* info_out,
                        S* debug_out, void* workspace, cudaStream_t st) {
  constexpr int NX = M::NX, NU = M::NU, NN = NX + NU;
  if (d->B == 0) return 0;
  if (!x0 || !Qd || !q || !xu || !lam || !rho || !factor || !info_out || !workspace)
    return set_err(-10, "null buffer");
  const int T = d->T, B = d->B;
  constexpr bool OBS = (V == 1), SE = (V == 2);
  const int n_obs = OBS ? d->n_obs : 0;
  const Layout<NX, NU> L(T, n_obs, SE);
  const LsLayout<NX, NU> LL(T, n_obs, SE);
  const LsGeom lg = ls_geom(B, T);
  size_t smem_fac = (size_t)kWPB * L.total * sizeof(S);
  if (const char* pad = getenv("ALMPC_FAC_SMEM_PAD")) smem_fac += (size_t)atoi(pad);  // occupancy experiments
  size_t smem_ls = 2 * (size_t)lg.pls * LL.total * sizeof(S);  // double-buffered staging
  if (smem_ls < (size_t)lg.threads * 3 * sizeof(double)) smem_ls = (size_t)lg.threads * 3 * sizeof(double);
  const size_t smem_fin = (size_t)kWPB * (L.N + L.neq) * sizeof(S);
  if (smem_fac + 1024 > 227 * 1024 || smem_ls + 1024 > 227 * 1024)
    return set_err(-11, "horizon too long for the shared-memory layout");
  auto kjac = k_jacobian<M, S, V>;
  auto kfac = k_factor_solve<NX, NU, S, kWPB, V>;
  auto kls = k_line_search<M, S, V>;
  auto kfin = k_final_dual<M, S, kWPB, V>;
  cudaError_t e;
  e = cudaFuncSetAttribute(kfac, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fac);
  if (e != cudaSuccess) return cuda_err(e, "cudaFuncSetAttribute(factor_solve)");
  e = cudaFuncSetAttribute(kls, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ls);
  if (e != cudaSuccess) return cuda_err(e, "cudaFuncSetAttribute(line_search)");
  e = cudaFuncSetAttribute(kfin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fin);
  if (e != cudaSuccess) return cuda_err(e, "cudaFuncSetAttribute(final_dual)");
  // persistent line-search grid: as many CTAs as fit on the device at once
  int ls_per_sm = 0, n_sm = 0, dev = 0;
  e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_err(e, "cudaGetDevice");
  e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return cuda_err(e, "cudaDeviceGetAttribute");
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ls_per_sm, kls, lg.threads, smem_ls);
  if (e != cudaSuccess) return cuda_err(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
  int ls_grid = n_sm * (ls_per_sm > 0 ? ls_per_sm : 1);
  if (ls_grid > lg.blocks) ls_grid = lg.blocks;
  // warp-synchronous line search when the stage items of a candidate fit a power-of-two lane group
  LsWarpCfg lw{nullptr, 0, 0, 1};
  if (!getenv("ALMPC_LS_CTA")) {
    switch (T - 1) {
      case 1: e = ls_warp_cfg<M, S, 1, V>(lw, n_sm, n_obs); break;
      case 2: e = ls_warp_cfg<M, S, 2, V>(lw, n_sm, n_obs); break;
      case 4: e = ls_warp_cfg<M, S, 4, V>(lw, n_sm, n_obs); break;
      case 8: e = ls_warp_cfg<M, S, 8, V>(lw, n_sm, n_obs); break;
      case 16: e = ls_warp_cfg<M, S, 16, V>(lw, n_sm, n_obs); break;
      case 32: e = ls_warp_cfg<M, S, 32, V>(lw, n_sm, n_obs); break;
      default: break;
    }
    if (e != cudaSuccess) return cuda_err(e, "line-search (warp) configuration");
  }
  KParams<S> P = make_params<S>(d);
  const bool prof_requested = g_profile_on != 0;
  const WsPlan w = ws_plan(B, T, NX, NU, sizeof(S));
  unsigned char* ws = (unsigned char*)workspace;
  Ctl* ctl = (Ctl*)(ws + w.ctl);
  double* partials = (double*)(ws + w.partials);
  S* merit = (S*)(ws + w.merit);
  double* rentry = (double*)(ws + w.rentry);
  int32_t* status = status_out ? status_out : (int32_t*)(ws + w.status);
  S* c_g = (S*)(ws + w.c);
  S* d_g = (S*)(ws + w.d);
  S* G_g = (S*)(ws + w.G);
  e = cudaMemsetAsync(ctl, 0, 256, st);
  if (e != cudaSuccess) return cuda_err(e, "cudaMemsetAsync(ctl)");
  e = cudaMemsetAsync(info_out, 0, ALMPC_INFO_INTS * sizeof(int32_t), st);
  if (e != cudaSuccess) return cuda_err(e, "cudaMemsetAsync(info)");
  // Optional (ALMPC_PARTS=2..4, default 1): process the batch in contiguous parts on separate streams so that
  // the LSU-bound factor kernel of one part could overlap the FMA-bound jacobian / line-search kernels of
  // another.  All parts feed the same batch-global statistics (one ticket counter) and the streams join
  // after every Newton step.  Measured on B200 (quadrotor, 64K): no gain — k_factor_solve fills the SM's
  // registers and shared memory, so CTAs of other kernels do not co-reside; kept as an experiment knob.
  int nparts = 1;
  {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    const char* env = getenv("ALMPC_PARTS");
    if (env) nparts = atoi(env);
    if (nparts < 1) nparts = 1;
    if (nparts > kMaxParts) nparts = kMaxParts;
    if (cs != cudaStreamCaptureStatusNone || prof_requested) nparts = 1;
  }
  static thread_local cudaStream_t aux[kMaxParts] = {nullptr};
  static thread_local cudaEvent_t ev_fork = nullptr, ev_part[kMaxParts] = {nullptr};
  static thread_local int aux_dev = -1;
  if (nparts > 1 && (aux_dev != dev || !ev_fork)) {
    for (int p = 0; p < kMaxParts; ++p) {
      if (p > 0) { e = cudaStreamCreateWithFlags(&aux[p], cudaStreamNonBlocking); if (e != cudaSuccess) return cuda_err(e, "cudaStreamCreate"); }
      e = cudaEventCreateWithFlags(&ev_part[p], cudaEventDisableTiming);
      if (e != cudaSuccess) return cuda_err(e, "cudaEventCreate");
    }
    e = cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming);
    if (e != cudaSuccess) return cuda_err(e, "cudaEventCreate");
    aux_dev = dev;
  }
  struct Part { int lo, n, ls_grid, nb; unsigned jac_blocks; cudaStream_t s; KParams<S> P; };
  Part parts[kMaxParts];
  const int chunk = ((B + nparts - 1) / nparts + 7) & ~7;
  int ticket_total = 0, np = 0;
  for (int p = 0; p < nparts; ++p) {
    const int lo = p * chunk;
    if (lo >= B) break;
    Part& pt = parts[np];
    pt.lo = lo;
    pt.n = (lo + chunk <= B) ? chunk : B - lo;
    if (lw.fn) {
      const int groups = (pt.n + lw.pw - 1) / lw.pw;
      const int ctas = (groups + kWLS - 1) / kWLS;
      pt.ls_grid = lw.grid_cap < ctas ? lw.grid_cap : ctas;
    } else {
      const int groups = (pt.n + lg.pls - 1) / lg.pls;
      pt.ls_grid = ls_grid < groups ? ls_grid : groups;
    }
    pt.nb = n_blocks(pt.n);
    pt.jac_blocks = (unsigned)(((long)pt.n * (T - 1) * (NN - M::NTRIV) + 127) / 128);
    pt.s = (np == 0) ? st : aux[np];
    pt.P = P;
    pt.P.B = pt.n;
    pt.P.part_off = ticket_total;
    if (pt.P.obs) pt.P.obs = P.obs + (size_t)lo * T * P.n_obs * 3;
    ticket_total += pt.ls_grid;
    ++np;
  }
  for (int p = 0; p < np; ++p) parts[p].P.ticket_total = ticket_total;
  const size_t FE = factor_elems<NX, NU>(T), GE = (size_t)(T - 1) * NX * NN, DBG = 2 * (size_t)L.N + 24;
  if (np > 1) {
    e = cudaEventRecord(ev_fork, st);
    if (e != cudaSuccess) return cuda_err(e, "cudaEventRecord");
    for (int p = 1; p < np; ++p) cudaStreamWaitEvent(parts[p].s, ev_fork, 0);
  }
  const int nb = n_blocks(B);
  const int launches = d->al_iter * d->max_newton;
  const bool pdl = !prof_requested && np == 1 && getenv("ALMPC_PDL");  // measured: no gain on B200, off by default
  ProfEv pe[3 * ALMPC_MAX_AL_ITER * 8 + 1];
  int npe = 0;
  const bool prof = prof_requested && launches * 3 + 1 <= (int)(sizeof(pe) / sizeof(pe[0]));
  auto tic = [&](int kind) {
    if (!prof) return;
    cudaEventCreate(&pe[npe].a);
    cudaEventCreate(&pe[npe].b);
    pe[npe].kind = kind;
    cudaEventRecord(pe[npe].a, st);
  };
  auto toc = [&]() {
    if (!prof) return;
    cudaEventRecord(pe[npe].b, st);
    ++npe;
  };
  for (int i = 0; i < launches; ++i) {
    for (int p = 0; p < np; ++p) {
      const Part& pt = parts[p];
      const size_t lo = (size_t)pt.lo;
      S* dbg = debug_out ? debug_out + lo * DBG : nullptr;
      tic(0);
      launch_k(kjac, pt.jac_blocks, 128, 0, pt.s, pdl, pt.P, x0 + lo * NX, xu + lo * L.N, G_g + lo * GE,
               c_g + lo * L.neq, ctl);
      toc();
      tic(1);
      launch_k(kfac, pt.nb, kWPB * 32, smem_fac, pt.s, pdl, pt.P, Qd + lo * L.N, q + lo * L.N, xu + lo * L.N,
               lam + lo * L.m, rho + lo, factor + lo * FE, G_g + lo * GE, c_g + lo * L.neq, d_g + lo * L.N,
               merit + lo, rentry + lo, status + lo, dbg, ctl);
      toc();
      tic(2);
      if (lw.fn) {
        const S *a_x0 = x0 + lo * NX, *a_Qd = Qd + lo * L.N, *a_q = q + lo * L.N, *a_lam = lam + lo * L.m;
        const S *a_rho = rho + lo, *a_d = d_g + lo * L.N;
        S *a_xu = xu + lo * L.N, *a_merit = merit + lo;
        const double* a_re = rentry + lo;
        int* a_st = status + lo;
        void* args[] = {(void*)&pt.P, &a_x0, &a_Qd, &a_q, &a_xu, &a_lam, &a_rho, &a_d, &a_merit, &a_re, &a_st,
                        &info_out, &dbg, &ctl, &partials};
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(pt.ls_grid);
        cfg.blockDim = dim3(kWLS * 32);
        cfg.dynamicSmemBytes = lw.smem;
        cfg.stream = pt.s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = pdl ? 1 : 0;
        e = cudaLaunchKernelExC(&cfg, lw.fn, args);
        if (e != cudaSuccess) return cuda_err(e, "k_line_search_warp launch");
      } else {
        launch_k(kls, pt.ls_grid, lg.threads, smem_ls, pt.s, pdl, pt.P, lg.pls, x0 + lo * NX, Qd + lo * L.N,
                 q + lo * L.N, xu + lo * L.N, lam + lo * L.m, rho + lo, d_g + lo * L.N, merit + lo, rentry + lo,
                 status + lo, info_out, dbg, ctl, partials);
      }
      toc();
    }
    if (np > 1) {  // join: every part must see the control block written by the globally last CTA
      for (int p = 0; p < np; ++p) cudaEventRecord(ev_part[p], parts[p].s);
      for (int p = 0; p < np; ++p)
        for (int o = 0; o < np; ++o)
          if (o != p) cudaStreamWaitEvent(parts[p].s, ev_part[o], 0);
    }
  }
  tic(3);
  launch_k(kfin, nb, kWPB * 32, smem_fin, st, pdl, P, x0, xu, lam, rho, x_out, u_out, dyn_res_out, info_out, ctl);
  toc();
  if (prof) {  // profiling mode synchronises (never used on the timed path)
    cudaStreamSynchronize(st);
    for (int k = 0; k < 8; ++k) g_profile[k] = 0;
    for (int i = 0; i < npe; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, pe[i].a, pe[i].b);
      if (ms > 0.02f || pe[i].kind == 3) {  // skip the no-op launches after the solve finished
        g_profile[pe[i].kind] += ms;
        g_profile[4 + pe[i].kind] += 1;
      }
      cudaEventDestroy(pe[i].a);
      cudaEventDestroy(pe[i].b);
    }
  }
  return cuda_err(cudaGetLastError(), "almpc_forward launch");
}

template <int NX, int NU, typename S>
static int backward_impl(const almpc_desc* d, const S* factor, const S* xu, const S* grad_xu, S* grad_Q, S* grad_q,
                         cudaStream_t st) {
  if (d->B == 0) return 0;
  if (!factor || !xu || !grad_xu || !grad_q) return set_err(-10, "null buffer");
  constexpr int NN = NX + NU;
  constexpr int LBb = NN | 1;
  const size_t smem = (size_t)kWPB * (((d->T * NN + 3) & ~3) + 2 * NN * LBb + 2 * NX * LBb) * sizeof(S);
  auto kern = k_backward<NX, NU, S, kWPB>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_err(e, "cudaFuncSetAttribute(backward)");
  }
  kern<<<n_blocks(d->B), kWPB * 32, smem, st>>>(d->B, d->T, factor, xu, grad_xu, grad_Q, grad_q);
  return cuda_err(cudaGetLastError(), "almpc_backward launch");
}

template <class M, typename S>
static int dynamics_impl(const almpc_desc* d, int64_t R, const S* x, const S* u, S* xn, S* Jx, S* Ju, cudaStream_t st) {
  if (R == 0) return 0;
  if (!x || !u) return set_err(-10, "null buffer");
  if ((Jx == nullptr) != (Ju == nullptr)) return set_err(-12, "Jx and Ju must both be given or both be NULL");
  if (!Jx && !xn) return set_err(-10, "null output");
  const KParams<S> P = make_params<S>(d);
  const int64_t threads = Jx ? R * (M::NX + M::NU) : R;
  const int bs = 128;
  k_dynamics<M, S><<<(unsigned)((threads + bs - 1) / bs), bs, 0, st>>>(P, R, x, u, xn, Jx, Ju);
  return cuda_err(cudaGetLastError(), "almpc_dynamics launch");
}

#define ALMPC_DISPATCH(dyn_id, CALL)                                    \
  switch (dyn_id) {                                                     \
    case ALMPC_DYN_PENDULUM: { using M = Pendulum; return CALL; }       \
    case ALMPC_DYN_CARTPOLE1L: { using M = Cartpole1L; return CALL; }   \
    case ALMPC_DYN_REX_QUADROTOR: { using M = RexQuad; return CALL; }   \
    case ALMPC_DYN_FLYING_CARTPOLE: { using M = FlyingCartpole; return CALL; } \
    case ALMPC_DYN_CARTPOLE2L: { using M = Cartpole2L; return CALL; }   \
    default: return set_err(-2, "unknown dyn_id");                      \
  }

// state-estimator variant: instantiated in its own translation units (ALMPC_BUILD_SE=1)
int forward_se_f64(const almpc_desc* d, const double* x0, const double* Qd, const double* q, double* xu, double* lam,
                   double* rho, double* factor, float* x_out, float* u_out, double* dyn_res_out, int32_t* status_out,
                   int32_t* info_out, double* debug_out, void* workspace, cudaStream_t st);
int forward_se_f32(const almpc_desc* d, const float* x0, const float* Qd, const float* q, float* xu, float* lam,
                   float* rho, float* factor, float* x_out, float* u_out, float* dyn_res_out, int32_t* status_out,
                   int32_t* info_out, float* debug_out, void* workspace, cudaStream_t st);
#if ALMPC_BUILD_SE
#if ALMPC_BUILD_F64
int forward_se_f64(const almpc_desc* d, const double* x0, const double* Qd, const double* q, double* xu, double* lam,
                   double* rho, double* factor, float* x_out, float* u_out, double* dyn_res_out, int32_t* status_out,
                   int32_t* info_out, double* debug_out, void* workspace, cudaStream_t st) {
  ALMPC_DISPATCH(d->dyn_id, (forward_impl<M, double, 2>(d, x0, Qd, q, xu, lam, rho, factor, x_out, u_out, dyn_res_out,
                                                        status_out, info_out, debug_out, workspace, st)))
}
#endif
#if ALMPC_BUILD_F32
int forward_se_f32(const almpc_desc* d, const float* x0, const float* Qd, const float* q, float* xu, float* lam,
                   float* rho, float* factor, float* x_out, float* u_out, float* dyn_res_out, int32_t* status_out,
                   int32_t* info_out, float* debug_out, void* workspace, cudaStream_t st) {
  ALMPC_DISPATCH(d->dyn_id, (forward_impl<M, float, 2>(d, x0, Qd, q, xu, lam, rho, factor, x_out, u_out, dyn_res_out,
                                                       status_out, info_out, debug_out, workspace, st)))
}
#endif
#else  // !ALMPC_BUILD_SE

template <typename S>
static int forward_any(const almpc_desc* d, const S* x0, const S* Qd, const S* q, S* xu, S* lam, S* rho, S* factor,
                       float* x_out, float* u_out, S* dyn_res_out, int32_t* status_out, int32_t* info_out,
                       S* debug_out, void* workspace, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (d->flags & ALMPC_FLAG_STATE_ESTIMATOR) {
    if (d->obstacles && d->n_obs > 0) return set_err(-8, "the state-estimator variant has no obstacle rows");
    if constexpr (sizeof(S) == 8)
      return forward_se_f64(d, x0, Qd, q, xu, lam, rho, factor, x_out, u_out, dyn_res_out, status_out, info_out,
                            debug_out, workspace, st);
    else
      return forward_se_f32(d, x0, Qd, q, xu, lam, rho, factor, x_out, u_out, dyn_res_out, status_out, info_out,
                            debug_out, workspace, st);
  }
  if (d->obstacles && d->n_obs > 0) {
    // sphere-avoidance rows act on the position x[0..2]; compiled for the one reference env that uses them
    // (FlyingCartpole_obstacles, deqmpc/policies.py:1180)
    if (d->dyn_id != ALMPC_DYN_FLYING_CARTPOLE)
      return set_err(-8, "obstacle constraints are compiled for the flying-cartpole model only");
    return forward_impl<FlyingCartpole, S, 1>(d, x0, Qd, q, xu, lam, rho, factor, x_out, u_out, dyn_res_out,
                                                 status_out, info_out, debug_out, workspace, st);
  }
  ALMPC_DISPATCH(d->dyn_id, (forward_impl<M, S, 0>(d, x0, Qd, q, xu, lam, rho, factor, x_out, u_out, dyn_res_out,
                                                       status_out, info_out, debug_out, workspace, st)))
}

template <typename S>
static int backward_any(const almpc_desc* d, const S* factor, const S* xu, const S* grad_xu, S* grad_Q, S* grad_q,
                        void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  ALMPC_DISPATCH(d->dyn_id, (backward_impl<M::NX, M::NU, S>(d, factor, xu, grad_xu, grad_Q, grad_q, st)))
}

template <typename S>
static int dynamics_any(const almpc_desc* d, int64_t R, const S* x, const S* u, S* xn, S* Jx, S* Ju, void* stream) {
  if (!d) return set_err(-1, "null descriptor");
  if (R < 0) return set_err(-3, "negative row count");
  cudaStream_t st = (cudaStream_t)stream;
  ALMPC_DISPATCH(d->dyn_id, (dynamics_impl<M, S>(d, R, x, u, xn, Jx, Ju, st)))
}


template <typename S>
static int blocktri(int B, int T, int nx, int nu, int do_factor, const S* G, const S* dD, const S* rho, S* factor,
                    const S* rhs, S* sol, int32_t* status, void* stream) {
  if (B < 0 || T < 2 || nx < 1 || nu < 0 || nx + nu > 32) return set_err(-3, "blocktri: need T >= 2, 1 <= nx, nx + nu <= 32");
  if (B == 0) return 0;
  if (!factor || !rhs || !sol || (do_factor && (!G || !dD || !rho))) return set_err(-10, "null buffer");
  const int n = nx + nu, LD = n | 1;
  const size_t smem = (size_t)kWPB * ((size_t)n * LD + 2 * (size_t)nx * LD + (size_t)T * n) * sizeof(S);
  if (smem + 1024 > 227 * 1024) return set_err(-11, "horizon too long for the shared-memory layout");
  auto kern = k_blocktri<S, kWPB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_err(e, "cudaFuncSetAttribute(blocktri)");
  kern<<<n_blocks(B), kWPB * 32, smem, (cudaStream_t)stream>>>(B, T, nx, nu, do_factor, G, dD, rho, factor, rhs, sol, status);
  return cuda_err(cudaGetLastError(), "almpc_blocktri launch");
}

template <typename S>
static int fma_peak(int blocks, int iters, S* sink, double* flops, void* stream) {
  if (blocks <= 0 || iters <= 0 || !sink) return set_err(-1, "bad fma_peak arguments");
  k_fma_peak<S><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, sink);
  if (flops) *flops = 2.0 * 8.0 * (double)iters * 256.0 * (double)blocks;
  return cuda_err(cudaGetLastError(), "almpc_fma_peak launch");
}

template <typename S>
static int tracking_cost(int B, int T, int nx, int nu, const S* Q0, const float* x_ref, const float* u_ref,
                         const float* scale, S* Qd, S* q, void* stream) {
  if (B < 0 || T < 1 || nx < 1 || nu < 0) return set_err(-3, "bad tracking_cost dimensions");
  if (B == 0) return 0;
  if (!Q0 || !x_ref || (nu > 0 && !u_ref) || !Qd || !q) return set_err(-10, "null buffer");
  const long total = (long)B * T * (nx + nu);
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_tracking_cost<S><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(B, T, nx, nu, Q0, x_ref, u_ref, scale, Qd, q);
  return cuda_err(cudaGetLastError(), "almpc_tracking_cost launch");
}

template <typename S>
static int tracking_cost_bwd(int B, int T, int nx, int nu, const S* Q0, const float* x_ref, const float* u_ref,
                             const float* scale, const S* gQ, const S* gq, float* gx, float* gu, float* gs,
                             void* stream) {
  if (B < 0 || T < 1 || nx < 1 || nu < 0) return set_err(-3, "bad tracking_cost dimensions");
  if (B == 0) return 0;
  if (!Q0 || !x_ref || (nu > 0 && !u_ref) || !gq) return set_err(-10, "null buffer");
  const long bt = (long)B * T;
  k_tracking_cost_bwd<S><<<(unsigned)((bt + 127) / 128), 128, 0, (cudaStream_t)stream>>>(B, T, nx, nu, Q0, x_ref, u_ref,
                                                                                     scale, gQ, gq, gx, gu, gs);
  return cuda_err(cudaGetLastError(), "almpc_tracking_cost_bwd launch");
}

#endif  // !ALMPC_BUILD_SE

}  // namespace almpc

#if !ALMPC_BUILD_SE
// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace almpc;

extern "C" {

#if ALMPC_BUILD_F64
int almpc_model_dims(int32_t dyn_id, int32_t* nx, int32_t* nu) {
  int x, u;
  switch (dyn_id) {
    case ALMPC_DYN_PENDULUM: x = Pendulum::NX; u = Pendulum::NU; break;
    case ALMPC_DYN_CARTPOLE1L: x = Cartpole1L::NX; u = Cartpole1L::NU; break;
    case ALMPC_DYN_REX_QUADROTOR: x = RexQuad::NX; u = RexQuad::NU; break;
    case ALMPC_DYN_FLYING_CARTPOLE: x = FlyingCartpole::NX; u = FlyingCartpole::NU; break;
    case ALMPC_DYN_CARTPOLE2L: x = Cartpole2L::NX; u = Cartpole2L::NU; break;
    default: return set_err(-2, "unknown dyn_id");
  }
  if (nx) *nx = x;
  if (nu) *nu = u;
  return 0;
}

size_t almpc_factor_elems(const almpc_desc* d) {
  int32_t nx, nu;
  if (!d || almpc_model_dims(d->dyn_id, &nx, &nu)) return 0;
  const size_t n = (size_t)nx + nu;
  return (size_t)d->T * n * n + (size_t)(d->T - 1) * nx * n;
}

size_t almpc_workspace_bytes(const almpc_desc* d) {
  int32_t nx, nu;
  if (!d || d->T < 2 || almpc_model_dims(d->dyn_id, &nx, &nu)) return 0;
  return ws_plan(d->B, d->T, nx, nu, sizeof(double)).total;
}

size_t almpc_debug_elems(const almpc_desc* d) {
  int32_t nx, nu;
  if (!d || almpc_model_dims(d->dyn_id, &nx, &nu)) return 0;
  return 2 * (size_t)d->T * (nx + nu) + 24;
}
#endif  // common entry points (fp64 unit)

#if ALMPC_BUILD_F64
int almpc_forward_f64(const almpc_desc* d, const double* x0, const double* Qd, const double* q, double* xu,
                      double* lam, double* rho, double* factor, float* x_out, float* u_out, double* dyn_res_out,
                      int32_t* status_out, int32_t* info_out, double* debug_out, void* workspace, void* stream) {
  return forward_any<double>(d, x0, Qd, q, xu, lam, rho, factor, x_out, u_out, dyn_res_out, status_out, info_out,
                             debug_out, workspace, stream);
}
#endif
#if ALMPC_BUILD_F32
int almpc_forward_f32(const almpc_desc* d, const float* x0, const float* Qd, const float* q, float* xu, float* lam,
                      float* rho, float* factor, float* x_out, float* u_out, float* dyn_res_out, int32_t* status_out,
                      int32_t* info_out, float* debug_out, void* workspace, void* stream) {
  return forward_any<float>(d, x0, Qd, q, xu, lam, rho, factor, x_out, u_out, dyn_res_out, status_out, info_out,
                            debug_out, workspace, stream);
}
#endif
#if ALMPC_BUILD_F64
int almpc_backward_f64(const almpc_desc* d, const double* factor, const double* xu, const double* grad_xu,
                       double* grad_Q, double* grad_q, void* stream) {
  return backward_any<double>(d, factor, xu, grad_xu, grad_Q, grad_q, stream);
}
#endif
#if ALMPC_BUILD_F32
int almpc_backward_f32(const almpc_desc* d, const float* factor, const float* xu, const float* grad_xu,
                       float* grad_Q, float* grad_q, void* stream) {
  return backward_any<float>(d, factor, xu, grad_xu, grad_Q, grad_q, stream);
}
#endif
#if ALMPC_BUILD_F64
int almpc_dynamics_f64(const almpc_desc* d, int64_t R, const double* x, const double* u, double* xn, double* Jx,
                       double* Ju, void* stream) {
  return dynamics_any<double>(d, R, x, u, xn, Jx, Ju, stream);
}
#endif
#if ALMPC_BUILD_F32
int almpc_dynamics_f32(const almpc_desc* d, int64_t R, const float* x, const float* u, float* xn, float* Jx,
                       float* Ju, void* stream) {
  return dynamics_any<float>(d, R, x, u, xn, Jx, Ju, stream);
}
#endif
#if ALMPC_BUILD_F64
int almpc_fma_peak_f64(int32_t blocks, int32_t iters, double* sink, double* flops, void* stream) {
  return fma_peak<double>(blocks, iters, sink, flops, stream);
}
#endif
#if ALMPC_BUILD_F32
int almpc_fma_peak_f32(int32_t blocks, int32_t iters, float* sink, double* flops, void* stream) {
  return fma_peak<float>(blocks, iters, sink, flops, stream);
}
#endif
#if ALMPC_BUILD_F64
int almpc_tracking_cost_f64(int32_t B, int32_t T, int32_t nx, int32_t nu, const double* Q0, const float* x_ref,
                            const float* u_ref, const float* scale, double* Qd, double* q, void* stream) {
  return tracking_cost<double>(B, T, nx, nu, Q0, x_ref, u_ref, scale, Qd, q, stream);
}
#endif
#if ALMPC_BUILD_F32
int almpc_tracking_cost_f32(int32_t B, int32_t T, int32_t nx, int32_t nu, const float* Q0, const float* x_ref,
                            const float* u_ref, const float* scale, float* Qd, float* q, void* stream) {
  return tracking_cost<float>(B, T, nx, nu, Q0, x_ref, u_ref, scale, Qd, q, stream);
}
#endif
#if ALMPC_BUILD_F64
int almpc_tracking_cost_bwd_f64(int32_t B, int32_t T, int32_t nx, int32_t nu, const double* Q0, const float* x_ref,
                                const float* u_ref, const float* scale, const double* grad_Q, const double* grad_q,
                                float* grad_x_ref, float* grad_u_ref, float* grad_scale, void* stream) {
  return tracking_cost_bwd<double>(B, T, nx, nu, Q0, x_ref, u_ref, scale, grad_Q, grad_q, grad_x_ref, grad_u_ref,
                                   grad_scale, stream);
}
#endif
#if ALMPC_BUILD_F32
int almpc_tracking_cost_bwd_f32(int32_t B, int32_t T, int32_t nx, int32_t nu, const float* Q0, const float* x_ref,
                                const float* u_ref, const float* scale, const float* grad_Q, const float* grad_q,
                                float* grad_x_ref, float* grad_u_ref, float* grad_scale, void* stream) {
  return tracking_cost_bwd<float>(B, T, nx, nu, Q0, x_ref, u_ref, scale, grad_Q, grad_q, grad_x_ref, grad_u_ref,
                                  grad_scale, stream);
}
#endif
#if ALMPC_BUILD_F64
int almpc_blocktri_f64(int32_t B, int32_t T, int32_t nx, int32_t nu, int32_t do_factor, const double* G,
                       const double* diagD, const double* rho, double* factor, const double* rhs, double* sol,
                       int32_t* status, void* stream) {
  return blocktri<double>(B, T, nx, nu, do_factor, G, diagD, rho, factor, rhs, sol, status, stream);
}
#endif
#if ALMPC_BUILD_F32
int almpc_blocktri_f32(int32_t B, int32_t T, int32_t nx, int32_t nu, int32_t do_factor, const float* G,
                       const float* diagD, const float* rho, float* factor, const float* rhs, float* sol,
                       int32_t* status, void* stream) {
  return blocktri<float>(B, T, nx, nu, do_factor, G, diagD, rho, factor, rhs, sol, status, stream);
}
#endif
#if ALMPC_BUILD_F64
void almpc_set_profiling(int32_t on) { g_profile_on = on; }
void almpc_get_profile(double* out8) {
  for (int k = 0; k < 8; ++k) out8[k] = g_profile[k];
}
const char* almpc_last_error(void) { return g_err; }
const char* almpc_version(void) { return "almpc 0.1 (sm_100a)"; }

#endif
}  // extern "C"
#endif  // !ALMPC_BUILD_SE
