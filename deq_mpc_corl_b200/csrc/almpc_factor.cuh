// almpc_factor.cuh — block-tridiagonal factor / solve kernels (k_factor_solve, k_backward, k_blocktri): part of libalmpc (see almpc_kernels.cu for the overview).
#pragma once
#include "almpc_common.cuh"

namespace almpc {

// ------------------------------------------------------------------------------------------------
// right-looking elimination on "one row per lane".
//   lanes 0..NN-1 : rows of the SPD block (lower part in a[0..lane])         -> rows of L (diag = 1/L_ii)
//   lanes >= NN   : extra rows b (rhs, sub-diagonal block rows)               -> b L^{-T}
// Factor and solve in one pass.
// ------------------------------------------------------------------------------------------------
// The pivot column travels through shared memory (one STS per lane, then 16-byte broadcast loads) instead
// of one shuffle pair per entry: the first version was bound by the SM's shuffle throughput.
//   colbuf : 2 x 32 scalars (double-buffered pivot column)
template <int NN, typename S>
__device__ __forceinline__ void eliminate(S (&a)[NN], int lane, S* colbuf) {
  {
    // software-pipelined pivots: as soon as column k+1 has its update, the next pivot is broadcast and its
    // reciprocal square root (the longest dependent chain of the step) runs under the remaining updates
    // a non-positive pivot turns into NaN / Inf here and poisons the whole solution of the problem, which is
    // what the caller tests (cheaper than a comparison per pivot)
    S d = __shfl_sync(FULL, a[0], 0);
    S inv = rsqrt_(d);
#pragma unroll
    for (int k = 0; k < NN; ++k) {
      const S lk = a[k] * inv;
      a[k] = (lane == k) ? inv : lk;
      if (k + 1 < NN) {
        // the next pivot lives in lane k+1 and needs only that lane's own multiplier (col[k+1] == its lk): it is
        // broadcast from registers, so the shared-memory round trip of the column is off the critical pivot chain
        d = __shfl_sync(FULL, a[k + 1] - lk * lk, k + 1);
        inv = rsqrt_(d);
        S* col = colbuf + (k & 1) * 32;
        if (lane < NN) col[lane] = lk;  // only rows k+1..NN-1 of the pivot column are read
        __syncwarp();
        a[k + 1] -= lk * col[k + 1];
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
  }
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
  if constexpr (sizeof(S) == 8 || ALMPC_F32_DMMA) {
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
  __shared__ int s_ring;
  grid_dep_wait();
  const int mode = ctl->mode;
  if (mode == MODE_DONE) return;
  if (threadIdx.x == 0) s_ring = ring_claim(P, ctl);
  __syncthreads();
  const int ring_slot = s_ring;
  const int T = P.T;
  const int n_obs = OBS ? P.n_obs : 0;
  const Layout<NX, NU> L(T, n_obs, SE);
  const int SI = OBS ? L.SI : 2 * NU;  // compile-time constant without obstacle rows
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int prob = blockIdx.x * WPB + warp;
  if (prob >= P.B) {  // (warp 0 always has a problem, so the release below is reached)
    return;
  }
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
  S* fac = (ring_slot >= 0) ? P.ring + ((size_t)ring_slot * 4 + warp) * factor_elems<NX, NU>(T)
                            : fac_g + (size_t)prob * factor_elems<NX, NU>(T);
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
          if (in_cost) cost += cost_term(z, Qv[u], qq[u]);
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
    eliminate<NN, S>(a, lane, scr);
    __syncwarp();
    {
      // write-outs with per-lane destinations: L rows -> global factor; rhs -> dv (y_t); E rows -> X + global factor
      const int ei = lane - NN - 1;
      const bool e_wr = (ei >= 0) && (ei < NX) && has_G;  // ei < EX0 by construction
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
    }
    if (NX > EX0 && has_G) {
      // Sub-diagonal rows that did not fit next to the L rows (flying cartpole: nx = 14, n = 18 -> one row):
      // x = (-rho G_t[i,:]) L^{-T} by a forward substitution ACROSS the lanes that still hold the rows of L in
      // registers: lane j owns element j; after x_k is known (lane k: b_k / L_kk) every lane j > k subtracts
      // L[j][k] x_k with its own a[k].  Registers and shuffles only — the first version ran a second full
      // elimination pass over shared memory for this one row.  Rows EX0.. of X still hold G_t (the rows written
      // above are 0..EX0-1).
      for (int i = EX0; i < NX; ++i) {
        S bj = (lane < NN) ? -rho * X[i * LB + lane] : S(0);
        S xres = S(0);
#pragma unroll
        for (int k = 0; k < NN; ++k) {
          const S xk = __shfl_sync(FULL, bj * a[k], k);  // lane k: a[k] = 1 / L_kk
          if (lane == k) xres = xk;
          else if (lane > k && lane < NN) bj -= xk * a[k];  // a[k] = L[lane][k]
        }
        if (lane < NN) {
          X[i * LB + lane] = xres;
          Eg[i * NN + lane] = xres;
        }
      }
    }
    __syncwarp();
    S* tmp = X; X = Y; Y = tmp;  // Y := E_t, X := G_{t+1}
  }
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
  bool ok = true;
  for (int i = lane; i < L.N; i += 32) ok = ok && finite_(dv[i]);
  ok = __all_sync(FULL, ok);
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
  if (ring_slot >= 0) {  // every warp of the CTA is done with the scratch factors
    __syncthreads();
    if (threadIdx.x == 0) ring_release(P, ring_slot);
  }
}

// ------------------------------------------------------------------------------------------------
// kernel 2/3, "pair" mapping for n <= 16 and nx + 1 <= 16 (pendulum, cartpoles, quadrotor):
//   1 HALF-warp = 1 problem, 2 problems per warp, TWO matrix rows per lane:
//     a[] = row hl of the SPD block (-> row of L)            b[] = rhs (hl = 0) / row hl-1 of -rho G_t (-> E_t)
// Everything that is per-pivot overhead in the row-per-lane kernel above (pivot broadcast through shared memory,
// the rsqrt chain, the diagonal shuffle, synchronisation) is issued once for two problems; the trailing updates
// keep the same DFMA count per problem.  Measured on the elimination alone (tools/exp/elim_pair.cu): 1.5x.
// Same inputs, outputs, factor layout and arithmetic per problem as k_factor_solve.
// ------------------------------------------------------------------------------------------------
template <int NN, typename S>
__device__ __forceinline__ void eliminate_pair(S (&a)[NN], S (&b)[NN], int hl, int half, S* colbuf) {
  // colbuf: this problem's 2 x 16 scalars (double-buffered pivot column)
  S d = __shfl_sync(FULL, a[0], half * 16);
  S inv = rsqrt_(d);
#pragma unroll
  for (int k = 0; k < NN; ++k) {
    const S lk = a[k] * inv, mk = b[k] * inv;
    a[k] = (hl == k) ? inv : lk;
    b[k] = mk;
    if (k + 1 < NN) {
      // next pivot straight from lane k+1's registers (col[k+1] is that lane's own lk): the shared-memory
      // round trip of the column stays off the critical pivot chain
      d = __shfl_sync(FULL, a[k + 1] - lk * lk, half * 16 + k + 1);
      inv = rsqrt_(d);
      S* col = colbuf + (k & 1) * 16;
      col[hl] = lk;
      __syncwarp();
      const S c1 = col[k + 1];
      a[k + 1] -= lk * c1;
      b[k + 1] -= mk * c1;
      const int j0 = k + 2;
      const int js = (j0 + 1) & ~1;
      if ((j0 & 1) && j0 < NN) {
        const S c = col[j0];
        a[j0] -= lk * c;
        b[j0] -= mk * c;
      }
#pragma unroll
      for (int j = 0; j < NN; j += 2) {
        if (j >= js && j + 1 < NN) {
          const Vec2<S> v = *reinterpret_cast<const Vec2<S>*>(col + j);
          a[j] -= lk * v.x;
          a[j + 1] -= lk * v.y;
          b[j] -= mk * v.x;
          b[j + 1] -= mk * v.y;
        }
      }
      if ((NN & 1) && (NN - 1) >= js) {
        const S c = col[NN - 1];
        a[NN - 1] -= lk * c;
        b[NN - 1] -= mk * c;
      }
    }
  }
}

template <typename S>
__device__ __forceinline__ S half_sum(S v) {  // sum over the 16 lanes of a half-warp
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

template <int ROWS, int NN, int LB, typename S>
__device__ __forceinline__ void half_copy_block_async(S* dst, const S* src, int hl) {
  for (int e = hl; e < ROWS * NN; e += 16) {
    const int r = e / NN, c = e - r * NN;
    if constexpr (sizeof(S) == 8) cp_async8(dst + r * LB + c, src + e);
    else cp_async4(dst + r * LB + c, src + e);
  }
}

// What the factor phase of one half-warp hands to the caller (the fused Newton-tail kernel continues with the line
// search on the same problems without a trip through global memory).
template <typename S>
struct PairCarry {
  S rho;        // penalty used by this Newton step (after the dual update of MODE_DUAL_NEWTON)
  S s2_entry;   // ||c+||^2 at the entry point
  S merit;      // merit at the entry point (valid when mode != MODE_CONT)
  int status;   // per-problem status bits after this factorisation
  int prob;     // problem of this half-warp (the neighbour's when !live)
  bool ok;      // every pivot positive / direction finite
  bool live;    // false: odd batch, this half-warp duplicates its neighbour and must store nothing
};

// The factor phase proper: one HALF-warp = one problem.  Returns false for a warp without problems.
//   FUSED = false: stand-alone kernel k_factor_solve_pair (direction, merit, entry residual and status go to global
//                  memory for k_line_search)
//   FUSED = true : called by k_newton_pair (almpc_fused.cuh); the direction stays in shared memory (dv, sign
//                  flipped by the caller), merit / entry residual / status stay in registers (PairCarry)
template <int NX, int NU, typename S, int WPB, int V, bool FUSED>
__device__ __forceinline__ bool factor_pair(const KParams<S>& P, const S* __restrict__ Qd_g, const S* __restrict__ q_g,
                                            const S* xu_g, S* lam_g, S* __restrict__ rho_g,
                                            S* fac_g, const S* __restrict__ G_g, const S* __restrict__ c_g,
                                            S* __restrict__ d_g, S* __restrict__ merit_g,
                                            double* __restrict__ rentry_g, int* __restrict__ status_g,
                                            S* __restrict__ debug_g, const int mode, unsigned char* smem_raw,
                                            const int smem_per_problem, const int pw, S* ring_cta, PairCarry<S>& out) {
  constexpr int NN = NX + NU;
  static_assert(NN <= 16 && NX + 1 <= 16, "pair mapping: both row sets must fit a half-warp");
  constexpr int LD = ScrLd<NN>::LD;
  constexpr int LB = NN | 1;
  constexpr bool SE = (V == 2);
  static_assert(V != 1, "obstacle rows are compiled for the row-per-lane kernel only");
  const int T = P.T;
  const Layout<NX, NU> L(T, 0, SE, true);
  constexpr int SI = 2 * NU;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
  if (pw >= P.B) return false;  // pw = first problem of this warp
  // an odd batch leaves the last half-warp without a problem: it recomputes its neighbour's and stores nothing
  const bool live = pw + half < P.B;
  const int prob = live ? pw + half : pw;
  // smem_per_problem >= L.total (the fused kernel may need more room for its line-search staging area)
  S* smw = reinterpret_cast<S*>(smem_raw) + (size_t)warp * 2 * smem_per_problem;
  S* sm = smw + (size_t)half * smem_per_problem;
  S* xu = sm + L.xu;
  S* dv = sm + L.dv;
  S* wv = sm + L.w;
  S* lin = sm + L.lin;
  S* qds = sm + L.qd;
  S* Gt = sm + L.Gt;
  S* Ep = sm + L.Ep;
  S* scr = sm + L.scr;
  const size_t pN = (size_t)prob * L.N;
  const S* Qd = Qd_g + pN;
  const S* qv = q_g + pN;
  S* lam = lam_g + (size_t)prob * L.m;
  const S* cres = c_g + (size_t)prob * L.neq;
  const S* Gp = G_g + (size_t)prob * (T - 1) * NX * NN;
  // ring_cta: this CTA's scratch slot (4 factors) when nobody reads the factor after this launch, else NULL
  S* fac = ring_cta ? ring_cta + (size_t)(warp * 2 + (live ? half : 0)) * factor_elems<NX, NU>(T)
                    : fac_g + (size_t)prob * factor_elems<NX, NU>(T);
  S* dbg = (debug_g && live) ? debug_g + (size_t)prob * (2 * L.N + 24) : nullptr;

  if (T > 1) half_copy_block_async<NX, NN, LB>(Gt, Gp, hl);
  cp_async_commit();
  S rho = rho_g[prob];
  // first round (64 entries per half-warp) of every input array is issued before anything is consumed: one HBM
  // round trip instead of one per array
  S p_c[4], p_l[4], p_li[4], p_Q[4], p_q[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = hl + 16 * u;
    p_c[u] = (i < L.neq) ? cres[i] : S(0);
    p_l[u] = (i < L.neq) ? lam[i] : S(0);
    p_li[u] = (i < L.nin) ? lam[L.neq + i] : S(0);
    p_Q[u] = (i < L.N) ? Qd[i] : S(0);
    p_q[u] = (i < L.N) ? qv[i] : S(0);
  }
  for (int base = 0; base < L.N; base += 64) {
    S xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + hl + 16 * u;
      xv[u] = (i < L.N) ? xu_g[pN + i] : S(0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + hl + 16 * u;
      if (i < L.N) xu[i] = xv[u];
    }
  }
  __syncwarp();

  // ---- dual / penalty update of the previous AL iteration, entry residual, merit, w = lam_eq + rho c, cost gradient
  const bool dual = (mode == MODE_DUAL_NEWTON);
  const S rho_old = rho;
  if (dual) rho = next_rho(rho, P.rho_mult, P.rho_f32);
  {
    S s2 = S(0), sl = S(0), cost = S(0);
    for (int base = 0; base < L.neq; base += 64) {
      S cv[4], lv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + hl + 16 * u;
        cv[u] = (base == 0) ? p_c[u] : ((i < L.neq) ? cres[i] : S(0));
        lv[u] = (base == 0) ? p_l[u] : ((i < L.neq) ? lam[i] : S(0));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + hl + 16 * u;
        if (i < L.neq) {
          S li = lv[u];
          if (dual) { li += rho_old * cv[u]; if (live) lam[i] = li; }
          s2 += cv[u] * cv[u];
          sl += li * cv[u];
          wv[i] = li + rho * cv[u];
        }
      }
    }
    for (int base = 0; base < L.nin; base += 64) {
      S lv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + hl + 16 * u;
        lv[u] = (base == 0) ? p_li[u] : ((i < L.nin) ? lam[L.neq + i] : S(0));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + hl + 16 * u;
        if (i < L.nin) {
          const int t = i / SI, sidx = i - t * SI;
          const S ci = ineq_res<NX, NU, S>(xu + t * NN, sidx, P.ulo, P.uhi, nullptr, S(0));
          S li = lv[u];
          if (dual) { li = fmax0(li + rho_old * ci); if (live) lam[L.neq + i] = li; }
          lin[i] = li;
          const S cp = fmax0(ci);
          s2 += cp * cp;
          sl += li * ci;
        }
      }
    }
    for (int base = 0; base < L.N; base += 64) {
      S Qv[4], qq[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + hl + 16 * u;
        Qv[u] = (base == 0) ? p_Q[u] : ((i < L.N) ? Qd[i] : S(0));
        qq[u] = (base == 0) ? p_q[u] : ((i < L.N) ? qv[i] : S(0));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + hl + 16 * u;
        if (i < L.N) {
          const S z = xu[i];
          const bool in_cost = !SE || (i % NN) < NX;
          if (in_cost) cost += cost_term(z, Qv[u], qq[u]);
          dv[i] = in_cost ? Qv[u] * z + qq[u] : S(0);
          qds[i] = Qv[u];
        }
      }
    }
    s2 = half_sum(s2);
    sl = half_sum(sl);
    cost = half_sum(cost);
    out.s2_entry = s2;
    out.merit = cost + S(0.5) * rho * s2 + sl;
    if (hl == 0 && live) {
      if (dual) rho_g[prob] = rho;
      if constexpr (!FUSED) {
        rentry_g[prob] = (double)s2;
        if (mode != MODE_CONT) merit_g[prob] = out.merit;
      }
    }
  }
  __syncwarp();

  // ---- forward sweep
  S* X = Gt;
  S* Y = Ep;
  const S rho_p[2] = {__shfl_sync(FULL, rho, 0), __shfl_sync(FULL, rho, 16)};
  for (int t = 0; t < T; ++t) {
    S* Lg = fac + (size_t)t * NN * NN;
    S* Eg = fac + (size_t)T * NN * NN + (size_t)t * NX * NN;
    const bool has_G = t < T - 1, has_E = t > 0;
    cp_async_wait_all();
    __syncwarp();
    if (hl < NN) {
      const int j = hl, idx = t * NN + j;
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
        else if (!SE) g += wv[(T - 1) * NX + j];
      } else if (!SE) {
        const int sidx = j - NX;
        const S u = xu[idx];
        const S cu = u - P.uhi[sidx], cl = P.ulo[sidx] - u;
        const S lu = lin[t * SI + sidx], ll = lin[t * SI + NU + sidx];
        g += (lu - ll) + rho * (fmax0(cu) - fmax0(cl));
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
    // Hessian blocks of both problems: the DMMA assembly is warp-wide, one problem after the other
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      S* smp = smw + (size_t)p * smem_per_problem;
      S* Xp = (X == Gt) ? smp + L.Gt : smp + L.Ep;
      S* Yp = (X == Gt) ? smp + L.Ep : smp + L.Gt;
      build_block<NX, NU, S>(smp + L.scr + 64, smp + L.xu + t * NN, Xp, Yp, has_G, has_E, lane, rho_p[p],
                             smp + L.qd + t * NN, P.ulo, P.uhi, !SE || has_E, !SE);
    }
    __syncwarp();
    S a[NN], b[NN];
    {
      const int ei = hl - 1;
      const bool e_act = (ei >= 0) && (ei < NX) && has_G;
      const S* arow = scr + 64 + (hl < NN ? hl : 0) * LD;
      const S* brow = (hl == 0) ? (dv + t * NN) : (X + (e_act ? ei : 0) * LB);
      const S bscale = (hl == 0) ? S(1) : (e_act ? -rho : S(0));
#pragma unroll
      for (int c = 0; c < NN; ++c) {
        a[c] = arow[c];
        b[c] = brow[c] * bscale;
      }
      if (NN < 16 && hl >= NN) {  // lanes beyond the block: a harmless identity row
#pragma unroll
        for (int c = 0; c < NN; ++c) a[c] = S(0);
      }
    }
    __syncwarp();
    if (t + 1 < T - 1) half_copy_block_async<NX, NN, LB>(Y, Gp + (size_t)(t + 1) * NX * NN, hl);
    cp_async_commit();
    eliminate_pair<NN, S>(a, b, hl, half, scr);
    __syncwarp();
    {
      const int ei = hl - 1;
      const bool e_wr = (ei >= 0) && (ei < NX) && has_G;
      if (hl < NN && live) {
        S* gdst = Lg + hl * NN;
        if constexpr (NN % 2 == 0) {
#pragma unroll
          for (int c = 0; c < NN; c += 2) {
            Vec2<S> v;
            v.x = (c <= hl) ? a[c] : S(0);
            v.y = (c + 1 <= hl) ? a[c + 1] : S(0);
            *reinterpret_cast<Vec2<S>*>(gdst + c) = v;
          }
        } else {
#pragma unroll
          for (int c = 0; c < NN; ++c) gdst[c] = (c <= hl) ? a[c] : S(0);
        }
      }
      S* sdst = (hl == 0) ? (dv + t * NN) : (e_wr ? (X + ei * LB) : nullptr);
      if (sdst) {
#pragma unroll
        for (int c = 0; c < NN; ++c) sdst[c] = b[c];
      }
      if (e_wr && live) {
        S* gdst = Eg + ei * NN;
        if constexpr (NN % 2 == 0) {
#pragma unroll
          for (int c = 0; c < NN; c += 2) {
            Vec2<S> v;
            v.x = b[c];
            v.y = b[c + 1];
            *reinterpret_cast<Vec2<S>*>(gdst + c) = v;
          }
        } else {
#pragma unroll
          for (int c = 0; c < NN; ++c) gdst[c] = b[c];
        }
      }
    }
    __syncwarp();
    S* tmp = X; X = Y; Y = tmp;
  }
  // ---- backward sweep, d = -H^{-1} g: L_tt / E_t come back from the streamed factor through cp.async.
  // A dead half-warp (odd batch) stored no factor: it reads its neighbour's, which is the same problem.
  cp_async_wait_all();
  __syncwarp();
  S* Lsm = scr + 64;
  S* Ecur = X;
  S* Enext = Y;
  constexpr int CW = CopyWidth<NX, NU, S>::W;
  // Optional (ALMPC_BULK_COPY, measured slower — almpc_common.cuh): blocks whose byte sizes are multiples of 16
  // (quadrotor: 2048 / 1536 B) come back by ONE bulk copy each (cp.async.bulk + mbarrier, issued by lane 0 of the
  // half-warp) instead of 8 + 12 per-lane cp.async.  The E block is then staged unpadded: the sweep reads it
  // column-per-lane, which is conflict-free at any row stride.
  constexpr bool BULK = (CW == 16) && ALMPC_BULK_COPY;
  constexpr int ES = BULK ? NN : LB;  // row stride of the staged E block
  __shared__ unsigned long long s_mbar[WPB * 2];
  unsigned long long* bar = &s_mbar[warp * 2 + half];
  unsigned phase = 0;
  if constexpr (BULK) {
    // every lane: its factor rows (ordinary global stores of the forward sweep) and its last shared-memory accesses
    // are ordered before the async-proxy copies that lane 0 issues after the warp barrier
    fence_proxy_async();
    if (hl == 0) mbar_init(bar, 1);
    fence_mbar_init();
    __syncwarp();
  }
  auto copy_L = [&](const S* src) {
    constexpr int CHUNKS = NN * NN * (int)sizeof(S) / CW;
    for (int ch = hl; ch < CHUNKS; ch += 16) {
      if constexpr (CW == 16) cp_async16((char*)Lsm + ch * 16, (const char*)src + ch * 16);
      else if constexpr (CW == 8) cp_async8((char*)Lsm + ch * 8, (const char*)src + ch * 8);
      else cp_async4((char*)Lsm + ch * 4, (const char*)src + ch * 4);
    }
  };
  // L_t -> Lsm and (with_E) E_t -> Enext
  auto fetch = [&](int t, bool with_E) {
    const S* srcL = fac + (size_t)t * NN * NN;
    const S* srcE = fac + (size_t)T * NN * NN + (size_t)t * NX * NN;
    if constexpr (BULK) {
      if (hl == 0) {
        constexpr unsigned bL = NN * NN * (unsigned)sizeof(S), bE = NX * NN * (unsigned)sizeof(S);
        mbar_expect_tx(bar, with_E ? bL + bE : bL);
        bulk_g2s(Lsm, srcL, bL, bar);
        if (with_E) bulk_g2s(Enext, srcE, bE, bar);
      }
    } else {
      copy_L(srcL);
      if (with_E) half_copy_block_async<NX, NN, LB>(Enext, srcE, hl);
    }
  };
  auto fetch_wait = [&]() {
    if constexpr (BULK) {
      mbar_wait(bar, phase);
      phase ^= 1u;
    } else {
      cp_async_wait_all();
    }
    __syncwarp();
  };
  fetch(T - 1, false);
  cp_async_commit();
  for (int t = T - 1; t >= 0; --t) {
    fetch_wait();
    S l[NN];
    if (hl < NN) {
#pragma unroll
      for (int r = 0; r < NN; ++r) l[r] = Lsm[r * NN + hl];
    } else {
#pragma unroll
      for (int r = 0; r < NN; ++r) l[r] = S(0);
    }
    S bb = (hl < NN) ? dv[t * NN + hl] : S(0);
    if (t < T - 1 && hl < NN) {
      S s0 = S(0), s1 = S(0);
#pragma unroll
      for (int i = 0; i + 1 < NX; i += 2) {
        s0 += Ecur[i * ES + hl] * dv[(t + 1) * NN + i];
        s1 += Ecur[(i + 1) * ES + hl] * dv[(t + 1) * NN + i + 1];
      }
      if (NX & 1) s0 += Ecur[(NX - 1) * ES + hl] * dv[(t + 1) * NN + NX - 1];
      bb -= s0 + s1;
    }
    __syncwarp();
    if (t > 0) fetch(t - 1, true);
    cp_async_commit();
#pragma unroll
    for (int r = NN - 1; r >= 0; --r) {
      const S dr = __shfl_sync(FULL, bb * l[r], half * 16 + r);
      bb = (hl == r) ? dr : (bb - l[r] * dr);
    }
    if (hl < NN) dv[t * NN + hl] = bb;
    S* tmp = Ecur; Ecur = Enext; Enext = tmp;
  }
  __syncwarp();
  bool okl = true;
  for (int i = hl; i < L.N; i += 16) okl = okl && finite_(dv[i]);
  const unsigned okm = __ballot_sync(FULL, okl);
  const bool ok = ((okm >> (half * 16)) & 0xffffu) == 0xffffu;
  int status = ALMPC_STATUS_FINITE | ALMPC_STATUS_CHOL_OK;
  if (mode != MODE_FIRST && hl == 0) status = status_g[prob];
  if (!ok) status &= ~ALMPC_STATUS_CHOL_OK;
  if (live) {
    if (!FUSED || dbg) {
      for (int i = hl; i < L.N; i += 16) {
        const S d = ok ? -dv[i] : S(0);
        if constexpr (!FUSED) d_g[pN + i] = d;
        if (dbg) dbg[L.N + i] = d;
      }
    }
    if (!FUSED && hl == 0) status_g[prob] = status;
  }
  out.rho = rho;
  out.status = status;  // meaningful on lane hl == 0 of each half-warp
  out.prob = prob;
  out.ok = ok;
  out.live = live;
  return true;
}

// small blocks (and fp32) keep little in shared memory, so their occupancy is set by registers: cap them for
// 12 (10) CTAs / SM; the fp64 quadrotor is limited to 6 by shared memory anyway and may use up to 170 registers
template <int NX, int NU, typename S, int WPB, int V>
__global__ void __launch_bounds__(WPB * 32, (NX + NU <= 8) ? 12 : (sizeof(S) == 4 ? 10 : 6))
k_factor_solve_pair(const KParams<S> P, const S* __restrict__ Qd_g, const S* __restrict__ q_g,
                    const S* __restrict__ xu_g, S* __restrict__ lam_g, S* __restrict__ rho_g, S* fac_g,
                    const S* __restrict__ G_g, const S* __restrict__ c_g, S* __restrict__ d_g,
                    S* __restrict__ merit_g, double* __restrict__ rentry_g, int* __restrict__ status_g,
                    S* __restrict__ debug_g, const Ctl* __restrict__ ctl) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_ring;
  static_assert(WPB * 2 <= 4, "a scratch slot holds 4 factors");
  grid_dep_wait();
  const int mode = ctl->mode;
  if (mode == MODE_DONE) return;
  if (threadIdx.x == 0) s_ring = ring_claim(P, ctl);
  __syncthreads();
  const int ring_slot = s_ring;
  PairCarry<S> carry;
  const Layout<NX, NU> L(P.T, 0, V == 2, true);
  S* ring_cta = (ring_slot >= 0) ? P.ring + (size_t)ring_slot * 4 * factor_elems<NX, NU>(P.T) : nullptr;
  factor_pair<NX, NU, S, WPB, V, false>(P, Qd_g, q_g, xu_g, lam_g, rho_g, fac_g, G_g, c_g, d_g, merit_g, rentry_g, status_g,
                                        debug_g, mode, smem_raw, L.total, (int)(blockIdx.x * WPB + (threadIdx.x >> 5)) * 2,
                                        ring_cta, carry);
  if (ring_slot >= 0) {
    __syncthreads();
    if (threadIdx.x == 0) ring_release(P, ring_slot);
  }
}

// ------------------------------------------------------------------------------------------------
// implicit backward: gbar = -H^{-1} grad ; grad_q = gbar ; grad_Q = gbar * xu     al_utils.py:578-615
// ------------------------------------------------------------------------------------------------
template <int NX, int NU, typename S, int WPB>
__global__ void __launch_bounds__(WPB * 32)
k_backward(int B, int T, const S* __restrict__ fac_g, const S* __restrict__ xu_g, const S* __restrict__ grad_g,
           const int* __restrict__ status_g, S* __restrict__ gQ_g, S* __restrict__ gq_g) {
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
  // A factor with non-positive pivots (flagged in the forward status) yields no gradient instead of NaNs.  With the
  // status the zeroing is limited to exactly those problems, so a non-finite UPSTREAM gradient propagates like in
  // the reference (al_utils.py:585-597) instead of turning into an exact zero.
  bool fin = true;
  if (status_g) {
    fin = (status_g[prob] & ALMPC_STATUS_CHOL_OK) != 0;
  } else {
    for (int i = lane; i < N; i += 32) fin = fin && finite_(dv[i]);
    fin = __all_sync(FULL, fin);
  }
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


}  // namespace almpc
