// almpc_fused.cuh — k_newton_pair: the tail of one Newton step in ONE kernel (part of libalmpc, see almpc_kernels.cu).
//
//   factor phase   (factor_pair, almpc_factor.cuh)   entry residual, dual / penalty update of the previous AL iteration
//                                                    (AL_mpc.py:315-325), merit, gradient, block-tridiagonal Cholesky,
//                                                    both sweeps -> Newton direction             al_utils.py:80-123,510-515
//   line search    (ls_item, almpc_linesearch.cuh)   20 candidates 2^-k, first argmin, strict-decrease accept, iterate
//                                                    update                                      al_utils.py:618-642
//   steering       (steer_solve)                     batch-global stop rule / rho_max exit by the last CTA
//                                                                                                al_utils.py:551-564
//
// Why it exists: north_star (3) asks for the line search and the dual / penalty updates "in the same persistent
// kernel loop"; the direction, the entry merit / residual and the problem's inputs then no longer cross HBM between
// the two phases.  Why it is NOT the default (ALMPC_SELECT_FUSE opts in): measured on B200 (quadrotor T=5, 64K
// problems, fp64) the fused kernel takes 1.44 ms against 0.83 + 0.40 ms for k_factor_solve_pair + k_line_search_warp.
// Both phases are bound by the latency of one warp's dependent chain times the warps an SM can hold; fusing them
// makes every warp carry both chains at the factor phase's occupancy (12 warps/SM at 168 registers, where the
// stand-alone line search runs 16 warps/SM at 128), and the hoped-for overlap of FP64-bound line-search warps with
// latency-bound factor warps does not materialise: staggering the phases of an SM's warps changed nothing
// (profiles/r02_fused_newton_tail.md).
//
// Mapping: the pair mapping of k_factor_solve_pair (1 half-warp = 1 problem, 2 problems per warp); the line search
// spreads the 2 x 20 x (T-1) (problem, candidate, stage) items of the warp over its 32 lanes, the T-1 stage items of
// a candidate in adjacent lanes (segmented shuffle reduction, fixed summation order).  Needs (T-1) | 32.
// The line-search inputs are restaged into the conflict-free LsLayout inside the shared-memory region the factor
// phase no longer needs (G / E / Hessian-block buffers).
#pragma once
#include "almpc_common.cuh"
#include "almpc_factor.cuh"
#include "almpc_linesearch.cuh"

namespace almpc {

// scalars of shared memory per problem of k_newton_pair
template <int NX, int NU>
__host__ __device__ inline int fused_smem_per_problem(int T, bool se) {
  const Layout<NX, NU> L(T, 0, se, true);
  const LsLayout<NX, NU> LL(T, 0, se);
  const int need = L.Gt + LL.part;  // LL.part = offset of the (unused) partial-sum area = staged scalars
  return need > L.total ? ((need + 3) & ~3) : L.total;
}

template <class M, typename S, int WPB, int V>
__global__ void __launch_bounds__(WPB * 32, (M::NX + M::NU <= 8) ? 12 : (sizeof(S) == 4 ? 10 : 6))
k_newton_pair(const KParams<S> P, const S* __restrict__ x0_g, const S* __restrict__ Qd_g, const S* __restrict__ q_g,
              S* xu_g, S* lam_g, S* __restrict__ rho_g, S* fac_g, const S* __restrict__ G_g,
              const S* __restrict__ c_g, S* __restrict__ merit_g, int* __restrict__ status_g, int* __restrict__ info_g,
              S* __restrict__ debug_g, Ctl* ctl, double* __restrict__ partials) {
  constexpr int NX = M::NX, NU = M::NU, NN = NX + NU, LS = NN | 1;
  constexpr bool SE = (V == 2);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double wred[WPB][3];
  __shared__ bool is_last;
  __shared__ int s_ring;
  grid_dep_wait();
  const int mode = ctl->mode;
  if (mode == MODE_DONE) return;
  const int al_a = ctl->a, nk = ctl->k;
  const int T = P.T, TM1 = T - 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, half = lane >> 4, hl = lane & 15;
  const int spp = fused_smem_per_problem<NX, NU>(T, SE);
  const Layout<NX, NU> L(T, 0, SE, true);
  const LsLayout<NX, NU> LL(T, 0, SE);

  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;  // this warp's statistics (identical on all lanes)
  // One warp = one pair of problems.  (Measured on B200, quadrotor 64K: persistent warps looping over pairs, with
  // or without a start-up stagger that spreads the phases of the warps of an SM, were 7-20 % slower than one CTA
  // per 4 problems — profiles/r02_fused_newton_tail.md.)
  if (tid == 0) s_ring = ring_claim(P, ctl);
  __syncthreads();
  const int ring_slot = s_ring;
  S* ring_cta = (ring_slot >= 0) ? P.ring + (size_t)ring_slot * 4 * factor_elems<NX, NU>(T) : nullptr;
  const int npairs = (P.B + 1) >> 1;
  const int gwarp = blockIdx.x * WPB + warp, nwarps = gridDim.x * WPB;
  for (int pair = gwarp; pair < npairs; pair += nwarps) {
    const int pw = pair * 2;
    __syncwarp();  // the previous pair's line search is done with the staging area before G_0 of this pair lands in it
    PairCarry<S> pc;
    factor_pair<NX, NU, S, WPB, V, true>(P, Qd_g, q_g, xu_g, lam_g, rho_g, fac_g, G_g, c_g, nullptr, merit_g, nullptr,
                                         status_g, debug_g, mode, smem_raw, spp, pw, ring_cta, pc);
    S* smw = reinterpret_cast<S*>(smem_raw) + (size_t)warp * 2 * spp;
    S* sm = smw + (size_t)half * spp;
    S* ls = sm + L.Gt;  // staging area of this half-warp's problem
    const int prob = pc.prob;
    // ---- restage: iterate, direction (sign, failed-pivot policy), cost diagonal from the factor phase's arrays;
    // q, x0 and the equality multipliers (as updated by this kernel: every lane reads back the entries it wrote
    // itself) from global memory
    __syncwarp();
    {
      const S* xu = sm + L.xu;
      const S* dv = sm + L.dv;
      const S* qds = sm + L.qd;
      const S* lin = sm + L.lin;
      const S* qv = q_g + (size_t)prob * L.N;
      const S* lam = lam_g + (size_t)prob * L.m;
      // global loads first (one round trip), shared-memory moves under them
      for (int i = hl; i < L.N; i += 16) {
        const int t = i / NN, o = t * LS + (i - t * NN);
        if constexpr (sizeof(S) == 8) cp_async8(ls + LL.q + o, qv + i);
        else cp_async4(ls + LL.q + o, qv + i);
      }
      if (hl < NX) {
        if constexpr (sizeof(S) == 8) cp_async8(ls + LL.x0 + hl, x0_g + (size_t)prob * NX + hl);
        else cp_async4(ls + LL.x0 + hl, x0_g + (size_t)prob * NX + hl);
      }
      cp_async_commit();
      // The staging area starts at L.Gt, i.e. ABOVE xu / dv / lin / qd of the factor layout: sources and
      // destinations never overlap.
      for (int i = hl; i < L.N; i += 16) {
        const int t = i / NN, o = t * LS + (i - t * NN);
        ls[LL.xu + o] = xu[i];
        ls[LL.d + o] = pc.ok ? -dv[i] : S(0);
        ls[LL.Qd + o] = qds[i];
      }
      for (int i = hl; i < L.nin; i += 16) ls[LL.lam + L.neq + i] = lin[i];
      // plain loads (not cp.async): in MODE_DUAL_NEWTON these entries were stored by this very lane in the prologue
      for (int base = 0; base < L.neq; base += 64) {
        S lv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = base + hl + 16 * u;
          lv[u] = (i < L.neq) ? lam[i] : S(0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = base + hl + 16 * u;
          if (i < L.neq) ls[LL.lam + i] = lv[u];
        }
      }
      cp_async_wait_all();
    }
    __syncwarp();

    // ---- per-problem scalars of the two problems of this warp, on every lane
    S rho_p[2], mcur_p[2];
    double rent_p[2];
    bool live_p[2];
    int prob_p[2];
    {
      S mcur = pc.merit;
      if (mode == MODE_CONT && hl == 0) mcur = merit_g[prob];  // merit <- m* of the previous line search (al_utils.py:569)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        rho_p[q] = __shfl_sync(FULL, pc.rho, q * 16);
        mcur_p[q] = __shfl_sync(FULL, mcur, q * 16);
        rent_p[q] = (double)__shfl_sync(FULL, pc.s2_entry, q * 16);
        live_p[q] = pw + q < P.B;
        prob_p[q] = pw + q;
      }
    }
    const int stat0 = __shfl_sync(FULL, pc.status, 0), stat1 = __shfl_sync(FULL, pc.status, 16);

    // ---- 20 x (T-1) items per problem over the 32 lanes
    const int IT = NLS * TM1;
    const int rounds = (2 * IT + 31) / 32;
    S best_v[2], best_r2[2];
    int best_i[2];
    bool nan_p[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) { best_v[q] = S(INFINITY); best_r2[q] = S(0); best_i[q] = NLS; nan_p[q] = false; }
#pragma unroll 1
    for (int r = 0; r < rounds; ++r) {
      const int w = r * 32 + lane;
      const int p = w / IT, it = w - p * IT;
      const int kk = it / TM1, t = it - kk * TM1;
      const bool valid = p < 2 && live_p[p & 1];
      S cost = S(0), s2 = S(0), sl = S(0);
      if (valid) ls_item<M, S, V>(smw + (size_t)p * spp + L.Gt, LL, P, T, kk, t, cost, s2, sl);
      // candidate totals: the T-1 stage items of a candidate are adjacent lanes; fixed summation order t = 0..T-2
      S c_tot = S(0), s_tot = S(0), l_tot = S(0);
      const int seg = lane - t;  // first lane of this candidate's items (valid lanes; others read themselves)
      for (int tt = 0; tt < TM1; ++tt) {
        const int src = valid ? seg + tt : lane;
        c_tot += __shfl_sync(FULL, cost, src);
        s_tot += __shfl_sync(FULL, s2, src);
        l_tot += __shfl_sync(FULL, sl, src);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (valid && p == q && t == 0) {
          const S mk = c_tot + S(0.5) * rho_p[q] * s_tot + l_tot;
          if (debug_g) debug_g[(size_t)prob_p[q] * (2 * L.N + 24) + 2 * L.N + kk] = mk;
          if (mk != mk) nan_p[q] = true;
          if (mk < best_v[q] || (mk == best_v[q] && kk < best_i[q])) { best_v[q] = mk; best_i[q] = kk; best_r2[q] = s_tot; }
        }
      }
    }
    // ---- per problem: first argmin over the lanes (torch.min NaN semantics), accept, iterate update
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (!live_p[q]) continue;  // uniform across the warp
      const int prq = prob_p[q];
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
      const S* xs = smw + (size_t)q * spp + L.Gt + LL.xu;
      const S* ds = smw + (size_t)q * spp + L.Gt + LL.d;
      bool fin = true;
      const size_t pN = (size_t)prq * L.N;
      for (int i = lane; i < L.N; i += 32) {
        const int ts = i / NN, si = ts * LS + (i - ts * NN);
        const S xo = xs[si];
        const S xnew = xo + alpha * ds[si];
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
        merit_g[prq] = bv;  // merit <- new_merit even when the step was rejected (al_utils.py:569)
        int st = q == 0 ? stat0 : stat1;
        if (!fin) st &= ~ALMPC_STATUS_FINITE;
        status_g[prq] = st;
        if (debug_g) {
          S* dbg = debug_g + (size_t)prq * (2 * L.N + 24) + 2 * L.N + NLS;
          dbg[0] = (S)bi;
          dbg[1] = accept ? S(1) : S(0);
          dbg[2] = mcur_p[q];
        }
      }
    }
  }
  // ---- batch-global statistics: per-CTA partials (fixed order), the last CTA steers the solve
  if (lane == 0) { wred[warp][0] = acc0; wred[warp][1] = acc1; wred[warp][2] = acc2; }
  __syncthreads();
  if (tid == 0 && ring_slot >= 0) ring_release(P, ring_slot);
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
  steer_solve<S>(P, ctl, info_g, partials, reinterpret_cast<double (*)[3]>(smem_raw), tid, blockDim.x, al_a, nk);
}

}  // namespace almpc
