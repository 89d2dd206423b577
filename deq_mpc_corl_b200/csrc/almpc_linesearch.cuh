// almpc_linesearch.cuh — line search, batch-global control block, final dual update: part of libalmpc (see almpc_kernels.cu for the overview).
#pragma once
#include "almpc_common.cuh"

namespace almpc {

// ------------------------------------------------------------------------------------------------
// kernel 3/3: 20-point line search + iterate update + batch-global stop statistics / control block.
// 1 thread = 1 (problem, candidate k, stage t); `pls` problems per CTA.
// ------------------------------------------------------------------------------------------------
// Executed by the LAST CTA of the line-search launch: deterministic reduction of the per-CTA statistics and the
// batch-global decisions of the reference (stop rule al_utils.py:560-564, rho_max exit AL_mpc.py:412).
// The reference's batch-global decisions from the batch statistics (s0 = sum ||c+||^2 at the entry point, s1 = at the
// exit point, s2 = max rho used): stop rule al_utils.py:560-564, rho_max exit AL_mpc.py:412.  One thread.
template <typename S>
__device__ __forceinline__ void steer_decide(const KParams<S>& P, Ctl* ctl, int* info_g, double s0, double s1, double s2,
                                             int al_a, int nk) {
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

// Executed by the LAST CTA of the kernel that ran the line search: deterministic reduction of the per-CTA statistics,
// then either the decisions (single shard) or the hand-over of this shard's totals to the caller's reduction
// (P.global_stop: k_steer decides after the all-reduce).
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
    if (P.global_stop) {
      P.stop_stats[0] = s0;
      P.stop_stats[1] = s1;
      P.stop_stats[2] = s2;
      ctl->ticket = 0;
    } else {
      steer_decide<S>(P, ctl, info_g, s0, s1, s2, al_a, nk);
    }
  }
}

// global-stop mode: the decisions, from the statistics the caller reduced over all shards (one thread)
template <typename S>
__global__ void k_steer(const KParams<S> P, Ctl* ctl, int* info_g) {
  if (threadIdx.x != 0 || ctl->mode == MODE_DONE) return;
  steer_decide<S>(P, ctl, info_g, P.stop_stats[0], P.stop_stats[1], P.stop_stats[2], ctl->a, ctl->k);
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
      cost += cost_term(z, Qd[t * LS + i], qv[t * LS + i]);
    }
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      const S z = xu[t * LS + NX + i] + alpha * dv[t * LS + NX + i];
      us[i] = z;
      if (!SE) {
        cost += cost_term(z, Qd[t * LS + NX + i], qv[t * LS + NX + i]);
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
      if (t == T - 2) cost += cost_term(z1, Qd[(t + 1) * LS + i], qv[(t + 1) * LS + i]);
    }
    if (!SE && t == T - 2) {
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        const int e = (T - 1) * LS + NX + i;
        const S z = xu[e] + alpha * dv[e];
        cost += cost_term(z, Qd[e], qv[e]);
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
             int* __restrict__ info_g, Ctl* ctl, const int ppw) {
  // a CTA owns WPB * ppw consecutive problems: the (problem, stage) dynamics evaluations are spread over all its
  // threads (one warp per problem left 28 of 32 lanes idle at T = 5), then each warp finishes ppw problems
  constexpr int NX = M::NX, NU = M::NU, NN = NX + NU;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int T = P.T;
  constexpr bool OBS = (V == 1), SE = (V == 2);
  const int n_obs = OBS ? P.n_obs : 0;
  const Layout<NX, NU> L(T, n_obs, SE);
  const int SI = OBS ? L.SI : 2 * NU;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int per_cta = WPB * ppw;
  const int prob0 = blockIdx.x * per_cta;
  const int np = (P.B - prob0 < per_cta) ? P.B - prob0 : per_cta;
  grid_dep_wait();
  if (blockIdx.x == 0 && tid == 0) {
    info_g[ALMPC_INFO_RHO_EXCEEDED] = (ctl->rho_final_max > P.rho_max) ? 1 : 0;
  }
  if (np <= 0) return;
  const int stride = L.N + L.neq;
  S* sm = reinterpret_cast<S*>(smem_raw);
  for (int e = tid; e < np * L.N; e += nthr) {
    const int p = e / L.N;
    sm[p * stride + (e - p * L.N)] = xu_g[(size_t)prob0 * L.N + e];
  }
  __syncthreads();
  for (int it = tid; it < np * (T - 1); it += nthr) {
    const int p = it / (T - 1), t = it - p * (T - 1);
    const S* xu = sm + p * stride;
    S* cres = sm + p * stride + L.N;
    S xs[NX], us[NU], xn[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) xs[i] = xu[t * NN + i];
#pragma unroll
    for (int i = 0; i < NU; ++i) us[i] = xu[t * NN + NX + i];
    rk4_step<M, S, S>(xs, us, xn, P.dyn);
#pragma unroll
    for (int i = 0; i < NX; ++i) cres[t * NX + i] = xu[(t + 1) * NN + i] - xn[i];
  }
  if (!SE) {
    for (int e = tid; e < np * NX; e += nthr) {
      const int p = e / NX, i = e - p * NX;
      sm[p * stride + L.N + (T - 1) * NX + i] = sm[p * stride + i] - x0_g[(size_t)(prob0 + p) * NX + i];
    }
  }
  __syncthreads();
  const bool upd = !(P.flags & ALMPC_FLAG_SKIP_FINAL_DUAL);
  for (int p = warp; p < np; p += WPB) {
    const int prob = prob0 + p;
    const S* xu = sm + p * stride;
    const S* cres = xu + L.N;
    const S rho = rho_g[prob];
    S* lamp = lam_g + (size_t)prob * L.m;
    S s2 = S(0);
    for (int i = lane; i < L.neq; i += 32) {
      const S ci = cres[i];
      if (upd) lamp[i] += rho * ci;
      s2 += ci * ci;
    }
    const S* obs = OBS ? P.obs + (size_t)prob * T * n_obs * 3 : nullptr;
    for (int i = lane; i < L.nin; i += 32) {
      const int t = i / SI, sx = i - t * SI;
      const S ci = ineq_res<NX, NU, S>(xu + t * NN, sx, P.ulo, P.uhi, OBS ? obs + (size_t)t * n_obs * 3 : nullptr, P.obs_r2);
      if (upd) lamp[L.neq + i] = fmax0(lamp[L.neq + i] + rho * ci);
      const S cp = fmax0(ci);
      s2 += cp * cp;
    }
    s2 = warp_sum(s2);
    if (lane == 0) {
      if (upd) rho_g[prob] = next_rho(rho, P.rho_mult, P.rho_f32);
      if (dyn_res_g) dyn_res_g[prob] = sqrt(s2);
    }
    if (x_out) {
      for (int i = lane; i < T * NX; i += 32) {
        const int t = i / NX, j = i - t * NX;
        x_out[(size_t)prob * T * NX + i] = (float)xu[t * NN + j];
      }
    }
    if (u_out) {
      for (int i = lane; i < T * NU; i += 32) {
        const int t = i / NU, j = i - t * NU;
        u_out[(size_t)prob * T * NU + i] = (float)xu[t * NN + NX + j];
      }
    }
  }
}


}  // namespace almpc
