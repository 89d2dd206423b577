// almpc_factor_mma.cuh — k_factor_solve_mma: the factor / solve step kernel with the block-tridiagonal Cholesky kept
// in FP64 tensor-core FRAGMENT layout from assembly to write-out (part of libalmpc, see almpc_kernels.cu).
//
// Same inputs, outputs and factor layout as k_factor_solve / k_factor_solve_pair (almpc_factor.cuh); what changes is
// where a problem's matrices live and which unit does the work:
//   * 1 warp = 1 problem.  The n x n block D_t (n = 16), the nx rows of -rho G_t below it and the right-hand side
//     are 7 accumulator tiles of mma.m8n8k4.f64 (14 doubles per lane) — not "one matrix row per lane" (32 doubles per
//     lane and a shared-memory round trip for every pivot column).
//   * blocked right-looking Cholesky, panel width 4 = the k of the FP64 MMA: per panel the 4x4 diagonal block is
//     broadcast once through shared memory (16 doubles) and factored + inverted redundantly by every lane in
//     registers; the panel solve  P = A21 W^T  (W = L11^{-1})  and the trailing update  A22 -= P P^T  are MMAs.
//   * a symmetric permutation inside every 8-tile (logical index pi(g) = g/2 + 4 (g%2) for physical row / column g)
//     makes the accumulator registers (c0, c1) of a tile its A/B fragments for the panel columns 0..3 / 4..7, and the
//     panel-solve output the A/B fragment of the trailing update: no shuffles and no shared-memory transposes between
//     the MMAs; E_{t-1} E_{t-1}^T, the coupling E_{t-1} y_{t-1} of the right-hand side, G_t^T G_t and G_t^T w are all
//     MMAs on registers that are already fragments.
// Shared memory per problem: gradient / cost diagonal / staged G (cp.async) / one L block for the backward sweep.
#pragma once
#include "almpc_factor.cuh"

namespace almpc {

#ifndef ALMPC_MMA_NO_SWZ
#define ALMPC_MMA_NO_SWZ 0
#endif
// chunk swizzle of a staged 128-byte row: a bit permutation (bit0 -> bit2, bit1 -> bit0, bit2 -> bit1), i.e. linear
// over XOR, so a lane's constant part and the compile-time part of a row index swizzle separately
__host__ __device__ constexpr int mma_swz(int r) { return 4 * (r & 1) + ((r >> 1) & 1) + 2 * ((r >> 2) & 1); }

template <int NX, int NU, typename S>
struct MmaLayout {
  static constexpr int NN = NX + NU;
  static constexpr bool COPY16 = (NN * sizeof(S)) % 16 == 0;  // rows of G are 16-byte multiples: chunked copies
  // 128-byte rows (16 doubles): staged UNPADDED with the 16-byte chunks of row r XOR-swizzled by swz(r mod 8) — both
  // fragment access patterns of the kernel and the cp.async writes are then bank-conflict free (with padded 144-byte
  // rows they are 2- to 3-way conflicts: 240 of 1260 shared wavefronts per problem)
  static constexpr bool SWZ = (NN * sizeof(S)) == 128 && !ALMPC_MMA_NO_SWZ;
  static constexpr int LDG = SWZ ? NN : (COPY16 ? NN + 16 / (int)sizeof(S) : NN + 1);  // row stride of the staged G block
  int N, neq, nin, m;
  int dv, dg, bc, gs0, gs1, lsm, total;  // byte offsets (every region 16-byte aligned)
  int xu, w, lin, obv, hob;              // arrays the forward sweep is done with before lsm is written (double)
  __host__ __device__ MmaLayout(int T, bool se, int nobs = 0) {
    const int SI = se ? 0 : 2 * NU + nobs;
    N = T * NN; neq = se ? (T - 1) * NX : T * NX; nin = T * SI; m = neq + nin;
    int o = 0;
    auto take = [&](int bytes) { int at = o; o += (bytes + 15) & ~15; return at; };
    dv = take(N * 8);
    dg = take(N * 8);
    bc = take(64 * 8);  // per panel parity: the 4x4 diagonal block, then the table of W / L11 entries
    gs0 = take(NX * LDG * (int)sizeof(S));
    gs1 = take(NX * LDG * (int)sizeof(S));
    // obstacle rows (Obstacle_MPC): per (stage, obstacle) J[3], lam + rho c+, active penalty; per stage the lower 3x3
    // of sum_k active_k rho J_k^T J_k
    const int nov = T * nobs * 5, nho = nobs ? T * 8 : 0;
    const int pro = (((N + 1) & ~1) + ((neq + 1) & ~1) + ((nin + 1) & ~1) + ((nov + 1) & ~1) + nho) * 8;
    const int lb = NN * NN * (int)sizeof(S);
    lsm = take(pro > lb ? pro : lb);
    xu = lsm; w = xu + ((N + 1) & ~1) * 8; lin = w + ((neq + 1) & ~1) * 8;
    obv = lin + ((nin + 1) & ~1) * 8; hob = obv + ((nov + 1) & ~1) * 8;
    total = o;
  }
};

#ifndef ALMPC_MMA3_MINB
#define ALMPC_MMA3_MINB 4  // 17..24-wide blocks (flying cartpole): 128 registers; 3 CTAs (164 registers) 0.606 vs 0.541 ms
#endif
#ifndef ALMPC_MMA_MINB
#define ALMPC_MMA_MINB 5  // 16-wide blocks: 5 CTAs x 4 warps per SM (96 registers)
#endif
// [ROWS][NN] block, contiguous in global memory -> shared rows of stride LD: 16-byte chunks where the rows are 16-byte
// multiples, element-wise otherwise
template <int ROWS, int NN, int LD, typename S>
__device__ __forceinline__ void warp_copy_rows_async(S* dst, const S* src, int lane) {
  if constexpr ((NN * sizeof(S)) % 16 == 0 && (LD * sizeof(S)) % 16 == 0) {
    constexpr int CPR = NN * (int)sizeof(S) / 16;
    constexpr int CH = ROWS * CPR;
    if constexpr (32 % CPR == 0) {  // a lane keeps its chunk column: one address computation, then constant steps
      const int r0 = lane / CPR, c = lane % CPR;
      char* d0 = (char*)(dst + r0 * LD) + c * 16;
      const char* s0 = (const char*)(src + r0 * NN) + c * 16;
#pragma unroll
      for (int k = 0; k < (CH + 31) / 32; ++k)
        if (CH % 32 == 0 || lane + 32 * k < CH)
          cp_async16(d0 + k * (32 / CPR) * LD * (int)sizeof(S), s0 + k * (32 / CPR) * NN * (int)sizeof(S));
    } else {
#pragma unroll
      for (int k = 0; k < (CH + 31) / 32; ++k) {
        const int ch = lane + 32 * k, r = ch / CPR, c = ch - r * CPR;
        if (ch < CH) cp_async16((char*)(dst + r * LD) + c * 16, (const char*)(src + r * NN) + c * 16);
      }
    }
  } else {
    warp_copy_block_async<ROWS, NN, LD>(dst, src, lane);
  }
}
// [ROWS][16 doubles] block -> unpadded shared rows with swizzled 16-byte chunks (MmaLayout::SWZ)
template <int ROWS, typename S>
__device__ __forceinline__ void warp_copy_rows_swz_async(S* dst, const S* src, int lane) {
  constexpr int CH = ROWS * 8;
  const int r0 = lane >> 3, c = lane & 7;
  const int cb = c ^ mma_swz(r0);  // rows r0 + 4k: swz(r) = swz(r0) ^ 2 (k & 1)
  char* d0 = (char*)dst + r0 * 128;
  const char* s0 = (const char*)src + lane * 16;
#pragma unroll
  for (int k = 0; k < (CH + 31) / 32; ++k)
    if (CH % 32 == 0 || lane + 32 * k < CH) cp_async16(d0 + k * 512 + ((cb ^ (2 * (k & 1))) << 4), s0 + k * 512);
}
// COUNT contiguous scalars, copy width W bytes (CopyWidth: the widest size that divides every block of the factor)
template <int COUNT, int W, typename S>
__device__ __forceinline__ void warp_copy_contig_async(S* dst, const S* src, int lane) {
  constexpr int CH = COUNT * (int)sizeof(S) / W;
  static_assert((COUNT * sizeof(S)) % W == 0, "copy width must divide the block");
#pragma unroll
  for (int k = 0; k < (CH + 31) / 32; ++k) {
    const int o = (lane + 32 * k) * W;
    if (CH % 32 == 0 || lane + 32 * k < CH) {
      if constexpr (W == 16) cp_async16((char*)dst + o, (const char*)src + o);
      else if constexpr (W == 8) cp_async8((char*)dst + o, (const char*)src + o);
      else cp_async4((char*)dst + o, (const char*)src + o);
    }
  }
}

// D = A B with a zero accumulator (the first MMA of a chain: no registers to clear first)
__device__ __forceinline__ void dmma884z(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1)
               : "d"(a), "d"(b), "d"(0.0), "d"(0.0));
}
// first MMA of a chain (compile-time flag after unrolling) or accumulation
__device__ __forceinline__ void dmma884f(bool first, double& c0, double& c1, double a, double b) {
  if (first) dmma884z(c0, c1, a, b);
  else dmma884(c0, c1, a, b);
}

constexpr __host__ __device__ int mma_dt(int rt, int ct) { return rt * (rt + 1) / 2 + ct; }  // lower tile (rt, ct) of D_t
// CTAs (of 4 warps) per SM the kernel is compiled for, by the number of 8-column tiles of a block
constexpr int mma_minb(int nt) { return nt == 1 ? 6 : (nt == 2 ? ALMPC_MMA_MINB : ALMPC_MMA3_MINB); }

template <int NX, int NU, typename S, int WPB, int V>
__global__ void __launch_bounds__(WPB * 32, mma_minb((NX + NU + 7) / 8) * 4 / WPB)
k_factor_solve_mma(const KParams<S> P, const S* __restrict__ Qd_g, const S* __restrict__ q_g, const S* __restrict__ xu_g,
                   S* __restrict__ lam_g, S* __restrict__ rho_g, S* fac_g, const S* __restrict__ G_g,
                   const S* __restrict__ c_g, S* __restrict__ d_g, S* __restrict__ merit_g,
                   double* __restrict__ rentry_g, int* __restrict__ status_g, S* __restrict__ debug_g,
                   const Ctl* __restrict__ ctl) {
  constexpr int NN = NX + NU;
  constexpr int NT = (NN + 7) / 8;     // 8-column tiles of a block (padding: identity)
  constexpr int NP = (NN + 3) / 4;     // panels of the elimination = k-chunks of a contraction over the columns
  constexpr int KG = (NX + 3) / 4;     // k-chunks of the contraction over the rows of G_t
  constexpr int NTE = (NX + 8) / 8;    // row tiles of the extra rows: nx rows of -rho G_t, then the right-hand side
  constexpr int NTX = (NX + 7) / 8;    // tiles of the x-x part of D_t (what E E^T touches)
  constexpr int XR = NX / 8, RHS = NX % 8;  // extra tile / logical in-tile row of the right-hand side
  static_assert(NT <= 3 && NTE <= 2 && NTE <= NT, "tile mapping: n <= 24, nx + 1 <= 16");
  constexpr bool OBS = (V == 1), SE = (V == 2);  // variant: 0 plain, 1 obstacle rows, 2 state estimator
  constexpr int LDG = MmaLayout<NX, NU, S>::LDG;
  constexpr bool SWZ = MmaLayout<NX, NU, S>::SWZ;
  constexpr int CW = CopyWidth<NX, NU, S>::W;
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
  const int SI = OBS ? 2 * NU + n_obs : 2 * NU;  // inequality rows per stage (a compile-time constant without obstacle rows)
  const MmaLayout<NX, NU, S> L(T, SE, n_obs);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int prob = blockIdx.x * WPB + warp;
  if (prob >= P.B) return;  // (warp 0 always has a problem, so the release below is reached)
  const int g = lane >> 2, q = lane & 3, r4 = g >> 1, odd = g & 1;
  const int pg = r4 + 4 * odd;  // logical in-tile index of physical row / column g
  const bool rhs_lane = (pg == RHS);
  unsigned char* sm = smem_raw + (size_t)warp * L.total;
  double* dv = reinterpret_cast<double*>(sm + L.dv);
  double* dg = reinterpret_cast<double*>(sm + L.dg);
  double* bc = reinterpret_cast<double*>(sm + L.bc);
  S* Gs0 = reinterpret_cast<S*>(sm + L.gs0);
  S* Gs1 = reinterpret_cast<S*>(sm + L.gs1);
  S* Lsm = reinterpret_cast<S*>(sm + L.lsm);
  double* xu = reinterpret_cast<double*>(sm + L.xu);
  double* wv = reinterpret_cast<double*>(sm + L.w);
  double* lin = reinterpret_cast<double*>(sm + L.lin);
  double* obv = reinterpret_cast<double*>(sm + L.obv);
  double* hob = reinterpret_cast<double*>(sm + L.hob);
  const S* obs = OBS ? P.obs + (size_t)prob * T * n_obs * 3 : nullptr;
  const size_t pN = (size_t)prob * L.N;
  const S* Qd = Qd_g + pN;
  const S* qv = q_g + pN;
  S* lam = lam_g + (size_t)prob * L.m;
  const S* cres = c_g + (size_t)prob * L.neq;
  const S* Gp = G_g + (size_t)prob * (T - 1) * NX * NN;
  S* fac = (ring_slot >= 0) ? P.ring + ((size_t)ring_slot * 4 + warp) * factor_elems<NX, NU>(T)
                            : fac_g + (size_t)prob * factor_elems<NX, NU>(T);
  S* dbg = debug_g ? debug_g + (size_t)prob * (2 * L.N + 24) : nullptr;

  if (T > 1) {
    if constexpr (SWZ) warp_copy_rows_swz_async<NX>(Gs0, Gp, lane);
    else warp_copy_rows_async<NX, NN, LDG>(Gs0, Gp, lane);
  }
  cp_async_commit();
  S rho_s = rho_g[prob];
  // ---- prologue: dual / penalty update of the previous AL iteration (AL_mpc.py:316-317,325), entry residual
  // ||c+||^2, merit at the entry point, w = lam_eq + rho c, and per z-entry the gradient without the G-terms
  // (al_utils.py:113-117, 438-446) and the diagonal Qd + rho (I_x | active bounds) of the Hessian block.
  constexpr int U = 3;
  S p_c[U], p_l[U], p_li[U], p_Q[U], p_q[U], p_x[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int i = lane + 32 * u;
    p_x[u] = (i < L.N) ? xu_g[pN + i] : S(0);
    p_c[u] = (i < L.neq) ? cres[i] : S(0);
    p_l[u] = (i < L.neq) ? lam[i] : S(0);
    p_li[u] = (i < L.nin) ? lam[L.neq + i] : S(0);
    p_Q[u] = (i < L.N) ? Qd[i] : S(0);
    p_q[u] = (i < L.N) ? qv[i] : S(0);
  }
  for (int base = 0; base < L.N; base += 32 * U) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + lane + 32 * u;
      if (i < L.N) xu[i] = (double)((base == 0) ? p_x[u] : xu_g[pN + i]);
    }
  }
  const bool dual = (mode == MODE_DUAL_NEWTON);
  const S rho_old = rho_s;
  if (dual) rho_s = next_rho(rho_s, P.rho_mult, P.rho_f32);
  const double rho = (double)rho_s;
  __syncwarp();
  {
    S s2 = S(0), sl = S(0), cost = S(0);
    for (int base = 0; base < L.neq; base += 32 * U) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + lane + 32 * u;
        if (i < L.neq) {
          const S cv = (base == 0) ? p_c[u] : cres[i];
          S li = (base == 0) ? p_l[u] : lam[i];
          if (dual) { li += rho_old * cv; lam[i] = li; }
          s2 += cv * cv;
          sl += li * cv;
          wv[i] = (double)(li + rho_s * cv);
        }
      }
    }
    for (int base = 0; base < L.nin; base += 32 * U) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + lane + 32 * u;
        if (i < L.nin) {
          const int t = i / SI, sidx = i - t * SI;
          const double* z_t = xu + t * NN;
          S ci, J[3];
          if (OBS && sidx >= 2 * NU) {  // obstacle row k of stage t: c = r^2 - |p - c_k|^2, J = -2 (p - c_k)^T
            const S pos[3] = {(S)z_t[0], (S)z_t[1], (S)z_t[2]};
            ci = obstacle_res<S>(pos, obs + ((size_t)t * n_obs + (sidx - 2 * NU)) * 3, P.obs_r2, J);
          } else {
            ci = (sidx < NU) ? (S)z_t[NX + sidx] - P.uhi[sidx] : P.ulo[sidx - NU] - (S)z_t[NX + sidx - NU];
          }
          S li = (base == 0) ? p_li[u] : lam[L.neq + i];
          if (dual) { li = fmax0(li + rho_old * ci); lam[L.neq + i] = li; }
          lin[i] = (double)li;
          const S cp = fmax0(ci);
          if (OBS && sidx >= 2 * NU) {
            double* ov = obv + (t * n_obs + (sidx - 2 * NU)) * 5;
            ov[0] = (double)J[0]; ov[1] = (double)J[1]; ov[2] = (double)J[2];
            ov[3] = (double)(li + rho_s * cp);                   // multiplies J in the gradient (al_utils.py:113-117)
            ov[4] = (ci >= S(0)) ? (double)rho_s : 0.0;          // multiplies J^T J in the Hessian (active rows only)
          }
          s2 += cp * cp;
          sl += li * ci;
        }
      }
    }
    __syncwarp();
    if constexpr (OBS) {  // + sum_k active_k rho J_k^T J_k on the position block (lower part), per stage
      for (int e = lane; e < T * 6; e += 32) {
        const int t = e / 6, w6 = e - t * 6;
        const int r = (w6 < 1) ? 0 : (w6 < 3 ? 1 : 2), c = w6 - (r * (r + 1)) / 2;
        double acc = 0.0;
        for (int k = 0; k < n_obs; ++k) {
          const double* ov = obv + (t * n_obs + k) * 5;
          acc += ov[4] * ov[r] * ov[c];
        }
        hob[t * 8 + w6] = acc;
      }
    }
    for (int base = 0; base < L.N; base += 32 * U) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + lane + 32 * u;
        if (i < L.N) {
          const S Qv = (base == 0) ? p_Q[u] : Qd[i];
          const S qq = (base == 0) ? p_q[u] : qv[i];
          const S z = (S)xu[i];
          const int t = i / NN, j = i - t * NN;
          const bool in_cost = !SE || j < NX;  // state estimator: the cost sees the states only (al_utils_se.py:272-305)
          if (in_cost) cost += cost_term(z, Qv, qq);
          S gv = in_cost ? Qv * z + qq : S(0);
          S dd = Qv;
          if (j < NX) {
            if (t > 0) gv += (S)wv[(t - 1) * NX + j];
            else if (!SE) gv += (S)wv[(T - 1) * NX + j];  // the x_0 - x0 block
            if (!SE || t > 0) dd += rho_s;
          } else if (!SE) {
            const int sidx = j - NX;
            const S cu = z - P.uhi[sidx], cl = P.ulo[sidx] - z;
            const S lu = (S)lin[t * SI + sidx], ll = (S)lin[t * SI + NU + sidx];
            gv += (lu - ll) + rho_s * (fmax0(cu) - fmax0(cl));
            dd += rho_s * ((cu >= S(0) ? S(1) : S(0)) + (cl >= S(0) ? S(1) : S(0)));
          }
          if (OBS && j < 3)
            for (int k = 0; k < n_obs; ++k) gv += (S)(obv[(t * n_obs + k) * 5 + 3] * obv[(t * n_obs + k) * 5 + j]);
          dv[i] = (double)gv;
          dg[i] = (double)dd;
        }
      }
    }
    s2 = warp_sum(s2);
    sl = warp_sum(sl);
    cost = warp_sum(cost);
    if (lane == 0) {
      if (dual) rho_g[prob] = rho_s;
      rentry_g[prob] = (double)s2;
      if (mode != MODE_CONT) merit_g[prob] = cost + S(0.5) * rho_s * s2 + sl;
    }
  }
  __syncwarp();

  // ---- forward sweep over the stages
  double Ep[NP][NTE];  // E_{t-1} (+ y_{t-1} in the right-hand-side row): [k-chunk][extra tile], A/B fragments
#pragma unroll
  for (int p = 0; p < NP; ++p)
#pragma unroll
    for (int xt = 0; xt < NTE; ++xt) Ep[p][xt] = 0.0;
  S* Gc = Gs0;
  S* Gn = Gs1;
  const bool isdiag = (q == r4);
  // lane-constant parts of the swizzled addresses of G_t (in scalars): fragment loads / accumulator-layout loads
  const int swz_f = q * 16 + (pg & 1) + 2 * ((pg >> 1) ^ mma_swz(q));
  const int swz_x = pg * 16 + (q & 1) + 2 * ((q >> 1) ^ mma_swz(pg));
  const int tl_idx = (q < r4) ? q * 4 + r4 : r4 * 5;  // table slot of L11[r4][q]
  for (int t = 0; t < T; ++t) {
    S* Lg = fac + (size_t)t * NN * NN;
    S* Eg = fac + (size_t)T * NN * NN + (size_t)t * NX * NN;
    const bool has_G = t < T - 1, has_E = t > 0;
    cp_async_wait_all();
    __syncwarp();
    double Dt[NT * (NT + 1) / 2][2];  // lower tiles of D_t
    double X[NTE][NT][2];             // extra rows: [extra tile][column tile][reg]: -rho G_t | rhs | 0
#pragma unroll
    for (int i = 0; i < NT * (NT + 1) / 2; ++i) Dt[i][0] = Dt[i][1] = 0.0;
    double rw[NT][2];  // w^T G_t in the right-hand-side row
#pragma unroll
    for (int ct = 0; ct < NT; ++ct) rw[ct][0] = rw[ct][1] = 0.0;
    if (has_G) {
      double f[NT][KG];
#pragma unroll
      for (int tl = 0; tl < NT; ++tl)
#pragma unroll
        for (int kk = 0; kk < KG; ++kk) {
          const int row = 4 * kk + q, col = 8 * tl + pg;
          // swizzled rows: element (row, col) sits in chunk (col / 2) ^ swz(row mod 8); row = 4 kk + q, col = 8 tl + pg
          const int at = SWZ ? (swz_f ^ (2 * ((4 * tl) ^ (2 * (kk & 1))))) + 64 * kk : row * LDG + col;
          f[tl][kk] = ((NX % 4 == 0 || row < NX) && (NN % 8 == 0 || col < NN)) ? (double)Gc[at] : 0.0;
        }
#pragma unroll
      for (int kk = 0; kk < KG; ++kk) {
#pragma unroll
        for (int rt = 0; rt < NT; ++rt)
#pragma unroll
          for (int ct = 0; ct <= rt; ++ct) dmma884f(kk == 0, Dt[mma_dt(rt, ct)][0], Dt[mma_dt(rt, ct)][1], f[rt][kk], f[ct][kk]);
        const double wa = (rhs_lane && (NX % 4 == 0 || 4 * kk + q < NX)) ? wv[t * NX + 4 * kk + q] : 0.0;
#pragma unroll
        for (int ct = 0; ct < NT; ++ct) dmma884f(kk == 0, rw[ct][0], rw[ct][1], wa, f[ct][kk]);
      }
    }
#pragma unroll
    for (int i = 0; i < NT * (NT + 1) / 2; ++i) { Dt[i][0] *= rho; Dt[i][1] *= rho; }
#pragma unroll
    for (int rt = 0; rt < NT; ++rt) {
      const double dd = (NN % 8 == 0 || 8 * rt + pg < NN) ? dg[t * NN + 8 * rt + pg] : 1.0;  // padding: identity
      if (isdiag) {
        if (odd) Dt[mma_dt(rt, rt)][1] += dd;
        else Dt[mma_dt(rt, rt)][0] += dd;
      }
    }
    if constexpr (OBS) {  // (the 3x3 position block sits inside the first 4x4 diagonal block, whose lower part is what is read)
      if (!odd && r4 < 3 && q <= r4) Dt[0][0] += hob[t * 8 + (r4 * (r4 + 1)) / 2 + q];
    }
#pragma unroll
    for (int xt = 0; xt < NTE; ++xt)
#pragma unroll
      for (int ct = 0; ct < NT; ++ct)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int row = 8 * xt + pg, col = 8 * ct + q + 4 * e;
          const bool cok = (NN % 8 == 0 || col < NN);
          double v = 0.0;
          // (swizzled rows: row = 8 xt + pg, col = 8 ct + 4 e + q)
          const int at = SWZ ? (swz_x ^ (2 * ((2 * e) ^ (4 * ct)))) + 128 * xt : row * LDG + col;
          if (has_G && row < NX && cok) v = -rho * (double)Gc[at];
          if (xt == XR && rhs_lane) {
            v = cok ? dv[t * NN + col] - rw[ct][e] : 0.0;
            if (dbg && cok) dbg[t * NN + col] = (S)v;
          }
          X[xt][ct][e] = v;
        }
    __syncwarp();  // G_t has no reader left: prefetch G_{t+1} behind the factorisation
    if (t + 1 < T - 1) {
      if constexpr (SWZ) warp_copy_rows_swz_async<NX>(Gn, Gp + (size_t)(t + 1) * NX * NN, lane);
      else warp_copy_rows_async<NX, NN, LDG>(Gn, Gp + (size_t)(t + 1) * NX * NN, lane);
    }
    cp_async_commit();
    if (has_E) {
      double sy[NTE][NTX][2];
#pragma unroll
      for (int rt = 0; rt < NTE; ++rt)
#pragma unroll
        for (int ct = 0; ct < NTX; ++ct) sy[rt][ct][0] = sy[rt][ct][1] = 0.0;
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        double bfr[NTX];
#pragma unroll
        for (int ct = 0; ct < NTX; ++ct) bfr[ct] = (8 * ct + pg < NX) ? Ep[p][ct] : 0.0;  // the y row (and the padding) is no column of E E^T
#pragma unroll
        for (int rt = 0; rt < NTE; ++rt)
#pragma unroll
          for (int ct = 0; ct < NTX; ++ct)
            if (ct <= rt) dmma884f(p == 0, sy[rt][ct][0], sy[rt][ct][1], Ep[p][rt], bfr[ct]);
      }
#pragma unroll
      for (int rt = 0; rt < NTE; ++rt)
#pragma unroll
        for (int ct = 0; ct < NTX; ++ct)
          if (ct <= rt) {
            if (rt == XR && rhs_lane) {  // row NX of the product is y_{t-1} E_{t-1}^T: the coupling of the right-hand side
              X[XR][ct][0] -= sy[rt][ct][0]; X[XR][ct][1] -= sy[rt][ct][1];
            } else {
              Dt[mma_dt(rt, ct)][0] -= sy[rt][ct][0]; Dt[mma_dt(rt, ct)][1] -= sy[rt][ct][1];
            }
          }
    }
    // ---- blocked elimination, panel width 4.  Look-ahead: as soon as a panel's solve is done, the tile holding the
    // NEXT diagonal block is updated and that block is published, so that its shared-memory round trip runs under
    // the remaining trailing updates and the write-outs of this panel.
    if (!odd) bc[r4 * 4 + q] = Dt[0][0];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int tc = p >> 1, pp = p & 1;
      const bool second_half = 4 * p + 4 < NN;  // the panel's tile has real columns behind the panel
      double* bcb = bc + pp * 32;
      __syncwarp();  // the diagonal block of this panel was published by the previous panel (look-ahead)
      const double a00 = bcb[0];
      const Vec2<double> r1 = *reinterpret_cast<const Vec2<double>*>(bcb + 4);
      const Vec2<double> r2 = *reinterpret_cast<const Vec2<double>*>(bcb + 8);
      const double a22 = bcb[10];
      const Vec2<double> r3a = *reinterpret_cast<const Vec2<double>*>(bcb + 12);
      const Vec2<double> r3b = *reinterpret_cast<const Vec2<double>*>(bcb + 14);
      // 4x4 Cholesky and its inverse W = L11^{-1}, redundantly in every lane; a non-positive pivot turns into
      // NaN / Inf and poisons the whole solution of the problem, which is what the end of the kernel tests
      const double i0 = rsqrt_(a00);
      const double l10 = r1.x * i0, l20 = r2.x * i0, l30 = r3a.x * i0;
      const double i1 = rsqrt_(fma(-l10, l10, r1.y));
      const double l21 = fma(-l20, l10, r2.y) * i1, l31 = fma(-l30, l10, r3a.y) * i1;
      const double i2 = rsqrt_(fma(-l21, l21, fma(-l20, l20, a22)));
      const double l32 = fma(-l31, l21, fma(-l30, l20, r3b.x)) * i2;
      const double i3 = rsqrt_(fma(-l32, l32, fma(-l31, l31, fma(-l30, l30, r3b.y))));
      const double w10 = -(l10 * i0) * i1;
      const double w20 = -fma(l21, w10, l20 * i0) * i2, w21 = -(l21 * i1) * i2;
      const double w30 = -fma(l32, w20, fma(l31, w10, l30 * i0)) * i3, w31 = -fma(l32, w21, l31 * i1) * i3,
                   w32 = -(l32 * i2) * i3;
      // every lane holds all of W and L11; the entry [r4][q] a lane needs (B fragment of the panel solve: physical
      // column 2c <-> column c of P; L11 for the write-out) is looked up through a 16-entry table in shared memory
      // (8 stores by one lane, 2 loads per lane) — a register select tree costs 70 instructions per panel
      {
        Vec2<double>* tb = reinterpret_cast<Vec2<double>*>(bcb + 16);  // lower: W (diag 1/L_ii), upper (c,r): L[r][c]
        if (lane == 0) {
          tb[0] = Vec2<double>{i0, l10}; tb[1] = Vec2<double>{l20, l30};
          tb[2] = Vec2<double>{w10, i1}; tb[3] = Vec2<double>{l21, l31};
          tb[4] = Vec2<double>{w20, w21}; tb[5] = Vec2<double>{i2, l32};
          tb[6] = Vec2<double>{w30, w31}; tb[7] = Vec2<double>{w32, i3};
        }
        __syncwarp();
      }
      const double wsel = bcb[16 + r4 * 4 + q];
      const double lraw = bcb[16 + tl_idx];
      const double lsel = (q <= r4) ? lraw : 0.0;
      const double wb = (odd || q > r4) ? 0.0 : wsel;
      // panel solve P = A W^T per row tile: c0 of the output is P[row][q], i.e. already the A/B fragment of the update
      double Pd[NT], PX[NTE], junk;
#pragma unroll
      for (int rt = 0; rt < NT; ++rt) {
        Pd[rt] = 0.0;
        if (rt > tc || (rt == tc && pp == 0 && second_half)) dmma884z(Pd[rt], junk, Dt[mma_dt(rt, tc)][pp], wb);
      }
#pragma unroll
      for (int xt = 0; xt < NTE; ++xt) dmma884z(PX[xt], junk, X[xt][tc][pp], wb);
      // trailing update A22 -= P P^T (columns inside the panel's own tile half are masked out of the B fragment):
      // first the tile with the next diagonal block, which is published at once
      const double nbm = odd ? -Pd[tc] : 0.0;
      if (p + 1 < NP) {
        double* bnx = bc + ((p + 1) & 1) * 32;
        if (pp == 0) {
          dmma884(Dt[mma_dt(tc, tc)][0], Dt[mma_dt(tc, tc)][1], Pd[tc], nbm);
          if (odd) bnx[r4 * 4 + q] = Dt[mma_dt(tc, tc)][1];
        } else {
          const int tn = tc + 1 < NT ? tc + 1 : tc;
          dmma884(Dt[mma_dt(tn, tn)][0], Dt[mma_dt(tn, tn)][1], Pd[tn], -Pd[tn]);
          if (!odd) bnx[r4 * 4 + q] = Dt[mma_dt(tn, tn)][0];
        }
      }
      if (pp == 0 && second_half) {
#pragma unroll
        for (int rt = 0; rt < NT; ++rt)
          if (rt > tc) dmma884(Dt[mma_dt(rt, tc)][0], Dt[mma_dt(rt, tc)][1], Pd[rt], nbm);
#pragma unroll
        for (int xt = 0; xt < NTE; ++xt) dmma884(X[xt][tc][0], X[xt][tc][1], PX[xt], nbm);
      }
#pragma unroll
      for (int ct = 0; ct < NT; ++ct)
        if (ct > tc) {
          const double nb = -Pd[ct];
#pragma unroll
          for (int rt = 0; rt < NT; ++rt)
            if (rt >= ct && !(pp == 1 && rt == tc + 1 && ct == tc + 1))  // (that tile went first)
              dmma884(Dt[mma_dt(rt, ct)][0], Dt[mma_dt(rt, ct)][1], Pd[rt], nb);
#pragma unroll
          for (int xt = 0; xt < NTE; ++xt) dmma884(X[xt][ct][0], X[xt][ct][1], PX[xt], nb);
        }
      // write-outs: L rows (zeros above the diagonal, 1/L_ii on it), E_t rows, y_t
      const bool pcol = (NN % 4 == 0 || 4 * p + q < NN);
#pragma unroll
      for (int rt = 0; rt < NT; ++rt) {
        double v;
        if (rt < tc) v = 0.0;
        else if (rt == tc) v = (pp == 0) ? (odd ? Pd[tc] : lsel) : (odd ? lsel : 0.0);
        else v = Pd[rt];
        if (pcol && (NN % 8 == 0 || 8 * rt + pg < NN)) Lg[(8 * rt + pg) * NN + 4 * p + q] = (S)v;
      }
      if (has_G) {
#pragma unroll
        for (int xt = 0; xt < NTE; ++xt)
          if (pcol && 8 * xt + pg < NX) Eg[(8 * xt + pg) * NN + 4 * p + q] = (S)PX[xt];
      }
      if (rhs_lane && pcol) dv[t * NN + 4 * p + q] = PX[XR];
#pragma unroll
      for (int xt = 0; xt < NTE; ++xt) Ep[p][xt] = PX[xt];
    }
    S* tmp = Gc; Gc = Gn; Gn = tmp;
  }
  // ---- backward sweep, d = -H^{-1} g.  L_tt / E_t come back from the streamed factor (L2) through cp.async: the
  // blocks of stage t-1 are in flight while the substitution of stage t runs.
  cp_async_wait_all();
  __syncwarp();
  S* Ecur = Gc;
  S* Enext = Gn;
  auto fetch = [&](int t, bool with_E) {
    warp_copy_contig_async<NN * NN, CW>(Lsm, fac + (size_t)t * NN * NN, lane);
    if (with_E) warp_copy_contig_async<NX * NN, CW>(Enext, fac + (size_t)T * NN * NN + (size_t)t * NX * NN, lane);
  };
  fetch(T - 1, false);
  cp_async_commit();
  for (int t = T - 1; t >= 0; --t) {
    cp_async_wait_all();
    __syncwarp();
    // lane j < n owns column j of L^T.  Its own 1/L_jj is taken out of the staged block (and replaced by 0), so the
    // column needs no masks: the stored factor has zeros above the diagonal.  Lanes >= n idle (no loads: a warp-wide
    // 8-byte load is two shared-memory wavefronts, one per half-warp).
    const bool act = lane < NN;
    const int jc = act ? lane : 0;
    const double inv = act ? (double)Lsm[jc * NN + jc] : 0.0;
    if (act) Lsm[jc * NN + jc] = S(0);
    double l[NN];
#pragma unroll
    for (int r = 0; r < NN; ++r) l[r] = act ? (double)Lsm[r * NN + jc] : 0.0;
    double bb = (lane < NN) ? dv[t * NN + lane] : 0.0;
    if (t < T - 1 && lane < NN) {
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int i = 0; i + 1 < NX; i += 2) {
        s0 += (double)Ecur[i * NN + lane] * dv[(t + 1) * NN + i];
        s1 += (double)Ecur[(i + 1) * NN + lane] * dv[(t + 1) * NN + i + 1];
      }
      if (NX & 1) s0 += (double)Ecur[(NX - 1) * NN + lane] * dv[(t + 1) * NN + NX - 1];
      bb -= s0 + s1;
    }
    __syncwarp();
    if (t > 0) fetch(t - 1, true);
    cp_async_commit();
#pragma unroll
    for (int r = NN - 1; r >= 0; --r) {
      const double dr = __shfl_sync(FULL, bb * inv, r);  // x_r, final on lane r once the rows below are done
      bb = fma(-l[r], dr, bb);                           // lanes j < r: L[r][j] x_r; l[r] = 0 on lanes j >= r
    }
    if (lane < NN) dv[t * NN + lane] = bb * inv;
    S* tmp = Ecur; Ecur = Enext; Enext = tmp;
  }
  __syncwarp();
  bool ok = true;
  for (int i = lane; i < L.N; i += 32) ok = ok && finite_(dv[i]);
  ok = __all_sync(FULL, ok);
  for (int i = lane; i < L.N; i += 32) {
    const S d = ok ? (S)(-dv[i]) : S(0);
    d_g[pN + i] = d;
    if (dbg) dbg[L.N + i] = d;
  }
  if (lane == 0) {
    int status = (mode == MODE_FIRST) ? (ALMPC_STATUS_FINITE | ALMPC_STATUS_CHOL_OK) : status_g[prob];
    if (!ok) status &= ~ALMPC_STATUS_CHOL_OK;
    status_g[prob] = status;
  }
  if (ring_slot >= 0) {
    __syncthreads();
    if (threadIdx.x == 0) ring_release(P, ring_slot);
  }
}

}  // namespace almpc
