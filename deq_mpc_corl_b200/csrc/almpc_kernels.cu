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
//   k_factor_solve(_pair)  1 warp = 1 problem (row per lane), or for blocks up to 16 wide 1 half-warp = 1 problem
//                   (two rows per lane, the default for pendulum / cartpoles / quadrotor): residual statistics,
//                   (optional) dual/penalty update of
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
#include "almpc_common.cuh"
#include "almpc_dynamics_kernels.cuh"
#include "almpc_factor.cuh"
#include "almpc_linesearch.cuh"

namespace almpc {

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

// the half-warp-per-problem factor kernel exists for blocks that fit a half-warp (n <= 16, nx + 1 <= 16) and
// for the variants without obstacle rows
constexpr int kWPP = 2;  // warps per CTA of the pair kernel (4 problems per CTA)
template <int NX, int NU, typename S, int V>
static auto pair_kernel() -> decltype(&k_factor_solve<NX, NU, S, kWPB, V>) {
  if constexpr (NX + NU <= 16 && NX + 1 <= 16 && V != 1) return &k_factor_solve_pair<NX, NU, S, kWPP, V>;
  else return nullptr;
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
  int fin_ppw = 32 / (T - 1);  // problems per warp of the final kernel: its threads share the (problem, stage) items
  fin_ppw = fin_ppw < 1 ? 1 : (fin_ppw > 8 ? 8 : fin_ppw);
  const size_t smem_fin = (size_t)kWPB * fin_ppw * (L.N + L.neq) * sizeof(S);
  if (smem_fac + 1024 > 227 * 1024 || smem_ls + 1024 > 227 * 1024)
    return set_err(-11, "horizon too long for the shared-memory layout");
  auto kjac = k_jacobian<M, S, V>;
  auto kfac = &k_factor_solve<NX, NU, S, kWPB, V>;
  int fac_threads = kWPB * 32, fac_per_cta = kWPB;
  if (auto kp = pair_kernel<NX, NU, S, V>(); kp && !getenv("ALMPC_FAC_SINGLE")) {  // env: A/B against the row-per-lane kernel
    // Both kernels are bound by (latency of one problem) / (problems resident per SM): take the pair kernel
    // unless its larger shared-memory footprint (long horizons) leaves it with clearly fewer resident problems.
    const size_t smem_pair = (size_t)2 * kWPP * Layout<NX, NU>(T, 0, SE, true).total * sizeof(S);
    int occ_single = 0, occ_pair = 0;
    bool use_pair = smem_pair + 1024 <= 227 * 1024;
    if (use_pair) {
      cudaError_t e1 = cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pair);
      cudaError_t e2 = cudaFuncSetAttribute(kfac, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fac);
      if (e1 == cudaSuccess && e2 == cudaSuccess &&
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_pair, kp, kWPP * 32, smem_pair) == cudaSuccess &&
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_single, kfac, kWPB * 32, smem_fac) == cudaSuccess)
        use_pair = occ_pair * 2 * kWPP * 5 >= occ_single * kWPB * 4;  // pair wins down to 0.8x the resident problems
      else
        (void)cudaGetLastError();
    }
    if (getenv("ALMPC_FAC_PAIR")) use_pair = smem_pair + 1024 <= 227 * 1024;
    if (use_pair) {
      kfac = kp;
      fac_threads = kWPP * 32;
      fac_per_cta = 2 * kWPP;
      smem_fac = smem_pair;
      if (const char* pad = getenv("ALMPC_FAC_SMEM_PAD")) smem_fac += (size_t)atoi(pad);
    }
  }
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
    pt.nb = (pt.n + fac_per_cta - 1) / fac_per_cta;
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
      launch_k(kfac, pt.nb, fac_threads, smem_fac, pt.s, pdl, pt.P, Qd + lo * L.N, q + lo * L.N, xu + lo * L.N,
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
  launch_k(kfin, (B + kWPB * fin_ppw - 1) / (kWPB * fin_ppw), kWPB * 32, smem_fin, st, pdl, P, x0, xu, lam, rho, x_out,
           u_out, dyn_res_out, info_out, ctl, fin_ppw);
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
