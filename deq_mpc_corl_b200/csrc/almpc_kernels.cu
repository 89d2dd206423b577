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
//   k_factor_solve(_mma|_pair)  1 warp = 1 problem.  Default for fp64 blocks wider than 8 (quadrotor, flying
//                   cartpole): k_factor_solve_mma, the blocked Cholesky kept in FP64 tensor-core fragment layout
//                   (almpc_factor_mma.cuh); for the small models and fp32 1 half-warp = 1 problem, two matrix rows
//                   per lane (_pair); the row-per-lane kernel is the fallback.  All three: residual statistics,
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
#include "almpc_factor_mma.cuh"
#include "almpc_linesearch.cuh"
#include "almpc_fused.cuh"
#include "almpc_kkt.cuh"

#include <mutex>

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
  P.global_stop = 0; P.stop_stats = nullptr;
  P.ring = nullptr; P.ring_mask = nullptr;
  P.n_obs = d->obstacles ? d->n_obs : 0;
  P.obs_r2 = (S)(d->obs_radius * d->obs_radius);
  P.obs = (const S*)d->obstacles;
  return P;
}

[[maybe_unused]] static int check_desc(const almpc_desc* d) {
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

// Per-call profiling (desc.profile_ms != NULL): events around every launch of this forward call.
struct ProfEv { cudaEvent_t a, b; int kind, slot; };

[[maybe_unused]] static int n_blocks(int B) { return (B + kWPB - 1) / kWPB; }

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
  size_t ctl, partials, merit, rentry, status, c, d, G, ring, total;
};
constexpr int kPdlMaxBatch = 16384;   // batches up to this size launch their step kernels with PDL
constexpr int kRingMinBatch = 8192;  // below this a whole batch of factors sits in L2 anyway
static WsPlan ws_plan(int B, int T, int nx, int nu, size_t es) {
  const size_t nn = (size_t)nx + nu;
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  WsPlan w;
  size_t o = 0;
  w.ctl = o; o += 1024;  // control block (256 B) + the scratch-ring slot masks (kRingSMs ints)
  // one slot of statistics per CTA of the kernel that runs the line search: the stand-alone kernels use at most
  // ls_geom().blocks CTAs, the fused Newton-tail kernel one CTA per 4 problems
  size_t slots = (size_t)ls_geom(B, T).blocks;
  if (slots < (size_t)(B + 3) / 4) slots = (size_t)(B + 3) / 4;
  w.partials = o; o = up(o + (slots + 4) * 4 * sizeof(double));
  w.merit = o; o = up(o + (size_t)B * es);
  w.rentry = o; o = up(o + (size_t)B * sizeof(double));
  w.status = o; o = up(o + (size_t)B * sizeof(int32_t));
  w.c = o; o = up(o + (size_t)B * T * nx * es);
  w.d = o; o = up(o + (size_t)B * T * nn * es);
  w.G = o; o = up(o + (size_t)B * (T - 1) * nx * nn * es);
  // scratch ring for the factors of Newton steps nobody differentiates through (KParams::ring): 4 factors per slot
  w.ring = 0;
  if (B >= kRingMinBatch) {
    w.ring = o;
    o = up(o + (size_t)kRingSMs * kRingSlots * 4 * ((size_t)T * nn * nn + (size_t)(T - 1) * nx * nn) * es);
  }
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

// Opt-in dynamic shared memory of a kernel, raised to everything the device allows next to the kernel's static
// shared memory.  Plans are cached per horizon while this limit is an attribute of the FUNCTION: with the maximum set
// once, a plan made for a short horizon never lowers the limit under a cached plan for a longer one.
template <typename F>
static cudaError_t allow_max_smem(F fn, int optin) {
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, (const void*)fn);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
}

constexpr int kWLS = 4;  // warps per CTA of the warp-synchronous line search

struct LsWarpCfg { const void* fn; int grid_cap; size_t smem; int pw; };

template <class M, typename S, int TM1, int V>
static cudaError_t ls_warp_cfg(LsWarpCfg& c, int n_sm, int n_obs, int smem_optin) {
  constexpr int PW = (8 / TM1) > 0 ? (8 / TM1) : 1;
  const LsLayout<M::NX, M::NU> LL(TM1 + 1, n_obs, V == 2);
  auto k = k_line_search_warp<M, S, TM1, kWLS, V>;
  size_t smem = (size_t)kWLS * PW * LL.part * sizeof(S);
  if (smem < (size_t)kWLS * 32 * 3 * sizeof(double)) smem = (size_t)kWLS * 32 * 3 * sizeof(double);
  if (smem > (size_t)smem_optin) return cudaErrorInvalidValue;
  cudaError_t e = allow_max_smem(k, smem_optin);
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
constexpr int kWPP = 2;  // warps per CTA of the pair kernels (4 problems per CTA)
template <int NX, int NU, typename S, int V>
static auto pair_kernel() -> decltype(&k_factor_solve<NX, NU, S, kWPB, V>) {
  if constexpr (NX + NU <= 16 && NX + 1 <= 16 && V != 1) return &k_factor_solve_pair<NX, NU, S, kWPP, V>;
  else return nullptr;
}
// the fragment-layout factor kernel (almpc_factor_mma.cuh): blocks up to 24 wide, nx + 1 <= 16
#ifndef ALMPC_MMA_WPB
#define ALMPC_MMA_WPB 4
#endif
constexpr int kWPM = ALMPC_MMA_WPB;  // warps = problems per CTA of the fragment-layout factor kernel
template <int NX, int NU>
constexpr bool mma_applies(int) { return NX + NU <= 24 && NX + 1 <= 16; }
template <int NX, int NU, typename S, int V>
static auto mma_kernel() -> decltype(&k_factor_solve<NX, NU, S, kWPB, V>) {
  if constexpr (mma_applies<NX, NU>(V)) return &k_factor_solve_mma<NX, NU, S, kWPM, V>;
  else return nullptr;
}
template <typename S>
using FusedFn = void (*)(const KParams<S>, const S*, const S*, const S*, S*, S*, S*, S*, const S*, const S*, S*, int*, int*,
                         S*, Ctl*, double*);
template <class M, typename S, int V>
static FusedFn<S> fused_kernel() {
  if constexpr (M::NX + M::NU <= 16 && M::NX + 1 <= 16 && V != 1) return &k_newton_pair<M, S, kWPP, V>;
  else return nullptr;
}

// Launch plan of one (model, dtype, variant) for a given (device, T, n_obs, kernel selection): kernel choice,
// shared-memory sizes, occupancy-derived grid caps.  Function attributes and occupancy queries cost ~10 us per
// forward call when redone every time (a fifth of a 200-problem solve), so they are resolved once and cached.
struct Plan {
  int dev = -1, T = 0, n_obs = -1, select = -1;
  int n_sm = 0;
  bool use_pair = false, fused = false, use_mma = false;
  size_t smem_fac = 0, smem_ls = 0, smem_fin = 0, smem_fused = 0;
  int fac_threads = 0, fac_per_cta = 0, fin_ppw = 1, ls_grid_cap = 0;
  LsWarpCfg lw{nullptr, 0, 0, 1};
};

template <class M, typename S, int V>
static int make_plan(Plan& pl, int dev, int T, int n_obs, int select) {
  constexpr int NX = M::NX, NU = M::NU;
  constexpr bool SE = (V == 2);
  const Layout<NX, NU> L(T, n_obs, SE);
  const LsLayout<NX, NU> LL(T, n_obs, SE);
  const LsGeom lg = ls_geom(1, T);
  pl.dev = dev; pl.T = T; pl.n_obs = n_obs; pl.select = select;
  cudaError_t e = cudaDeviceGetAttribute(&pl.n_sm, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return cuda_err(e, "cudaDeviceGetAttribute");
  // every kernel gets the device maximum of opt-in shared memory once (allow_max_smem); the occupancy queries below
  // use the actual launch sizes
  int optin = 0;
  e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e != cudaSuccess) return cuda_err(e, "cudaDeviceGetAttribute");
  pl.smem_fac = (size_t)kWPB * L.total * sizeof(S);
  pl.smem_ls = 2 * (size_t)lg.pls * LL.total * sizeof(S);  // double-buffered staging
  if (pl.smem_ls < (size_t)lg.threads * 3 * sizeof(double)) pl.smem_ls = (size_t)lg.threads * 3 * sizeof(double);
  pl.fin_ppw = 32 / (T - 1);  // problems per warp of the final kernel: its threads share the (problem, stage) items
  pl.fin_ppw = pl.fin_ppw < 1 ? 1 : (pl.fin_ppw > 8 ? 8 : pl.fin_ppw);
  pl.smem_fin = (size_t)kWPB * pl.fin_ppw * (L.N + L.neq) * sizeof(S);
  if (pl.smem_fac + 1024 > 227 * 1024 || pl.smem_ls + 1024 > 227 * 1024 || pl.smem_fin > (size_t)optin)
    return set_err(-11, "horizon too long for the shared-memory layout");
  auto kfac = &k_factor_solve<NX, NU, S, kWPB, V>;
  pl.fac_threads = kWPB * 32; pl.fac_per_cta = kWPB;
  pl.use_pair = pl.fused = pl.use_mma = false;
  if (auto km = mma_kernel<NX, NU, S, V>(); km && !(select & (ALMPC_SELECT_FAC_PAIR | ALMPC_SELECT_FAC_SINGLE | ALMPC_SELECT_FUSE)) &&
                                               ((sizeof(S) == 8 && NX + NU > 8) || (select & ALMPC_SELECT_FAC_MMA))) {
    // Default for fp64 problems with blocks wider than 8 (measured per launch, B200): quadrotor 64K 0.63 vs 0.88 ms (8K:
    // 0.095 vs 0.137 ms), flying cartpole T=10 16K 0.61 vs 0.95 ms.  Blocks that fit one 8x8 tile (pendulum, cartpoles)
    // keep the half-warp kernel (64K problems: 0.170 / 0.226 / 0.297 ms against 0.184 / 0.271 / 0.311 ms; for single-wave
    // batches this kernel is faster, 13 vs 17 us per launch at 200 pendulum problems, but the last iteration of the pendulum
    // DEQ chain (rho = 1e12, cond(H) beyond 1e12) then sits 4.4e-5 from the oracle instead of 7e-6, outside the budget of
    // test_deq_chain_without_resync_matches_oracle; tools/chain_accuracy.py), and so do
    // fp32 problems: it factors in fp32 registers at 10 CTAs/SM (quadrotor 0.553 ms) while this kernel keeps its blocks
    // in FP64 accumulators whatever the interface type (0.620 ms) — selectable for its accuracy (no lost pivots).
    if constexpr (mma_applies<NX, NU>(V)) {
      const size_t smem_mma = (size_t)kWPM * MmaLayout<NX, NU, S>(T, SE, n_obs).total;
      if (smem_mma + 1024 <= 227 * 1024 && allow_max_smem(km, optin) == cudaSuccess) {
        pl.use_mma = true;
        pl.fac_threads = kWPM * 32;
        pl.fac_per_cta = kWPM;
        pl.smem_fac = smem_mma;
      } else {
        (void)cudaGetLastError();
      }
    }
  }
  if (auto kp = pair_kernel<NX, NU, S, V>(); kp && !pl.use_mma && !(select & ALMPC_SELECT_FAC_SINGLE)) {
    // Both kernels are bound by (latency of one problem) / (problems resident per SM): take the pair kernel
    // unless its larger shared-memory footprint (long horizons) leaves it with clearly fewer resident problems.
    const size_t smem_pair = (size_t)2 * kWPP * Layout<NX, NU>(T, 0, SE, true).total * sizeof(S);
    int occ_single = 0, occ_pair = 0;
    bool use_pair = smem_pair + 1024 <= 227 * 1024;
    if (use_pair) {
      cudaError_t e1 = allow_max_smem(kp, optin);
      cudaError_t e2 = allow_max_smem(kfac, optin);
      if (e1 == cudaSuccess && e2 == cudaSuccess &&
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_pair, kp, kWPP * 32, smem_pair) == cudaSuccess &&
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_single, kfac, kWPB * 32, pl.smem_fac) == cudaSuccess)
        use_pair = occ_pair * 2 * kWPP * 5 >= occ_single * kWPB * 4;  // pair wins down to 0.8x the resident problems
      else
        (void)cudaGetLastError();
    }
    if (use_pair) {
      pl.use_pair = true;
      pl.fac_threads = kWPP * 32;
      pl.fac_per_cta = 2 * kWPP;
      pl.smem_fac = smem_pair;
      // fused Newton tail (factor + line search in one kernel, opt-in): needs the warp line-search item layout, (T-1) | 32
      if (auto kf = fused_kernel<M, S, V>(); kf && (select & ALMPC_SELECT_FUSE) && (32 % (T - 1) == 0)) {
        const size_t smem_fused = (size_t)2 * kWPP * fused_smem_per_problem<NX, NU>(T, SE) * sizeof(S);
        if (smem_fused + 1024 <= 227 * 1024 &&
            allow_max_smem(kf, optin) == cudaSuccess) {
          pl.fused = true;
          pl.smem_fused = smem_fused;
        } else {
          (void)cudaGetLastError();
        }
      }
    }
  }
  if (!pl.use_pair && !pl.use_mma) {
    e = allow_max_smem(kfac, optin);
    if (e != cudaSuccess) return cuda_err(e, "cudaFuncSetAttribute(factor_solve)");
  }
  auto kls = k_line_search<M, S, V>;
  auto kfin = k_final_dual<M, S, kWPB, V>;
  e = allow_max_smem(kls, optin);
  if (e != cudaSuccess) return cuda_err(e, "cudaFuncSetAttribute(line_search)");
  e = allow_max_smem(kfin, optin);
  if (e != cudaSuccess) return cuda_err(e, "cudaFuncSetAttribute(final_dual)");
  // persistent line-search grid: as many CTAs as fit on the device at once
  int ls_per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ls_per_sm, kls, lg.threads, pl.smem_ls);
  if (e != cudaSuccess) return cuda_err(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
  pl.ls_grid_cap = pl.n_sm * (ls_per_sm > 0 ? ls_per_sm : 1);
  // warp-synchronous line search when the stage items of a candidate fit a power-of-two lane group
  pl.lw = LsWarpCfg{nullptr, 0, 0, 1};
  if (!(select & ALMPC_SELECT_LS_CTA) && !pl.fused) {
    switch (T - 1) {
      case 1: e = ls_warp_cfg<M, S, 1, V>(pl.lw, pl.n_sm, n_obs, optin); break;
      case 2: e = ls_warp_cfg<M, S, 2, V>(pl.lw, pl.n_sm, n_obs, optin); break;
      case 4: e = ls_warp_cfg<M, S, 4, V>(pl.lw, pl.n_sm, n_obs, optin); break;
      case 8: e = ls_warp_cfg<M, S, 8, V>(pl.lw, pl.n_sm, n_obs, optin); break;
      case 16: e = ls_warp_cfg<M, S, 16, V>(pl.lw, pl.n_sm, n_obs, optin); break;
      case 32: e = ls_warp_cfg<M, S, 32, V>(pl.lw, pl.n_sm, n_obs, optin); break;
      default: break;
    }
    if (e != cudaSuccess) return cuda_err(e, "line-search (warp) configuration");
  }
  return 0;
}

template <class M, typename S, int V>
static int get_plan(Plan& out, int T, int n_obs, int select) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_err(e, "cudaGetDevice");
  // a handful of entries per instantiation, looked up under a lock: re-entrant, no per-call CUDA queries
  static std::mutex mu;
  static Plan cache[8];
  static int next = 0;
  std::lock_guard<std::mutex> lk(mu);
  for (const Plan& p : cache)
    if (p.dev == dev && p.T == T && p.n_obs == n_obs && p.select == select) { out = p; return 0; }
  Plan pl;
  const int rc = make_plan<M, S, V>(pl, dev, T, n_obs, select);
  if (rc) return rc;
  cache[next] = pl;
  next = (next + 1) % 8;
  out = pl;
  return 0;
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
  Plan pl;
  if (int rc = get_plan<M, S, V>(pl, T, n_obs, d->select)) return rc;
  const bool global_stop = d->stop_reduce != nullptr;
  if (global_stop && !d->stop_stats) return set_err(-13, "stop_reduce needs the device buffer stop_stats");
  auto kjac = k_jacobian<M, S, V>;
  auto kfac = &k_factor_solve<NX, NU, S, kWPB, V>;
  if (pl.use_pair) kfac = pair_kernel<NX, NU, S, V>();
  if (pl.use_mma) kfac = mma_kernel<NX, NU, S, V>();
  auto kfused = fused_kernel<M, S, V>();
  auto kls = k_line_search<M, S, V>;
  auto kfin = k_final_dual<M, S, kWPB, V>;
  const LsGeom lg = ls_geom(B, T);
  KParams<S> P = make_params<S>(d);
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
  static_assert(256 + kRingSMs * sizeof(int) <= 1024, "slot masks must fit the control region");
  if (w.ring && !(d->select & ALMPC_SELECT_NO_RING) && !debug_out) {
    P.ring = (S*)(ws + w.ring);
    P.ring_mask = (int*)(ws + w.ctl + 256);
  }
  cudaError_t e = cudaMemsetAsync(ctl, 0, 1024, st);
  if (e != cudaSuccess) return cuda_err(e, "cudaMemsetAsync(ctl)");
  e = cudaMemsetAsync(info_out, 0, ALMPC_INFO_INTS * sizeof(int32_t), st);
  if (e != cudaSuccess) return cuda_err(e, "cudaMemsetAsync(info)");
  // launch geometry of this batch
  int ls_grid;
  if (pl.fused) {
    ls_grid = (B + 2 * kWPP - 1) / (2 * kWPP);
  } else if (pl.lw.fn) {
    const int groups = (B + pl.lw.pw - 1) / pl.lw.pw;
    const int ctas = (groups + kWLS - 1) / kWLS;
    ls_grid = pl.lw.grid_cap < ctas ? pl.lw.grid_cap : ctas;
  } else {
    const int groups = (B + lg.pls - 1) / lg.pls;
    ls_grid = pl.ls_grid_cap < groups ? pl.ls_grid_cap : groups;
  }
  P.part_off = 0;
  P.ticket_total = ls_grid;
  P.global_stop = global_stop ? 1 : 0;
  P.stop_stats = d->stop_stats;
  const int nb = (B + pl.fac_per_cta - 1) / pl.fac_per_cta;
  const unsigned jac_blocks = (unsigned)(((long)B * (T - 1) * (NN - M::NTRIV) + 127) / 128);
  const int launches = d->al_iter * d->max_newton;
  // Programmatic dependent launch (the next kernel's CTAs are scheduled while the previous kernel drains; every kernel
  // starts with griddepcontrol.wait).  Measured on B200, quadrotor fp64 forward: 64K problems no gain, 16K -1.6 %,
  // 8K -3.9 % (the drain / ramp-up of 24 dependent launches is a visible share of a 1.8 ms solve): on for small
  // batches, selectable otherwise.
  const bool pdl = !d->profile_ms && ((d->select & ALMPC_SELECT_PDL) || (B <= kPdlMaxBatch && !(d->select & ALMPC_SELECT_NO_PDL)));
  ProfEv pe[3 * ALMPC_MAX_AL_ITER * 8 + 1];
  int npe = 0;
  const bool prof = d->profile_ms && launches * 3 + 1 <= (int)(sizeof(pe) / sizeof(pe[0]));
  int slot = 0;
  auto tic = [&](int kind) {
    if (!prof) return;
    cudaEventCreate(&pe[npe].a);
    cudaEventCreate(&pe[npe].b);
    pe[npe].kind = kind;
    pe[npe].slot = slot;
    cudaEventRecord(pe[npe].a, st);
  };
  auto toc = [&]() {
    if (!prof) return;
    cudaEventRecord(pe[npe].b, st);
    ++npe;
  };
  for (int i = 0; i < launches; ++i) {
    slot = i;
    tic(0);
    e = launch_k(kjac, jac_blocks, 128, 0, st, pdl, P, x0, xu, G_g, c_g, ctl);
    if (e != cudaSuccess) return cuda_err(e, "k_jacobian launch");
    toc();
    if (pl.fused) {
      if constexpr (NX + NU <= 16 && NX + 1 <= 16 && V != 1) {
        tic(1);
        e = launch_k(kfused, (unsigned)ls_grid, kWPP * 32, pl.smem_fused, st, pdl, P, x0, Qd, q, xu, lam, rho, factor, G_g,
                     c_g, merit, status, info_out, debug_out, ctl, partials);
        if (e != cudaSuccess) return cuda_err(e, "k_newton_pair launch");
        toc();
      }
    } else {
      tic(1);
      e = launch_k(kfac, nb, pl.fac_threads, pl.smem_fac, st, pdl, P, Qd, q, xu, lam, rho, factor, G_g, c_g, d_g, merit,
                   rentry, status, debug_out, ctl);
      if (e != cudaSuccess) return cuda_err(e, "k_factor_solve launch");
      toc();
      tic(2);
      if (pl.lw.fn) {
        const S *a_x0 = x0, *a_Qd = Qd, *a_q = q, *a_lam = lam, *a_rho = rho, *a_d = d_g;
        S *a_xu = xu, *a_merit = merit;
        const double* a_re = rentry;
        int* a_st = status;
        void* args[] = {(void*)&P, &a_x0, &a_Qd, &a_q, &a_xu, &a_lam, &a_rho, &a_d, &a_merit, &a_re, &a_st,
                        &info_out, &debug_out, &ctl, &partials};
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(ls_grid);
        cfg.blockDim = dim3(kWLS * 32);
        cfg.dynamicSmemBytes = pl.lw.smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = pdl ? 1 : 0;
        e = cudaLaunchKernelExC(&cfg, pl.lw.fn, args);
        if (e != cudaSuccess) return cuda_err(e, "k_line_search_warp launch");
      } else {
        e = launch_k(kls, ls_grid, lg.threads, pl.smem_ls, st, pdl, P, lg.pls, x0, Qd, q, xu, lam, rho, d_g, merit, rentry,
                     status, info_out, debug_out, ctl, partials);
        if (e != cudaSuccess) return cuda_err(e, "k_line_search launch");
      }
      toc();
    }
    if (global_stop) {
      // The last CTA of the line search left this rank's (sum ||c+||^2 entry, sum exit, max rho) in stop_stats
      // instead of steering.  The caller's hook enqueues the reduction over the ranks on this stream (two SUMs and
      // a MAX: 24 bytes, latency only), then one tiny kernel takes the reference's batch-GLOBAL decisions
      // (al_utils.py:551-564, AL_mpc.py:412) identically on every rank.
      const int rc = d->stop_reduce(d->stop_stats, (void*)st, d->stop_reduce_user);
      if (rc) return set_err(-14, "stop_reduce hook failed");
      e = launch_k(k_steer<S>, 1u, 32u, 0, st, false, P, ctl, info_out);
      if (e != cudaSuccess) return cuda_err(e, "k_steer launch");
    }
  }
  tic(3);
  e = launch_k(kfin, (B + kWPB * pl.fin_ppw - 1) / (kWPB * pl.fin_ppw), kWPB * 32, pl.smem_fin, st, pdl, P, x0, xu, lam, rho,
               x_out, u_out, dyn_res_out, info_out, ctl, pl.fin_ppw);
  if (e != cudaSuccess) return cuda_err(e, "k_final_dual launch");
  toc();
  if (prof) {  // profiling mode synchronises (never used on a timed or captured path)
    cudaStreamSynchronize(st);
    int executed = 0;  // Newton steps the solve realised: the launches of the later slots found MODE_DONE and did nothing
    cudaMemcpy(&executed, info_out + ALMPC_INFO_TOTAL_NEWTON, sizeof(int), cudaMemcpyDeviceToHost);
    double acc[8] = {0};
    for (int i = 0; i < npe; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, pe[i].a, pe[i].b);
      if (pe[i].slot < executed || pe[i].kind == 3) {
        acc[pe[i].kind] += ms;
        acc[4 + pe[i].kind] += 1;
      }
      cudaEventDestroy(pe[i].a);
      cudaEventDestroy(pe[i].b);
    }
    for (int k = 0; k < 8; ++k) d->profile_ms[k] = acc[k];
  }
  return cuda_err(cudaGetLastError(), "almpc_forward launch");
}

template <int NX, int NU, typename S>
static int backward_impl(const almpc_desc* d, const S* factor, const S* xu, const S* grad_xu, const int32_t* status,
                         S* grad_Q, S* grad_q, cudaStream_t st) {
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
  kern<<<n_blocks(d->B), kWPB * 32, smem, st>>>(d->B, d->T, factor, xu, grad_xu, status, grad_Q, grad_q);
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
static int backward_any(const almpc_desc* d, const S* factor, const S* xu, const S* grad_xu, const int32_t* status,
                        S* grad_Q, S* grad_q, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  ALMPC_DISPATCH(d->dyn_id, (backward_impl<M::NX, M::NU, S>(d, factor, xu, grad_xu, status, grad_Q, grad_q, st)))
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
static int kkt_solve(int B, int nz, int nineq, int neq, const S* Q, const S* G, const S* A, const S* s, const S* z, double eps,
                     int n_refine, const S* rx, const S* rs, const S* rz, const S* ry, S* dx, S* ds, S* dz, S* dy, int32_t* status,
                     void* stream) {
  if (B < 0 || nz < 1 || nineq < 0 || neq < 0 || n_refine < 0)
    return set_err(-3, "kkt_solve: need nz >= 1, nineq >= 0, neq >= 0, n_refine >= 0");
  if (B == 0) return 0;
  if (!Q || !rx || !dx || (nineq && (!G || !s || !z || !rs || !rz || !ds || !dz)) || (neq && (!A || !ry || !dy)))
    return set_err(-10, "null buffer");
  const size_t smem = kkt_smem_elems(nz, nineq, neq) * sizeof(S);
  int dev = 0, optin = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_err(e, "cudaGetDevice");
  e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e != cudaSuccess) return cuda_err(e, "cudaDeviceGetAttribute");
  if (smem > (size_t)optin)
    return set_err(-11, "kkt_solve: (nz + neq)^2 scalars do not fit the shared memory of one CTA");
  auto kern = k_kkt_solve<S>;
  e = allow_max_smem(kern, optin);
  if (e != cudaSuccess) return cuda_err(e, "cudaFuncSetAttribute(kkt_solve)");
  kern<<<(unsigned)B, kKktThreads, smem, (cudaStream_t)stream>>>(B, nz, nineq, neq, Q, G, A, s, z, (S)eps, n_refine, rx, rs, rz, ry,
                                                              dx, ds, dz, dy, status);
  return cuda_err(cudaGetLastError(), "almpc_kkt_solve launch");
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
                       const int32_t* status, double* grad_Q, double* grad_q, void* stream) {
  return backward_any<double>(d, factor, xu, grad_xu, status, grad_Q, grad_q, stream);
}
#endif
#if ALMPC_BUILD_F32
int almpc_backward_f32(const almpc_desc* d, const float* factor, const float* xu, const float* grad_xu,
                       const int32_t* status, float* grad_Q, float* grad_q, void* stream) {
  return backward_any<float>(d, factor, xu, grad_xu, status, grad_Q, grad_q, stream);
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
int almpc_kkt_solve_f64(int32_t B, int32_t nz, int32_t nineq, int32_t neq, const double* Q, const double* G, const double* A,
                        const double* s, const double* z, double eps, int32_t n_refine, const double* rx, const double* rs,
                        const double* rz, const double* ry, double* dx, double* ds, double* dz, double* dy, int32_t* status, void* stream) {
  return kkt_solve<double>(B, nz, nineq, neq, Q, G, A, s, z, eps, n_refine, rx, rs, rz, ry, dx, ds, dz, dy, status, stream);
}
#endif
#if ALMPC_BUILD_F32
int almpc_kkt_solve_f32(int32_t B, int32_t nz, int32_t nineq, int32_t neq, const float* Q, const float* G, const float* A,
                        const float* s, const float* z, double eps, int32_t n_refine, const float* rx, const float* rs,
                        const float* rz, const float* ry, float* dx, float* ds, float* dz, float* dy, int32_t* status, void* stream) {
  return kkt_solve<float>(B, nz, nineq, neq, Q, G, A, s, z, eps, n_refine, rx, rs, rz, ry, dx, ds, dz, dy, status, stream);
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
const char* almpc_last_error(void) { return g_err; }
const char* almpc_version(void) { return "almpc 0.1 (sm_100a)"; }

#endif
}  // extern "C"
#endif  // !ALMPC_BUILD_SE
