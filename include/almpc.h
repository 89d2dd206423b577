/*
 * libalmpc — C ABI of the B200-native batched augmented-Lagrangian MPC solver.
 *
 * This is the drop-in boundary for the reference's AL-MPC hot path.  Every entry point takes plain
 * device pointers + sizes + a CUDA stream; there are no torch types, no allocations and no host
 * synchronisation inside the library (all calls are CUDA-graph capturable), and no global state: everything a
 * call needs travels in its descriptor (thread-safe, re-entrant).  Return value: 0 = ok,
 * <0 = bad argument, >0 = cudaError_t of the failing runtime call; almpc_last_error() gives text.
 *
 * What each entry point replaces in the reference (paths relative to the reference repo):
 *
 *   almpc_forward_{f64,f32}   qpth/AL_mpc.py:260-339 (MPC.al_solve) and :342-423 (al_solve_stream):
 *                             the whole al_iter x (<=4 Newton steps) loop, i.e. every
 *                             al_utils.NewtonAL.apply call (qpth/al_utils.py:451-576) with its
 *                             merit_grad_hessian (:80-123), cholesky/cholesky_solve (:510-515),
 *                             line_search_newton (:618-642) and the dual / penalty update
 *                             (AL_mpc.py:315-325), plus the fp32 output cast (AL_mpc.py:337-338).
 *   almpc_backward_{f64,f32}  al_utils.NewtonAL.backward (qpth/al_utils.py:578-615).
 *   almpc_dynamics_{f64,f32}  the dx / dx_jac plug-ins: RexQuadrotor_dynamics(_jac).forward
 *                             (deqmpc/rex_quadrotor.py:109-110,136-144), FlyingCartpole_dynamics(_jac)
 *                             (deqmpc/flying_cartpole2d.py:93-94,139-148) and the native
 *                             pkg.dynamics / pkg.derivatives extensions
 *                             (deqmpc/my_envs/<pkg>/src/dynamics.cpp:13-54).
 *
 * Layouts (all contiguous, row-major, S = double or float):
 *   n = nx+nu, N = T*n, neq = T*nx, nineq = T*(2*nu + n_obs), m = neq+nineq
 *   x0[B][nx]   Qd[B][T][n]   q[B][T][n]   xu[B][T][n] (x_t then u_t per stage)
 *   lam[B][m]   eq rows first: (T-1) dynamics blocks of nx, then the x_0 - x0 block (al_utils.py:224);
 *               then per stage t: nu upper-bound rows, nu lower-bound rows (al_utils.py:293), n_obs obstacle rows
 *   rho[B]      penalty per problem, stored in S
 *   factor[B][almpc_factor_elems]  block-tridiagonal Cholesky factor of the last Newton Hessian:
 *               T lower-triangular blocks L_tt[n][n] (diagonal holds 1/L_ii) followed by
 *               (T-1) sub-diagonal blocks E_t[nx][n] (rows = x_{t+1}, the only non-zero rows)
 */
#ifndef ALMPC_H_
#define ALMPC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* native dynamics models (device functors fused into the solver) */
enum almpc_dyn_id {
  ALMPC_DYN_PENDULUM = 0,        /* my_envs/pendulum1l   nx=2  nu=1 */
  ALMPC_DYN_CARTPOLE1L = 1,      /* my_envs/cartpole1l   nx=4  nu=1 */
  ALMPC_DYN_REX_QUADROTOR = 2,   /* rex_quadrotor.py     nx=12 nu=4 */
  ALMPC_DYN_FLYING_CARTPOLE = 3, /* flying_cartpole2d.py nx=14 nu=4 */
  ALMPC_DYN_CARTPOLE2L = 4,      /* my_envs/cartpole2l   nx=6  nu=1 */
  ALMPC_DYN_COUNT = 5
};

#define ALMPC_MAX_PARAMS 48
#define ALMPC_MAX_NU 8
#define ALMPC_MAX_AL_ITER 32
#define ALMPC_N_LS 20 /* line-search candidates 2^-k, k=0..19 (al_utils.py:619) */

/* desc.flags: leave lam/rho untouched at the end (NewtonAL-only use: al_utils.py:449 has no dual update) */
#define ALMPC_FLAG_SKIP_FINAL_DUAL 1
/* desc.flags: MPC(state_estimator=True) (qpth/AL_mpc.py:186-199,426-442,508-509 -> qpth/al_utils_se.py): the
 * constraints are the T-1 dynamics blocks only (lam has (T-1)*nx entries), the controls are data (their
 * Jacobian columns are zeroed, al_utils_se.py:153; u_lower/u_upper are ignored) and the merit's cost and
 * its gradient see the states only (al_utils_se.py:272-305); the Hessian keeps diag(Q) on the controls. */
#define ALMPC_FLAG_STATE_ESTIMATOR 2

/* desc.select: kernel-selection switches for A/B measurements and differential tests (0 = the library's choice) */
#define ALMPC_SELECT_FAC_SINGLE 1 /* row-per-lane factor kernel even where the half-warp-per-problem one applies */
#define ALMPC_SELECT_FUSE 2       /* factor + line search + steering in ONE kernel (k_newton_pair) where it applies */
#define ALMPC_SELECT_LS_CTA 4     /* CTA line-search kernel even where the warp-synchronous one applies */
#define ALMPC_SELECT_PDL 8        /* launch the step kernels with programmatic stream serialization (default for B <= 16384) */
#define ALMPC_SELECT_NO_PDL 32    /* ... never */
#define ALMPC_SELECT_FAC_PAIR 64  /* half-warp-per-problem factor kernel even where the fragment-layout one (k_factor_solve_mma: blocked Cholesky on FP64 MMAs, the default for 16-wide blocks) applies */
#define ALMPC_SELECT_FAC_MMA 128  /* the fragment-layout factor kernel for fp32 problems too (blocks kept in FP64 accumulators) */
#define ALMPC_SELECT_NO_RING 16   /* every Newton step streams its block factor to `factor` (no L2-resident scratch ring) */

/* Hook for the reference's batch-GLOBAL decisions under batch sharding (qpth/al_utils.py:486,552,560-564: the
 * Newton stop rule is a statistic of the WHOLE batch; qpth/AL_mpc.py:412: max(rho) over the batch).  Called on the
 * host, at enqueue time, once per Newton-step slot: it must enqueue on `stream` an in-place reduction over all
 * shards of stats[0], stats[1] (SUM) and stats[2] (MAX) — e.g. two NCCL all-reduces — and return 0.  No host
 * synchronisation is required or performed. */
typedef int (*almpc_stop_reduce_fn)(double* stats_dev, void* stream, void* user);

typedef struct almpc_desc {
  int32_t B;          /* problems in the batch */
  int32_t T;          /* horizon */
  int32_t dyn_id;     /* enum almpc_dyn_id */
  int32_t al_iter;    /* AL iterations (MPC.al_iter, AL_mpc.py:290) */
  int32_t max_newton; /* Newton steps per AL iteration, reference value 4 (al_utils.py:485) */
  int32_t stream_mode;/* 1 = al_solve_stream semantics: stop when max(rho) > rho_max (AL_mpc.py:412) */
  int32_t rho_f32;    /* 1 = round rho to fp32 after each x10 (rho_prev is an fp32 tensor, AL_mpc.py:573) */
  int32_t flags;      /* ALMPC_FLAG_* */
  double rho_mult;    /* 10 (AL_mpc.py:325) */
  double rho_max;     /* 1e8 (AL_mpc.py:193) */
  double dyn_params[ALMPC_MAX_PARAMS]; /* model parameters, layout in DESIGN.md / dynamics.py */
  double u_lower[ALMPC_MAX_NU];
  double u_upper[ALMPC_MAX_NU];
  /* Obstacle_MPC (qpth/AL_mpc_custom.py:22-135): n_obs sphere constraints r^2 - |p_t - c_k|^2 <= 0 per stage on
   * the position x[0..2]; rows follow the 2*nu control rows of each stage (al_utils.py:313-323).  obstacles is a
   * DEVICE pointer to [B][T][n_obs][3] scalars of the solver dtype (NULL / n_obs = 0: plain MPC). */
  int32_t n_obs;
  int32_t select;     /* ALMPC_SELECT_* */
  double obs_radius;
  const void* obstacles;
  /* Optional per-kernel timing of THIS call: HOST pointer to 8 doubles or NULL.  When set, almpc_forward_* brackets
   * every launch with CUDA events and SYNCHRONISES at the end (never set it on a timed or graph-captured path):
   * [0..3] = total ms of k_jacobian, factor (k_factor_solve* or the fused k_newton_pair), k_line_search*,
   * k_final_dual (no-op launches after the solve finished excluded), [4..7] = launches counted. */
  double* profile_ms;
  /* Batch-global stop rule across shards (NULL = every shard applies the rule to its own problems):
   * stop_stats = DEVICE buffer of 4 doubles owned by the caller, stop_reduce = the reduction hook above. */
  double* stop_stats;
  almpc_stop_reduce_fn stop_reduce;
  void* stop_reduce_user;
} almpc_desc;

/* info_out layout (int32, device memory, ALMPC_INFO_INTS entries) */
#define ALMPC_INFO_AL_DONE 0       /* AL iterations executed */
#define ALMPC_INFO_RHO_EXCEEDED 1  /* final max(rho) > rho_max (stream-mode status flag, AL_mpc.py:420) */
#define ALMPC_INFO_TOTAL_NEWTON 2  /* Newton steps executed in total */
#define ALMPC_INFO_NEWTON0 4       /* [4 + a] = Newton steps of AL iteration a */
#define ALMPC_INFO_INTS (4 + ALMPC_MAX_AL_ITER)

/* per-problem status bits written to status_out */
#define ALMPC_STATUS_FINITE 1   /* iterate stayed finite (al_utils.py:545-549) */
#define ALMPC_STATUS_CHOL_OK 2  /* every pivot of every factorisation was positive */

int almpc_model_dims(int32_t dyn_id, int32_t* nx, int32_t* nu);
size_t almpc_factor_elems(const almpc_desc* d);          /* scalars per problem in `factor` */
size_t almpc_workspace_bytes(const almpc_desc* d);       /* device scratch the caller must provide */
size_t almpc_debug_elems(const almpc_desc* d);           /* scalars per problem in `debug_out` */

/* Forward solve.  In/out: xu (warm start -> solution), lam, rho (dual warm start -> updated duals).
 * Out: factor (for backward), x_out/u_out fp32 (may be NULL), dyn_res_out[B] = ||c+|| per problem
 * (may be NULL), status_out[B] (may be NULL), info_out[ALMPC_INFO_INTS] (device),
 * debug_out (may be NULL): per problem [grad N | direction N | merits 20 | argmin | accept | merit_in]
 * of the last executed Newton step. */
int almpc_forward_f64(const almpc_desc* d, const double* x0, const double* Qd, const double* q,
                      double* xu, double* lam, double* rho, double* factor, float* x_out, float* u_out,
                      double* dyn_res_out, int32_t* status_out, int32_t* info_out, double* debug_out,
                      void* workspace, void* stream);
int almpc_forward_f32(const almpc_desc* d, const float* x0, const float* Qd, const float* q,
                      float* xu, float* lam, float* rho, float* factor, float* x_out, float* u_out,
                      float* dyn_res_out, int32_t* status_out, int32_t* info_out, float* debug_out,
                      void* workspace, void* stream);

/* Implicit backward: gbar = -H^{-1} grad_xu with the cached factor; grad_q = gbar,
 * grad_Q = gbar * xu (grad_Q may be NULL when Q needs no gradient).
 * status = the forward's status_out or NULL.  With it, a problem whose factor is flagged broken
 * (ALMPC_STATUS_CHOL_OK cleared) gets a zero gradient and every other problem the plain result — a NaN / Inf in
 * grad_xu propagates as in the reference.  Without it, any non-finite result is replaced by zero. */
int almpc_backward_f64(const almpc_desc* d, const double* factor, const double* xu, const double* grad_xu,
                       const int32_t* status, double* grad_Q, double* grad_q, void* stream);
int almpc_backward_f32(const almpc_desc* d, const float* factor, const float* xu, const float* grad_xu,
                       const int32_t* status, float* grad_Q, float* grad_q, void* stream);

/* Stand-alone dynamics: xn[R][nx] = f(x[R][nx], u[R][nu]); optional Jx[R][nx][nx], Ju[R][nx][nu]
 * (Jx[r][i][j] = d xn_i / d x_j, al_utils.py:237-248).  Uses d->dyn_id / d->dyn_params only. */
int almpc_dynamics_f64(const almpc_desc* d, int64_t R, const double* x, const double* u, double* xn,
                       double* Jx, double* Ju, void* stream);
int almpc_dynamics_f32(const almpc_desc* d, int64_t R, const float* x, const float* u, float* xn,
                       float* Jx, float* Ju, void* stream);

/* Tracking-cost build of the caller (deqmpc/policies.py:1285-1297 compute_pf, with the optional per-stage
 * q_scaling of :1245-1248), fused: Qd[b,t,j] = Q0[j] (1 + scale[b,t]), q = -Qd * [x_ref | u_ref].  x_ref / u_ref /
 * scale are the fp32 network outputs (scale may be NULL).  The backward maps the solver's grad_Q / grad_q to
 * grad_x_ref, grad_u_ref, grad_scale (any of the three outputs, and grad_Q, may be NULL). */
int almpc_tracking_cost_f64(int32_t B, int32_t T, int32_t nx, int32_t nu, const double* Q0, const float* x_ref,
                            const float* u_ref, const float* scale, double* Qd, double* q, void* stream);
int almpc_tracking_cost_f32(int32_t B, int32_t T, int32_t nx, int32_t nu, const float* Q0, const float* x_ref,
                            const float* u_ref, const float* scale, float* Qd, float* q, void* stream);
int almpc_tracking_cost_bwd_f64(int32_t B, int32_t T, int32_t nx, int32_t nu, const double* Q0, const float* x_ref,
                                const float* u_ref, const float* scale, const double* grad_Q, const double* grad_q,
                                float* grad_x_ref, float* grad_u_ref, float* grad_scale, void* stream);
int almpc_tracking_cost_bwd_f32(int32_t B, int32_t T, int32_t nx, int32_t nu, const float* Q0, const float* x_ref,
                                const float* u_ref, const float* scale, const float* grad_Q, const float* grad_q,
                                float* grad_x_ref, float* grad_u_ref, float* grad_scale, void* stream);

/* Generic-callable dynamics mode: when `dx` / `dx_jac` are arbitrary host-language callables
 * (qpth/al_utils.py:237-248 convention), the host evaluates them itself and calls this for the part of the Newton
 * step the reference does with dense [N,N] LAPACK calls (al_utils.py:510-513 forward, :586-590 backward):
 *   H = diag(diagD) + rho * (G_t^T G_t on the stage blocks, -rho G_t below them),   sol = -H^{-1} rhs
 * Run-time dimensions, nx + nu <= 32.  G [B][T-1][nx][n] with G_t = [d f/d x | d f/d u] at stage t;
 * diagD [B][T][n] carries EVERY diagonal term (diag(Q), rho on the states coupled by an identity block, rho per
 * active bound); factor has the layout of almpc_forward_* (almpc_factor_elems with the same nx, nu, T).
 * do_factor = 1: build + factor + solve (G, diagD, rho required; status bit ALMPC_STATUS_CHOL_OK is cleared on
 * a non-positive pivot and sol = 0);  do_factor = 0: solve with an existing factor (backward pass). */
int almpc_blocktri_f64(int32_t B, int32_t T, int32_t nx, int32_t nu, int32_t do_factor, const double* G,
                       const double* diagD, const double* rho, double* factor, const double* rhs, double* sol,
                       int32_t* status, void* stream);
int almpc_blocktri_f32(int32_t B, int32_t T, int32_t nx, int32_t nu, int32_t do_factor, const float* G,
                       const float* diagD, const float* rho, float* factor, const float* rhs, float* sol,
                       int32_t* status, void* stream);

/* Interior-point path: the KKT solve of the reference's primal-dual interior-point QP solver,
 * `solve_kkt` of qpth/solvers/pdipm/batch_LU.py:212-244 (called from its forward loop :70,155,175 and from
 * DenseQPFunction.backward, qpth/qp.py:249-253).  Unknowns (x[nz], s[nineq], z[nineq], y[neq]) of
 *     Ktilde l = -(rx, rs, rz, ry),   K = [Q 0 G^T A^T; 0 D(z) D(s) 0; G I 0 0; A 0 0 0],
 *     Ktilde = K + eps diag(+1 on the x and s rows, -1 on the z and y rows)        (batch_LU.py:42,58-60,113-116)
 * followed by n_refine steps of iterative refinement against K (the reference takes one).  The slack / multiplier blocks are
 * eliminated in closed form and the quasi-definite reduced system of size nz + neq is factored as L D L^T in shared
 * memory by one CTA per problem (instead of a pivoted LU of the dense (nz + 2 nineq + neq)^2 matrix).  The backward pass
 * (reference: an LU of K itself) uses eps = 1e-7 with three refinement steps: eliminating the slack block of the
 * unregularised K would put z_i / s_i ~ 1e12 of the active rows into the reduced matrix.  Q [B][nz][nz] symmetric, G [B][nineq][nz], A [B][neq][nz], s, z [B][nineq];
 * (nz + neq)^2 scalars must fit the opt-in shared memory (error -11 otherwise).  status[B] (may be NULL): 1 = all
 * pivots finite and non-zero. */
int almpc_kkt_solve_f64(int32_t B, int32_t nz, int32_t nineq, int32_t neq, const double* Q, const double* G, const double* A,
                        const double* s, const double* z, double eps, int32_t n_refine, const double* rx, const double* rs,
                        const double* rz, const double* ry, double* dx, double* ds, double* dz, double* dy, int32_t* status, void* stream);
int almpc_kkt_solve_f32(int32_t B, int32_t nz, int32_t nineq, int32_t neq, const float* Q, const float* G, const float* A,
                        const float* s, const float* z, double eps, int32_t n_refine, const float* rx, const float* rs,
                        const float* rz, const float* ry, float* dx, float* ds, float* dz, float* dy, int32_t* status, void* stream);

/* FMA-pipe micro-benchmark used by bench.py to measure the FP64 / FP32 roofline denominators:
 * launches `blocks` x 256 threads, each doing `iters` x 8 independent FMAs; returns via *flops the
 * number of floating-point operations performed (2 per FMA). */
int almpc_fma_peak_f64(int32_t blocks, int32_t iters, double* sink, double* flops, void* stream);
int almpc_fma_peak_f32(int32_t blocks, int32_t iters, float* sink, double* flops, void* stream);

const char* almpc_last_error(void);
const char* almpc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ALMPC_H_ */
