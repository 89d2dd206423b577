// almpc_common.cuh: build switches, control block, kernel parameters, small device helpers and the
// shared-memory layouts used by every kernel of libalmpc (see almpc_kernels.cu for the overview).
#pragma once
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
// The fp32 solver assembles its Hessian blocks on the FP64 tensor-core path too (operands widened, fp64 accumulation,
// result rounded to fp32): no full-precision fp32 MMA exists, and the scalar triangle loop it replaces cost 5x the
// instructions.  Measured (quadrotor 64K fp32): factor kernel 0.715 -> 0.553 ms, 7.38 -> 8.46 M solves/s; the block
// is MORE accurate than the reference's fp32 bmm, and the results stay inside the reference's own fp32 noise floor
// (q_grad vs the fp32 golden 5.4e-4, before 7.0e-4).  0 restores the scalar fp32 assembly.
#ifndef ALMPC_F32_DMMA
#define ALMPC_F32_DMMA 1
#endif
// Backward sweep of the pair factor kernel: L / E blocks of stage t-1 by cp.async.bulk + mbarrier (one instruction
// per 2 KB block, issued by one lane) instead of 8 + 12 per-lane cp.async.  Built, bit-identical, and measured on B200
// (quadrotor 64K fp64): the per-lane 16-byte cp.async costs 240 shared-memory wavefronts per problem where 80 are
// ideal (ncu source page, profiles/r02_factor_pair_smem_by_line.txt) and the bulk copy removes them — but the kernel
// goes from 0.885 to 1.10 ms (fp32 0.553 -> 0.625): the copy has to hide behind ONE 16-step substitution (~700
// cycles), and the round trip of a bulk copy through the TMA unit is longer than that of the L2-hitting cp.async it
// replaces.  Off by default; -DALMPC_BULK_COPY=1 builds it.
#ifndef ALMPC_BULK_COPY
#define ALMPC_BULK_COPY 0
#endif
#ifndef ALMPC_LS_MINB
#define ALMPC_LS_MINB 4  // CTAs/SM the warp line search is compiled for
#endif
#ifndef ALMPC_JAC_MINB
#define ALMPC_JAC_MINB 2  // CTAs/SM the Jacobian kernel is compiled for (3 forces spills: measured slower)
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
  int part_off, ticket_total;  // line search: slot offset of the CTA partials / CTAs that take a ticket
  // Scratch ring for block factors that only this launch reads (the backward sweep of the same kernel): every Newton
  // step of an AL iteration that cannot be the last one.  A CTA claims one of kRingSlots slots of its SM (bit mask in
  // ring_mask[smid % kRingSMs]) and writes its problems' factors there instead of into the caller's `factor` buffer;
  // the slots are overwritten ~20 times per launch while their lines are still in L2, so most of these 16 KB per
  // problem never reach HBM.  NULL: every step streams its factor to `factor` (what round 1 did).
  S* ring;
  int* ring_mask;
  int global_stop;             // 1: the last CTA publishes this shard's statistics instead of steering (k_steer does)
  double* stop_stats;          // device [4]: sum ||c+||^2 at entry, at exit, max rho — reduced over the shards by the caller
  int n_obs;                   // obstacle rows per stage (Obstacle_MPC), 0 = none
  S obs_r2;                    // obstacle radius squared
  const S* obs;                // [B][T][n_obs][3] obstacle centres (device)
};

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
// 1/sqrt(x): the FP64 MUFU seed (rsqrt.approx.ftz.f64: ~20 correct bits, one instruction, no conversions) + one
// cubically convergent correction in fp64 (~1 ulp):   r = 1 - x y^2 ;  x^-1/2 = y (1 + r/2 + 3 r^2/8 + O(r^3)).
// A non-positive pivot yields NaN / Inf, which the callers detect on the finished solution.
__device__ __forceinline__ double rsqrt_(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double r = fma(-x * y, y, 1.0);
  const double p = fma(0.375, r, 0.5);
  return fma(y * r, p, y);
}
__device__ __forceinline__ float rsqrt_(float x) { return rsqrtf(x); }
// max(x, 0) that propagates NaN like torch.clamp(min=0) (AL_mpc.py:317, al_utils.py:419)
template <typename S> __device__ __forceinline__ S fmax0(S x) { return x < S(0) ? S(0) : x; }
// one entry of the quadratic cost (al_utils.py:425-435), 1/2 Q z^2 + q z as z (1/2 (Q z) + q): one multiply and two
// fused multiply-adds.  Every kernel that evaluates a merit uses THIS form, so entry merit and candidates agree.
template <typename S> __device__ __forceinline__ S cost_term(S z, S Q, S q) { return z * (S(0.5) * (Q * z) + q); }
template <typename S> __device__ __forceinline__ bool finite_(S x) { return (x - x) == S(0); }

constexpr int kRingSMs = 160;   // mask words (>= SMs of the device; %smid is folded into this range)
constexpr int kRingSlots = 8;   // scratch slots per mask word

// Claims a scratch slot for this CTA (one thread calls it): slot index, or -1 when the ring is off for this launch or
// all slots of this SM are taken (the CTA then simply uses the caller's factor buffer — never waits).
template <typename S>
__device__ __forceinline__ int ring_claim(const KParams<S>& P, const Ctl* ctl) {
  if (!P.ring || P.stream_mode || ctl->a + 1 >= P.al_iter) return -1;  // this AL iteration may be the last: keep the factor
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  int* m = P.ring_mask + (smid % kRingSMs);
  for (int tries = 0; tries < 4; ++tries) {
    const int old = *(volatile int*)m;
    const int bit = __ffs(~old) - 1;
    if (bit < 0 || bit >= kRingSlots) return -1;
    if (atomicCAS(m, old, old | (1 << bit)) == old) return (int)(smid % kRingSMs) * kRingSlots + bit;
  }
  return -1;
}
template <typename S>
__device__ __forceinline__ void ring_release(const KParams<S>& P, int slot) {
  atomicAnd(P.ring_mask + slot / kRingSlots, ~(1 << (slot % kRingSlots)));
}

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
// Bulk asynchronous copies (sm_90+ / sm_100a: cp.async.bulk, executed by the TMA unit, completion through an mbarrier):
// ONE instruction moves a whole contiguous block global -> shared with full-line shared-memory writes, where the
// per-lane 16-byte cp.async above costs one instruction per lane and 16 bytes and — measured with ncu on the factor
// kernel — three shared-memory wavefronts per ideal one (its data lands sector by sector).
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// orders this thread's earlier generic-proxy accesses (ordinary loads / stores, shared and global) before later
// async-proxy operations (the bulk copies)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

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
  int xu, dv, w, lin, qd, Gt, Ep, scr, total;
  __host__ __device__ Layout(int T, int nobs = 0, bool se = false, bool staged = false) {
    // inequality rows per stage: control bounds, then obstacle rows.  State-estimator variant
    // (al_utils_se.py): the T-1 dynamics blocks only
    SI = se ? 0 : 2 * NU + nobs;
    N = T * NN; neq = se ? (T - 1) * NX : T * NX; nin = T * SI; m = neq + nin;
    int o = 0;
    auto take = [&](int cnt) { int at = o; o += (cnt + 3) & ~3; return at; };  // regions 16-byte aligned for float and double
    xu = take(N);
    dv = take(N);
    w = take(neq);
    // pair kernel only: inequality multipliers and the cost diagonal are read once from HBM in the prologue and
    // then per stage from shared memory (a global load inside the stage loop stalls the whole chain)
    lin = staged ? take(nin) : 0;
    qd = staged ? take(N) : 0;
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

}  // namespace almpc
