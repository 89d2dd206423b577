"""Interior-point path: drop-in for ``qpth.qp.DenseQPFunction`` (reference qpth/qp.py:187-270) and the
primal-dual interior-point loop behind it (qpth/solvers/pdipm/batch_LU.py:29-244) — SURVEY.md §8(f) row 4.

Same factory and call signature as the reference::

    zhat = DenseQPFunction(eps=1e-12, notImprovedLim=3, maxIter=20)(Q, p, G, h, A, b, dyn_res, cost_grad=None)

``Q [B,nz,nz]`` (symmetric), ``p [B,nz]``, ``G [B,nineq,nz]``, ``h [B,nineq]``, ``A [B,neq,nz]``, ``b [B,neq]``;
``dyn_res(x) -> [B,neq]`` is the (possibly NON-linear) equality residual the reference evaluates at every iterate
(batch_LU.py:95; qp_wrapper.py:316-342 passes the true dynamics residual, not ``A x - b``), ``cost_grad(x)`` an
optional replacement of ``Q x + p`` (batch_LU.py:90-92).  Gradients: ``(dQ, dp, dG, dh, dA, db, None, None)`` as in
qp.py:255-268.

What runs where: the reference assembles the dense ``(nz + 2 nineq + neq)^2`` KKT matrix, LU-factors its regularised
copy with partial pivoting at every ``solve_kkt`` (three per iteration) and refines once.  Here every ``solve_kkt`` is
ONE launch of ``almpc_kkt_solve_*`` (csrc/almpc_kkt.cuh): slack and multiplier blocks eliminated in closed form, the
quasi-definite reduced system of size ``nz + neq`` factored as ``L D L^T`` in shared memory by one CTA per problem, the
same refinement step.  The iteration itself (residuals, step lengths, centring, best-iterate bookkeeping and the
batch-global exit tests, which need the host exactly where the reference needs it) stays a handful of torch ops on
the GPU, like the callables ``dyn_res`` / ``cost_grad`` the caller supplies.  CPU tensors raise: no CPU fallback.

Not mirrored: the SQP outer loop of ``qpth/qp_wrapper.py::MPC`` around this function (per-problem Python loops and
roll-out line searches, qp_wrapper.py:345-420) — it holds no arithmetic of its own.
"""
import ctypes

import torch

from . import _lib

KKT_EPS = 1e-7  # regularisation of the KKT matrix (batch_LU.py:42)


def solve_kkt(Q, G, A, s, z, eps, rx, rs, rz, ry, n_refine=1):
    """`solve_kkt(K, Ktilde, rx, rs, rz, ry)` of batch_LU.py:212-244 with K given by its blocks (Q, G, A) and the
    current slacks / multipliers (s, z): -> (dx, ds, dz, dy).  eps = KKT_EPS, n_refine = 1: the regularised Ktilde
    and the single refinement step of the forward loop."""
    lib = _lib.load()
    _lib.require_cuda(Q, "solve_kkt(Q)")
    B, nz = rx.shape
    nineq, neq = G.shape[1], A.shape[1]
    dt = Q.dtype
    cont = lambda t: t.to(dt).contiguous()
    Q, G, A, s, z, rx, rs, rz, ry = map(cont, (Q, G, A, s, z, rx, rs, rz, ry))
    dx, ds, dz, dy = torch.empty_like(rx), torch.empty_like(rs), torch.empty_like(rz), torch.empty_like(ry)
    status = torch.empty(B, dtype=torch.int32, device=Q.device)
    with torch.cuda.device(Q.device):
        rc = getattr(lib, "almpc_kkt_solve_" + _lib.suffix(dt))(
            B, nz, nineq, neq, _lib.ptr(Q), _lib.ptr(G), _lib.ptr(A), _lib.ptr(s), _lib.ptr(z), float(eps), int(n_refine),
            _lib.ptr(rx),
            _lib.ptr(rs), _lib.ptr(rz), _lib.ptr(ry), _lib.ptr(dx), _lib.ptr(ds), _lib.ptr(dz), _lib.ptr(dy),
            _lib.ptr(status), _lib.stream_ptr())
    _lib.check(rc, "almpc_kkt_solve")
    return dx, ds, dz, dy


def get_step(v, dv):
    """batch_LU.py:204-210, including its batch-global `a.max()`."""
    a = -v / dv
    a[dv == 0] = 1.0
    a[dv > 0] = max(1.0, a.max().item())
    return a.min(1)[0]


def pdipm_forward(Q, p, G, h, A, b, dyn_res, cost_grad=None, eps=1e-12, notImprovedLim=3, maxIter=20, trace=None):
    """The loop of batch_LU.py:29-198: Mehrotra predictor-corrector with the reference's initial point, step rule,
    best-iterate bookkeeping and batch-global exits.  -> (x, y, z, s) of the best iterate per problem."""
    B, nineq, nz = G.shape
    neq = A.shape[1]
    zeros = lambda n: torch.zeros(B, n, dtype=Q.dtype, device=Q.device)
    ones = torch.ones(B, nineq, dtype=Q.dtype, device=Q.device)
    GT, AT = G.transpose(1, 2), A.transpose(1, 2)
    bmv = lambda M, v: torch.bmm(M, v.unsqueeze(-1)).squeeze(-1)
    # initial point (batch_LU.py:46-80): K with S = Z = I
    x, s, z, y = solve_kkt(Q, G, A, ones, ones, KKT_EPS, p, zeros(nineq), -h, -b)
    for v in (s, z):
        M = torch.min(v, 1)[0][:, None].repeat(1, nineq)
        I = M < 0
        v[I] -= M[I] - 1
    best = None
    nNotImproved = 0
    for i in range(maxIter):
        rx = bmv(AT, y) + bmv(GT, z) + (bmv(Q, x) + p if cost_grad is None else cost_grad(x))
        rs = s * z
        rz = bmv(G, x) + s - h
        ry = dyn_res(x)
        mu = torch.abs((s * z).sum(1) / nineq)
        pri_resid = torch.norm(ry, 2, 1) + torch.norm(rz, 2, 1)
        dual_resid = torch.norm(rx, 2, 1)
        resids = pri_resid + dual_resid + nineq * mu
        if best is None:
            best = dict(resids=resids, x=x.clone(), z=z.clone(), s=s.clone(), y=y.clone())
            nNotImproved = 0
        else:
            I = resids < best["resids"]
            if I.sum().item() > 0:
                nNotImproved = 0
            else:
                nNotImproved += 1
            best["resids"][I] = resids[I]
            for k, cur in (("x", x), ("z", z), ("s", s), ("y", y)):
                best[k][I] = cur[I]
        if trace is not None:
            trace.append(dict(resids=resids.detach().clone(), mu=mu.detach().clone()))
        if nNotImproved == notImprovedLim or best["resids"].max().item() < eps or mu.min().item() > 1e32:
            break
        if trace is not None:
            trace[-1]["solved"] = True  # this iterate is followed by the two KKT solves of a full iteration
        dx_aff, ds_aff, dz_aff, dy_aff = solve_kkt(Q, G, A, s, z, KKT_EPS, rx, rs, rz, ry)
        alpha = torch.min(torch.min(get_step(z, dz_aff), get_step(s, ds_aff)), torch.ones(B, dtype=Q.dtype, device=Q.device))
        a_in = alpha[:, None]
        t3 = torch.sum((s + a_in * ds_aff) * (z + a_in * dz_aff), 1)
        t4 = torch.sum(s * z, 1)
        sig = (t3 / t4) ** 3
        rs_cor = (-mu * sig)[:, None] + ds_aff * dz_aff
        dx_cor, ds_cor, dz_cor, dy_cor = solve_kkt(Q, G, A, s, z, KKT_EPS, zeros(nz), rs_cor, zeros(nineq), zeros(neq))
        dx, ds, dz, dy = dx_aff + dx_cor, ds_aff + ds_cor, dz_aff + dz_cor, dy_aff + dy_cor
        alpha = torch.min(0.999 * torch.min(get_step(z, dz), get_step(s, ds)), torch.ones(B, dtype=Q.dtype, device=Q.device))
        a_in = alpha[:, None]
        x = x + a_in * dx
        s = s + a_in * ds
        z = z + a_in * dz
        y = y + a_in * dy
    return best["x"], best["y"], best["z"], best["s"]


def bger(x, y):
    return x.unsqueeze(2).bmm(y.unsqueeze(1))


def DenseQPFunction(bsz=1, eps=1e-12, verbose=0, notImprovedLim=3, maxIter=20):
    """Factory with the reference's keywords (qp.py:187-193; `bsz` and `verbose` are unused there too)."""

    class Solver(torch.autograd.Function):
        @staticmethod
        def forward(ctx, Q, p, G, h, A, b, dyn_res, cost_grad=None):
            _lib.require_cuda(Q, "DenseQPFunction(Q)")
            with torch.no_grad():
                zhats, nus, lams, slacks = pdipm_forward(Q, p, G, h, A, b, dyn_res, cost_grad, eps, notImprovedLim, maxIter)
            ctx.save_for_backward(zhats, nus, lams, slacks, Q, G, A)
            return zhats

        @staticmethod
        def backward(ctx, dl_dzhat):
            zhats, nus, lams, slacks, Q, G, A = ctx.saved_tensors
            B, nineq, neq = Q.shape[0], G.shape[1], A.shape[1]
            z0 = lambda n: torch.zeros(B, n, dtype=Q.dtype, device=Q.device)
            # qp.py:249-253 solves with K itself at the best iterate's (z, s).  Same solution here from the regularised
            # factorisation + three refinement steps against K (csrc/almpc_kkt.cuh explains why not eps = 0).
            dx, _, dlam, dnu = solve_kkt(Q, G, A, slacks, lams, KKT_EPS, dl_dzhat.to(Q.dtype), z0(nineq), z0(nineq),
                                         z0(neq), n_refine=3)
            dps = dx
            dGs = bger(dlam, zhats) + bger(lams, dx)
            dhs = -dlam
            dAs = bger(dnu, zhats) + bger(nus, dx)
            dbs = -dnu
            dQs = 0.5 * (bger(dx, zhats) + bger(zhats, dx))
            return dQs, dps, dGs, dhs, dAs, dbs, None, None

    return Solver.apply
