"""Generic-callable dynamics mode: the AL-MPC solve for an arbitrary user ``dx`` / ``dx_jac``.

The fused kernels are compiled per dynamics functor; a Python callable cannot run inside them.  This mode keeps
the drop-in valid for such callables: the host loop below follows ``NewtonAL.forward`` / ``MPC.al_solve``
(qpth/al_utils.py:451-576, qpth/AL_mpc.py:260-423) step by step, evaluates the user's callables with torch on the
GPU exactly where the reference does (one ``dx_jac`` call per Newton step, one batched ``dx`` call over the 20
line-search candidates, al_utils.py:618-642), never materialises the dense ``J[B,m,N]`` / ``H[B,N,N]`` — gradient
and Hessian are assembled block-wise — and hands the linear algebra (the reference's ``cholesky_ex`` +
``cholesky_solve`` on ``[N,N]``, al_utils.py:510-513, 586-590) to ``almpc_blocktri_*``, the run-time-dimension
block-tridiagonal CUDA kernel.  Like the reference it synchronises with the host once per Newton step (the
batch-global stop rule, al_utils.py:560-564).

CUDA tensors only; there is no CPU path.  ``dx(x[R,nx], u[R,nu]) -> x+[R,nx]`` and
``dx_jac(x, u) -> (x+, (Jx[R,nx,nx], Ju[R,nx,nu]))`` with ``Jx[r,i,j] = d x+_i / d x_j`` (al_utils.py:237-248).
"""
from dataclasses import dataclass

import torch

from . import _lib

N_LS = 20  # al_utils.py:619


@dataclass
class GenericConfig:
    dx: object
    dx_jac: object
    nx: int
    nu: int
    T: int
    u_lower: torch.Tensor
    u_upper: torch.Tensor
    al_iter: int = 2
    max_newton: int = 4
    stream_mode: bool = False
    rho_f32: bool = False
    rho_mult: float = 10.0
    rho_max: float = 1e8
    state_estimator: bool = False


def _residuals(cfg, xu, x0):
    """c and c+ for xu [..., B, T, n] (leading candidate dims allowed).  al_utils.py:209-226, 288-326, 407-421."""
    nx, nu, T = cfg.nx, cfg.nu, cfg.T
    x, u = xu[..., :nx], xu[..., nx:]
    lead = xu.shape[:-2]
    xn = cfg.dx(x[..., :-1, :].reshape(-1, nx), u[..., :-1, :].reshape(-1, nu)).to(xu.dtype).reshape(*lead, T - 1, nx)
    c_dyn = (x[..., 1:, :] - xn).reshape(*lead, -1)
    if cfg.state_estimator:  # al_utils_se.py:247-258
        return c_dyn, c_dyn
    c_init = x[..., 0, :] - x0
    c_in = torch.cat([u - cfg.u_upper, -u + cfg.u_lower], dim=-1).reshape(*lead, -1)
    return torch.cat([c_dyn, c_init, c_in], -1), torch.cat([c_dyn, c_init, c_in.clamp(min=0)], -1)


def _merit(cfg, xu, x0, Qd, q, lam, rho):
    """al_utils.py:52-77 (the constant f is not part of the merit); xu [..., B, T, n] -> [..., B]."""
    c, cp = _residuals(cfg, xu, x0)
    z, Qz, qz = xu, Qd, q
    if cfg.state_estimator:  # al_utils_se.py:272-289: states only
        z, Qz, qz = xu[..., : cfg.nx], Qd[..., : cfg.nx], q[..., : cfg.nx]
    cost = (0.5 * (z * Qz * z).sum(-1) + (qz * z).sum(-1)).sum(-1)
    return cost + 0.5 * rho[:, 0] * (cp * cp).sum(-1) + (lam * c).sum(-1)


def _gradient_blocks(cfg, xu, x0, Qd, q, lam, rho):
    """Block-wise al_utils.py:80-123: gradient [B,T,n], G [B,T-1,nx,n], Hessian diagonal terms [B,T,n]."""
    nx, nu, T = cfg.nx, cfg.nu, cfg.T
    B = xu.shape[0]
    x, u = xu[..., :nx], xu[..., nx:]
    xn, (Jx, Ju) = cfg.dx_jac(x[:, :-1].reshape(-1, nx), u[:, :-1].reshape(-1, nu))
    Jx = Jx.to(xu.dtype).reshape(B, T - 1, nx, nx)
    Ju = Ju.to(xu.dtype).reshape(B, T - 1, nx, nu)
    if cfg.state_estimator:
        Ju = Ju * 0.0  # al_utils_se.py:153
    G = torch.cat([Jx, Ju], dim=-1).contiguous()
    c_dyn = x[:, 1:] - xn.to(xu.dtype).reshape(B, T - 1, nx)
    r3 = rho[:, :, None]
    w = lam[:, : (T - 1) * nx].reshape(B, T - 1, nx) + r3 * c_dyn
    g = Qd * xu + q
    dD = Qd.clone()
    if cfg.state_estimator:
        g[..., nx:] = 0
    g[:, :-1] -= torch.einsum("btij,bti->btj", G, w)
    g[:, 1:, :nx] += w
    dD[:, 1:, :nx] += r3
    if not cfg.state_estimator:
        neq = T * nx
        g[:, 0, :nx] += lam[:, (T - 1) * nx : neq] + rho * (x[:, 0] - x0)
        dD[:, 0, :nx] += rho
        cu, cl = u - cfg.u_upper, cfg.u_lower - u
        li = lam[:, neq:].reshape(B, T, 2 * nu)
        g[..., nx:] += (li[..., :nu] - li[..., nu:]) + r3 * (cu.clamp(min=0) - cl.clamp(min=0))
        dD[..., nx:] += r3 * ((cu >= 0).to(xu.dtype) + (cl >= 0).to(xu.dtype))
    return g.contiguous(), G, dD.contiguous()


def _blocktri(cfg, B, do_factor, G, dD, rho, factor, rhs, status):
    lib = _lib.load()
    sol = torch.empty_like(rhs)
    with torch.cuda.device(rhs.device):
        rc = getattr(lib, "almpc_blocktri_" + _lib.suffix(rhs.dtype))(
            B, cfg.T, cfg.nx, cfg.nu, int(do_factor), _lib.ptr(G), _lib.ptr(dD), _lib.ptr(rho), _lib.ptr(factor),
            _lib.ptr(rhs), _lib.ptr(sol), _lib.ptr(status), _lib.stream_ptr())
    _lib.check(rc, "almpc_blocktri")
    return sol


def _factor_elems(cfg):
    n = cfg.nx + cfg.nu
    return cfg.T * n * n + (cfg.T - 1) * cfg.nx * n


def forward(cfg, x0, Qd, q, xu, lam, rho):
    """One MPC.al_solve / al_solve_stream with user callables.  rho is [B] in the solver dtype.
    Returns the same dictionary of outputs as ``solver.raw_forward`` plus the updated (xu, lam, rho)."""
    for t, name in ((x0, "x0"), (Qd, "Qd"), (q, "q"), (xu, "xu"), (lam, "lam"), (rho, "rho")):
        _lib.require_cuda(t, name)
    B, T, n = xu.shape
    nx = cfg.nx
    if T != cfg.T or n != nx + cfg.nu:
        raise _lib.AlmpcError(f"xu has shape {tuple(xu.shape)}, expected [B,{cfg.T},{nx + cfg.nu}]")
    if n > 32:
        raise _lib.AlmpcError("generic-callable mode supports n_state + n_ctrl <= 32")
    dev, dt = xu.device, xu.dtype
    rho = rho.reshape(B, 1)
    factor = torch.empty(B, _factor_elems(cfg), dtype=dt, device=dev)
    if B == 0:
        z = lambda *sh, d=dt: torch.empty(*sh, dtype=d, device=dev)
        return dict(factor=factor, x=z(0, T, nx, d=torch.float32), u=z(0, T, cfg.nu, d=torch.float32), dyn_res=z(0),
                    status=z(0, d=torch.int32), info=torch.zeros(_lib.INFO_INTS, dtype=torch.int32, device=dev),
                    xu=xu, lam=lam, rho=rho.reshape(-1))
    status = torch.full((B,), _lib.STATUS_FINITE | _lib.STATUS_CHOL_OK, dtype=torch.int32, device=dev)
    alphas = (2.0 ** (-torch.arange(N_LS, device=dev).float())).to(dt)  # exact powers of two (al_utils.py:622-627)
    bidx = torch.arange(B, device=dev)
    steps, total, exceeded, al_done = [], 0, False, 0
    cp = None
    with torch.no_grad():
        for _a in range(cfg.al_iter):
            merit = _merit(cfg, xu, x0, Qd, q, lam, rho)
            res_old = torch.norm(_residuals(cfg, xu, x0)[1]).item()
            k = 0
            while k < cfg.max_newton:
                k += 1
                g, G, dD = _gradient_blocks(cfg, xu, x0, Qd, q, lam, rho)
                d = _blocktri(cfg, B, True, G, dD, rho.reshape(-1).contiguous(), factor, g, status)
                cand = xu[None] + alphas[:, None, None, None] * d[None]
                vals = _merit(cfg, cand, x0, Qd, q, lam, rho)
                best = torch.min(vals, dim=0)  # first argmin, NaN propagates like the reference (al_utils.py:633)
                chosen = cand[best.indices, bidx]
                acc = (best.values < merit).to(dt)[:, None, None]
                xu = acc * chosen + (1 - acc) * xu  # al_utils.py:640-641, including its NaN behaviour
                fin = torch.isfinite(xu.reshape(B, -1)).all(dim=1)
                status = torch.where(fin, status, status & ~_lib.STATUS_FINITE)
                res_new = torch.norm(_residuals(cfg, xu, x0)[1]).item()
                merit = best.values  # even when the step was rejected (al_utils.py:569)
                if res_new < 1e-3 or abs(res_old - res_new) / res_new < 1e-3:  # al_utils.py:560-564
                    break
                res_old = res_new
            steps.append(k)
            total += k
            al_done += 1
            c, cp = _residuals(cfg, xu, x0)
            neq = c.shape[1] if cfg.state_estimator else T * nx
            lam = lam + rho * c
            lam = torch.cat([lam[:, :neq], lam[:, neq:].clamp(min=0)], dim=1)
            rho = rho * cfg.rho_mult
            if cfg.rho_f32:
                rho = rho.float().to(dt)
            if cfg.stream_mode and rho.max().item() > cfg.rho_max:  # AL_mpc.py:412
                exceeded = True
                break
    info = torch.zeros(_lib.INFO_INTS, dtype=torch.int32)
    info[_lib.INFO_AL_DONE] = al_done
    info[_lib.INFO_RHO_EXCEEDED] = int(exceeded or (rho.max().item() > cfg.rho_max))
    info[_lib.INFO_TOTAL_NEWTON] = total
    for i, s in enumerate(steps):
        info[_lib.INFO_NEWTON0 + i] = s
    return dict(factor=factor, x=xu[..., :nx].float().contiguous(), u=xu[..., nx:].float().contiguous(),
                dyn_res=cp.norm(dim=-1), status=status, info=info.to(dev), xu=xu.contiguous(), lam=lam.contiguous(),
                rho=rho.reshape(-1).contiguous())


class GenericALSolve(torch.autograd.Function):
    """Same contract as ``solver.ALSolve`` for a ``GenericConfig``: (Qd, q, x0, xu_init, lam, rho, cfg) ->
    (x f32, u f32, xu, lam_new, rho_new, dyn_res, status, info, debug); gradients to Qd and q only."""

    @staticmethod
    def forward(ctx, Qd, q, x0, xu_init, lam, rho, cfg):
        out = forward(cfg, x0.detach().contiguous(), Qd.detach().contiguous(), q.detach().contiguous(),
                      xu_init.detach().clone().contiguous(), lam.detach().clone().contiguous(), rho.detach().clone())
        ctx.cfg = cfg
        ctx.save_for_backward(out["factor"], out["xu"])
        dbg = torch.empty(0, device=x0.device)
        ctx.mark_non_differentiable(out["xu"], out["lam"], out["rho"], out["dyn_res"], out["status"], out["info"], dbg)
        return out["x"], out["u"], out["xu"], out["lam"], out["rho"], out["dyn_res"], out["status"], out["info"], dbg

    @staticmethod
    def backward(ctx, gx, gu, *unused):
        factor, xu = ctx.saved_tensors
        cfg = ctx.cfg
        B = xu.shape[0]
        if gx is None:
            gx = torch.zeros(B, cfg.T, cfg.nx, dtype=xu.dtype, device=xu.device)
        if gu is None:
            gu = torch.zeros(B, cfg.T, cfg.nu, dtype=xu.dtype, device=xu.device)
        grad_xu = torch.cat([gx.to(xu.dtype), gu.to(xu.dtype)], dim=2).contiguous()
        gbar = _blocktri(cfg, B, False, None, None, None, factor, grad_xu, None)  # -H^{-1} grad (al_utils.py:586-590)
        gQ = gbar * xu if ctx.needs_input_grad[0] else None
        return gQ, (gbar if ctx.needs_input_grad[1] else None), None, None, None, None, None
