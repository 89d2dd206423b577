"""Drop-in pieces of ``qpth.al_utils`` (reference qpth/al_utils.py): ``QuadCost`` / ``LinDx``
(:8-13) and the ``NewtonAL`` autograd.Function (:449-615) with the reference's 15-argument order.

``NewtonAL.apply(meritfn, dyn_fn, cost_fn, merit_grad_hessfn, xi, x0, lam, rho, Q, q, f, threshold,
eps, ls, verbose) -> (x_est, status)`` runs the <=4 damped Gauss-Newton steps of ONE AL iteration on
the GPU (no dual update).  The four callables are not evaluated: the fused kernel needs to know
which native dynamics functor and which control bounds they close over, so they must either carry
an ``almpc_context`` attribute (``NewtonContext``) or be the reference's own lambdas from
``MPC.al_solve`` (AL_mpc.py:300-303), whose closure (``self``, ``dx_``) is inspected.
"""
from collections import namedtuple
from dataclasses import dataclass

import torch

from . import _lib

QuadCost = namedtuple("QuadCost", "C c f")
LinDx = namedtuple("LinDx", "F f")
QuadCost.__new__.__defaults__ = (None,) * len(QuadCost._fields)
LinDx.__new__.__defaults__ = (None,) * len(LinDx._fields)


def compute_cost(xu, Q, q, f=None, diag_cost=True):
    """Quadratic cost of a trajectory batch, same conventions as qpth/al_utils.py:425-435 (diagonal ``Q[B,T,n]``,
    linear term ``q``, optional per-stage constant ``f[B,T]``).  A plain torch expression for callers that want
    the cost of an iterate (``MPC.get_cost``); the solver's own merit evaluation lives in the CUDA kernels."""
    if not diag_cost:
        raise ValueError("only the diagonal cost is supported (diag_cost=True is the only mode the reference's callers use)")
    c = (0.5 * (xu * Q * xu).sum(-1) + (q * xu).sum(-1)).sum(dim=-1)
    return c if f is None else c + f.sum(dim=-1)


def compute_cost_gradient(xu, Q, q, diag_cost=True):
    """qpth/al_utils.py:438-446 for the diagonal cost."""
    if not diag_cost:
        raise ValueError("only the diagonal cost is supported")
    return Q * xu + q


@dataclass
class NewtonContext:
    dynamics: object
    u_lower: object
    u_upper: object
    max_newton: int = 4


def _context_from(*fns):
    from .dynamics import resolve

    for fn in fns:
        ctx = getattr(fn, "almpc_context", None)
        if ctx is not None:
            return ctx
    # reference lambdas: closure holds `self` (the MPC module: u_lower/u_upper) and `dx_` (dynamics)
    dyn = bounds = None
    for fn in fns:
        for cell in getattr(fn, "__closure__", None) or ():
            try:
                obj = cell.cell_contents
            except ValueError:
                continue
            if bounds is None and hasattr(obj, "u_lower") and hasattr(obj, "u_upper") and hasattr(obj, "n_ctrl"):
                bounds = (obj.u_lower, obj.u_upper)
            if dyn is None:
                try:
                    dyn = resolve(obj)
                except Exception:
                    pass
    if dyn is None or bounds is None:
        raise _lib.AlmpcError(
            "NewtonAL: cannot tell which native dynamics / control bounds the callables stand for; attach "
            "`almpc_context = NewtonContext(dynamics, u_lower, u_upper)` to the merit function")
    return NewtonContext(dyn, bounds[0], bounds[1])


class NewtonAL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, meritfn, dyn_fn, cost_fn, merit_grad_hessfn, xi, x0, lam, rho, Q, q, f, threshold, eps, ls,
                verbose):
        from .dynamics import resolve
        from .solver import SolveConfig, raw_forward

        if not ls:
            raise _lib.AlmpcError("NewtonAL without line search is not part of the fused path")
        nctx = _context_from(meritfn, dyn_fn, merit_grad_hessfn)
        dyn = resolve(nctx.dynamics)
        dt = xi.dtype
        B = xi.shape[0]
        rho_t = rho if torch.is_tensor(rho) else torch.full((B,), float(rho))
        cfg = SolveConfig(
            dyn=dyn, T=xi.shape[1],
            u_lower=torch.as_tensor(nctx.u_lower).double().reshape(-1).cpu().tolist(),
            u_upper=torch.as_tensor(nctx.u_upper).double().reshape(-1).cpu().tolist(),
            al_iter=1, max_newton=nctx.max_newton, skip_final_dual=True)
        xu = xi.detach().clone().contiguous()
        lam_w = lam.detach().to(dt).clone().contiguous()
        rho_w = rho_t.detach().to(xi.device).to(dt).reshape(-1).clone().contiguous()
        out = raw_forward(cfg, x0.detach().to(dt).contiguous(), Q.detach().to(dt).contiguous(),
                          q.detach().to(dt).contiguous(), xu, lam_w, rho_w)
        ctx.cfg = cfg
        ctx.save_for_backward(out["factor"], xu)
        status = (out["status"] & _lib.STATUS_FINITE).bool()
        ctx.mark_non_differentiable(status)
        return xu, status

    @staticmethod
    def backward(ctx, x_grad, status_grad):
        from .solver import raw_backward

        factor, xu = ctx.saved_tensors
        gQ, gq = raw_backward(ctx.cfg, factor, xu, x_grad, need_Q=ctx.needs_input_grad[8])
        return (None, None, None, None, None, None, None, None, gQ, gq if ctx.needs_input_grad[9] else None, None,
                None, None, None, None)
