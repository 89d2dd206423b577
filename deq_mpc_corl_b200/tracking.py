"""`Tracking_MPC` — the reference's only production caller of the AL-MPC solver — with the cost build and
its chain rule fused into two small kernels around the solve.

Mirror of deqmpc/policies.py:1117-1315 (`Tracking_MPC`): same constructor (`args`, `env`), same
`forward(x0, xu_ref, x_ref, u_ref, q_scaling=None, model_call=None, al_iters=2)` returning
`(nominal_states, nominal_actions, status)`, same `reinitialize(x, mask)` /
`warm_start_initialize(x_ref, u_ref)`.  What changes underneath:

* the reference materialises `Q` as a diag-embedded `[B,T,n,n]` fp64 tensor (policies.py:1172) and builds
  `p = -(Q * xu_ref).sum(-1)`, `f` with a handful of elementwise kernels (policies.py:1285-1297); here
  `almpc_tracking_cost_*` writes `Qd[B,T,n]` and `q[B,T,n]` in one pass from the fp32 network outputs, and the
  backward maps the solver's `(Q_grad, q_grad)` straight to `x_ref.grad`, `u_ref.grad`, `q_scaling.grad`;
* only the AL solver (`solver_type == "al"`) exists on this path; `"obstacles" in args.env` selects
  `Obstacle_MPC` (policies.py:1180-1198), `state_estimator=True` (policies.py:321) the dynamics-rows-only
  variant of qpth/al_utils_se.py.

`env` needs `nx, nu, dt, dynamics` (anything `dynamics.resolve` accepts) and `action_space.low/high`;
`args` needs `T, bsz, Q, R, dtype ("double"|"float"), device` and, for streaming, `rho_init_max`.
"""
import ctypes

import torch

from . import _lib
from .al_mpc import MPC, Obstacle_MPC
from .al_utils import QuadCost
from .dynamics import resolve


class _TrackingCost(torch.autograd.Function):
    """(x_ref, u_ref, scale|None, Q0) -> (Qd, q) in the solver dtype; backward through the fused kernel."""

    @staticmethod
    def forward(ctx, x_ref, u_ref, scale, Q0):
        lib = _lib.load()
        _lib.require_cuda(x_ref, "Tracking_MPC(x_ref)")
        B, T, nx = x_ref.shape
        nu = u_ref.shape[-1]
        xr = x_ref.detach().float().contiguous()
        ur = u_ref.detach().float().contiguous()
        sc = None if scale is None else scale.detach().float().contiguous()
        Qd = torch.empty(B, T, nx + nu, dtype=Q0.dtype, device=x_ref.device)
        q = torch.empty_like(Qd)
        with torch.cuda.device(x_ref.device):
            rc = getattr(lib, "almpc_tracking_cost_" + _lib.suffix(Q0.dtype))(
                B, T, nx, nu, _lib.ptr(Q0), _lib.ptr(xr), _lib.ptr(ur), _lib.ptr(sc), _lib.ptr(Qd), _lib.ptr(q),
                _lib.stream_ptr())
        _lib.check(rc, "almpc_tracking_cost")
        ctx.save_for_backward(xr, ur, sc if sc is not None else torch.empty(0, device=x_ref.device), Q0)
        ctx.has_scale = sc is not None
        ctx.in_dtypes = (x_ref.dtype, u_ref.dtype, None if scale is None else scale.dtype)
        return Qd, q

    @staticmethod
    def backward(ctx, gQ, gq):
        lib = _lib.load()
        xr, ur, sc, Q0 = ctx.saved_tensors
        B, T, nx = xr.shape
        nu = ur.shape[-1]
        if gq is None:
            gq = torch.zeros(B, T, nx + nu, dtype=Q0.dtype, device=xr.device)
        gq = gq.to(Q0.dtype).contiguous()
        gQc = None if gQ is None else gQ.to(Q0.dtype).contiguous()
        need = ctx.needs_input_grad
        gx = torch.empty_like(xr) if need[0] else None
        gu = torch.empty_like(ur) if need[1] else None
        gs = torch.empty(B, T, dtype=torch.float32, device=xr.device) if (ctx.has_scale and need[2]) else None
        with torch.cuda.device(xr.device):
            rc = getattr(lib, "almpc_tracking_cost_bwd_" + _lib.suffix(Q0.dtype))(
                B, T, nx, nu, _lib.ptr(Q0), _lib.ptr(xr), _lib.ptr(ur), _lib.ptr(sc if ctx.has_scale else None),
                _lib.ptr(gQc), _lib.ptr(gq), _lib.ptr(gx), _lib.ptr(gu), _lib.ptr(gs), _lib.stream_ptr())
        _lib.check(rc, "almpc_tracking_cost_bwd")
        dx, du, ds = ctx.in_dtypes
        return (None if gx is None else gx.to(dx), None if gu is None else gu.to(du),
                None if gs is None else gs.to(ds), None)


class Tracking_MPC(torch.nn.Module):
    def __init__(self, args, env, state_estimator=False):
        super().__init__()
        if getattr(args, "solver_type", "al") != "al":
            raise _lib.AlmpcError("only the AL solver (solver_type == 'al') is on the fused path")
        self.args = args
        self.nu, self.nx = env.nu, env.nx
        self.nq = getattr(env, "nq", None)
        self.dt = env.dt
        self.T = args.T
        self.dyn = env.dynamics
        self.dyn_jac = getattr(env, "dynamics_derivatives", env.dynamics)
        self.native = resolve(self.dyn)
        self.device = args.device
        self.u_upper = torch.as_tensor(env.action_space.high).to(self.device)
        self.u_lower = torch.as_tensor(env.action_space.low).to(self.device)
        self.bsz = args.bsz
        self.dtype = torch.float64 if args.dtype == "double" else torch.float32
        if getattr(args, "Q", None) is None:
            Q = torch.ones(self.nx)
            R = torch.ones(self.nu)
        else:
            Q, R = args.Q, args.R
        # diagonal of the reference's [B,T,n,n] Q (policies.py:1148,1172), kept as one [n] vector
        self.Q = torch.cat([torch.as_tensor(Q), torch.as_tensor(R)], dim=0).to(self.dtype).to(self.device).contiguous()
        self.u_init = torch.randn(self.bsz, self.T, self.nu, dtype=self.dtype, device=self.device)
        self.x_init = None
        kw = dict(u_lower=self.u_lower, u_upper=self.u_upper, exit_unconverged=False, eps=1e-2, n_batch=self.bsz,
                  backprop=False, verbose=0, u_init=self.u_init, solver_type="dense", dtype=self.dtype,
                  state_estimator=state_estimator)
        if "obstacles" in str(getattr(args, "env", "")):  # policies.py:1180-1198
            self.ctrl = Obstacle_MPC(self.nx, self.nu, self.T, env=env, **kw)
        else:
            self.ctrl = MPC(self.nx, self.nu, self.T, **kw)

    def forward(self, x0, xu_ref, x_ref, u_ref, q_scaling=None, model_call=None, al_iters=2):
        if self.x_init is None:  # policies.py:1242-1244
            self.x_init = self.ctrl.x_init = x_ref.detach().clone()
            self.u_init = self.ctrl.u_init = u_ref.detach().clone()
        Qd, q = _TrackingCost.apply(x_ref, u_ref, q_scaling, self.Q)
        self.p = q
        self.ctrl.al_iter = al_iters
        x, u, status = self.ctrl(x0, QuadCost(Qd, q, None), self.native, self.native, None)
        self.u_init = u.clone().detach()
        return x, u, status

    def compute_pf(self, x_ref, Q=None):
        """Reference-shaped helper (policies.py:1285-1297): p and f for xu_ref = [x_ref|u_ref] [B,T,n]."""
        Qv = self.Q if Q is None else Q
        p = -(Qv * x_ref)
        f = 0.5 * (x_ref * Qv * x_ref).sum(-1)
        self.p, self.f = p, f
        return p, f

    def reinitialize(self, x, mask=None):
        self.u_init = torch.randn(self.bsz, self.T, self.nu, dtype=x.dtype, device=x.device)
        self.x_init = None
        self.ctrl.reinitialize(x, mask)

    def warm_start_initialize(self, x_ref, u_ref):
        self.u_init = self.ctrl.u_init
        self.x_init = self.ctrl.x_init
        self.u_init[:, -1:] = u_ref[:, -1:]
        self.x_init[:, -1:] = x_ref[:, -1:]
        self.ctrl.warm_start_initialize(self.x_init, self.u_init, self.args)
