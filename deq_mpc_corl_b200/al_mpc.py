"""Drop-in for ``qpth.AL_mpc.MPC`` (reference qpth/AL_mpc.py:53-592) on top of the fused CUDA solver.

Same constructor keywords, same ``forward(x0, cost, dx, dx_jac, compute_Qq=None, u_init=None,
x_init=None) -> (x, u, status)`` and the same mutable state the caller touches
(``al_iter``, ``x_init`` / ``u_init``, ``lamda_prev`` / ``rho_prev``, ``warm_starting``,
``reinitialize``, ``warm_start_initialize``) — see deqmpc/policies.py:1200-1216,1236-1310 for the
call sites.  Differences, all deliberate:

* a dynamics object that resolves to a native CUDA functor (``dynamics.resolve``) runs on the fused kernels;
  any other pair of callables ``dx`` / ``dx_jac`` (qpth/al_utils.py:237-248 convention) runs in the
  generic-callable mode (``generic.py``: the callables are evaluated by torch on the GPU, the linear algebra by
  the block-tridiagonal CUDA kernel); CPU tensors raise — there is no CPU fallback;
* the al_iter x Newton loop runs as one stream-ordered sequence of kernel launches with the
  batch-global decisions taken on the device, so ``forward`` only synchronises with the host in
  streaming mode, where the reference's own ``status`` is a Python bool (AL_mpc.py:420-423);
* ``cost.C`` may be the reference's diag-embedded [B,T,n,n] tensor or already its diagonal [B,T,n].
"""
import torch
from torch.nn import Module

from . import _lib
from .al_utils import LinDx, QuadCost  # noqa: F401  (re-exported like the reference module)
from .dynamics import resolve
from .generic import GenericALSolve, GenericConfig
from .solver import ALSolve, SolveConfig


def _detach_maybe(x):
    if x is None:
        return None
    return x.detach() if torch.is_tensor(x) and x.requires_grad else x


class MPC(Module):
    def __init__(self, n_state, n_ctrl, T, u_lower=None, u_upper=None, u_init=None, x_init=None, al_iter=2,
                 verbose=0, eps=1e-7, back_eps=1e-7, n_batch=None, linesearch_decay=0.2, max_linesearch_iter=10,
                 exit_unconverged=True, detach_unconverged=True, backprop=True, slew_rate_penalty=None,
                 solver_type="dense", add_goal_constraint=False, x_goal=None, diag_cost=True, ineqG=None, ineqh=None,
                 state_estimator=False, dtype=torch.float64):
        super().__init__()
        assert (u_lower is None) == (u_upper is None)
        assert max_linesearch_iter > 0
        if u_lower is None:
            raise _lib.AlmpcError("MPC needs control bounds (the reference dereferences them unconditionally, AL_mpc.py:152)")
        if add_goal_constraint or ineqG is not None or not diag_cost:
            raise _lib.AlmpcError("goal constraint / ineqG / dense cost are outside the fused AL-MPC path")
        self.dtype = dtype
        self.n_state, self.n_ctrl, self.T = n_state, n_ctrl, T
        self.u_lower = _detach_maybe(torch.as_tensor(u_lower).to(dtype))
        self.u_upper = _detach_maybe(torch.as_tensor(u_upper).to(dtype))
        if self.u_lower.dim() == 0:
            self.u_lower = self.u_lower.repeat(n_ctrl)
            self.u_upper = self.u_upper.repeat(n_ctrl)
        self.x_lower = self.x_upper = None
        self.u_init = _detach_maybe(u_init)
        self.x_init = _detach_maybe(x_init)
        self.verbose, self.eps, self.back_eps, self.n_batch = verbose, eps, back_eps, n_batch
        self.linesearch_decay, self.max_linesearch_iter = linesearch_decay, max_linesearch_iter
        self.exit_unconverged, self.detach_unconverged, self.backprop = exit_unconverged, detach_unconverged, backprop
        self.slew_rate_penalty, self.solver_type = slew_rate_penalty, solver_type
        self.state_estimator, self.diag_cost = state_estimator, diag_cost
        self.al_iter = al_iter
        # state_estimator=True (AL_mpc.py:186-199; merit / Hessian of qpth/al_utils_se.py): the T-1 dynamics blocks
        # only, no control bounds — the controls are data and only the states are estimated
        self.neq = n_state * (T - 1) if state_estimator else n_state * T
        self.nineq = 0 if state_estimator else 2 * n_ctrl * T
        self.rho_prev = 1.0
        self.rho_max = 1e8
        self.max_newton_steps = 4
        self.lamda_prev = None if n_batch is None else torch.zeros(n_batch, self.neq + self.nineq, dtype=dtype)
        self.dyn_res_prev = 1000000
        self.warm_starting = False
        self.just_initialized = True
        self.cost_hist_stream = [[], []]
        self.mask = None
        # opt-in: replay the forward launch sequence as one CUDA graph (saves ~0.25 ms per call; pays for batches of
        # a few hundred problems, e.g. the reference's bsz = 200 training runs).  Fused native path only.
        self.cuda_graph = False
        self._graph_cache = None
        # opt-in: a sharding.GlobalStop -> the reference's batch-GLOBAL stop rule / rho_max exit across the ranks of a
        # sharded batch (default None: every shard applies the rule to its own problems)
        self.global_stop = None
        self.last_info = None      # device int32[INFO_INTS] of the latest solve (no sync to keep it)
        self.last_status = None    # device int32[B] per-problem status bits
        self._bounds_host = (self.u_lower.double().cpu().tolist(), self.u_upper.double().cpu().tolist())

    # ------------------------------------------------------------------------------------------
    def forward(self, x0, cost, dx, dx_jac, compute_Qq=None, u_init=None, x_init=None):
        n_batch = self.n_batch if self.n_batch is not None else cost.C.size(0)
        assert x0.ndimension() == 2 and x0.size(0) == n_batch
        _lib.require_cuda(x0, "MPC.forward(x0)")
        dyn = self._resolve(dx, dx_jac)
        if dyn is not None and (dyn.nx != self.n_state or dyn.nu != self.n_ctrl):
            raise _lib.AlmpcError("dynamics dimensions do not match n_state / n_ctrl")
        dev = x0.device

        # ---- primal warm start (AL_mpc.py:222-246)
        if u_init is not None:
            u = u_init
        elif self.u_init is None:
            u = torch.zeros(n_batch, self.T, self.n_ctrl, dtype=x0.dtype, device=dev)
        else:
            u = self.u_init
        if u.ndimension() == 2:
            u = u.unsqueeze(0).expand(n_batch, self.T, -1).clone()
        u = u.to(dev).type_as(x0.data)
        if x_init is not None:
            x = x_init
        elif self.x_init is None:
            x = self.rollout(x0, u, dyn if dyn is not None else dx)
        else:
            x = self.x_init
        if x.ndimension() == 2:
            x = x.unsqueeze(0).expand(n_batch, self.T, -1).clone()
        x = x.to(dev).type_as(x0.data)

        # ---- cost (AL_mpc.py:249-250): diagonal of C, linear term c
        C = cost.C
        assert C.ndimension() in (3, 4)
        Qd = C.diagonal(dim1=-2, dim2=-1) if C.ndimension() == 4 else C
        Qd = Qd.to(self.dtype)
        q = cost.c.to(self.dtype)

        # ---- dual warm start
        m = self.neq + self.nineq
        lam = self.lamda_prev
        if lam is None:
            lam = torch.zeros(n_batch, m, dtype=self.dtype, device=dev)
        elif lam.shape != (n_batch, m):
            # the reference would fail inside its first residual expression; silently restarting the duals from zero
            # (wrong batch size, missing reinitialize(), a subclass that changed nineq) would change results unnoticed
            raise _lib.AlmpcError(f"lamda_prev has shape {tuple(lam.shape)}, expected {(n_batch, m)}: call reinitialize() "
                                  "for a new batch")
        rho = self.rho_prev
        if not torch.is_tensor(rho):
            rho = torch.full((n_batch, 1), float(rho), dtype=self.dtype, device=dev)
        rho_dtype = rho.dtype
        extra = self._extra_constraints(dev)
        if dyn is not None:
            solve = ALSolve
            cfg = self._solve_config(dyn, rho_dtype, extra)
        else:
            if extra:
                raise _lib.AlmpcError("obstacle constraints need a native dynamics functor")
            solve = GenericALSolve
            cfg = GenericConfig(dx=dx, dx_jac=dx_jac, nx=self.n_state, nu=self.n_ctrl,
                                u_lower=self.u_lower.to(dev).to(self.dtype), u_upper=self.u_upper.to(dev).to(self.dtype),
                                T=self.T, al_iter=int(self.al_iter), max_newton=self.max_newton_steps,
                                stream_mode=bool(self.warm_starting),
                                rho_f32=(rho_dtype == torch.float32 and self.dtype == torch.float64), rho_max=self.rho_max,
                                state_estimator=bool(self.state_estimator))
        # one fresh [B,T,n] buffer, filled by two converting copies (instead of two casts + a concatenation)
        xu0 = torch.empty(n_batch, self.T, self.n_state + self.n_ctrl, dtype=self.dtype, device=dev)
        xu0[:, :, :self.n_state].copy_(x)
        xu0[:, :, self.n_state:].copy_(u)
        cfg.__dict__["_owns_xu"] = True
        xo, uo, xu, lam_new, rho_new, dyn_res, status, info, _ = solve.apply(
            Qd, q, x0.to(self.dtype), xu0, lam.to(dev).to(self.dtype), rho.to(dev).to(self.dtype).reshape(-1), cfg)

        # ---- state kept for the next call (AL_mpc.py:333-336, 256-257).  Written through __dict__: these are plain
        # attributes, and nn.Module.__setattr__ costs ~5 us each (0.1 ms per call, a fifth of a small-batch solve)
        self.__dict__.update(
            lamda_prev=lam_new, rho_prev=rho_new.to(rho_dtype).reshape(n_batch, 1), dyn_res_prev=dyn_res,
            just_initialized=False, last_info=info, last_status=status, last_xu=xu,
            x_init=xo.detach().clone(), u_init=uo.detach().clone())
        flag = False
        if self.warm_starting:
            flag = bool(info[_lib.INFO_RHO_EXCEEDED].item())  # the reference syncs here too (AL_mpc.py:412,420)
        return xo, uo, flag

    def _extra_constraints(self, dev):
        """Hook for subclasses (Obstacle_MPC): extra SolveConfig fields."""
        return {}

    def _solve_config(self, dyn, rho_dtype, extra):
        """The SolveConfig of this call.  Cached per (dynamics object, al_iter, mode): building it — and the C
        descriptor it carries — on every forward cost a fifth of a 200-problem solve."""
        cache = self.__dict__.setdefault("_cfg_cache", {})
        key = (id(dyn), int(self.al_iter), bool(self.warm_starting), rho_dtype, self.max_newton_steps, self.rho_max)
        cfg = cache.get(key)
        if cfg is None or cfg.dyn is not dyn:
            if len(cache) > 32:
                cache.clear()
            cfg = cache[key] = SolveConfig(
                dyn=dyn, u_lower=self._bounds_host[0], u_upper=self._bounds_host[1], T=self.T, al_iter=int(self.al_iter),
                max_newton=self.max_newton_steps, stream_mode=bool(self.warm_starting),
                rho_f32=(rho_dtype == torch.float32 and self.dtype == torch.float64), rho_max=self.rho_max,
                state_estimator=bool(self.state_estimator))
        cfg.obstacles = extra.get("obstacles")
        cfg.obs_radius = extra.get("obs_radius", 0.0)
        cfg.global_stop = self.global_stop
        cfg.graph_cache = None
        if self.cuda_graph and self.global_stop is None:
            if self._graph_cache is None:
                from .solver import GraphCache

                self.__dict__["_graph_cache"] = GraphCache()
            cfg.graph_cache = self._graph_cache
        return cfg

    @staticmethod
    def _resolve(dx, dx_jac):
        """Native functor descriptor, or None -> generic-callable mode (both callables are then required)."""
        try:
            return resolve(dx)
        except _lib.AlmpcError:
            if callable(dx) and callable(dx_jac):
                return None
            raise

    # ------------------------------------------------------------------------------------------
    def rollout(self, x, actions, dynamics):
        """AL_mpc.py:521-534 with the native functor (or the user's callable in generic mode); a ``LinDx(F, f)``
        rolls out the time-varying linear model ``x+ = F_t [x; u] + f_t`` like the reference (:527-529)."""
        if isinstance(dynamics, LinDx):
            xs = [x]
            for t in range(self.T - 1):
                z = torch.cat([xs[t], actions[:, t]], dim=-1)
                xs.append(torch.bmm(dynamics.F[:, t], z.unsqueeze(-1)).squeeze(-1) + dynamics.f[:, t])
            return torch.stack(xs, 1)
        try:
            dyn = resolve(dynamics)
        except _lib.AlmpcError:
            if not callable(dynamics):
                raise
            dyn = dynamics
        xs = [x]
        for t in range(self.T - 1):
            xs.append(dyn(xs[t].contiguous(), actions[:, t].contiguous()))
        return torch.stack(xs, 1)

    def get_xu(self):
        return torch.cat((self.x_init, self.u_init), dim=2)

    def get_cost(self, cost):
        """AL_mpc.py:558-563: cost of the current warm start under ``cost`` (C may be diag-embedded or its diagonal)."""
        from .al_utils import compute_cost

        C = cost.C.diagonal(dim1=-2, dim2=-1) if cost.C.ndimension() == 4 else cost.C
        xu = self.get_xu()
        return compute_cost(xu, C.to(xu.dtype), cost.c.to(xu.dtype), None if cost.f is None else cost.f.to(xu.dtype))

    def reinitialize(self, x, mask=None):
        """AL_mpc.py:569-579: cold start — lam = 0, rho = 1 (in the dtype of ``x``), drop the warm start."""
        n_batch = x.size(0)
        self.__dict__.update(
            u_init=None, x_init=None, n_batch=n_batch,
            rho_prev=torch.ones((n_batch, 1), device=x.device, dtype=x.dtype),
            lamda_prev=torch.zeros(n_batch, self.neq + self.nineq, device=x.device, dtype=x.dtype),
            cost_hist_stream=[[], []], dyn_res_prev=1000000, just_initialized=True, warm_starting=False, mask=mask)

    def warm_start_initialize(self, x, u, args):
        """AL_mpc.py:581-592: keep the primal warm start, zero the duals, clamp rho, switch to streaming."""
        self.u_init = u
        self.x_init = x
        self.lamda_prev = self.lamda_prev * 0
        rho = self.rho_prev
        if not torch.is_tensor(rho):
            rho = torch.full((self.n_batch, 1), float(rho), dtype=self.dtype, device=x.device)
        self.rho_prev = torch.clamp(rho, max=args.rho_init_max)
        self.just_initialized = True
        self.warm_starting = True

    def newton_steps(self):
        """Host copy of the realised Newton-step counts of the latest solve (synchronises)."""
        info = self.last_info.cpu().tolist()
        return info[_lib.INFO_NEWTON0:_lib.INFO_NEWTON0 + info[_lib.INFO_AL_DONE]]


class Obstacle_MPC(MPC):
    """Drop-in for ``qpth.AL_mpc_custom.Obstacle_MPC`` (qpth/AL_mpc_custom.py:22-135): the AL-MPC solve with
    ``n_obstacle_constraints = 4`` sphere-avoidance rows per stage, ``r^2 - |p_t - c_k|^2 <= 0`` on the position
    ``x[0:3]``, against the 4 obstacles nearest to the reference trajectory (chosen in ``reinitialize`` and
    shifted in ``warm_start_initialize`` exactly like the reference).  ``env`` supplies ``obstacle_positions``
    ``[n_obstacles,3]`` and ``obstacle_radius``."""

    def __init__(self, n_state, n_ctrl, T, *args, env=None, **kwargs):
        super().__init__(n_state, n_ctrl, T, *args, **kwargs)
        if self.state_estimator:
            raise _lib.AlmpcError("Obstacle_MPC has no state-estimator variant")
        self.n_obstacle_constraints = 4
        self.nineq += self.n_obstacle_constraints * T
        # the reference turns the radius into a default-dtype (fp32) tensor first (AL_mpc_custom.py:60)
        self.obstacle_radius = float(torch.tensor(0.2 if env is None else float(env.obstacle_radius)))
        self.obstacle_positions = None if env is None else torch.as_tensor(env.obstacle_positions)
        self.n_obstacles = 0 if self.obstacle_positions is None else self.obstacle_positions.shape[0]
        self.obstacles = None
        if self.n_batch is not None:
            self.lamda_prev = torch.zeros(self.n_batch, self.neq + self.nineq, dtype=self.dtype)

    def _nearest(self, xyz):
        """xyz [B,t,3] -> positions of the n_obstacle_constraints nearest obstacles [B,t,4,3]
        (AL_mpc_custom.py:110-112: argsort of the distance to every obstacle)."""
        obs = self.obstacle_positions.to(xyz.device)
        d = (obs[None, None] - xyz[:, :, None, :]).norm(dim=-1)
        ids = torch.argsort(d, dim=-1)[..., : self.n_obstacle_constraints]
        return obs[ids]

    def reinitialize(self, x, mask=None):
        super().reinitialize(x, mask)
        self.obstacles = (self._nearest(x[:, :, :3]), self.obstacle_radius)

    def warm_start_initialize(self, x, u, args):
        super().warm_start_initialize(x, u, args)
        shifted = self.obstacles[0][:, 1:]
        last = self._nearest(x[:, -1:, :3])
        self.obstacles = (torch.cat([shifted, last], dim=1), self.obstacle_radius)

    def _extra_constraints(self, dev):
        if self.obstacles is None:
            raise _lib.AlmpcError("Obstacle_MPC.forward before reinitialize(): no obstacles selected")
        return dict(obstacles=self.obstacles[0].to(dev).to(self.dtype).contiguous(), obs_radius=self.obstacle_radius)
