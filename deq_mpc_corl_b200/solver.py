"""Fused AL-MPC solve as one torch.autograd.Function over the C ABI.

``ALSolve.apply`` replaces the whole loop of ``MPC.al_solve`` / ``al_solve_stream``
(qpth/AL_mpc.py:260-423): every ``NewtonAL.apply`` (qpth/al_utils.py:451-576), the dual / penalty
update (AL_mpc.py:315-325) and the fp32 output cast (AL_mpc.py:337-338) happen inside
``almpc_forward_*``; the backward is ``NewtonAL.backward`` (al_utils.py:578-615) via
``almpc_backward_*``.  No host synchronisation happens here.
"""
import ctypes
from dataclasses import dataclass

import torch

from . import _lib
from .dynamics import NativeDynamics


@dataclass
class SolveConfig:
    dyn: NativeDynamics
    T: int
    u_lower: list
    u_upper: list
    al_iter: int = 2
    max_newton: int = 4
    stream_mode: bool = False
    rho_f32: bool = False
    rho_mult: float = 10.0
    rho_max: float = 1e8
    skip_final_dual: bool = False
    debug: bool = False
    obstacles: object = None      # device tensor [B,T,n_obs,3] in the solver dtype (Obstacle_MPC)
    obs_radius: float = 0.0
    state_estimator: bool = False  # MPC(state_estimator=True): dynamics rows only, controls are data (al_utils_se.py)
    graph_cache: object = None     # a GraphCache: replay the forward launch sequence as one CUDA graph
    global_stop: object = None     # a sharding.GlobalStop: batch-global stop rule across shards (NCCL all-reduce hook)
    profile: bool = False          # per-kernel CUDA-event timing of the next forward (synchronises): raw_forward's "profile_ms"

    def desc(self, B):
        """The C descriptor of this configuration.  Filling the 70-field ctypes struct costs ~25 us, so the filled
        struct is cached per batch size and only the per-call fields are refreshed."""
        cache = self.__dict__.setdefault("_desc_cache", {})
        key = (B, self.al_iter, self.max_newton, self.stream_mode, self.rho_f32, self.skip_final_dual)
        d = cache.get(key)
        if d is None:
            if len(cache) > 16:
                cache.clear()
            d = cache[key] = self._fill_desc(B)
        d.select = _lib.env_select()
        if self.obstacles is not None:
            d.n_obs = int(self.obstacles.shape[2])
            d.obs_radius = float(self.obs_radius)
            d.obstacles = self.obstacles.data_ptr()
        d.profile_ms = None
        if self.global_stop is not None:
            d.stop_stats = self.global_stop.stats.data_ptr()
            d.stop_reduce = ctypes.cast(self.global_stop.hook, ctypes.c_void_p)
        else:
            d.stop_stats = None
            d.stop_reduce = None
        return d

    def _fill_desc(self, B):
        d = _lib.Desc()
        d.B, d.T, d.dyn_id = B, self.T, self.dyn.dyn_id
        d.al_iter, d.max_newton = self.al_iter, self.max_newton
        d.stream_mode, d.rho_f32 = int(self.stream_mode), int(self.rho_f32)
        d.flags = (_lib.FLAG_SKIP_FINAL_DUAL if self.skip_final_dual else 0) | (
            _lib.FLAG_STATE_ESTIMATOR if self.state_estimator else 0)
        d.rho_mult, d.rho_max = self.rho_mult, self.rho_max
        for i, v in enumerate(self.dyn.param_block()):
            d.dyn_params[i] = v
        if len(self.u_lower) != self.dyn.nu or len(self.u_upper) != self.dyn.nu:
            raise _lib.AlmpcError("u_lower / u_upper must have n_ctrl entries")
        for i in range(self.dyn.nu):
            d.u_lower[i] = float(self.u_lower[i])
            d.u_upper[i] = float(self.u_upper[i])
        return d


def raw_forward(cfg: SolveConfig, x0, Qd, q, xu, lam, rho):
    """In-place forward on caller-owned buffers (xu, lam, rho are updated).  Returns a dict of the
    freshly allocated outputs.  All tensors: CUDA, contiguous, solver dtype."""
    lib = _lib.load()
    for t, name in ((x0, "x0"), (Qd, "Qd"), (q, "q"), (xu, "xu"), (lam, "lam"), (rho, "rho")):
        _lib.require_cuda(t, name)
        if not t.is_contiguous():
            raise _lib.AlmpcError(f"{name} must be contiguous")
        if t.dtype != xu.dtype:
            raise _lib.AlmpcError(f"{name} has dtype {t.dtype}, expected {xu.dtype}")
    B, T, n = xu.shape
    nx, nu = cfg.dyn.nx, cfg.dyn.nu
    if T != cfg.T or n != nx + nu:
        raise _lib.AlmpcError(f"xu has shape {tuple(xu.shape)}, expected [B,{cfg.T},{nx + nu}]")
    n_obs = 0
    if cfg.obstacles is not None:
        ob = cfg.obstacles
        _lib.require_cuda(ob, "obstacles")
        if ob.dtype != xu.dtype or not ob.is_contiguous() or ob.dim() != 4 or ob.shape[0] != B or ob.shape[1] != T or ob.shape[3] != 3:
            raise _lib.AlmpcError("obstacles must be a contiguous [B,T,n_obs,3] tensor of the solver dtype")
        n_obs = ob.shape[2]
    m = (T - 1) * nx if cfg.state_estimator else T * nx + T * (2 * nu + n_obs)
    if cfg.state_estimator and n_obs:
        raise _lib.AlmpcError("the state-estimator variant has no obstacle rows")
    if x0.shape != (B, nx) or Qd.shape != (B, T, n) or q.shape != (B, T, n) or lam.shape != (B, m) or rho.numel() != B:
        raise _lib.AlmpcError("inconsistent AL-MPC buffer shapes")
    dev, dt = xu.device, xu.dtype
    d = cfg.desc(B)
    prof = None
    if cfg.profile:
        prof = (ctypes.c_double * 8)()
        d.profile_ms = ctypes.cast(prof, ctypes.POINTER(ctypes.c_double))
    out = dict(
        factor=torch.empty(B, lib.almpc_factor_elems(ctypes.byref(d)), dtype=dt, device=dev),
        x=torch.empty(B, T, nx, dtype=torch.float32, device=dev),
        u=torch.empty(B, T, nu, dtype=torch.float32, device=dev),
        dyn_res=torch.empty(B, dtype=dt, device=dev),
        status=torch.empty(B, dtype=torch.int32, device=dev),
        info=torch.empty(_lib.INFO_INTS, dtype=torch.int32, device=dev),
        debug=(torch.zeros(B, lib.almpc_debug_elems(ctypes.byref(d)), dtype=dt, device=dev) if cfg.debug else None),
    )
    ws = torch.empty(lib.almpc_workspace_bytes(ctypes.byref(d)), dtype=torch.uint8, device=dev)
    if B > 0:
        fn = getattr(lib, "almpc_forward_" + _lib.suffix(dt))
        args = (ctypes.byref(d), _lib.ptr(x0), _lib.ptr(Qd), _lib.ptr(q), _lib.ptr(xu), _lib.ptr(lam), _lib.ptr(rho),
                _lib.ptr(out["factor"]), _lib.ptr(out["x"]), _lib.ptr(out["u"]), _lib.ptr(out["dyn_res"]),
                _lib.ptr(out["status"]), _lib.ptr(out["info"]), _lib.ptr(out["debug"]), _lib.ptr(ws))
        if torch.cuda.current_device() == dev.index:  # the device guard costs a few us: only when it is needed
            rc = fn(*args, _lib.stream_ptr())
        else:
            with torch.cuda.device(dev):
                rc = fn(*args, _lib.stream_ptr())
        if cfg.global_stop is not None and cfg.global_stop.error is not None:
            err, cfg.global_stop.error = cfg.global_stop.error, None
            raise _lib.AlmpcError(f"stop_reduce hook failed: {err!r}")
        _lib.check(rc, "almpc_forward")
    else:
        out["info"].zero_()
    if prof is not None:
        out["profile_ms"] = list(prof)
    return out


class GraphCache:
    """Opt-in CUDA-graph replay of ``raw_forward`` for small batches, where a solve is bound by its ~27 kernel
    launches (``MPC.cuda_graph = True``).  One entry per (shape, dtype, configuration): static input / output
    buffers and the captured launch sequence; a call copies its inputs in, replays, and hands back fresh clones
    (the factor is saved for backward by autograd, so the static buffers must not leak out).  The library
    itself neither allocates nor synchronises, which is what makes the capture legal
    (tests/test_gpu_parity.py::test_forward_backward_are_cuda_graph_capturable)."""

    MAX_ENTRIES = 8  # every entry pins static inputs + a factor buffer of B * factor_elems scalars

    def __init__(self):
        self.entries = {}

    @staticmethod
    def _key(cfg, xu):
        ob = cfg.obstacles
        return (tuple(xu.shape), xu.dtype, xu.device, cfg.dyn.dyn_id, tuple(cfg.dyn.param_block()), cfg.T,
                tuple(cfg.u_lower), tuple(cfg.u_upper), cfg.al_iter, cfg.max_newton, cfg.stream_mode, cfg.rho_f32,
                cfg.rho_mult, cfg.rho_max, cfg.skip_final_dual, cfg.state_estimator,
                None if ob is None else tuple(ob.shape), cfg.obs_radius)

    def forward(self, cfg, x0, Qd, q, xu, lam, rho):
        import copy

        key = self._key(cfg, xu)
        ent = self.entries.get(key)
        names = ("x0", "Qd", "q", "xu", "lam", "rho")
        live = dict(zip(names, (x0, Qd, q, xu, lam, rho)))
        if ent is None:
            static = {k: v.clone() for k, v in live.items()}
            scfg = copy.copy(cfg)
            scfg.debug = False
            if cfg.obstacles is not None:
                scfg.obstacles = cfg.obstacles.clone()
            side = torch.cuda.Stream(device=xu.device)
            side.wait_stream(torch.cuda.current_stream(xu.device))
            with torch.cuda.stream(side):  # warm-up outside capture (lazy loading, function attributes)
                raw_forward(scfg, *(static[k].clone() for k in names))
            torch.cuda.current_stream(xu.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = raw_forward(scfg, *(static[k] for k in names))
            while len(self.entries) >= self.MAX_ENTRIES:  # least recently used first (dicts keep insertion order)
                self.entries.pop(next(iter(self.entries)))
            ent = self.entries[key] = dict(static=static, cfg=scfg, graph=graph, out=out)
        else:
            self.entries[key] = self.entries.pop(key)  # mark as most recently used
        for k in names:
            ent["static"][k].copy_(live[k])
        if cfg.obstacles is not None:
            ent["cfg"].obstacles.copy_(cfg.obstacles)
        ent["graph"].replay()
        # results leave as fresh tensors; the caller-owned in/out buffers are updated like raw_forward does
        xu.copy_(ent["static"]["xu"])
        lam.copy_(ent["static"]["lam"])
        rho.copy_(ent["static"]["rho"])
        return {k: (v.clone() if torch.is_tensor(v) else v) for k, v in ent["out"].items()}


def raw_backward(cfg: SolveConfig, factor, xu, grad_xu, need_Q=True, status=None):
    """status = the forward's per-problem status bits (int32[B]) or None: with it, only problems whose factor is
    flagged broken (STATUS_CHOL_OK cleared) get a zero gradient; a non-finite upstream gradient propagates like in
    the reference instead of being hidden."""
    lib = _lib.load()
    _lib.require_cuda(grad_xu, "grad_xu")
    B = xu.shape[0]
    d = cfg.desc(B)
    grad_xu = grad_xu.to(xu.dtype).contiguous()
    gq = torch.empty_like(xu)
    gQ = torch.empty_like(xu) if need_Q else None
    if B > 0:
        with torch.cuda.device(xu.device):
            rc = getattr(lib, "almpc_backward_" + _lib.suffix(xu.dtype))(
                ctypes.byref(d), _lib.ptr(factor), _lib.ptr(xu), _lib.ptr(grad_xu), _lib.ptr(status), _lib.ptr(gQ),
                _lib.ptr(gq), _lib.stream_ptr())
        _lib.check(rc, "almpc_backward")
    return gQ, gq


class ALSolve(torch.autograd.Function):
    """(Qd, q, x0, xu_init, lam, rho, cfg) -> (x f32, u f32, xu, lam_new, rho_new, dyn_res, status, info, debug).
    Gradients flow to Qd and q only (al_utils.py:599-615)."""

    @staticmethod
    def forward(ctx, Qd, q, x0, xu_init, lam, rho, cfg):
        # the solver updates xu / lam / rho in place: work on copies unless the caller handed over a buffer it built
        # for this call (MPC.forward does, `_owns_xu`)
        if cfg.__dict__.pop("_owns_xu", False) and xu_init.is_contiguous() and not xu_init.requires_grad:
            xu = xu_init
        else:
            xu = xu_init.detach().clone().contiguous()
        lam = lam.detach().clone().contiguous()
        rho = rho.detach().clone().contiguous()
        fwd = raw_forward if cfg.graph_cache is None or cfg.debug else cfg.graph_cache.forward
        out = fwd(cfg, x0.detach().contiguous(), Qd.detach().contiguous(), q.detach().contiguous(), xu, lam, rho)
        ctx.cfg = cfg
        ctx.save_for_backward(out["factor"], xu, out["status"])
        dbg = out["debug"] if out["debug"] is not None else torch.empty(0, device=xu.device)
        ctx.mark_non_differentiable(xu, lam, rho, out["dyn_res"], out["status"], out["info"], dbg)
        return out["x"], out["u"], xu, lam, rho, out["dyn_res"], out["status"], out["info"], dbg

    @staticmethod
    def backward(ctx, gx, gu, *unused):
        factor, xu, status = ctx.saved_tensors
        nx = ctx.cfg.dyn.nx
        if gx is None:
            gx = torch.zeros(xu.shape[0], xu.shape[1], nx, dtype=xu.dtype, device=xu.device)
        if gu is None:
            gu = torch.zeros(xu.shape[0], xu.shape[1], xu.shape[2] - nx, dtype=xu.dtype, device=xu.device)
        grad_xu = torch.cat([gx.to(xu.dtype), gu.to(xu.dtype)], dim=2)
        need_Q = ctx.needs_input_grad[0]
        gQ, gq = raw_backward(ctx.cfg, factor, xu, grad_xu, need_Q=need_Q, status=status)
        return gQ, (gq if ctx.needs_input_grad[1] else None), None, None, None, None, None
