"""Native dynamics plug-ins: the host-side descriptors of the device functors fused into the solver.

A ``NativeDynamics`` object plays the role of the reference's ``dx`` / ``dx_jac`` callables
(qpth/al_utils.py:215,237): it can be called as ``dx(x, u) -> x+`` and as
``dx.jacobian(x, u) -> (x+, (Jx, Ju))`` (both run the CUDA functor through the C ABI), and it tells
``MPC.forward`` which device functor (``dyn_id``) and parameter block to fuse into the solve.

``resolve(dx)`` also recognises the REFERENCE's own dynamics objects by class name / attributes
(RexQuadrotor_dynamics, FlyingCartpole_dynamics — eager or TorchScript — and the my_envs
``Dynamics`` wrappers of the pendulum1l / cartpole1l / cartpole2l extensions), so a reference user's
``env.dynamics`` can be passed unchanged.

Parameter block layout: see csrc/almpc_models.cuh.  Values the reference stores as FloatTensors
(J, inv(J), g, 4*bf, ss/|ss|, u_hover) are rounded to fp32 exactly where the reference rounds them
(rex_quadrotor.py:17-21,29,45-46; flying_cartpole2d.py:41).
"""
import ctypes

import torch

from . import _lib

_J_REX = [[0.01566089, 0.00000318037, 0.0], [0.00000318037, 0.01562078, 0.0], [0.0, 0.0, 0.02226868]]


class NativeDynamics(torch.nn.Module):
    dyn_id = -1
    nx = 0
    nu = 0

    def __init__(self, dt):
        super().__init__()
        self.dt = float(dt)

    # ---- parameter block -------------------------------------------------------------------
    def param_block(self):
        """The functor's parameter block (csrc/almpc_models.cuh).  Memoised: the quadrotor family inverts its inertia
        tensor and rounds a dozen values through fp32 tensors here, which costs more than a small-batch solve when
        it is redone on every call.  The model parameters are constructor arguments; change them -> build a new
        object (or call ``invalidate()``)."""
        pb = self.__dict__.get("_param_block")
        if pb is None:
            p = [0.0] * _lib.MAX_PARAMS
            p[0] = self.dt
            self._fill_params(p)
            pb = self.__dict__["_param_block"] = tuple(p)
        return list(pb)

    def invalidate(self):
        self.__dict__.pop("_param_block", None)
        self.__dict__.pop("_dyn_desc", None)

    def _fill_params(self, p):
        pass

    def _desc(self):
        d = self.__dict__.get("_dyn_desc")
        if d is None:
            d = _lib.Desc()
            d.dyn_id = self.dyn_id
            d.T = 2
            d.al_iter = 1
            d.max_newton = 1
            for i, v in enumerate(self.param_block()):
                d.dyn_params[i] = v
            self.__dict__["_dyn_desc"] = d
        return d

    # ---- dx / dx_jac -----------------------------------------------------------------------
    def _run(self, x, u, want_jac):
        lib = _lib.load()
        _lib.require_cuda(x, "NativeDynamics")
        sfx = _lib.suffix(x.dtype)
        x = x.contiguous()
        u = u.to(x.dtype).contiguous()
        R = x.shape[0]
        assert x.shape == (R, self.nx) and u.shape == (R, self.nu)
        xn = torch.empty_like(x)
        Jx = torch.empty(R, self.nx, self.nx, dtype=x.dtype, device=x.device) if want_jac else None
        Ju = torch.empty(R, self.nx, self.nu, dtype=x.dtype, device=x.device) if want_jac else None
        d = self._desc()
        with torch.cuda.device(x.device):
            rc = getattr(lib, "almpc_dynamics_" + sfx)(
                ctypes.byref(d), R, _lib.ptr(x), _lib.ptr(u), _lib.ptr(xn), _lib.ptr(Jx), _lib.ptr(Ju), _lib.stream_ptr()
            )
        _lib.check(rc, "almpc_dynamics")
        return xn, Jx, Ju

    def forward(self, x, u):
        return self._run(x, u, False)[0]

    def jacobian(self, x, u):
        xn, Jx, Ju = self._run(x, u, True)
        return xn, (Jx, Ju)


class _QuadFamily(NativeDynamics):
    def _quad_params(self, p, mass, J, gravity, motor_dist, kf_force, kf_moment, bf, km, act_scale, u_offset):
        J32 = torch.tensor(J, dtype=torch.float32)
        if J32.dim() == 1:
            J32 = torch.diag(J32)
        Jinv32 = torch.linalg.inv(J32)
        g32 = torch.tensor(gravity, dtype=torch.float32)
        mg32 = mass * g32  # fp32 product, rex_quadrotor.py:66
        bf32 = torch.zeros(3, dtype=torch.float32)
        bf32[2] = 4 * bf
        ss = torch.tensor([[1.0, 1, 0], [1.0, -1, 0], [-1.0, -1, 0], [-1.0, 1, 0]], dtype=torch.float32)
        ss = ss / ss.norm(dim=-1, keepdim=True)
        arm = motor_dist * ss.double()
        p[1] = act_scale
        p[2] = u_offset
        p[3] = kf_force
        p[4] = kf_moment
        p[5] = bf
        p[6] = km
        p[7] = 1.0 / mass
        for i in range(3):
            p[8 + i] = float(mg32[i])
            p[11 + i] = float(bf32[i])
        for i in range(3):
            for j in range(3):
                p[14 + 3 * i + j] = float(J32[i, j])
                p[23 + 3 * i + j] = float(Jinv32[i, j])
        for i in range(4):
            p[32 + 2 * i] = float(arm[i, 0])
            p[32 + 2 * i + 1] = float(arm[i, 1])


class RexQuadrotorDynamics(_QuadFamily):
    """deqmpc/rex_quadrotor.py:7-127 (constructor argument names kept)."""

    dyn_id = _lib.DYN_REX_QUADROTOR
    nx, nu = 12, 4

    def __init__(self, mass=2.0, J=_J_REX, gravity=(0, 0, -9.81), motor_dist=0.28, kf=0.0244101, bf=-30.48576,
                 km=0.00029958, dt=0.05, act_scale=100.0):
        super().__init__(dt)
        self.mass, self.J, self.gravity, self.motor_dist = mass, J, list(gravity), motor_dist
        self.kf, self.bf, self.km, self.act_scale = kf, bf, km, act_scale

    def _fill_params(self, p):
        # forces() hard-codes kf = 0.0244101 (rex_quadrotor.py:55); moments() uses self.kf (:72,83)
        self._quad_params(p, self.mass, self.J, self.gravity, self.motor_dist, 0.0244101, self.kf, self.bf, self.km,
                          self.act_scale, 0.0)


class FlyingCartpoleDynamics(_QuadFamily):
    """deqmpc/flying_cartpole2d.py:11-133 with the env-level defaults of FlyingCartpole.__init__ (:151)."""

    dyn_id = _lib.DYN_FLYING_CARTPOLE
    nx, nu = 14, 4

    def __init__(self, mass_q=2.0, mass_p=0.1, J=((0.0023, 0.0, 0.0), (0.0, 0.0023, 0.0), (0.0, 0.0, 0.004)), L=0.5,
                 gravity=(0, 0, -9.81), motor_dist=0.175, kf=1.0, km=0.025, dt=0.05, act_scale=10.0):
        super().__init__(dt)
        self.mass = mass_q + mass_p
        self.J, self.gravity, self.motor_dist = [list(r) for r in J], list(gravity), motor_dist
        self.kf, self.km, self.act_scale, self.L = kf, km, act_scale, L
        # fp32 tensor in the reference (flying_cartpole2d.py:41)
        self.u_hover = float(torch.tensor([(-self.mass * gravity[2]) / act_scale / kf / 4])[0])

    def _fill_params(self, p):
        self._quad_params(p, self.mass, self.J, self.gravity, self.motor_dist, self.kf, self.kf, 0.0, self.km,
                          self.act_scale, self.u_hover)
        p[40] = float(torch.tensor(self.gravity, dtype=torch.float32)[2])
        p[41] = 1.0 / float(torch.tensor(self.L))


class PendulumDynamics(NativeDynamics):
    """my_envs/pendulum1l (CasADi): qdd = 4 u - 19.62 sin q."""

    dyn_id = _lib.DYN_PENDULUM
    nx, nu = 2, 1

    def __init__(self, dt=0.05):
        super().__init__(dt)


class Cartpole1LDynamics(NativeDynamics):
    """my_envs/cartpole1l (CasADi)."""

    dyn_id = _lib.DYN_CARTPOLE1L
    nx, nu = 4, 1

    def __init__(self, dt=0.05):
        super().__init__(dt)


class Cartpole2LDynamics(NativeDynamics):
    """my_envs/cartpole2l (CasADi)."""

    dyn_id = _lib.DYN_CARTPOLE2L
    nx, nu = 6, 1

    def __init__(self, dt=0.03):
        super().__init__(dt)


BY_NAME = {
    "pendulum": PendulumDynamics,
    "cartpole1l": Cartpole1LDynamics,
    "cartpole2l": Cartpole2LDynamics,
    "rex_quadrotor": RexQuadrotorDynamics,
    "flying_cartpole": FlyingCartpoleDynamics,
}


def make(name, **kw):
    return BY_NAME[name](**kw)


def _attr(obj, name, default=None):
    try:
        return getattr(obj, name)
    except Exception:
        return default


_RESOLVED = {}  # id(reference dynamics object) -> (weakref to it, NativeDynamics): resolve() runs on every MPC.forward


def resolve(dx):
    """Map a dynamics object (ours or the reference's) to a NativeDynamics descriptor, or raise.  The descriptor of
    a reference object is built once (reading its parameter tensors synchronises with the device) and cached for
    the lifetime of that object."""
    if isinstance(dx, NativeDynamics):
        return dx
    native = _attr(dx, "almpc_native")
    if isinstance(native, NativeDynamics):
        return native
    hit = _RESOLVED.get(id(dx))
    if hit is not None and hit[0]() is dx:
        return hit[1]
    native = _resolve_uncached(dx)
    try:
        import weakref

        _RESOLVED[id(dx)] = (weakref.ref(dx, lambda _r, k=id(dx): _RESOLVED.pop(k, None)), native)
    except TypeError:  # object without weak-reference support: resolve every time
        pass
    return native


def _resolve_uncached(dx):
    cls = _attr(dx, "original_name") or type(dx).__name__
    dt = _attr(dx, "dt")
    if cls in ("RexQuadrotor_dynamics", "RexQuadrotor_dynamics_jac"):
        J = _attr(dx, "J")
        return RexQuadrotorDynamics(
            mass=float(dx.m), J=J.reshape(3, 3).double().tolist() if torch.is_tensor(J) else _J_REX,
            gravity=dx.g.reshape(3).double().tolist(), motor_dist=float(dx.motor_dist), kf=float(dx.kf),
            bf=float(dx.bf), km=float(dx.km), dt=float(dt), act_scale=float(dx.act_scale))
    if cls in ("FlyingCartpole_dynamics", "FlyingCartpole_dynamics_jac"):
        J = dx.J.reshape(3, 3).double().tolist()
        return FlyingCartpoleDynamics(
            mass_q=float(dx.mq), mass_p=float(dx.mp), J=J, L=float(dx.L), gravity=dx.g.reshape(3).double().tolist(),
            motor_dist=float(dx.motor_dist), kf=float(dx.kf), km=float(dx.km), dt=float(dt),
            act_scale=float(dx.act_scale))
    pkg = _attr(dx, "package")
    pkg_name = _attr(pkg, "__name__", "") if pkg is not None else ""
    if pkg_name.startswith("pendulum1l"):
        return PendulumDynamics(dt=float(dt))
    if pkg_name.startswith("cartpole1l"):
        return Cartpole1LDynamics(dt=float(dt))
    if pkg_name.startswith("cartpole2l"):
        return Cartpole2LDynamics(dt=float(dt))
    raise _lib.AlmpcError(
        f"no native CUDA functor registered for dynamics object {cls!r}: pass a deq_mpc_corl_b200.dynamics."
        "NativeDynamics for the fused kernels, or callable dx AND dx_jac for the generic-callable mode (torch evaluates "
        "them on the GPU, the block-tridiagonal linear algebra runs in CUDA); there is no CPU fallback"
    )
