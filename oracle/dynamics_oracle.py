"""CPU oracle for the discrete dynamics models on the AL-MPC hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
file; the product package never does (it fails loudly when its CUDA library is missing).

Every model is "one explicit RK4 step, zero-order-hold control" (SURVEY.md §7.4).  The continuous
ODEs restate:

* rex-quadrotor      deqmpc/rex_quadrotor.py:52-127   (helpers: oracle/rexquad_standin.py, unpinned)
* flying cartpole    deqmpc/flying_cartpole2d.py:50-133
* pendulum1l         deqmpc/my_envs/pendulum1l/src/generated_dynamics.c (CasADi) -> qdd = 4u - 19.62 sin q
* cartpole1l         deqmpc/my_envs/cartpole1l/src/generated_dynamics.c (CasADi)
                     M = [[11,-cos th],[-cos th,2]],  b = [sin th * thd^2, -9.81 sin th],  M qdd + b = [u,0]
* cartpole2l         deqmpc/my_envs/cartpole2l/src/generated_dynamics.c (CasADi) — Lagrangian restated
                     with autograd (cart 10 kg + two unit links, relative angles)

Pinning: pendulum1l / cartpole1l / cartpole2l are checked against the reference's own generated C
compiled into ``oracle/_ref`` (oracle/Makefile) and cartpole2l also against the shipped
``my_envs/traj.npz``; quadrotor / flying cartpole are checked against the reference's TorchScript
modules imported in the build container (tests/golden, oracle/gen_golden.py).

Jacobians: ``step_jac`` uses forward-mode ``torch.func.jacfwd`` under ``vmap`` (the reference
replicates rows and calls ``autograd.grad`` once, rex_quadrotor.py:136-144; both are exact
derivatives of the same RK4 map).  Convention: Jx[r,i,j] = d x+_i / d x_j, Ju[r,i,j] = d x+_i / d u_j
(al_utils.py:237-248).
"""
import math

import numpy as np
import torch


def _cross(a, b):
    ax, ay, az = a[..., 0], a[..., 1], a[..., 2]
    bx, by, bz = b[..., 0], b[..., 1], b[..., 2]
    return torch.stack([ay * bz - az * by, az * bx - ax * bz, ax * by - ay * bx], dim=-1)


def _mrp2quat(p):
    n2 = (p * p).sum(-1, keepdim=True)
    return torch.cat([1.0 - n2, 2.0 * p], -1) / (1.0 + n2)


def _quatrot(q, v):
    w, r = q[..., 0:1], q[..., 1:]
    v = v.to(q.dtype).expand_as(r)
    rv = _cross(r, v)
    return v + 2.0 * (w * rv) + 2.0 * _cross(r, rv)


def _mrp_kinematics(p, w):
    n2 = (p * p).sum(-1, keepdim=True)
    pw = (p * w).sum(-1, keepdim=True)
    return 0.25 * ((1.0 - n2) * w + 2.0 * _cross(p, w) + 2.0 * pw * p)


class _Model:
    """Base: RK4 discretisation (rex_quadrotor.py:98-107) + batched Jacobian."""

    nx: int
    nu: int
    dt: float
    name: str

    def f(self, x, u):  # continuous dynamics, x[..., nx], u[..., nu]
        raise NotImplementedError

    def step(self, x, u):
        dt = self.dt
        dt2 = dt / 2.0
        k1 = self.f(x, u)
        k2 = self.f(x + dt2 * k1, u)
        k3 = self.f(x + dt2 * k2, u)
        k4 = self.f(x + dt * k3, u)
        return x + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)

    __call__ = step

    def step_jac(self, x, u):
        jx, ju = torch.func.vmap(torch.func.jacfwd(self.step, argnums=(0, 1)))(x, u)
        return self.step(x, u), (jx, ju)


class RexQuadrotor(_Model):
    """rex_quadrotor.py:7-127.  Parameters that the reference stores as FloatTensors keep their
    fp32-rounded VALUES (J, Jinv=inv_fp32(J), g, 4*bf, ss/|ss|) — rex_quadrotor.py:17-21,29,45-46."""

    name = "rex_quadrotor"
    nx, nu = 12, 4

    def __init__(self, dt=0.05):
        J = [[0.01566089, 0.00000318037, 0.0], [0.00000318037, 0.01562078, 0.0], [0.0, 0.0, 0.02226868]]
        self.mass = 2.0
        self.J = torch.tensor(J, dtype=torch.float32)
        self.Jinv = torch.linalg.inv(self.J)
        self.g = torch.tensor([0.0, 0.0, -9.81], dtype=torch.float32)
        self.motor_dist = 0.28
        self.kf = 0.0244101
        self.bf = -30.48576
        self.km = 0.00029958
        self.Bf = torch.zeros(3, dtype=torch.float32)
        self.Bf[2] = 4 * self.bf
        ss = torch.tensor([[1.0, 1, 0], [1.0, -1, 0], [-1.0, -1, 0], [-1.0, 1, 0]], dtype=torch.float32)
        self.ss = ss / ss.norm(dim=-1, keepdim=True)
        self.dt = dt
        self.act_scale = 100.0
        self.u_hover_offset = 0.0
        # env-level constants (rex_quadrotor.py:164-168)
        self.Qlqr = [10.0] * 6 + [1.0] * 6
        self.Rlqr = [1e-8] * 4
        self.u_lower = [11.5] * 4
        self.u_upper = [18.3] * 4
        self.x_window = [5.0, 5.0, 5.0] + [math.radians(70)] * 3 + [0.5] * 3 + [0.25] * 3

    def _wrench(self, m, u):
        dt_ = m.dtype
        q_inv = _mrp2quat(-m)
        Fz = (self.kf * u).sum(-1)
        zero = torch.zeros_like(Fz)
        F = torch.stack([zero, zero, Fz], -1)
        # (self.m * self.g) is an fp32 product in the reference (rex_quadrotor.py:66)
        F = F + _quatrot(q_inv, (self.mass * self.g)) + self.Bf.to(dt_)
        tau3 = self.km * (u[..., 0] - u[..., 1] + u[..., 2] - u[..., 3])
        tau = torch.stack([zero, zero, tau3], -1)
        arm = self.motor_dist * self.ss.to(dt_)  # [4,3]
        thrust = self.kf * u + self.bf  # [...,4]
        zeros4 = torch.zeros_like(thrust)
        fvec = torch.stack([zeros4, zeros4, thrust], -1)  # [...,4,3]
        tau = tau + _cross(arm.expand_as(fvec), fvec).sum(-2)
        return F, tau

    def f(self, x, u):
        dt_ = x.dtype
        u = self.act_scale * (u + self.u_hover_offset)
        m, v, w = x[..., 3:6], x[..., 6:9], x[..., 9:12]
        q = _mrp2quat(m)
        F, tau = self._wrench(m, u)
        mdot = _mrp_kinematics(m, w)
        pdot = _quatrot(q, v)
        vdot = F / self.mass - _cross(w, v)
        J, Jinv = self.J.to(dt_), self.Jinv.to(dt_)
        Jw = (J * w.unsqueeze(-2)).sum(-1)
        wdot = (Jinv * (tau - _cross(w, Jw)).unsqueeze(-2)).sum(-1)
        return torch.cat([pdot, mdot, vdot, wdot], -1)


class FlyingCartpole(RexQuadrotor):
    """flying_cartpole2d.py:11-133 with the env-level constructor values of
    FlyingCartpole.__init__ (flying_cartpole2d.py:151): mass_q=2, mass_p=0.1, J=diag(.0023,.0023,.004),
    L=0.5, motor_dist=0.175, kf=1, km=0.025, no bf, u <- 10 (u + u_hover)."""

    name = "flying_cartpole"
    nx, nu = 14, 4

    def __init__(self, dt=0.05):
        self.mass = 2.0 + 0.1
        self.J = torch.tensor([[0.0023, 0.0, 0.0], [0.0, 0.0023, 0.0], [0.0, 0.0, 0.004]], dtype=torch.float32)
        self.Jinv = torch.linalg.inv(self.J)
        self.g = torch.tensor([0.0, 0.0, -9.81], dtype=torch.float32)
        self.motor_dist = 0.175
        self.kf = 1.0
        self.bf = 0.0
        self.km = 0.025
        self.Bf = torch.zeros(3, dtype=torch.float32)
        self.L = torch.tensor(0.5)
        ss = torch.tensor([[1.0, 1, 0], [1.0, -1, 0], [-1.0, -1, 0], [-1.0, 1, 0]], dtype=torch.float32)
        self.ss = ss / ss.norm(dim=-1, keepdim=True)
        self.dt = dt
        self.act_scale = 10.0
        # u_hover is an fp32 tensor in the reference (flying_cartpole2d.py:41)
        self.u_hover_offset = float(torch.tensor([(self.mass * 9.81) / self.act_scale / self.kf / 4]).item())
        self.Qlqr = [10.0] * 6 + [80.0] + [1.0] * 7
        self.Rlqr = [1e-8] * 4
        ub = float(np.array([0.3 * torch.tensor([self.u_hover_offset])[0]])[0])
        self.u_lower = [-ub] * 4
        self.u_upper = [ub] * 4
        self.x_window = [5.0] * 3 + [math.radians(45.0)] * 3 + [math.pi] + [1.0] * 7

    def f(self, x, u):
        dt_ = x.dtype
        u = self.act_scale * (u + self.u_hover_offset)
        m, v, w = x[..., 3:6], x[..., 7:10], x[..., 10:13]
        theta, theta_dot = x[..., 6:7], x[..., 13:14]
        q = _mrp2quat(m)
        F, tau = self._wrench(m, u)
        mdot = _mrp_kinematics(m, w)
        pdot = _quatrot(q, v)
        vdot = F / self.mass - _cross(w, v)
        J, Jinv = self.J.to(dt_), self.Jinv.to(dt_)
        Jw = (J * w.unsqueeze(-2)).sum(-1)
        wdot = (Jinv * (tau - _cross(w, Jw)).unsqueeze(-2)).sum(-1)
        x_ddot = _quatrot(q, vdot)[..., 0:1]
        theta_ddot = (self.g[2] * torch.sin(theta) + x_ddot * torch.cos(theta)) / self.L
        return torch.cat([pdot, mdot, theta_dot, vdot, wdot, theta_ddot], -1)


class Pendulum1L(_Model):
    """pendulum1l generated_dynamics.c: qdd = 4 tau - 2*9.81 sin q  (q=0 hanging)."""

    name = "pendulum"
    nx, nu = 2, 1

    def __init__(self, dt=0.05):
        self.dt = dt
        self.Qlqr = [1.0, 1.0]
        self.Rlqr = [1e-3]
        self.u_lower = [-2.0]
        self.u_upper = [2.0]

    def f(self, x, u):
        q, qd = x[..., 0:1], x[..., 1:2]
        return torch.cat([qd, 4.0 * u - 2.0 * (9.81 * torch.sin(q))], -1)


class Cartpole1L(_Model):
    """cartpole1l generated_dynamics.c: M(q) qdd + b(q,qd) = [u, 0]."""

    name = "cartpole1l"
    nx, nu = 4, 1

    def __init__(self, dt=0.05):
        self.dt = dt
        self.Qlqr = [1.0] * 4
        self.Rlqr = [1e-10]
        self.u_lower = [-100.0]
        self.u_upper = [100.0]

    def f(self, x, u):
        th, xd, thd = x[..., 1], x[..., 2], x[..., 3]
        s, c = torch.sin(th), torch.cos(th)
        r0 = u[..., 0] - s * thd * thd
        r1 = 9.81 * s
        det = 22.0 - c * c
        xdd = (2.0 * r0 + c * r1) / det
        thdd = (c * r0 + 11.0 * r1) / det
        return torch.stack([xd, thd, xdd, thdd], -1)


class Cartpole2L(_Model):
    """cartpole2l generated_dynamics.c: cart (10 kg) + 2 links (m=1, I=1 about COM, COM at distance 1,
    joint-to-joint length 1), relative angles, g=9.81, no damping; Lagrangian via autograd."""

    name = "cartpole2l"
    nx, nu = 6, 1

    def __init__(self, dt=0.03):
        self.dt = dt
        self.Qlqr = [1.0] * 6
        self.Rlqr = [1e-10]
        self.u_lower = [-250.0]
        self.u_upper = [250.0]

    @staticmethod
    def _mass_bias(q, qd):
        th1, th2 = q[..., 1], q[..., 2]
        thd1, thd2 = qd[..., 1], qd[..., 2]
        a = th1 + th2
        ad = thd1 + thd2
        c1, s1, ca, sa = torch.cos(th1), torch.sin(th1), torch.cos(a), torch.sin(a)
        c2 = torch.cos(th2)
        s2 = torch.sin(th2)
        one = torch.ones_like(c1)
        # p1 = (x - s1, c1), p2 = (x - s1 - sa, c1 + ca); omega1 = thd1, omega2 = thd1+thd2
        M00 = 12.0 * one
        M01 = -2.0 * c1 - ca
        M02 = -ca
        M11 = 4.0 + 2.0 * c2 + one
        M12 = 2.0 + c2
        M22 = 2.0 * one
        M = torch.stack(
            [torch.stack([M00, M01, M02], -1), torch.stack([M01, M11, M12], -1), torch.stack([M02, M12, M22], -1)], -2
        )
        # Coriolis/centrifugal + gravity (bias), derived from d/dt dL/dqd - dL/dq
        b0 = 2.0 * s1 * thd1 * thd1 + sa * ad * ad
        b1 = -s2 * (2.0 * thd1 * thd2 + thd2 * thd2) - 9.81 * (2.0 * s1 + sa)
        b2 = s2 * thd1 * thd1 - 9.81 * sa
        return M, torch.stack([b0, b1, b2], -1)

    def f(self, x, u):
        q, qd = x[..., :3], x[..., 3:]
        M, b = self._mass_bias(q, qd)
        tau = torch.cat([u, torch.zeros_like(qd[..., :2])], -1)
        # explicit 3x3 solve by Cramer's rule (torch.linalg.solve under vmap(jacfwd) returned wrong
        # tangents for batches > 1 with torch 2.11 — caught by the finite-difference self-check)
        r = tau - b
        a, bb, c = M[..., 0, 0], M[..., 0, 1], M[..., 0, 2]
        d, e, f_ = M[..., 1, 1], M[..., 1, 2], M[..., 2, 2]
        c00, c01, c02 = d * f_ - e * e, c * e - bb * f_, bb * e - c * d
        c11, c12, c22 = a * f_ - c * c, bb * c - a * e, a * d - bb * bb
        det = a * c00 + bb * c01 + c * c02
        r0, r1, r2 = r[..., 0], r[..., 1], r[..., 2]
        qdd = torch.stack([c00 * r0 + c01 * r1 + c02 * r2, c01 * r0 + c11 * r1 + c12 * r2,
                           c02 * r0 + c12 * r1 + c22 * r2], -1) / det.unsqueeze(-1)
        return torch.cat([qd, qdd], -1)


MODELS = {
    "pendulum": Pendulum1L,
    "cartpole1l": Cartpole1L,
    "cartpole2l": Cartpole2L,
    "rex_quadrotor": RexQuadrotor,
    "flying_cartpole": FlyingCartpole,
}


def make_model(name, dt=None):
    cls = MODELS[name]
    return cls() if dt is None else cls(dt=dt)


# --------------------------------------------------------------------------------------------------
# ctypes access to the reference's own CasADi C, compiled by oracle/Makefile into oracle/_ref/
# --------------------------------------------------------------------------------------------------
class CasadiRef:
    """Calls eval_forward_dynamics / eval_forward_derivatives of the reference's generated C
    (my_envs/*/src/generated_dynamics.c, generated_derivatives.c) one sample at a time."""

    def __init__(self, pkg, nq):
        import ctypes
        import os

        here = os.path.dirname(os.path.abspath(__file__))
        path = os.path.join(here, "_ref", f"lib{pkg}_casadi.so")
        self.lib = ctypes.CDLL(path)
        self.nq = nq
        self.ct = ctypes

    def _call(self, fn, q, qd, tau, h, n_out, out_sizes):
        ct = self.ct
        dbl_p = ct.POINTER(ct.c_double)
        ins = [np.ascontiguousarray(a, dtype=np.float64) for a in (q, qd, tau, np.array([h]))]
        outs = [np.zeros(s, dtype=np.float64) for s in out_sizes]
        arg = (dbl_p * 4)(*[a.ctypes.data_as(dbl_p) for a in ins])
        res = (dbl_p * n_out)(*[o.ctypes.data_as(dbl_p) for o in outs])
        iw = (ct.c_longlong * 64)()
        w = (ct.c_double * 4096)()
        fn.restype = ct.c_int
        fn(arg, res, iw, w, 0)
        return outs

    def step(self, x, u, h):
        """x[nx], u scalar -> x+[nx] (tau only on the first joint, my_envs/dynamics.py:55-56)."""
        nq = self.nq
        tau = np.zeros(nq)
        tau[0] = u
        qn, qdn = self._call(self.lib.eval_forward_dynamics, x[:nq], x[nq:], tau, h, 2, [nq, nq])
        return np.concatenate([qn, qdn])

    def jac(self, x, u, h):
        """-> Jx[nx,nx], Ju[nx,1] in the al_utils convention; raw CasADi blocks are column-major
        (my_envs/dynamics.py:97-108 does the cat/transposes)."""
        nq = self.nq
        tau = np.zeros(nq)
        tau[0] = u
        blocks = self._call(self.lib.eval_forward_derivatives, x[:nq], x[nq:], tau, h, 6, [nq * nq] * 6)
        # column-major [nq,nq] == row-major transposed: B[k][in, out]
        qq, qqd, qt, qdq, qdqd, qdt = [b.reshape(nq, nq) for b in blocks]  # indexed [in_j, out_i]
        Jx = np.zeros((2 * nq, 2 * nq))
        Jx[:nq, :nq] = qq.T
        Jx[:nq, nq:] = qqd.T
        Jx[nq:, :nq] = qdq.T
        Jx[nq:, nq:] = qdqd.T
        Ju = np.concatenate([qt.T[:, 0:1], qdt.T[:, 0:1]], 0)
        return Jx, Ju
