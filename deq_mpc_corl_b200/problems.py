"""Synthetic AL-MPC problem batches — pure torch, no solver, no oracle.

Used by bench.py, tests and the golden generator so that every leg sees identical inputs.
Constants restate the reference environments:

* rex-quadrotor     deqmpc/rex_quadrotor.py:164-168 (Qlqr, Rlqr, action bounds [11.5,18.3], x_window)
* flying cartpole   deqmpc/flying_cartpole2d.py:166-173
* cartpole 1-/2-link deqmpc/my_envs/cartpole.py:60-79 (bounds +-100 / +-250, Q=1, R=1e-10)
* pendulum          env not shipped; Gym-style +-2 torque bound, Q=1, R=1e-3 (assumption)
* cost build        deqmpc/policies.py:1285-1297 (p = -Q*xu_ref, f = 1/2 xu_ref^T Q xu_ref)
* x_ref recipe      SURVEY.md §8d (x_ref[:,0]=x0, decaying copy + 0.05 N(0,1)); u_ref = 0 (deq_layer.py:90)
"""
import math

import torch

MODEL_SPECS = {
    # name: nx, nu, dt, Q diag, R diag, u_lower, u_upper
    "pendulum": dict(nx=2, nu=1, dt=0.05, Q=[1.0, 1.0], R=[1e-3], u_lower=[-2.0], u_upper=[2.0]),
    "cartpole1l": dict(nx=4, nu=1, dt=0.05, Q=[1.0] * 4, R=[1e-10], u_lower=[-100.0], u_upper=[100.0]),
    "cartpole2l": dict(nx=6, nu=1, dt=0.03, Q=[1.0] * 6, R=[1e-10], u_lower=[-250.0], u_upper=[250.0]),
    "rex_quadrotor": dict(nx=12, nu=4, dt=0.05, Q=[10.0] * 6 + [1.0] * 6, R=[1e-8] * 4,
                          u_lower=[11.5] * 4, u_upper=[18.3] * 4),
    "flying_cartpole": dict(nx=14, nu=4, dt=0.05, Q=[10.0] * 6 + [80.0] + [1.0] * 7, R=[1e-8] * 4,
                            u_lower=None, u_upper=None),  # +-0.3*u_hover, filled below
}


def _flying_u_bound():
    # u_hover is an fp32 tensor in the reference (flying_cartpole2d.py:41); bound = 0.3*u_hover (:172)
    u_hover = torch.tensor([(2.1 * 9.81) / 10.0 / 1.0 / 4.0])
    return float((0.3 * u_hover)[0])


_fb = _flying_u_bound()
MODEL_SPECS["flying_cartpole"]["u_lower"] = [-_fb] * 4
MODEL_SPECS["flying_cartpole"]["u_upper"] = [_fb] * 4


def euler_to_mrp(e):
    """ZYX Euler [roll,pitch,yaw] -> scalar-first quaternion -> MRP (reset_torch, rex_quadrotor.py:270-271)."""
    roll, pitch, yaw = e[..., 0], e[..., 1], e[..., 2]
    cr, sr = torch.cos(roll / 2), torch.sin(roll / 2)
    cp, sp = torch.cos(pitch / 2), torch.sin(pitch / 2)
    cy, sy = torch.cos(yaw / 2), torch.sin(yaw / 2)
    qw = cr * cp * cy + sr * sp * sy
    qv = torch.stack([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy], -1)
    return qv / (1.0 + qw[..., None])


def sample_x0(model, B, gen):
    spec = MODEL_SPECS[model]
    nx = spec["nx"]
    r = torch.rand(B, nx, generator=gen, dtype=torch.float64) * 2 - 1
    if model == "rex_quadrotor":
        win = torch.tensor([5.0] * 3 + [math.radians(70)] * 3 + [0.5] * 3 + [0.25] * 3, dtype=torch.float64)
        x = r * win * 0.3
        return torch.cat([x[:, :3], euler_to_mrp(x[:, 3:6]), x[:, 6:]], -1)
    if model == "flying_cartpole":
        win = torch.tensor([5.0] * 3 + [math.radians(45)] * 3 + [math.pi] + [1.0] * 7, dtype=torch.float64)
        x = r * win * 0.3
        return torch.cat([x[:, :3], euler_to_mrp(x[:, 3:6]), math.pi + x[:, 6:7], x[:, 7:]], -1)
    if model == "pendulum":
        return r * torch.tensor([math.pi, 1.0], dtype=torch.float64)
    return r * math.pi  # cartpoles: U(-pi,pi) on all dims (cartpole.py:122-131)


def make_inputs(model, T, B, seed):
    """-> dict(x0[B,nx] f32, x_ref[B,T,nx] f32, u_ref[B,T,nu] f32, Qd[n] f64, u_lower, u_upper)."""
    spec = MODEL_SPECS[model]
    nx, nu = spec["nx"], spec["nu"]
    gen = torch.Generator().manual_seed(seed)
    x0 = sample_x0(model, B, gen)
    x_ref = torch.empty(B, T, nx, dtype=torch.float64)
    x_ref[:, 0] = x0
    for t in range(1, T):
        x_ref[:, t] = x0 * (1.0 - t / (2.0 * (T - 1))) + 0.05 * torch.randn(B, nx, generator=gen, dtype=torch.float64)
    return dict(
        x0=x0.float(),
        x_ref=x_ref.float(),
        u_ref=torch.zeros(B, T, nu, dtype=torch.float32),
        Qd=torch.tensor(spec["Q"] + spec["R"], dtype=torch.float64),
        u_lower=torch.tensor(spec["u_lower"], dtype=torch.float64),
        u_upper=torch.tensor(spec["u_upper"], dtype=torch.float64),
    )


def tracking_cost(Qd, x_ref, u_ref, dtype=torch.float64):
    """Diagonal tracking cost of Tracking_MPC.compute_pf (policies.py:1285-1297):
    -> Qd[B,T,n], q[B,T,n] = -Qd*xu_ref, f[B,T]."""
    xu_ref = torch.cat([x_ref, u_ref], -1)
    B, T, n = xu_ref.shape
    Q = Qd.to(dtype).to(xu_ref.device).expand(B, T, n).contiguous()
    q = -(Q * xu_ref)
    f = 0.5 * (xu_ref * Q * xu_ref).sum(-1)
    return Q, q, f
