"""Stand-in for the reference's UN-SHIPPED ``rexquad_utils`` module.  TEST INFRASTRUCTURE ONLY.

The reference imports ``rexquad_utils`` (deqmpc/rex_quadrotor.py:3, deqmpc/flying_cartpole2d.py:3)
but the file is not in the tree, so the rotation helpers below are *parity unpinned*: they are the
textbook modified-Rodrigues-parameter / scalar-first-quaternion formulas, chosen to satisfy every
constraint visible at the call sites:

* ``mrp2quat(-m)`` must be the inverse rotation of ``mrp2quat(m)``      (rex_quadrotor.py:54,115)
* ``quatrot(q, v)`` must broadcast ``v[1,3]`` against ``q[...,4]`` of rank 2 *and* 3 and accept an
  fp32 ``v`` with an fp64 ``q`` (``self.g`` is a FloatTensor)            (rex_quadrotor.py:21,66,118)
* ``Spaces_np(shape, high, low)`` argument order                          (rex_quadrotor.py:167)
* everything used inside ``dynamics_`` must be TorchScript-compilable     (rex_quadrotor.py:153)

``oracle/gen_golden.py`` registers this module as ``sys.modules['rexquad_utils']`` so that the
reference's own ``rex_quadrotor.py`` / ``flying_cartpole2d.py`` can be imported unmodified.
Nothing in the product package may import this file.
"""
import math

import numpy as np
import torch


def deg2rad(deg: float) -> float:
    return deg * math.pi / 180.0


def mrp2quat(p: torch.Tensor) -> torch.Tensor:
    """MRP -> unit quaternion, scalar first: q = [1-|p|^2, 2p] / (1+|p|^2)."""
    n2 = (p * p).sum(dim=-1, keepdim=True)
    return torch.cat([1.0 - n2, 2.0 * p], dim=-1) / (1.0 + n2)


def quat2mrp(q: torch.Tensor) -> torch.Tensor:
    """unit quaternion (scalar first) -> MRP: p = q_v / (1 + q_w)."""
    return q[..., 1:] / (1.0 + q[..., 0:1])


def quatrot(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """Rotate v by unit quaternion q: v' = v + 2 w (r x v) + 2 r x (r x v)."""
    w = q[..., 0:1]
    r = q[..., 1:]
    v = v.to(q.dtype).expand_as(r)
    rv = torch.cross(r, v, dim=-1)
    return v + 2.0 * (w * rv) + 2.0 * torch.cross(r, rv, dim=-1)


def w2pdotkinematics_mrp(p: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """MRP kinematics: pdot = 1/4 ((1-|p|^2) w + 2 p x w + 2 (p.w) p)."""
    n2 = (p * p).sum(dim=-1, keepdim=True)
    pw = (p * w).sum(dim=-1, keepdim=True)
    return 0.25 * ((1.0 - n2) * w + 2.0 * torch.cross(p, w, dim=-1) + 2.0 * pw * p)


def mrp2rot(p: torch.Tensor) -> torch.Tensor:
    """MRP -> rotation matrix (columns = rotated basis vectors)."""
    q = mrp2quat(p)
    eye = torch.eye(3, dtype=p.dtype, device=p.device)
    cols = [quatrot(q, eye[i].expand(p.shape[:-1] + (3,))) for i in range(3)]
    return torch.stack(cols, dim=-1)


def euler_to_quaternion(r: torch.Tensor) -> torch.Tensor:
    """ZYX Euler angles [roll, pitch, yaw] -> scalar-first quaternion."""
    roll, pitch, yaw = r[..., 0], r[..., 1], r[..., 2]
    cr, sr = torch.cos(roll / 2), torch.sin(roll / 2)
    cp, sp = torch.cos(pitch / 2), torch.sin(pitch / 2)
    cy, sy = torch.cos(yaw / 2), torch.sin(yaw / 2)
    return torch.stack(
        [
            cr * cp * cy + sr * sp * sy,
            sr * cp * cy - cr * sp * sy,
            cr * sp * cy + sr * cp * sy,
            cr * cp * sy - sr * sp * cy,
        ],
        dim=-1,
    )


def rk4(f, x, u, dt):
    k1 = f(x, u)
    k2 = f(x + 0.5 * dt * k1, u)
    k3 = f(x + 0.5 * dt * k2, u)
    k4 = f(x + dt * k3, u)
    return x + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)


class Spaces_np:
    """Box space; argument order (shape, high, low) fixed by rex_quadrotor.py:167."""

    def __init__(self, shape, high=None, low=None):
        self.shape = shape
        self.high = np.full(shape, np.inf) if high is None else np.asarray(high)
        self.low = np.full(shape, -np.inf) if low is None else np.asarray(low)

    def sample(self):
        return np.random.uniform(self.low, self.high)


class Spaces(Spaces_np):
    pass
