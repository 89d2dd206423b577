"""CPU: pin the oracle dynamics models (oracle/dynamics_oracle.py) to the reference's own CasADi C
(compiled into oracle/_ref by oracle/Makefile; skipped where it is absent), to the shipped
2-link trajectory, and to finite differences for batches > 1."""
import os

import numpy as np
import pytest
import torch

from oracle.dynamics_oracle import CasadiRef
from oracle.problem_gen import make_oracle_model

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(os.path.dirname(HERE), "oracle", "_ref")


@pytest.mark.parametrize("model", ["pendulum", "cartpole1l", "cartpole2l", "rex_quadrotor", "flying_cartpole"])
def test_batched_jacobian_matches_finite_differences(model):
    om = make_oracle_model(model)
    g = torch.Generator().manual_seed(5)
    R = 33
    x = (torch.rand(R, om.nx, generator=g, dtype=torch.float64) * 2 - 1) * 0.8
    lo, hi = torch.tensor(om.u_lower), torch.tensor(om.u_upper)
    u = lo + (hi - lo) * torch.rand(R, om.nu, generator=g, dtype=torch.float64)
    _, (jx, ju) = om.step_jac(x, u)
    h = 1e-6
    for j in range(om.nx):
        e = torch.zeros(om.nx, dtype=torch.float64)
        e[j] = h
        fd = (om.step(x + e, u) - om.step(x - e, u)) / (2 * h)
        assert (fd - jx[:, :, j]).abs().max() < 1e-6 * max(1.0, jx.abs().max().item())
    for j in range(om.nu):
        e = torch.zeros(om.nu, dtype=torch.float64)
        e[j] = h
        fd = (om.step(x, u + e) - om.step(x, u - e)) / (2 * h)
        assert (fd - ju[:, :, j]).abs().max() < 1e-6 * max(1.0, ju.abs().max().item())


@pytest.mark.parametrize("model,pkg,nq", [("pendulum", "pendulum1l", 1), ("cartpole1l", "cartpole1l", 2), ("cartpole2l", "cartpole2l", 3)])
def test_matches_reference_casadi_c(model, pkg, nq):
    if not os.path.exists(os.path.join(REF_DIR, f"lib{pkg}_casadi.so")):
        pytest.skip("oracle/_ref not built (needs /root/reference: `make -C oracle`)")
    ref = CasadiRef(pkg, nq)
    om = make_oracle_model(model)
    rng = np.random.default_rng(0)
    xs = rng.uniform(-3, 3, (12, 2 * nq))
    us = rng.uniform(om.u_lower[0], om.u_upper[0], (12, 1))
    y, (jx, ju) = om.step_jac(torch.from_numpy(xs), torch.from_numpy(us))
    for i in range(12):
        xn = ref.step(xs[i], us[i, 0], om.dt)
        Jx, Ju = ref.jac(xs[i], us[i, 0], om.dt)
        assert np.abs(y[i].numpy() - xn).max() < 1e-12 * max(1.0, np.abs(xn).max())
        assert np.abs(jx[i].numpy() - Jx).max() < 1e-10 * max(1.0, np.abs(Jx).max())
        assert np.abs(ju[i].numpy() - Ju).max() < 1e-10 * max(1.0, np.abs(Ju).max())


def test_cartpole2l_reproduces_shipped_trajectory():
    """my_envs/traj.npz: X[k+1] = f(X[k], U[k]) at dt = 0.03 (test_trajopt_altro.jl:82)."""
    z = np.load(os.path.join(HERE, "golden", "cartpole2l_traj.npz"))
    X, U = torch.from_numpy(z["X_np"]), torch.from_numpy(z["U_np"])
    om = make_oracle_model("cartpole2l")
    assert (om.step(X[:-1], U) - X[1:]).abs().max().item() < 2e-6
