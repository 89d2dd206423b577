"""Interior-point path: `DenseQPFunction` -> pdipm `batch_LU` (reference qpth/qp.py:187-270,
qpth/solvers/pdipm/batch_LU.py:29-244).

CPU  : the oracle restatement (oracle/qp_oracle.py) against goldens of the UNMODIFIED reference (tests/golden/ip_*.npz).
GPU  : the CUDA path (deq_mpc_corl_b200/qp.py -> almpc_kkt_solve_*) against the same goldens, against the oracle on
       seeded problems, and the KKT kernel alone against a dense LU of the full KKT matrix.
Tolerance: 1e-6 relative (north_star, fp64) on the solution; the gradients divide by the complementarity of the best
iterate (mu ~ 1e-10 at exit), so their budget is 1e-5.
"""
import os

import numpy as np
import pytest
import torch

from golden_util import GOLDEN_DIR, rel_err
from oracle import qp_oracle

cuda = torch.cuda.is_available()
DEV = "cuda:0"
IP_GOLDENS = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.startswith("ip_") and f.endswith(".npz"))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files if not k.startswith("meta_")}, str(z["meta_kind"])


def quad_dyn_res(step, x0, B, T=5, nx=12, nu=4):
    """qp_wrapper.py:316-342: [f(x_t,u_t) - x_{t+1} ; x_0 - x0] on the stacked (x_t | u_t) vector."""
    n = nx + nu

    def dyn_res(z):
        zz = z.reshape(B, T, n)
        x, u = zz[:, :, :nx], zz[:, :, nx:]
        xn = step(x[:, :-1].reshape(-1, nx).contiguous(), u[:, :-1].reshape(-1, nu).contiguous()).reshape(B, T - 1, nx)
        return torch.cat(((xn - x[:, 1:]).reshape(B, -1), x[:, 0] - x0), 1)

    return dyn_res


@pytest.mark.parametrize("name", IP_GOLDENS)
def test_oracle_matches_reference_ip(name):
    g, kind = load(name)
    B = int(g["B"])
    if kind == "random":
        dyn_res = lambda x: (g["A"] @ x.unsqueeze(-1)).squeeze(-1) - g["b"]
    else:
        from oracle.problem_gen import make_oracle_model

        dyn_res = quad_dyn_res(make_oracle_model("rex_quadrotor").step, g["x0"], B)
    trace = []
    x, y, z, s = qp_oracle.pdipm_forward(g["Q"], g["p"], g["G"], g["h"], g["A"], g["b"], dyn_res, trace=trace)
    assert 1 + 2 * sum(1 for t in trace if t.get("solved")) == int(g["n_solve_kkt"]), "interior-point iteration count differs"
    assert rel_err(x, g["zhat"]) < 1e-8
    grads = qp_oracle.qp_backward(g["Q"], g["G"], g["A"], x, y, z, s, g["dl_dzhat"])
    for got, key in zip(grads, ("dQ", "dp", "dG", "dh", "dA", "db")):
        assert rel_err(got, g[key]) < 1e-5, key


# ------------------------------------------------------------------------------------------------------------------
gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("nz,nineq,neq,B,eps", [(10, 6, 3, 17, 1e-7), (80, 40, 60, 5, 1e-7), (12, 0, 4, 3, 1e-7),
                                                (9, 5, 0, 4, 1e-7), (80, 40, 60, 3, 0.0), (33, 17, 11, 2, 0.0)])
def test_kkt_kernel_matches_dense_lu(nz, nineq, neq, B, eps):
    """almpc_kkt_solve (closed-form elimination + L D L^T of the reduced system + 1 refinement) against the
    reference's route on the dense (nz + 2 nineq + neq)^2 matrix (oracle: pivoted LU + 1 refinement)."""
    from deq_mpc_corl_b200 import qp

    g = torch.Generator().manual_seed(nz * 131 + nineq)
    rnd = lambda *sh: torch.randn(*sh, generator=g, dtype=torch.float64)
    L = rnd(B, nz, nz)
    Q = L @ L.transpose(1, 2) / nz + 0.3 * torch.eye(nz, dtype=torch.float64)
    G, A = rnd(B, nineq, nz), rnd(B, neq, nz)
    s, z = 0.05 + torch.rand(B, nineq, generator=g, dtype=torch.float64) * 3, 0.05 + torch.rand(B, nineq, generator=g, dtype=torch.float64) * 3
    rx, rs, rz, ry = rnd(B, nz), rnd(B, nineq), rnd(B, nineq), rnd(B, neq)
    ref = qp_oracle.solve_kkt(Q, G, A, s, z, eps, rx, rs, rz, ry)
    got = qp.solve_kkt(*(t.to(DEV) for t in (Q, G, A, s, z)), eps, *(t.to(DEV) for t in (rx, rs, rz, ry)))
    for a, b, nm in zip(got, ref, ("dx", "ds", "dz", "dy")):
        if b.numel():
            assert rel_err(a.cpu(), b) < 1e-9, (nm, rel_err(a.cpu(), b))


@gpu
def test_kkt_kernel_backward_solve_with_active_constraints():
    """The backward pass solves with K at a converged iterate: active rows have s ~ 1e-12, inactive rows z ~ 1e-12.
    Regularised factorisation + three refinement steps against K must reproduce the reference's LU of K itself."""
    from deq_mpc_corl_b200 import qp

    nz, nineq, neq, B = 14, 9, 4, 6
    g = torch.Generator().manual_seed(5)
    rnd = lambda *sh: torch.randn(*sh, generator=g, dtype=torch.float64)
    L = rnd(B, nz, nz)
    Q = L @ L.transpose(1, 2) / nz + 0.3 * torch.eye(nz, dtype=torch.float64)
    G, A = rnd(B, nineq, nz), rnd(B, neq, nz)
    active = torch.rand(B, nineq, generator=g) < 0.4
    big = 0.2 + torch.rand(B, nineq, generator=g, dtype=torch.float64)
    s = torch.where(active, 1e-12 * big, big)
    z = torch.where(active, big, 1e-12 * big)
    rx = rnd(B, nz)
    zero = lambda n: torch.zeros(B, n, dtype=torch.float64)
    ref = qp_oracle.solve_kkt(Q, G, A, s, z, 0.0, rx, zero(nineq), zero(nineq), zero(neq))
    got = qp.solve_kkt(*(t.to(DEV) for t in (Q, G, A, s, z)), 1e-7, *(t.to(DEV) for t in (rx, zero(nineq), zero(nineq), zero(neq))),
                       n_refine=3)
    for a, b, nm in zip(got, ref, ("dx", "ds", "dz", "dy")):
        assert rel_err(a.cpu(), b) < 1e-7, (nm, rel_err(a.cpu(), b))


@gpu
@pytest.mark.parametrize("name", IP_GOLDENS if cuda else [])
def test_dense_qp_function_matches_reference_golden(name):
    from deq_mpc_corl_b200.dynamics import make as make_dyn
    from deq_mpc_corl_b200.qp import DenseQPFunction, pdipm_forward

    g, kind = load(name)
    B = int(g["B"])
    d = {k: g[k].to(DEV) for k in ("Q", "p", "G", "h", "A", "b")}
    if kind == "random":
        dyn_res = lambda x: (d["A"] @ x.unsqueeze(-1)).squeeze(-1) - d["b"]
    else:
        dyn_res = quad_dyn_res(make_dyn("rex_quadrotor"), g["x0"].to(DEV), B)
    trace = []
    pdipm_forward(d["Q"], d["p"], d["G"], d["h"], d["A"], d["b"], dyn_res, trace=trace)
    assert 1 + 2 * sum(1 for t in trace if t.get("solved")) == int(g["n_solve_kkt"]), "interior-point iteration count differs"
    leaves = {k: v.clone().requires_grad_(True) for k, v in d.items()}
    zhat = DenseQPFunction()(leaves["Q"], leaves["p"], leaves["G"], leaves["h"], leaves["A"], leaves["b"], dyn_res)
    e = rel_err(zhat.detach().cpu(), g["zhat"])
    grads = torch.autograd.grad(zhat, [leaves[k] for k in ("Q", "p", "G", "h", "A", "b")], g["dl_dzhat"].to(DEV))
    errs = {key: rel_err(got.cpu(), g[key]) for got, key in zip(grads, ("dQ", "dp", "dG", "dh", "dA", "db"))}
    print(f"{name}: zhat {e:.1e} " + " ".join(f"{k} {v:.1e}" for k, v in errs.items()))
    assert e < 1e-6
    assert all(v < 1e-5 for v in errs.values()), errs


@gpu
def test_dense_qp_function_matches_oracle_on_seeded_batch_and_refuses_cpu():
    from deq_mpc_corl_b200 import _lib
    from deq_mpc_corl_b200.qp import DenseQPFunction

    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    from oracle.gen_golden import ip_problem

    prob, _ = ip_problem("random", 77, 64, torch.float64)
    dyn_o = lambda x: (prob["A"] @ x.unsqueeze(-1)).squeeze(-1) - prob["b"]
    xo, _, _, _ = qp_oracle.pdipm_forward(prob["Q"], prob["p"], prob["G"], prob["h"], prob["A"], prob["b"], dyn_o)
    d = {k: v.to(DEV) for k, v in prob.items()}
    dyn_c = lambda x: (d["A"] @ x.unsqueeze(-1)).squeeze(-1) - d["b"]
    xc = DenseQPFunction()(d["Q"], d["p"], d["G"], d["h"], d["A"], d["b"], dyn_c)
    assert rel_err(xc.cpu(), xo) < 1e-6
    # feasibility of the returned point
    assert ((d["G"] @ xc.unsqueeze(-1)).squeeze(-1) - d["h"]).max().item() < 1e-6
    assert dyn_c(xc).abs().max().item() < 1e-6
    with pytest.raises(_lib.AlmpcError):
        DenseQPFunction()(prob["Q"], prob["p"], prob["G"], prob["h"], prob["A"], prob["b"], dyn_o)
