"""CPU oracle of the interior-point path: a dense torch restatement of the reference's batched primal-dual
interior-point QP solver.  TEST INFRASTRUCTURE ONLY (tests/, never the product).

What it restates (paths relative to /root/reference):
* KKT matrix layout and its regularised copy      qpth/qp.py:196-220 (preprocess), qpth/solvers/pdipm/batch_LU.py:42-60,113-116
* solve_kkt: pivoted LU of Ktilde + 1 refinement  qpth/solvers/pdipm/batch_LU.py:212-244
* the predictor-corrector loop, best iterate, exits   qpth/solvers/pdipm/batch_LU.py:29-198, get_step :204-210
* gradients of the QP solution                    qpth/qp.py:239-268 (DenseQPFunction.backward)

Like the reference it factors the dense (nz + 2 nineq + neq)^2 matrix with torch's CPU LAPACK.  Pinned by
tests/golden/ip_*.npz (outputs of the UNMODIFIED reference, oracle/gen_golden.py::run_ip_scenario).
"""
import torch

KKT_EPS = 1e-7


def kkt_matrix(Q, G, A, s, z, eps):
    """K (eps = 0) or Ktilde (eps > 0): rows/cols ordered (x, s, z, y)."""
    B, nineq, nz = G.shape
    neq = A.shape[1]
    n = nz + 2 * nineq + neq
    K = torch.zeros(B, n, n, dtype=Q.dtype)
    ix, is_, iz, iy = slice(0, nz), slice(nz, nz + nineq), slice(nz + nineq, nz + 2 * nineq), slice(nz + 2 * nineq, n)
    K[:, ix, ix] = Q
    K[:, ix, iz] = G.transpose(1, 2)
    K[:, ix, iy] = A.transpose(1, 2)
    K[:, is_, is_] = torch.diag_embed(z)
    K[:, is_, iz] = torch.diag_embed(s)
    K[:, iz, ix] = G
    K[:, iz, is_] = torch.eye(nineq, dtype=Q.dtype)
    K[:, iy, ix] = A
    if eps:
        sign = torch.ones(n, dtype=Q.dtype)
        sign[nz + nineq:] = -1
        K = K + eps * torch.diag_embed(sign.expand(B, n))
    return K


def solve_kkt(Q, G, A, s, z, eps, rx, rs, rz, ry):
    K, Kt = kkt_matrix(Q, G, A, s, z, 0.0), kkt_matrix(Q, G, A, s, z, eps)
    nz, nineq = rx.shape[1], rz.shape[1]
    r = -torch.cat((rx, rs, rz, ry), 1)
    lu = torch.linalg.lu_factor(Kt)
    sol = torch.linalg.lu_solve(*lu, r.unsqueeze(-1)).squeeze(-1)
    res = r - torch.bmm(K, sol.unsqueeze(-1)).squeeze(-1)
    sol = sol + torch.linalg.lu_solve(*lu, res.unsqueeze(-1)).squeeze(-1)
    return sol[:, :nz], sol[:, nz:nz + nineq], sol[:, nz + nineq:nz + 2 * nineq], sol[:, nz + 2 * nineq:]


def get_step(v, dv):
    a = -v / dv
    a[dv == 0] = 1.0
    a[dv > 0] = max(1.0, a.max().item())
    return a.min(1)[0]


def pdipm_forward(Q, p, G, h, A, b, dyn_res, cost_grad=None, eps=1e-12, notImprovedLim=3, maxIter=20, trace=None):
    B, nineq, nz = G.shape
    neq = A.shape[1]
    zeros = lambda n: torch.zeros(B, n, dtype=Q.dtype)
    one_v = torch.ones(B, nineq, dtype=Q.dtype)
    bmv = lambda M, v: torch.bmm(M, v.unsqueeze(-1)).squeeze(-1)
    x, s, z, y = (t.clone() for t in solve_kkt(Q, G, A, one_v, one_v, KKT_EPS, p, zeros(nineq), -h, -b))
    for v in (s, z):
        m = v.min(1)[0]
        shift = torch.where(m < 0, 1 - m, torch.zeros_like(m))
        v += shift[:, None]
    best, n_bad = None, 0
    for _ in range(maxIter):
        rx = bmv(A.transpose(1, 2), y) + bmv(G.transpose(1, 2), z) + (bmv(Q, x) + p if cost_grad is None else cost_grad(x))
        rs, rz, ry = s * z, bmv(G, x) + s - h, dyn_res(x)
        mu = ((s * z).sum(1) / nineq).abs()
        resids = ry.norm(dim=1) + rz.norm(dim=1) + rx.norm(dim=1) + nineq * mu
        if best is None:
            best = dict(resids=resids.clone(), x=x.clone(), y=y.clone(), z=z.clone(), s=s.clone())
        else:
            imp = resids < best["resids"]
            n_bad = 0 if imp.any() else n_bad + 1
            best["resids"] = torch.where(imp, resids, best["resids"])
            for k, cur in (("x", x), ("y", y), ("z", z), ("s", s)):
                best[k] = torch.where(imp[:, None], cur, best[k])
        if trace is not None:
            trace.append(dict(resids=resids.clone(), mu=mu.clone()))
        if n_bad == notImprovedLim or best["resids"].max().item() < eps or mu.min().item() > 1e32:
            break
        if trace is not None:
            trace[-1]["solved"] = True  # this iterate is followed by the two KKT solves of a full iteration
        aff = solve_kkt(Q, G, A, s, z, KKT_EPS, rx, rs, rz, ry)
        alpha = torch.minimum(torch.minimum(get_step(z, aff[2]), get_step(s, aff[1])), torch.ones(B, dtype=Q.dtype))[:, None]
        sig = (((s + alpha * aff[1]) * (z + alpha * aff[2])).sum(1) / (s * z).sum(1)) ** 3
        cor = solve_kkt(Q, G, A, s, z, KKT_EPS, zeros(nz), (-mu * sig)[:, None] + aff[1] * aff[2], zeros(nineq), zeros(neq))
        dx, ds, dz, dy = (a + c for a, c in zip(aff, cor))
        alpha = torch.minimum(0.999 * torch.minimum(get_step(z, dz), get_step(s, ds)), torch.ones(B, dtype=Q.dtype))[:, None]
        x, s, z, y = x + alpha * dx, s + alpha * ds, z + alpha * dz, y + alpha * dy
    return best["x"], best["y"], best["z"], best["s"]


def qp_backward(Q, G, A, zhat, nus, lams, slacks, dl_dzhat):
    """-> (dQ, dp, dG, dh, dA, db)"""
    B, nineq, neq = Q.shape[0], G.shape[1], A.shape[1]
    z0 = lambda n: torch.zeros(B, n, dtype=Q.dtype)
    dx, _, dlam, dnu = solve_kkt(Q, G, A, slacks, lams, 0.0, dl_dzhat, z0(nineq), z0(nineq), z0(neq))
    outer = lambda a, c: a.unsqueeze(2) * c.unsqueeze(1)
    return (0.5 * (outer(dx, zhat) + outer(zhat, dx)), dx, outer(dlam, zhat) + outer(lams, dx), -dlam,
            outer(dnu, zhat) + outer(nus, dx), -dnu)
