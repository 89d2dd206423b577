"""CPU oracle: a dense torch restatement of the reference's batched augmented-Lagrangian MPC solve.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of
``bench.py`` may import this file.  The product (``deq_mpc_corl_b200``) never routes through it.

What it restates (all citations relative to /root/reference):

* residuals               qpth/al_utils.py:209-226 (eq), :288-326 (ineq), :407-421
* constraint Jacobian     qpth/al_utils.py:233-284 (eq), :390-404 + :329-347 (ineq, active mask ``res >= 0``)
* merit                   qpth/al_utils.py:52-77, cost :425-435 (constant ``f`` is NOT in the merit)
* gradient / GN Hessian   qpth/al_utils.py:80-123, :196-206
* Newton loop             qpth/al_utils.py:451-576 (<=4 steps, dense Cholesky, batch-global stop rule)
* 20-point line search    qpth/al_utils.py:618-642 (first-argmin, strict-decrease accept)
* implicit backward       qpth/al_utils.py:578-615
* AL outer loop           qpth/AL_mpc.py:260-339 (al_solve), :342-423 (al_solve_stream)
* solver state            qpth/AL_mpc.py:569-592 (reinitialize / warm_start_initialize)

Like the reference it is dense (J[B,m,N], H[B,N,N]) and uses torch's CPU LAPACK
(``linalg.cholesky_ex`` / ``cholesky_solve``).  Pinning: ``tests/golden/*.npz`` hold outputs of the
UNMODIFIED reference (``oracle/gen_golden.py`` imports it from /root/reference in the build
container); ``tests/test_oracle_golden.py`` checks this file against them.  The quadrotor /
flying-cartpole rotation helpers are the stand-ins of ``oracle/rexquad_standin.py`` (the reference
does not ship ``rexquad_utils``), so those two models are "parity unpinned" at the helper level.
"""
from dataclasses import dataclass, field
from typing import List, Optional

import torch


@dataclass
class StepTrace:
    """Per-Newton-step record (golden comparison / debugging)."""

    grad: torch.Tensor
    direction: torch.Tensor
    merits: torch.Tensor  # [n_ls, B]
    ls_index: torch.Tensor  # [B]
    accept: torch.Tensor  # [B] bool
    merit_in: torch.Tensor  # [B] merit the candidates were compared with
    res_new: float
    H: Optional[torch.Tensor] = None


@dataclass
class SolveTrace:
    newton_steps: List[int] = field(default_factory=list)  # per AL iteration
    steps: List[StepTrace] = field(default_factory=list)
    xu_al: List[torch.Tensor] = field(default_factory=list)  # iterate after each AL iteration


class Problem:
    """Static description of one batched problem family (dims, dynamics, bounds)."""

    def __init__(self, model, T, u_lower, u_upper, dtype=torch.float64, obstacles=None, state_estimator=False):
        """obstacles = (positions [B,T,K,3], radius tensor) for Obstacle_MPC (qpth/AL_mpc_custom.py), else None.
        state_estimator = the ``MPC(state_estimator=True)`` variant (qpth/al_utils_se.py, AL_mpc.py:186-199,
        426-442, 508-509): only the T-1 dynamics rows, controls are data (their Jacobian columns are zeroed,
        al_utils_se.py:153), the cost and its gradient see the states only (al_utils_se.py:272-305)."""
        self.model = model
        self.T = T
        self.nx, self.nu = model.nx, model.nu
        self.n = self.nx + self.nu
        self.N = T * self.n
        self.se = state_estimator
        self.neq = (T - 1) * self.nx if state_estimator else T * self.nx
        self.obstacles = obstacles
        self.n_obs = 0 if obstacles is None else obstacles[0].shape[2]
        self.nineq = 0 if state_estimator else T * (2 * self.nu + self.n_obs)
        self.m = self.neq + self.nineq
        self.dtype = dtype
        self.u_lower = torch.as_tensor(u_lower).to(dtype)
        self.u_upper = torch.as_tensor(u_upper).to(dtype)


# --------------------------------------------------------------------------------------------------
# residuals / merit
# --------------------------------------------------------------------------------------------------
def residuals(pb: Problem, xu, x0):
    """c(z) [B,m] and its clamped version (ineq part max(.,0)).  al_utils.py:407-421."""
    B = xu.shape[0]
    nx = pb.nx
    x, u = xu[..., :nx], xu[..., nx:]
    xn = pb.model.step(x[:, :-1].reshape(-1, nx), u[:, :-1].reshape(-1, pb.nu)).view(B, pb.T - 1, nx)
    c_dyn = (x[:, 1:] - xn).reshape(B, -1)
    if pb.se:  # al_utils_se.py:247-258
        return c_dyn, c_dyn
    c_init = x[:, 0] - x0
    c_in = _ineq_rows(pb, x, u).reshape(B, -1)
    c = torch.cat([c_dyn, c_init, c_in], dim=1)
    c_plus = torch.cat([c_dyn, c_init, c_in.clamp(min=0)], dim=1)
    return c, c_plus


def _ineq_rows(pb, x, u):
    """[B,T,2nu(+K)]: control bounds (al_utils.py:293) and obstacle rows r^2 - |p - c_k|^2 (al_utils.py:313-323)."""
    rows = torch.cat([u - pb.u_upper, -u + pb.u_lower], dim=2)
    if pb.obstacles is not None:
        pos, radius = pb.obstacles
        B = x.shape[0]
        if pos.shape[0] != B:  # candidate-stacked batch of the line search
            pos = pos.repeat(B // pos.shape[0], 1, 1, 1)
        d2 = ((x[:, :, None, :3] - pos) ** 2).sum(-1)
        rows = torch.cat([rows, radius ** 2 - d2], dim=2)
    return rows


def cost_value(xu, Qd, q, nx=None):
    """al_utils.py:425-435 with diag cost and f=0 (nx given: states only, al_utils_se.py:272-289)."""
    if nx is not None:
        xu, Qd, q = xu[..., :nx], Qd[..., :nx], q[..., :nx]
    return (0.5 * (xu * Qd * xu).sum(-1) + (q * xu).sum(-1)).sum(-1)


def merit(pb: Problem, xu, x0, Qd, q, lam, rho):
    """M(z) per problem; accepts a leading candidate dim [K,B,T,n] (al_utils.py:52-77)."""
    if xu.dim() == 4:
        K, B = xu.shape[:2]
        rep = lambda a: a[None].expand(K, *a.shape).reshape(K * B, *a.shape[1:])
        return merit(pb, xu.reshape(K * B, pb.T, pb.n), rep(x0), rep(Qd), rep(q), rep(lam), rep(rho)).view(K, B)
    c, c_plus = residuals(pb, xu, x0)
    return cost_value(xu, Qd, q, pb.nx if pb.se else None) + 0.5 * rho[:, 0] * (c_plus * c_plus).sum(1) + (lam * c).sum(1)


# --------------------------------------------------------------------------------------------------
# gradient / Gauss-Newton Hessian (dense, like the reference)
# --------------------------------------------------------------------------------------------------
def constraint_jacobian(pb: Problem, xu, x0):
    """Dense J[B,m,N], active-masked Jc, residuals.  al_utils.py:183-206, 233-284, 329-347, 390-404."""
    B, T, n, nx, nu = xu.shape[0], pb.T, pb.n, pb.nx, pb.nu
    x, u = xu[..., :nx], xu[..., nx:]
    xn, (Jx, Ju) = pb.model.step_jac(x[:, :-1].reshape(-1, nx), u[:, :-1].reshape(-1, nu))
    G = torch.cat([Jx.reshape(B, T - 1, nx, nx), Ju.reshape(B, T - 1, nx, nu) * (0.0 if pb.se else 1.0)], dim=-1)
    J = torch.zeros(B, pb.m, pb.N, dtype=xu.dtype)
    eye = torch.eye(nx, dtype=xu.dtype)
    for t in range(T - 1):
        J[:, t * nx : (t + 1) * nx, t * n : (t + 1) * n] = -G[:, t]
        J[:, t * nx : (t + 1) * nx, (t + 1) * n : (t + 1) * n + nx] = eye
    if pb.se:  # al_utils_se.py:97-109,143-187: dynamics rows only
        c_dyn = (x[:, 1:] - xn.view(B, T - 1, nx)).reshape(B, -1)
        return c_dyn, c_dyn, J, J
    J[:, (T - 1) * nx : T * nx, :nx] = eye
    c_in = _ineq_rows(pb, x, u).reshape(B, -1)
    SI = 2 * nu + pb.n_obs
    for t in range(T):
        r0 = pb.neq + t * SI
        for j in range(nu):
            J[:, r0 + j, t * n + nx + j] = 1.0
            J[:, r0 + nu + j, t * n + nx + j] = -1.0
        if pb.n_obs:
            # the reference builds these rows in a default-dtype (fp32) buffer: al_utils.py:376-379
            jac = (-2 * (x[:, t, None, :3] - pb.obstacles[0][:, t])).float().to(xu.dtype)
            J[:, r0 + 2 * nu : r0 + SI, t * n : t * n + 3] = jac
    active = (c_in >= 0).to(xu.dtype)
    Jc = J.clone()
    Jc[:, pb.neq :] = Jc[:, pb.neq :] * active[..., None]
    c_dyn = (x[:, 1:] - xn.view(B, T - 1, nx)).reshape(B, -1)
    c_init = x[:, 0] - x0
    c = torch.cat([c_dyn, c_init, c_in], 1)
    c_plus = torch.cat([c_dyn, c_init, c_in.clamp(min=0)], 1)
    return c, c_plus, J, Jc


def grad_hess(pb: Problem, xu, x0, Qd, q, lam, rho):
    """al_utils.py:80-123: grad = Qd*z + q + J^T lam + rho Jc^T c+ ; H = diag(Qd) + rho Jc^T Jc."""
    B = xu.shape[0]
    c, c_plus, J, Jc = constraint_jacobian(pb, xu, x0)
    gc = Qd * xu + q
    if pb.se:  # al_utils_se.py:292-305: the control entries of the cost gradient are zeroed
        gc = torch.cat([gc[..., : pb.nx], torch.zeros_like(gc[..., pb.nx :])], -1)
    g = (
        gc.reshape(B, -1)
        + (lam[..., None] * J).sum(-2)
        + rho * (c_plus[..., None] * Jc).sum(-2)
    )
    H = torch.diag_embed(Qd.reshape(B, -1)) + rho[:, :, None] * torch.bmm(Jc.transpose(1, 2), Jc)
    return g, H


# --------------------------------------------------------------------------------------------------
# Newton loop with line search
# --------------------------------------------------------------------------------------------------
N_LS = 20
MAX_NEWTON = 4


def line_search(pb, d, xu, x0, Qd, q, lam, rho, merit_in):
    """al_utils.py:618-642."""
    alphas = 2.0 ** (-torch.arange(N_LS).float())
    cand = xu[None] + alphas[:, None, None, None] * d[None]
    vals = merit(pb, cand, x0, Qd, q, lam, rho)
    best = torch.min(vals, dim=0)
    bidx = torch.arange(xu.shape[0])
    chosen = cand[best.indices, bidx]
    accept = best.values < merit_in
    acc = accept.to(xu.dtype)[:, None, None]
    return acc * chosen + (1 - acc) * xu, best.values, best.indices, accept, vals


def newton_al(pb: Problem, xu, x0, lam, rho, Qd, q, trace: Optional[SolveTrace] = None, keep_H=False,
              force_steps: Optional[int] = None):
    """NewtonAL.forward, al_utils.py:451-576.  Returns (xu, status[B], saved) where saved holds the
    factor of the LAST iteration's pre-step Hessian for the implicit backward."""
    B = xu.shape[0]
    m_cur = merit(pb, xu, x0, Qd, q, lam, rho)
    res_old = torch.norm(residuals(pb, xu, x0)[1]).item()
    status = torch.ones(B, dtype=torch.bool)
    chol_fail = False
    nstep = 0
    H = L = None
    while nstep < MAX_NEWTON:
        nstep += 1
        g, H = grad_hess(pb, xu, x0, Qd, q, lam, rho)
        if not chol_fail:
            L, _info = torch.linalg.cholesky_ex(H)
            d = -torch.cholesky_solve(g.reshape(B, -1, 1), L).reshape(B, pb.T, pb.n)
            if d.isnan().any() or d.isinf().any():
                d = -torch.linalg.solve(H, g).reshape(B, pb.T, pb.n)
                chol_fail = True
        else:
            L = H
            d = -torch.linalg.solve(H, g).reshape(B, pb.T, pb.n)
        xu1, m_new, idx, accept, vals = line_search(pb, d, xu, x0, Qd, q, lam, rho, m_cur)
        stepsz = (2.0 ** (-idx.float())).mean().item()
        flat = xu1.reshape(B, -1)
        if flat.isnan().any() or flat.isinf().any():
            status = status & ~(flat.isnan().any(-1) | flat.isinf().any(-1))
        res_new = torch.norm(residuals(pb, xu1, x0)[1]).item()
        if trace is not None:
            trace.steps.append(
                StepTrace(g.clone(), d.clone(), vals.clone(), idx.clone(), accept.clone(), m_cur.clone(), res_new,
                          H.clone() if keep_H else None)
            )
        xu = xu1
        if force_steps is not None:
            # test hook: replay the step count another (larger) batch realised — the reference's stop
            # rule is a whole-batch statistic (al_utils.py:486,552), so a sub-batch cannot reproduce it
            if nstep >= force_steps:
                break
        elif abs(res_old - res_new) / res_new < 1e-3 or res_new < 1e-3:
            break
        res_old = res_new
        m_cur = m_new
        if not stepsz > 1e-8:
            break
    if trace is not None:
        trace.newton_steps.append(nstep)
    return xu, status, (H, L, chol_fail)


def newton_backward(saved, xu, grad_xu):
    """NewtonAL.backward, al_utils.py:578-615 -> (Q_grad, q_grad)."""
    H, L, chol_fail = saved
    B = grad_xu.shape[0]
    if chol_fail:
        gbar = -torch.linalg.solve(H, grad_xu.reshape(B, -1)).reshape(grad_xu.shape)
    else:
        gbar = -torch.cholesky_solve(grad_xu.reshape(B, -1, 1), L).reshape(grad_xu.shape)
    return gbar * xu, gbar


# --------------------------------------------------------------------------------------------------
# AL outer loop + solver state
# --------------------------------------------------------------------------------------------------
class OracleMPC:
    """State-carrying mirror of qpth.AL_mpc.MPC (same attribute names for the state the caller
    touches: al_iter, x_init, u_init, lamda_prev, rho_prev, warm_starting)."""

    rho_max = 1e8

    def __init__(self, model, T, u_lower, u_upper, n_batch, dtype=torch.float64, al_iter=2, obstacle_positions=None,
                 obstacle_radius=None, state_estimator=False):
        self.pb = Problem(model, T, u_lower, u_upper, dtype, state_estimator=state_estimator)
        self.obstacle_positions = obstacle_positions  # [n_obstacles,3]: Obstacle_MPC (AL_mpc_custom.py:22-135)
        self.obstacle_radius = None if obstacle_radius is None else torch.tensor(obstacle_radius).to(dtype)
        if obstacle_positions is not None:
            self.pb.n_obs = 4
            self.pb.nineq = T * (2 * self.pb.nu + 4)
            self.pb.m = self.pb.neq + self.pb.nineq
        self.dtype = dtype
        self.n_batch = n_batch
        self.al_iter = al_iter
        self.x_init = None
        self.u_init = None
        self.warm_starting = False
        self.lamda_prev = torch.zeros(n_batch, self.pb.m, dtype=dtype)
        self.rho_prev = 1.0
        self.saved = None
        self.trace = None
        self.force_steps = None  # optional list: Newton steps per AL iteration (see newton_al)

    def reinitialize(self, x, mask=None):  # AL_mpc.py:569-579
        self.x_init = self.u_init = None
        self.n_batch = x.size(0)
        self.rho_prev = torch.ones((self.n_batch, 1), dtype=x.dtype)
        self.lamda_prev = torch.zeros(self.n_batch, self.pb.m, dtype=x.dtype)
        self.warm_starting = False
        if self.obstacle_positions is not None:  # AL_mpc_custom.py:110-113
            self.pb.obstacles = (self._nearest(x[:, :, :3]), self.obstacle_radius)

    def _nearest(self, xyz):
        d = (self.obstacle_positions[None, None] - xyz[:, :, None, :]).norm(dim=-1)
        return self.obstacle_positions[torch.argsort(d, dim=-1)[..., :4]]

    def warm_start_initialize(self, x, u, rho_init_max):  # AL_mpc.py:581-592
        self.x_init, self.u_init = x, u
        self.lamda_prev = self.lamda_prev * 0
        self.rho_prev = torch.clamp(self.rho_prev, max=rho_init_max)
        self.warm_starting = True
        if self.obstacle_positions is not None:  # AL_mpc_custom.py:129-135
            pos = torch.cat([self.pb.obstacles[0][:, 1:], self._nearest(x[:, -1:, :3])], dim=1)
            self.pb.obstacles = (pos, self.obstacle_radius)

    def solve(self, x0, Qd, q, trace: Optional[SolveTrace] = None, keep_H=False):
        """MPC.forward + al_solve / al_solve_stream (AL_mpc.py:207-423).  Qd, q are [B,T,n].
        Returns (x fp32, u fp32, status_flag, xu in solver dtype)."""
        pb = self.pb
        x = self.x_init.type_as(x0).to(self.dtype)
        u = self.u_init.type_as(x0).to(self.dtype)
        x0 = x0.to(self.dtype)
        lam = self.lamda_prev.to(self.dtype)
        rho = self.rho_prev
        xu = torch.cat([x, u], 2)
        exceeded = False
        for a in range(self.al_iter):
            xu = xu.detach().clone()
            force = None if self.force_steps is None else self.force_steps[a]
            xu, status, saved = newton_al(pb, xu, x0, lam, rho, Qd, q, trace, keep_H, force)
            c, c_plus = residuals(pb, xu, x0)
            lam = lam + rho * c
            lam = torch.cat([lam[:, : pb.neq], lam[:, pb.neq :].clamp(min=0)], 1)
            rho = rho * 10
            if trace is not None:
                trace.xu_al.append(xu.clone())
            if self.warm_starting and rho.max().item() > self.rho_max:
                exceeded = True
                break
        self.saved = saved
        self.xu = xu
        self.lamda_prev, self.rho_prev = lam, rho
        self.dyn_res_prev = c_plus.norm(dim=-1)
        xf, uf = xu[..., : pb.nx].float(), xu[..., pb.nx :].float()
        self.x_init, self.u_init = xf.clone(), uf.clone()
        return xf, uf, (exceeded if self.warm_starting else False), xu

    def backward(self, grad_x, grad_u):
        """Chain through the .float() cast (AL_mpc.py:337-338) then NewtonAL.backward."""
        g = torch.cat([grad_x, grad_u], 2).to(self.dtype)
        return newton_backward(self.saved, self.xu, g)
