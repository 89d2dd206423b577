"""CPU: pin the oracle (oracle/almpc_oracle.py) against golden outputs of the UNMODIFIED reference
(tests/golden/*.npz, produced by oracle/gen_golden.py in the build container)."""
import pytest
import torch

from golden_util import Golden, list_golden, list_tracking_golden, rel_err
from oracle import almpc_oracle as orc
from oracle.problem_gen import make_oracle_model


def replay(gold, resync=True):
    """Run the oracle through the same call chain as the golden scenario, re-synchronising nothing:
    the oracle carries its own (lam, rho, warm start) state exactly like the reference's MPC does."""
    model = make_oracle_model(gold.model)
    obs_kw = {}
    if "obstacle_positions" in gold.z.files:  # Obstacle_MPC scenario
        obs_kw = dict(obstacle_positions=gold.t("obstacle_positions"), obstacle_radius=float(gold.z["obstacle_radius"]))
    mpc = orc.OracleMPC(model, gold.T, gold.t("u_lower"), gold.t("u_upper"), gold.B, dtype=gold.dtype,
                        state_estimator=gold.state_estimator, **obs_kw)
    x0 = gold.t("x0")
    Qd1 = gold.t("Qd").to(gold.dtype)
    out = []
    for ci in range(gold.n_calls):
        c = gold.call(ci)
        if ci == 0:
            mpc.reinitialize(c["x_ref"])
        if gold.stream_after == ci:
            xi, ui = mpc.x_init, mpc.u_init
            ui[:, -1:] = c["u_ref"][:, -1:]
            xi[:, -1:] = c["x_ref"][:, -1:]
            mpc.warm_start_initialize(xi, ui, gold.rho_init_max)
        if mpc.x_init is None:
            mpc.x_init, mpc.u_init = c["x_ref"].clone(), c["u_ref"].clone()
        if obs_kw:
            # nearest-obstacle selection (AL_mpc_custom.py:110-113,129-135) must pick the reference's obstacles
            assert torch.equal(mpc.pb.obstacles[0], c["obs_pos"]), f"call {ci}: obstacle selection"
        if resync:
            # call-level parity: start every call from the reference's own (warm start, lam, rho)
            nx = model.nx
            assert bool(c["warm"]) == mpc.warm_starting
            mpc.x_init, mpc.u_init = c["xu_init"][..., :nx].clone(), c["xu_init"][..., nx:].clone()
            mpc.lamda_prev, mpc.rho_prev = c["lam_in"].clone(), c["rho_in"].clone()
        mpc.al_iter = int(c["al_iter"])
        Qd = Qd1.expand(gold.B, gold.T, -1).contiguous()
        tr = orc.SolveTrace()
        x, u, status, xu = mpc.solve(x0, Qd, c["q"], trace=tr)
        Qg, qg = mpc.backward(c["gx"], c["gu"])
        out.append(dict(x=x, u=u, status=status, xu=xu, trace=tr, Q_grad=Qg, q_grad=qg,
                        lam=mpc.lamda_prev, rho=mpc.rho_prev))
    return out


def tolerances(c, dtype):
    """cond(H) ~ rho/R, so the attainable agreement between two correct implementations degrades
    with the largest penalty used in the call (measured: oracle-vs-reference table in DESIGN.md)."""
    if dtype != torch.float64:
        return dict(xu=1e-3, lam=5e-3, grad=5e-3)
    rho_used = float(c["rho_out"].double().max()) / 10.0
    return dict(xu=1e-8, lam=max(1e-7, 1e-11 * rho_used), grad=max(1e-7, 3e-12 * rho_used))


def check_line_search(trace, c, dtype, ci):
    """argmin / accept must agree except at numerical ties (two candidates whose reference merits
    differ by less than the rounding noise of the merit itself)."""
    tie = 1e-11 if dtype == torch.float64 else 1e-4
    for k, s in enumerate(trace.steps):
        gm = c["step_merits"][k].double()
        bidx = torch.arange(gm.shape[1])
        mine, ref = s.ls_index, c["step_idx"][k]
        diff = (gm[mine, bidx] - gm[ref, bidx]).abs() / gm[ref, bidx].abs().clamp_min(1e-300)
        assert (diff[mine != ref] < tie).all(), f"call {ci} step {k}: argmin differs beyond a tie"
        gap = (gm[ref, bidx] - c["step_merit_in"][k].double()).abs() / c["step_merit_in"][k].double().abs().clamp_min(1e-300)
        bad = s.accept != c["step_accept"][k]
        assert (gap[bad] < tie).all(), f"call {ci} step {k}: accept differs beyond a tie"


@pytest.mark.parametrize("name", list_golden())
def test_oracle_matches_reference(name):
    gold = Golden(name)
    res = replay(gold)
    for ci, r in enumerate(res):
        c = gold.call(ci)
        assert r["trace"].newton_steps == c["nsteps"].tolist(), f"call {ci}: Newton step counts differ"
        assert bool(r["status"]) == bool(c["status"])
        check_line_search(r["trace"], c, gold.dtype, ci)
        tol = tolerances(c, gold.dtype)
        assert rel_err(r["xu"], c["xu64"]) < tol["xu"], f"call {ci}: xu"
        assert rel_err(r["lam"], c["lam_out"]) < tol["lam"]
        assert torch.equal(r["rho"].double(), c["rho_out"].double())
        assert rel_err(r["q_grad"], c["q_grad"]) < tol["grad"], f"call {ci}: q_grad"
        assert rel_err(r["Q_grad"], c["Q_grad"]) < tol["grad"]
        assert rel_err(r["x"], c["x"]) < max(tol["xu"], 1e-6)


@pytest.mark.parametrize("name", list_tracking_golden())
def test_oracle_matches_reference_tracking_mpc(name):
    """The reference's caller `Tracking_MPC.forward` (deqmpc/policies.py:1236-1297) restated on the oracle: cost
    build Q (1 + q_scaling), p = -Q xu_ref (:1245-1248,1295) and its chain rule to the fp32 network outputs, with
    the solver state carried over the whole chain (cold calls, warm_start_initialize :1305-1310, streaming calls)."""
    gold = Golden(name)
    model = make_oracle_model(gold.model)
    nx = model.nx
    mpc = orc.OracleMPC(model, gold.T, gold.t("u_lower"), gold.t("u_upper"), gold.B, dtype=gold.dtype)
    Q0 = torch.cat([gold.t("Q"), gold.t("R")]).to(gold.dtype)
    x0 = gold.t("x0")
    mpc.reinitialize(gold.call(0)["x_ref"])
    for ci in range(gold.n_calls):
        c = gold.call(ci)
        if gold.stream_after == ci:
            xi, ui = mpc.x_init, mpc.u_init
            ui[:, -1:] = c["u_ref"][:, -1:]
            xi[:, -1:] = c["x_ref"][:, -1:]
            mpc.warm_start_initialize(xi, ui, gold.rho_init_max)
        if mpc.x_init is None:
            mpc.x_init, mpc.u_init = c["x_ref"].clone(), c["u_ref"].clone()
        xu_ref = torch.cat([c["x_ref"], c["u_ref"]], -1).to(gold.dtype)
        scale = 1.0 + c["q_scaling"].to(gold.dtype) if "q_scaling" in c else torch.ones(gold.B, gold.T, dtype=gold.dtype)
        Qd = Q0 * scale[..., None]
        mpc.al_iter = int(c["al_iters"])
        x, u, status, xu = mpc.solve(x0, Qd, -Qd * xu_ref)
        Qg, qg = mpc.backward(c["gx"], c["gu"])
        # chain rule of p = -Q0 (1+s) xu_ref, Qd = Q0 (1+s)
        g_ref = -(Qd * qg)
        g_s = ((Qg - xu_ref * qg) * Q0).sum(-1)
        rho_used = float(c["rho_out"].double().max()) / 10.0
        tol_x, tol_g = max(1e-6, 1e-14 * rho_used), max(2e-6, 3e-12 * rho_used)
        assert bool(status) == bool(c["status"]), f"call {ci}"
        assert torch.equal(mpc.rho_prev.double(), c["rho_out"].double())
        assert rel_err(x, c["x"]) < tol_x and rel_err(u, c["u"]) < 10 * tol_x, f"call {ci}"
        assert rel_err(g_ref[..., :nx].float(), c["g_x_ref"]) < tol_g, f"call {ci}: x_ref grad"
        assert rel_err(g_ref[..., nx:].float(), c["g_u_ref"]) < tol_g, f"call {ci}: u_ref grad"
        if "q_scaling" in c:
            assert rel_err(g_s.float(), c["g_q_scaling"]) < tol_g, f"call {ci}: q_scaling grad"
