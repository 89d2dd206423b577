#!/usr/bin/env python
"""Headline benchmark: AL-MPC solves/sec (forward + implicit backward), rex-quadrotor T=5, 64K problems
per GPU, fp64 (BASELINE.json).  One *step* = one cold-start `MPC.forward` (al_iter=2) plus its
`NewtonAL.backward` for every problem of the batch.

    python bench.py [--gpus N --steps K --warmup W] [--dtype f64|f32] [--model ...] [--T ..] [--batch ..]
    torchrun --nproc-per-node N bench.py --gpus N ...          (one rank per GPU, batch sharded, no collective
                                                                 on the data path -> "scaling": "weak")
    python bench.py --impl reference ...                        (the CPU oracle port on the host cores)

Prints ONE JSON line (contract in the task statement): value = device-timed whole-job solves/s with
inputs resident in HBM; e2e = the same through the reference-facing `MPC` module with pinned HOST
buffers and the H2D/D2H copies inside the timed region; roofline = FP64 (FP32) FMA-pipe fraction of
the dominant kernel of the Newton step (k_jacobian / k_factor_solve / k_line_search, timed per launch with
CUDA events inside the library) from algorithmic flops and bytes, against the FMA peak
measured in this run by almpc_fma_peak (MEASURED_PEAKS.json has no FP64 figure), plus the HBM view;
cpu_baseline = the oracle port timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from deq_mpc_corl_b200.problems import MODEL_SPECS, make_inputs, tracking_cost  # noqa: E402

# F_dyn: flops of one RK4 step of each model's ODE as written in csrc/almpc_models.cuh, counted by
# tools/count_flops.cpp (add/sub/mul = 1, fma = 2, reciprocal = 1, sincos = 2).  Frozen here.
F_DYN = {"pendulum": 50, "cartpole1l": 120, "cartpole2l": 322, "rex_quadrotor": 980, "flying_cartpole": 1150}


def algorithmic_flops(model, T):
    """SURVEY.md §8(d) formulas.  F_jac = forward mode, n tangents at 2x primal each."""
    spec = MODEL_SPECS[model]
    nx, nu = spec["nx"], spec["nu"]
    n, m = nx + nu, T * nx + 2 * T * nu
    f_dyn = F_DYN[model]
    f_jac = 2 * n * f_dyn
    subst = 2 * (T * n * n + 2 * (T - 1) * nx * n)
    f_newton = ((T - 1) * f_jac + (T - 1) * nx * n * (n + 1) + 4 * (T - 1) * nx * n + 4 * T * nu
                + T * n ** 3 / 3 + (T - 1) * (nx * n * n + nx * nx * n) + subst
                + 20 * ((T - 1) * f_dyn + 6 * T * n + 2 * m) + (T - 1) * f_dyn + 2 * m)
    f_dual = (T - 1) * f_dyn + 3 * m
    f_bwd = subst + T * n
    return dict(newton=f_newton, dual=f_dual, bwd=f_bwd)


def kernel_work(model, T, esize):
    """Algorithmic flops / HBM bytes per problem of the three kernels of one Newton step (same SURVEY §8d
    terms as algorithmic_flops, split by kernel; G and c cross HBM between the kernels)."""
    spec = MODEL_SPECS[model]
    nx, nu = spec["nx"], spec["nu"]
    n, N, m = nx + nu, T * (nx + nu), T * nx + 2 * T * nu
    f_dyn = F_DYN[model]
    G = (T - 1) * nx * n
    factor = T * n * n + (T - 1) * nx * n
    subst = 2 * (T * n * n + 2 * (T - 1) * nx * n)
    return {
        "k_jacobian": dict(flops=(T - 1) * 2 * n * f_dyn, bytes=(N + nx + G + T * nx) * esize),
        "k_factor_solve": dict(
            flops=(T - 1) * nx * n * (n + 1) + 4 * (T - 1) * nx * n + 4 * T * nu + T * n ** 3 / 3
            + (T - 1) * (nx * n * n + nx * nx * n) + subst + 6 * N + 3 * m,
            bytes=(G + T * nx + 3 * N + m + 1 + factor + N + 2) * esize),
        "k_line_search": dict(flops=20 * ((T - 1) * f_dyn + 6 * T * n + 2 * m) + 2 * N,
                              bytes=(5 * N + m + nx + 3 + N) * esize),
    }


def algorithmic_bytes(model, T, esize):
    """forward reads x0+Qd+q+xu+lam+rho, writes xu+lam+rho+factor; backward reads
    factor+xu+grad, writes 2N."""
    spec = MODEL_SPECS[model]
    nx, nu = spec["nx"], spec["nu"]
    n, N, m = nx + nu, T * (nx + nu), T * nx + 2 * T * nu
    factor = T * n * n + (T - 1) * nx * n
    fwd = (nx + 3 * N + m + 1) + (N + m + 1 + factor)
    bwd = factor + 2 * N + 2 * N
    return dict(fwd=fwd * esize, bwd=bwd * esize, newton_launch=(nx + 3 * N + m + 2 + N + factor + 2) * esize)


# DRAM traffic per launch measured once with `ncu --set full` (dram__bytes_read.sum + dram__bytes_write.sum),
# profiles/r01_v6_three_kernels_ncu_full_summary.csv — only valid for the exact workload it was captured on.
NCU_TRAFFIC = {("rex_quadrotor", 5, 65536, "f64"): {"k_jacobian": 48.37e6 + 375.15e6, "k_factor_solve": 629.38e6 + 1063.53e6,  # k_factor_solve_pair, profiles/r01_v11_*
                                                    "k_line_search": 228.87e6 + 31.84e6}}


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.QUERY}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_oracle_run(model, T, B, dtype, seed, repeats=1):
    """Times the oracle port (oracle/almpc_oracle.py) forward+backward on the host: returns
    (solves/s, newton step counts)."""
    from oracle import almpc_oracle as orc
    from oracle.problem_gen import make_oracle_model

    torch.set_num_threads(os.cpu_count())
    inp = make_inputs(model, T, B, seed)
    Qd, q, _ = tracking_cost(inp["Qd"], inp["x_ref"], inp["u_ref"], dtype)
    om = make_oracle_model(model)
    best, steps = None, None
    for _ in range(repeats):
        mpc = orc.OracleMPC(om, T, inp["u_lower"], inp["u_upper"], B, dtype=dtype, al_iter=2)
        mpc.reinitialize(inp["x_ref"])
        mpc.x_init, mpc.u_init = inp["x_ref"].clone(), inp["u_ref"].clone()
        tr = orc.SolveTrace()
        t0 = time.perf_counter()
        x, u, _, _ = mpc.solve(inp["x0"], Qd, q, trace=tr)
        mpc.backward(torch.ones_like(x), torch.ones_like(u))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        steps = tr.newton_steps
    return B / best, steps, best


def run_reference_arm(args, rank):
    """`--impl reference`: the reference's CPU path (oracle port) on the host cores, rank 0 only."""
    if rank != 0:
        return
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    B = args.cpu_sample
    for _ in range(args.warmup):
        cpu_oracle_run(args.model, args.T, min(B, 64), dtype, 1003)
    t0 = time.perf_counter()
    steps = None
    for _ in range(args.steps):
        _, steps, _ = cpu_oracle_run(args.model, args.T, B, dtype, 1003)
    el = time.perf_counter() - t0
    val = args.steps * B / el
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": "AL-MPC solves/sec (fwd+implicit bwd)", "value": val, "unit": "solves/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": workload_config(args, B, 1),
        "cpu_baseline": {"value": val, "unit": "solves/s", "cores": cores, "kind": "port",
                         "sample": f"{B} problems per step x {args.steps} steps, oracle port of qpth AL_mpc (torch {torch.__version__} CPU, {cores} threads); newton steps {steps}"},
        "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, batch_per_gpu, n_gpus):
    return {
        "workload": f"{args.model} T={args.T} al_iter=2 cold start (lam=0, rho=1), max 4 Newton steps/AL iter, "
                    f"20-pt line search; {batch_per_gpu} problems per GPU",
        "batch_per_gpu": batch_per_gpu, "global_batch": batch_per_gpu * n_gpus, "T": args.T,
        "parallelism": f"batch-sharded x{n_gpus} (independent problems, no data-path collective)",
        "l2": "inputs+factor per step exceed the 126 MB L2 (no explicit flush)" if batch_per_gpu >= 32768 else "small batch: L2-resident",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="rex_quadrotor", choices=list(MODEL_SPECS))
    ap.add_argument("--T", type=int, default=5)
    ap.add_argument("--batch", type=int, default=65536, help="problems per GPU")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--cpu-sample", type=int, default=None, help="problems in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.cpu_sample is None:
        args.cpu_sample = 6144 if args.impl == "b200" else 1024

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the solver has no CPU fallback; use --impl reference for the CPU arm)")
    import torch.distributed as dist
    from deq_mpc_corl_b200 import _lib
    from deq_mpc_corl_b200.al_mpc import MPC
    from deq_mpc_corl_b200.al_utils import QuadCost
    from deq_mpc_corl_b200.dynamics import make as make_dyn
    from deq_mpc_corl_b200.solver import SolveConfig, raw_backward, raw_forward

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    esize = 8 if dtype == torch.float64 else 4
    model, T, B = args.model, args.T, args.batch
    spec = MODEL_SPECS[model]
    nx, nu = spec["nx"], spec["nu"]
    n, m = nx + nu, T * nx + 2 * T * nu

    # ---- synthetic inputs (seed differs per rank: every GPU solves different problems)
    inp = make_inputs(model, T, B, 1003 + rank)
    Qd_h, q_h, _ = tracking_cost(inp["Qd"], inp["x_ref"], inp["u_ref"], dtype)
    dyn = make_dyn(model, dt=spec["dt"])
    cfg = SolveConfig(dyn=dyn, T=T, u_lower=inp["u_lower"].tolist(), u_upper=inp["u_upper"].tolist(), al_iter=2,
                      rho_f32=(dtype == torch.float64))
    x0 = inp["x0"].to(dtype).to(dev)
    Qd, q = Qd_h.to(dev), q_h.to(dev)
    xu_init = torch.cat([inp["x_ref"], inp["u_ref"]], -1).to(dtype).to(dev).contiguous()
    lam0 = torch.zeros(B, m, dtype=dtype, device=dev)
    rho0 = torch.ones(B, dtype=dtype, device=dev)
    g = torch.Generator().manual_seed(7 + rank)
    grad_xu = torch.randn(B, T, n, generator=g).to(dtype).to(dev)

    xu, lam, rho = torch.empty_like(xu_init), torch.empty_like(lam0), torch.empty_like(rho0)

    def device_step():
        xu.copy_(xu_init)
        lam.copy_(lam0)
        rho.copy_(rho0)
        out = raw_forward(cfg, x0, Qd, q, xu, lam, rho)
        raw_backward(cfg, out["factor"], xu, grad_xu, need_Q=True)
        return out

    # ---- FMA-pipe peak (roofline denominator), measured here because MEASURED_PEAKS.json has none
    lib = _lib.load()
    import ctypes

    def fma_peak():
        sink = torch.zeros(256 * 148 * 16, dtype=dtype, device=dev)
        fl = ctypes.c_double()
        fn = getattr(lib, "almpc_fma_peak_" + _lib.suffix(dtype))
        best = 0.0
        for i in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(fn(148 * 16, 8192, _lib.ptr(sink), ctypes.byref(fl), _lib.stream_ptr()), "fma_peak")
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                best = max(best, fl.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        return best

    peak_tflops = fma_peak()

    # ---- warm-up
    for _ in range(args.warmup):
        out = device_step()
    torch.cuda.synchronize()
    info = out["info"].cpu().tolist()
    newton = info[_lib.INFO_NEWTON0:_lib.INFO_NEWTON0 + info[_lib.INFO_AL_DONE]]
    n_newton = sum(newton)

    # ---- kernel-level timing for the roofline: CUDA events around every launch inside the library
    # (almpc_set_profiling), on the stream the kernels are launched on; median of 3 forward solves
    prof_runs = []
    lib.almpc_set_profiling(1)
    for _ in range(3):
        xu.copy_(xu_init); lam.copy_(lam0); rho.copy_(rho0)
        raw_forward(cfg, x0, Qd, q, xu, lam, rho)
        torch.cuda.synchronize()
        buf = (ctypes.c_double * 8)()
        lib.almpc_get_profile(buf)
        prof_runs.append(list(buf))
    lib.almpc_set_profiling(0)
    prof_runs.sort(key=lambda r: r[0] + r[1] + r[2])
    pr = prof_runs[len(prof_runs) // 2]
    kernel_ms = {"k_jacobian": pr[0] / max(pr[4], 1), "k_factor_solve": pr[1] / max(pr[5], 1),
                 "k_line_search": pr[2] / max(pr[6], 1)}
    newton_launch_ms = sum(kernel_ms.values())

    # ---- timed region: K steps, device resident
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        device_step()
    e1.record()
    barrier()
    el_ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(el_ms, op=dist.ReduceOp.MAX)
    el_ms = el_ms.item()

    # ---- e2e: reference-facing MPC module, pinned host buffers, H2D + D2H inside the timed region
    ctrl = MPC(nx, nu, T, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=B, dtype=dtype)
    pin = lambda t: t.contiguous().pin_memory()
    h = dict(x0=pin(inp["x0"]), Qd=pin(Qd_h), q=pin(q_h), x_ref=pin(inp["x_ref"]), u_ref=pin(inp["u_ref"]),
             gx=pin(torch.randn(B, T, nx, generator=g)), gu=pin(torch.randn(B, T, nu, generator=g)))
    out_h = dict(x=torch.empty(B, T, nx).pin_memory(), u=torch.empty(B, T, nu).pin_memory(),
                 qg=torch.empty(B, T, n, dtype=dtype).pin_memory())
    h2d = sum(t.numel() * t.element_size() for t in h.values())
    d2h = sum(t.numel() * t.element_size() for t in out_h.values())

    def pipelined_ms(solve, hin, hout, k_steps):
        """K end-to-end steps; every step copies ITS inputs from the pinned host buffers and its results back.
        Double-buffered device inputs: the H2D of step j+1 (copy stream) and the D2H of step j-1 (second copy
        stream) overlap the solve of step j; all three are inside the timed region.  Returns device ms."""
        cur = torch.cuda.current_stream()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        bufs = [{k: torch.empty_like(v, device=dev) for k, v in hin.items()} for _ in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def h2d(i):
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_free[i])
                for k, v in hin.items():
                    bufs[i][k].copy_(v, non_blocking=True)
                ev_in[i].record(s_in)

        for e in ev_free:
            e.record(cur)
        t0.record(cur)
        s_in.wait_event(t0)
        s_out.wait_event(t0)
        h2d(0)
        for j in range(k_steps):
            i = j & 1
            if j + 1 < k_steps:
                h2d(i ^ 1)
            cur.wait_event(ev_in[i])
            res = solve(bufs[i])
            ev_free[i].record(cur)
            done = torch.cuda.Event()
            done.record(cur)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                for k, t in res.items():
                    hout[k].copy_(t, non_blocking=True)
                    t.record_stream(s_out)
        fin = torch.cuda.Event()
        fin.record(s_out)
        cur.wait_event(fin)
        t1.record(cur)
        torch.cuda.synchronize()
        return t0.elapsed_time(t1)

    def e2e_solve(d):
        ctrl.reinitialize(d["x_ref"], None)
        ctrl.x_init, ctrl.u_init = d["x_ref"], d["u_ref"]
        ql = d["q"].detach().requires_grad_(True)
        x, u, _ = ctrl(d["x0"], QuadCost(d["Qd"], ql, None), dyn, dyn)
        qg, = torch.autograd.grad((x, u), ql, (d["gx"], d["gu"]))
        return dict(x=x.detach(), u=u.detach(), qg=qg)

    def e2e_step():  # serial variant: copy in, solve, copy out on one stream
        res = e2e_solve({k: v.to(dev, non_blocking=True) for k, v in h.items()})
        for k, t in res.items():
            out_h[k].copy_(t, non_blocking=True)

    for _ in range(2):
        e2e_step()
    barrier()
    k_e2e = max(3, args.steps // 2)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(k_e2e):
        e2e_step()
    e3.record()
    barrier()
    e2e_serial_ms = torch.tensor([e2.elapsed_time(e3)], device=dev, dtype=torch.float64)
    torch.cuda.synchronize()
    serial_out = {k: v.clone() for k, v in out_h.items()}
    for v in out_h.values():
        v.zero_()
    pipelined_ms(e2e_solve, h, out_h, 6)  # warm-up: the caching allocator needs a few steps to reach steady state
    barrier()
    k_pipe = max(4, args.steps)
    e2e_ms = torch.tensor([pipelined_ms(e2e_solve, h, out_h, k_pipe)], device=dev, dtype=torch.float64)
    # the overlapped copies must deliver exactly what the serial path delivers (same inputs every step)
    e2e_verified = all(torch.equal(out_h[k], serial_out[k]) for k in out_h)
    if not e2e_verified:
        raise RuntimeError("pipelined e2e results differ from the serial e2e results")
    barrier()
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_serial_ms, op=dist.ReduceOp.MAX)
    e2e_ms = e2e_ms.item()
    e2e_serial_ms = e2e_serial_ms.item()
    # ---- e2e one level up: Tracking_MPC (the reference's actual caller, policies.py:1117-1315): host inputs are
    # the fp32 network outputs (x0, x_ref, u_ref) + output grads; the cost build and its chain rule run fused
    import types as _types
    from deq_mpc_corl_b200.tracking import Tracking_MPC

    targs = _types.SimpleNamespace(T=T, bsz=B, Q=torch.tensor(spec["Q"]), R=torch.tensor(spec["R"]),
                                   dtype="double" if dtype == torch.float64 else "float", device=dev,
                                   solver_type="al", env=model, rho_init_max=1e4)
    tenv = _types.SimpleNamespace(nx=nx, nu=nu, nq=None, dt=spec["dt"], dynamics=dyn, dynamics_derivatives=dyn,
                                  action_space=_types.SimpleNamespace(low=inp["u_lower"].numpy(), high=inp["u_upper"].numpy()))
    tm = Tracking_MPC(targs, tenv)
    th = dict(x0=h["x0"], x_ref=h["x_ref"], u_ref=h["u_ref"], gx=h["gx"], gu=h["gu"])
    tout = dict(x=out_h["x"], u=out_h["u"], gxr=torch.empty(B, T, nx).pin_memory(), gur=torch.empty(B, T, nu).pin_memory())
    t_h2d = sum(t.numel() * t.element_size() for t in th.values())
    t_d2h = sum(t.numel() * t.element_size() for t in tout.values())

    def e2e_tracking_solve(d):
        xr, ur = d["x_ref"].detach().requires_grad_(True), d["u_ref"].detach().requires_grad_(True)
        tm.reinitialize(xr.detach(), None)
        x, u, _ = tm(d["x0"], None, xr, ur, al_iters=2)
        gxr, gur = torch.autograd.grad((x, u), (xr, ur), (d["gx"], d["gu"]))
        return dict(x=x.detach(), u=u.detach(), gxr=gxr, gur=gur)

    pipelined_ms(e2e_tracking_solve, th, tout, 6)
    barrier()
    e2e_t_ms = torch.tensor([pipelined_ms(e2e_tracking_solve, th, tout, k_pipe)], device=dev, dtype=torch.float64)
    barrier()
    if world > 1:
        dist.all_reduce(e2e_t_ms, op=dist.ReduceOp.MAX)
    e2e_t_ms = e2e_t_ms.item()
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        fl = algorithmic_flops(model, T)
        by = algorithmic_bytes(model, T, esize)
        fma_name = "fp64_fma" if dtype == torch.float64 else "fp32_fma"
        ach_tflops = fl["newton"] * B / (newton_launch_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_ach = by["newton_launch"] * B / (newton_launch_ms * 1e-3) / 1e9
        kw = kernel_work(model, T, esize)
        kernels = {}
        for kn, ms in kernel_ms.items():
            tf = kw[kn]["flops"] * B / (ms * 1e-3) / 1e12
            gb = kw[kn]["bytes"] * B / (ms * 1e-3) / 1e9
            kernels[kn] = {"avg_launch_ms": ms, "flops_per_problem": kw[kn]["flops"], "bytes_per_problem": kw[kn]["bytes"],
                           "fma_tflops": tf, "fma_frac": tf / peak_tflops, "hbm_gbs": gb, "hbm_frac": gb / hbm_peak,
                           "algorithmic_bytes_per_launch": kw[kn]["bytes"] * B,
                           "traffic": NCU_TRAFFIC.get((model, T, B, args.dtype), {}).get(kn)}
        dom = max(kernel_ms, key=kernel_ms.get)
        dk = kernels[dom]
        dom_fma = dk["fma_frac"] >= dk["hbm_frac"]
        value = args.steps * B * world / (el_ms * 1e-3)
        line = {
            "metric": "AL-MPC solves/sec (fwd+implicit bwd)", "value": value, "unit": "solves/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": el_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(args, B, world),
            "newton_steps_per_al_iter": newton,
            "clocks": clocks,
            "e2e": {"value": k_pipe * B * world / (e2e_ms * 1e-3), "unit": "solves/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / k_pipe, "steps": k_pipe, "matches_serial": e2e_verified,
                    "api": "deq_mpc_corl_b200.al_mpc.MPC.forward + autograd backward; every step copies its inputs "
                           "from pinned host buffers and its results back, double-buffered: the H2D of step j+1 and the "
                           "D2H of step j-1 run on copy streams under the solve of step j",
                    "serial": {"value": k_e2e * B * world / (e2e_serial_ms * 1e-3), "ms_per_step": e2e_serial_ms / k_e2e,
                               "note": "same step with copy-in, solve, copy-out serialised on one stream"}},
            "e2e_tracking": {"value": k_pipe * B * world / (e2e_t_ms * 1e-3), "unit": "solves/s", "h2d_bytes_per_step": t_h2d,
                             "d2h_bytes_per_step": t_d2h, "ms_per_step": e2e_t_ms / k_pipe,
                             "api": "deq_mpc_corl_b200.tracking.Tracking_MPC.forward (fused cost build) + autograd backward to x_ref/u_ref; same double-buffered copies"},
            "gpu_launches": args.steps * (cfg.al_iter * cfg.max_newton * 3 + 2),  # 3 kernels per Newton-step slot + final dual + backward
            "roofline": {
                # dominant kernel by time, against whichever of its two rooflines it sits closer to
                "kernel": dom, "bound": (fma_name if dom_fma else "hbm"),
                "achieved": dk["fma_tflops"] if dom_fma else dk["hbm_gbs"],
                "peak": peak_tflops if dom_fma else hbm_peak, "unit": "TFLOP/s" if dom_fma else "GB/s",
                "frac": dk["fma_frac"] if dom_fma else dk["hbm_frac"], "avg_launch_ms": dk["avg_launch_ms"],
                "traffic": dk["traffic"],
                "peak_source": {"fma": "measured in this run by almpc_fma_peak (8 independent FMA chains/thread); MEASURED_PEAKS.json has no FP64/FP32 FMA figure",
                                "hbm": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s"},
                "kernels": kernels,
                # the whole Newton step (3 launches) against the FMA roofline that binds the path
                "newton_step": {"bound": fma_name, "achieved": ach_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
                                "frac": ach_tflops / peak_tflops, "flops_per_problem": fl["newton"],
                                "ms": newton_launch_ms,
                                "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                                        "bytes_per_problem": by["newton_launch"]}},
            },
            "flops_per_solve": fl["newton"] * n_newton + 2 * fl["dual"] + fl["bwd"],
        }
        if not args.no_cpu_baseline and world == 1:
            v, steps, secs = cpu_oracle_run(model, T, args.cpu_sample, dtype, 1003)
            line["cpu_baseline"] = {
                "value": v, "unit": "solves/s", "cores": os.cpu_count(), "kind": "port",
                "sample": f"{args.cpu_sample} of the {B} problems, 1 fwd+bwd, {secs:.1f} s, oracle port of qpth AL_mpc "
                          f"(torch {torch.__version__} CPU, {os.cpu_count()} threads), newton steps {steps}"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
