"""Locate and import the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY (tests/, bench.py's CPU arm, smoke()).

Search order: $ALMPC_REFERENCE, /root/reference (the build container), oracle/_ref/py (a verbatim copy staged
by `make -C oracle pyref`; git-ignored, travels to the GPU box with the snapshot, never part of the history).

Two modules the reference imports do not exist anywhere: ``ipdb`` (debugger; an empty stub) and
``rexquad_utils`` (un-shipped helper module of deqmpc/rex_quadrotor.py and flying_cartpole2d.py; replaced by
oracle/rexquad_standin.py — textbook MRP / quaternion formulas, *unpinned* by the reference).
No reference file is edited.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref", "py")


def reference_root():
    """Directory holding `qpth/` and `deqmpc/` of the reference, or None."""
    for cand in (os.environ.get("ALMPC_REFERENCE"), "/root/reference", STAGED):
        if cand and os.path.isfile(os.path.join(cand, "qpth", "AL_mpc.py")):
            return cand
    return None


def available():
    return reference_root() is not None


def import_reference(with_policies=False):
    """-> namespace with AL_mpc, al_utils, rex_quadrotor, flying_cartpole2d (and policies, qp, qp_wrapper)."""
    root = reference_root()
    if root is None:
        raise ImportError("reference not found: neither /root/reference nor oracle/_ref/py (make -C oracle pyref)")
    sys.modules.setdefault("ipdb", types.ModuleType("ipdb"))
    import oracle.rexquad_standin as standin

    sys.modules["rexquad_utils"] = standin
    for p in (root, os.path.join(root, "deqmpc")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import contextlib
    import io
    import warnings

    # the reference prints at import time ("cvxpy not installed, skipping test"): keep stdout clean for callers that
    # print machine-readable output (bench.py's one JSON line)
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        from qpth import AL_mpc, al_utils
        import rex_quadrotor
        import flying_cartpole2d

        ns = types.SimpleNamespace(root=root, AL_mpc=AL_mpc, al_utils=al_utils, rex_quadrotor=rex_quadrotor,
                                   flying_cartpole2d=flying_cartpole2d)
        if with_policies:
            import policies
            from qpth import qp, qp_wrapper

            ns.policies, ns.qp, ns.qp_wrapper = policies, qp, qp_wrapper
    return ns
