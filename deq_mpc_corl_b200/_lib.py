"""ctypes binding of libalmpc.so — the thin host shim above the C ABI (include/almpc.h).

Only raw device pointers, sizes and the current CUDA stream cross this boundary.  There is NO CPU
fallback: if the library is missing or no CUDA device is present the solver raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ALMPC_LIB", os.path.join(_HERE, "libalmpc.so"))

MAX_PARAMS = 48
MAX_NU = 8
MAX_AL_ITER = 32
N_LS = 20
INFO_AL_DONE, INFO_RHO_EXCEEDED, INFO_TOTAL_NEWTON, INFO_NEWTON0 = 0, 1, 2, 4
INFO_INTS = 4 + MAX_AL_ITER
STATUS_FINITE, STATUS_CHOL_OK = 1, 2

DYN_PENDULUM, DYN_CARTPOLE1L, DYN_REX_QUADROTOR, DYN_FLYING_CARTPOLE, DYN_CARTPOLE2L = range(5)


class Desc(ctypes.Structure):
    """struct almpc_desc (include/almpc.h)."""

    _fields_ = [
        ("B", ctypes.c_int32),
        ("T", ctypes.c_int32),
        ("dyn_id", ctypes.c_int32),
        ("al_iter", ctypes.c_int32),
        ("max_newton", ctypes.c_int32),
        ("stream_mode", ctypes.c_int32),
        ("rho_f32", ctypes.c_int32),
        ("flags", ctypes.c_int32),
        ("rho_mult", ctypes.c_double),
        ("rho_max", ctypes.c_double),
        ("dyn_params", ctypes.c_double * MAX_PARAMS),
        ("u_lower", ctypes.c_double * MAX_NU),
        ("u_upper", ctypes.c_double * MAX_NU),
        ("n_obs", ctypes.c_int32),
        ("select", ctypes.c_int32),
        ("obs_radius", ctypes.c_double),
        ("obstacles", ctypes.c_void_p),
        ("profile_ms", ctypes.POINTER(ctypes.c_double)),
        ("stop_stats", ctypes.c_void_p),
        ("stop_reduce", ctypes.c_void_p),
        ("stop_reduce_user", ctypes.c_void_p),
    ]


# int (*almpc_stop_reduce_fn)(double* stats_dev, void* stream, void* user)
STOP_REDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p)

SELECT_FAC_SINGLE, SELECT_FUSE, SELECT_LS_CTA, SELECT_PDL, SELECT_NO_RING, SELECT_NO_PDL, SELECT_FAC_PAIR, SELECT_FAC_MMA = 1, 2, 4, 8, 16, 32, 64, 128


def env_select():
    """desc.select from the environment (A/B measurements, differential tests): ALMPC_FAC_SINGLE, ALMPC_FUSE,
    ALMPC_LS_CTA, ALMPC_PDL, ALMPC_NO_PDL, ALMPC_NO_RING, ALMPC_FAC_PAIR, ALMPC_FAC_MMA.  Read per call on the host side; the library itself has no global state."""
    e = os.environ
    return ((SELECT_FAC_SINGLE if e.get("ALMPC_FAC_SINGLE") else 0) | (SELECT_FUSE if e.get("ALMPC_FUSE") else 0)
            | (SELECT_LS_CTA if e.get("ALMPC_LS_CTA") else 0) | (SELECT_PDL if e.get("ALMPC_PDL") else 0)
            | (SELECT_NO_RING if e.get("ALMPC_NO_RING") else 0) | (SELECT_NO_PDL if e.get("ALMPC_NO_PDL") else 0)
            | (SELECT_FAC_PAIR if e.get("ALMPC_FAC_PAIR") else 0) | (SELECT_FAC_MMA if e.get("ALMPC_FAC_MMA") else 0))


FLAG_SKIP_FINAL_DUAL = 1
FLAG_STATE_ESTIMATOR = 2

_lib = None


class AlmpcError(RuntimeError):
    pass


def load():
    """dlopen libalmpc.so (built in-tree by deq_mpc_corl_b200/build.py).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AlmpcError(
            f"{LIB_PATH} not found: build it with `python -m deq_mpc_corl_b200.build` "
            "(there is no CPU fallback for the AL-MPC solver)"
        )
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
    dp = ctypes.POINTER(Desc)
    lib.almpc_last_error.restype = ctypes.c_char_p
    lib.almpc_version.restype = ctypes.c_char_p
    lib.almpc_model_dims.argtypes = [i32, ctypes.POINTER(i32), ctypes.POINTER(i32)]
    for fn in (lib.almpc_factor_elems, lib.almpc_workspace_bytes, lib.almpc_debug_elems):
        fn.argtypes = [dp]
        fn.restype = ctypes.c_size_t
    for sfx in ("f64", "f32"):
        f = getattr(lib, "almpc_forward_" + sfx)
        f.argtypes = [dp] + [vp] * 15
        f.restype = ctypes.c_int
        b = getattr(lib, "almpc_backward_" + sfx)
        b.argtypes = [dp] + [vp] * 7
        b.restype = ctypes.c_int
        d = getattr(lib, "almpc_dynamics_" + sfx)
        d.argtypes = [dp, i64] + [vp] * 6
        d.restype = ctypes.c_int
        p = getattr(lib, "almpc_fma_peak_" + sfx)
        p.argtypes = [i32, i32, vp, ctypes.POINTER(dbl), vp]
        p.restype = ctypes.c_int
    for sfx in ("f64", "f32"):
        f = getattr(lib, "almpc_tracking_cost_" + sfx)
        f.argtypes = [i32] * 4 + [vp] * 7
        f.restype = ctypes.c_int
        f = getattr(lib, "almpc_tracking_cost_bwd_" + sfx)
        f.argtypes = [i32] * 4 + [vp] * 10
        f.restype = ctypes.c_int
    for sfx in ("f64", "f32"):
        f = getattr(lib, "almpc_blocktri_" + sfx)
        f.argtypes = [i32] * 5 + [vp] * 8
        f.restype = ctypes.c_int
        f = getattr(lib, "almpc_kkt_solve_" + sfx)
        f.argtypes = [i32] * 4 + [vp] * 5 + [dbl, i32] + [vp] * 10
        f.restype = ctypes.c_int
    _lib = lib
    return lib


EXPORTS = [
    "almpc_model_dims", "almpc_factor_elems", "almpc_workspace_bytes", "almpc_debug_elems",
    "almpc_forward_f64", "almpc_forward_f32", "almpc_backward_f64", "almpc_backward_f32",
    "almpc_dynamics_f64", "almpc_dynamics_f32", "almpc_fma_peak_f64", "almpc_fma_peak_f32",
    "almpc_tracking_cost_f64", "almpc_tracking_cost_f32", "almpc_tracking_cost_bwd_f64", "almpc_tracking_cost_bwd_f32",
    "almpc_blocktri_f64", "almpc_blocktri_f32", "almpc_kkt_solve_f64", "almpc_kkt_solve_f32",
    "almpc_last_error", "almpc_version",
]


def check(rc, what):
    if rc != 0:
        msg = load().almpc_last_error().decode()
        raise AlmpcError(f"{what} failed (code {rc}): {msg}")


def require_cuda(t, what):
    if not t.is_cuda:
        raise AlmpcError(
            f"{what}: tensor is on {t.device}; the AL-MPC solver runs on CUDA only (no CPU fallback)"
        )


def ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def suffix(dtype):
    if dtype == torch.float64:
        return "f64"
    if dtype == torch.float32:
        return "f32"
    raise AlmpcError(f"unsupported solver dtype {dtype}")
