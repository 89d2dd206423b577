"""Helpers shared by the oracle-vs-golden (CPU) and CUDA-vs-golden/oracle (GPU) parity tests."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def list_golden():
    """Solver-level goldens (MPC.forward chains).  `tracking_*` files are goldens of the reference's caller."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR)
                  if f.endswith(".npz") and not f.endswith("_traj.npz") and not f.startswith(("tracking_", "ip_")))


def list_tracking_golden():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f.startswith("tracking_"))


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.z = z
        self.name = name
        self.model = str(z["meta_model"])
        self.dtype = torch.float64 if str(z["meta_dtype"]) == "f64" else torch.float32
        self.T, self.B = int(z["T"]), int(z["B"])
        self.n_calls = int(z["n_calls"])
        self.stream_after = int(z["stream_after"])
        self.rho_init_max = float(z["rho_init_max"])
        self.state_estimator = bool(z["state_estimator"]) if "state_estimator" in z.files else False

    def t(self, key):
        return torch.from_numpy(np.asarray(self.z[key]))

    def call(self, ci):
        pfx = f"c{ci}_"
        return {k[len(pfx):]: torch.from_numpy(np.asarray(self.z[k])) for k in self.z.files if k.startswith(pfx)}


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def max_rel_err_rows(a, b):
    """max over problems of ||a_i-b_i|| / max(||b_i||, tiny)."""
    a, b = a.double().reshape(a.shape[0], -1), b.double().reshape(b.shape[0], -1)
    return ((a - b).norm(dim=1) / b.norm(dim=1).clamp_min(1e-300)).max().item()
