"""Glue between the shared synthetic-input generator and the oracle models.  TEST INFRASTRUCTURE ONLY."""
from deq_mpc_corl_b200.problems import MODEL_SPECS, make_inputs, tracking_cost  # noqa: F401
from oracle.dynamics_oracle import make_model


def make_oracle_model(name):
    m = make_model(name, dt=MODEL_SPECS[name]["dt"])
    spec = MODEL_SPECS[name]
    m.u_lower, m.u_upper = list(spec["u_lower"]), list(spec["u_upper"])
    return m
