"""B200-native batched augmented-Lagrangian MPC solver (drop-in for qpth.AL_mpc.MPC / al_utils.NewtonAL)."""
__version__ = "0.1.0"
