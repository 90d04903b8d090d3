"""Drop-in for the reference's ``lbfgs`` module."""
from fastoptsolver_b200.lbfgs import LBFGSSolver, grad_call_times, reset_metrics  # noqa: F401
