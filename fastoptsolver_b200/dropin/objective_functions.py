"""Drop-in for the reference's ``objective_functions`` module."""
from fastoptsolver_b200.operators import compute_objective  # noqa: F401
