"""Drop-in for the reference's ``prox_operators`` module."""
from fastoptsolver_b200.operators import prox_elastic_net, prox_l1  # noqa: F401
