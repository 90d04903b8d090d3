"""Drop-in for the reference's ``iterative_solvers`` module: put this directory first on
``sys.path`` and ``import iterative_solvers`` resolves to the B200 implementation."""
from fastoptsolver_b200.iterative_solvers import *  # noqa: F401,F403
from fastoptsolver_b200.iterative_solvers import (  # noqa: F401
    C, estimate_lipschitz, fista, fista_delta, get_metrics, grad_call_times, ista, last_run,
    ls_call_iters, ls_call_times, reset_metrics)
import fastoptsolver_b200.iterative_solvers as _impl
import sys as _sys


class _Forward(_sys.modules[__name__].__class__):
    """Keep the writable module global ``C`` (iterative_solvers.py:11) in sync with the
    implementation module, which reads it at call time."""

    def __setattr__(self, name, value):
        if name == "C":
            setattr(_impl, "C", value)
        super().__setattr__(name, value)


_sys.modules[__name__].__class__ = _Forward
