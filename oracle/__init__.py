"""CPU oracle for the ISTA / FISTA / FISTA-delta / L-BFGS hot path.

TEST INFRASTRUCTURE ONLY.  This package is a numpy restatement of the reference
algorithms (ElBaldo1/FastOptSolver: iterative_solvers.py, prox_operators.py,
objective_functions.py, lbfgs.py).  It exists so that the CUDA path can be
checked on a box where /root/reference is absent.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it; the product package ``fastoptsolver_b200`` never
does (tests/test_library_cpu.py::test_product_never_imports_oracle enforces that).
``oracle.lbfgs_model`` / ``oracle.gram_model`` restate two pieces of the PRODUCT's own algorithms
(the device L-BFGS driver, the Gram-matrix formulation) so that their equivalence with the
reference's formulation can be pinned on the CPU.

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md section 8c), so the pins are outputs of the unmodified reference,
generated in the dev container by ``tests/golden/make_golden.py`` (which
imports /root/reference) and committed as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this restatement against every one.
"""
from . import ref_numpy  # noqa: F401
from .ref_numpy import (  # noqa: F401
    prox_l1,
    prox_elastic_net,
    compute_objective,
    estimate_lipschitz,
    ista,
    fista,
    fista_delta,
    LBFGSSolver,
    smooth_value_and_grad,
    METRICS,
    ARMIJO_C,
)
