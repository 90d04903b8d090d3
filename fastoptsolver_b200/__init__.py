"""fastoptsolver_b200 -- B200-native solver core behind FastOptSolver's module API.

The hot path (fused A^T(Ay-b) gradient, prox / momentum epilogue, power iteration,
L-BFGS vector kernels) lives in ``csrc/`` as hand-written sm_100a CUDA behind the
C ABI declared in ``include/fos.h`` and is loaded with ctypes (``_lib.py``).  There
is no CPU fallback: any solver call raises if ``libfos_b200.so`` is missing.

Drop-in modules with the reference's names live in ``fastoptsolver_b200/dropin``.
"""
__version__ = "0.1.0"
