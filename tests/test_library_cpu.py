"""CPU: the C-ABI library builds, loads and exports every symbol include/fos.h declares;
compute entry points fail loudly without a GPU; the product never imports the oracle."""
import ast
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="module")
def lib():
    from fastoptsolver_b200 import _lib, build
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    header = open(os.path.join(ROOT, "include", "fos.h")).read()
    declared = set(re.findall(r"\b(fos_[a-z0-9_]+)\s*\(", header))
    declared -= {"fos_status", "fos_dtype"}
    from fastoptsolver_b200 import _lib
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in fos.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.fos_abi_version() == 1


def test_struct_layout_matches_header():
    from fastoptsolver_b200 import _lib
    # field order of the ctypes mirrors == field order in the header
    header = open(os.path.join(ROOT, "include", "fos.h")).read()
    for cname, cls in (("fos_pg_params", _lib.PGParams), ("fos_pg_result", _lib.PGResult),
                       ("fos_lbfgs_params", _lib.LbfgsParams), ("fos_lbfgs_result", _lib.LbfgsResult),
                       ("fos_path_params", _lib.PathParams), ("fos_path_result", _lib.PathResult)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), header, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            parts = decl.split(",")
            first = parts[0].split()[-1].lstrip("*")
            names.append(first)
            names.extend(p.strip().lstrip("*") for p in parts[1:])
        assert names == [f[0] for f in cls._fields_], cname


@pytest.mark.skipif(_has_cuda(), reason="CPU-only check")
def test_no_cpu_fallback(lib):
    from fastoptsolver_b200 import iterative_solvers as S
    from fastoptsolver_b200 import _lib
    from fastoptsolver_b200.operators import prox_l1
    assert lib.fos_device_count() == 0
    with pytest.raises(_lib.FosError, match="no CPU fallback"):
        S.fista(np.ones((4, 2)), np.ones(4), "lasso", 0.1, 0.0, max_iter=2)
    with pytest.raises(_lib.FosError, match="no CPU fallback"):
        prox_l1(np.ones(3), 0.5)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fastoptsolver_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith(".py"):
                continue
            tree = ast.parse(open(os.path.join(dirpath, f)).read())
            for node in ast.walk(tree):
                mods = []
                if isinstance(node, ast.Import):
                    mods = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    mods = [node.module or ""]
                for m in mods:
                    assert not m.split(".")[0] == "oracle", f"{f} imports the oracle"
    for f in os.listdir(os.path.join(pkg, "csrc")):
        if f.endswith((".cu", ".cuh")):
            assert "oracle" not in open(os.path.join(pkg, "csrc", f)).read()


def test_dropin_modules_have_reference_names():
    import importlib
    import sys
    d = os.path.join(ROOT, "fastoptsolver_b200", "dropin")
    sys.path.insert(0, d)
    try:
        for m in ("iterative_solvers", "prox_operators", "objective_functions", "lbfgs", "easy_boston_data"):
            sys.modules.pop(m, None)
        IS = importlib.import_module("iterative_solvers")
        for name in ("C", "grad_call_times", "ls_call_times", "ls_call_iters", "reset_metrics", "get_metrics",
                     "estimate_lipschitz", "ista", "fista", "fista_delta"):
            assert hasattr(IS, name)
        assert IS.C == 1e-2
        import inspect
        sig = inspect.signature(IS.fista)
        assert list(sig.parameters) == ["A", "b", "reg_type", "alpha1", "alpha2", "backtracking", "eta",
                                        "t_init_factor", "max_iter", "tol", "tol_ratio", "adaptive_restart",
                                        "restart_threshold", "return_history"]
        assert sig.parameters["max_iter"].default == 500 and sig.parameters["eta"].default == 0.5
        sig = inspect.signature(IS.fista_delta)
        assert list(sig.parameters) == ["A", "b", "reg_type", "alpha1", "alpha2", "delta", "backtracking", "eta",
                                        "t_init_factor", "max_iter", "tol", "tol_ratio", "return_history"]
        sig = inspect.signature(IS.ista)
        assert list(sig.parameters) == ["x0", "g", "grad_g", "prox_h", "L", "backtracking", "eta",
                                        "t_init_factor", "max_iter", "tol", "return_history"]
        assert list(inspect.signature(IS.estimate_lipschitz).parameters) == ["A", "n_iter", "tol"]
        LB = importlib.import_module("lbfgs")
        lb = inspect.signature(LB.LBFGSSolver.__init__).parameters
        positional = [k for k, v in lb.items() if v.kind == v.POSITIONAL_OR_KEYWORD]
        assert positional == ["self", "reg_type", "alpha1", "alpha2", "max_iter", "tol", "eps"]
        # extras must be keyword-only so that reference call sites keep working
        assert all(v.kind == v.KEYWORD_ONLY for k, v in lb.items() if k not in positional)
        assert LB.grad_call_times is IS.grad_call_times
        PO = importlib.import_module("prox_operators")
        assert list(inspect.signature(PO.prox_l1).parameters)[:2] == ["v", "tau"]
        assert list(inspect.signature(PO.prox_elastic_net).parameters)[:4] == ["v", "tau", "alpha1", "alpha2"]
        OF = importlib.import_module("objective_functions")
        assert list(inspect.signature(OF.compute_objective).parameters) == ["x", "A", "b", "reg_type", "alpha1",
                                                                            "alpha2"]
        EB = importlib.import_module("easy_boston_data")
        assert list(inspect.signature(EB.generate_correlated_boston_like_data).parameters) == [
            "m", "seed", "noise_std", "rho1", "rho2"]
        # writable module global C reaches the implementation
        import fastoptsolver_b200.iterative_solvers as impl
        IS.C = 0.5
        assert impl.C == 0.5
        IS.C = 1e-2
        assert impl.C == 1e-2
    finally:
        sys.path.remove(d)
        for m in ("iterative_solvers", "prox_operators", "objective_functions", "lbfgs", "easy_boston_data"):
            sys.modules.pop(m, None)


def test_struct_offsets_match_the_c_compiler(tmp_path):
    """sizeof / offsetof of every ABI struct as gcc lays them out from include/fos.h == the ctypes
    mirrors (a C caller and the Python binding see the same bytes)."""
    import ctypes
    import shutil
    import subprocess
    from fastoptsolver_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    structs = {"fos_pg_params": _lib.PGParams, "fos_pg_result": _lib.PGResult,
               "fos_lbfgs_params": _lib.LbfgsParams, "fos_lbfgs_result": _lib.LbfgsResult,
               "fos_path_params": _lib.PathParams, "fos_path_result": _lib.PathResult}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "fos.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    got = {}
    for ln in out.splitlines():
        c, f, v = ln.split()
        got[(c, f)] = int(v)
    for cname, cls in structs.items():
        assert got[(cname, "size")] == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)
