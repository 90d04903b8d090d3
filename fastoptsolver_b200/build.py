"""Build libfos_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m fastoptsolver_b200.build [--force] [--verbose]

The shared object lands next to this file so that it travels to the GPU box with the
repo snapshot; it is git-ignored.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libfos_b200.so")
OBJDIR = os.path.join(HERE, "csrc", "_obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libfos_b200.so cannot be built")
    return exe


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, variant=None, defines=()):
    """variant="name" + defines=("-DX=1", ...): an A/B build written to libfos_b200_<name>.so (own
    object directory); select it at run time with FOS_LIB_PATH."""
    global OUT, OBJDIR
    nvcc = _nvcc()
    extra = list(defines)
    if variant:
        OUT = os.path.join(HERE, f"libfos_b200_{variant}.so")
        OBJDIR = os.path.join(HERE, "csrc", f"_obj_{variant}")
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "fos.h"))
    headers.append(os.path.abspath(__file__))
    jobs = []
    objs = []
    for src in sources():
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        p = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, p.returncode, p.stdout + p.stderr

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for cmd, rc, out in ex.map(run, jobs):
                if verbose or rc != 0:
                    sys.stderr.write(" ".join(cmd) + "\n" + out + "\n")
                if rc != 0:
                    raise RuntimeError(f"nvcc failed on {cmd[-3]}")
    if jobs or force or _stale(OUT, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT, *objs,
               "-Xlinker", "-rpath,/usr/local/cuda/lib64"]
        cmd, rc, out = run(cmd)
        if rc != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + out + "\n")
            raise RuntimeError("link of libfos_b200.so failed")
    return OUT


if __name__ == "__main__":
    variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, variant=variant,
                 defines=[a for a in sys.argv[1:] if a.startswith("-D")])
    print(path)
