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


def build(force=False, verbose=False):
    nvcc = _nvcc()
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
            cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
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
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
