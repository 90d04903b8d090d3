"""CPU: the shipped library really contains the Blackwell code paths the design claims (cuobjdump -sass of
fastoptsolver_b200/libfos_b200.so; no GPU needed).  Guards against a build that silently lost them -- a kernel
recompiled without the bulk-copy ring or the tensor-map staging would still pass parity, only slower."""
import os
import re
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(os.path.dirname(HERE), "fastoptsolver_b200", "libfos_b200.so")


@pytest.fixture(scope="module")
def kernels():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(SO):
        import __graft_entry__
        __graft_entry__.build()
    p = subprocess.run([exe, "-sass", SO], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-500:]
    out, name = {}, None
    for line in p.stdout.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            out[name] = []
        elif name and re.search(r"/\*[0-9a-f]{4}\*/", line):
            out[name].append(line)
        elif ".target" in line or "arch =" in line:
            out.setdefault("__arch__", []).append(line)
    return out


def _pick(kernels, pattern):
    hits = {k: v for k, v in kernels.items() if re.search(pattern, k)}
    assert hits, f"no kernel matching {pattern}"
    return hits


def _count(lines, mnemonic):
    return sum(1 for ln in lines if re.search(r"\b" + re.escape(mnemonic), ln))


def test_built_for_sm_100a(kernels):
    assert any("sm_100" in ln for ln in kernels.get("__arch__", [])), kernels.get("__arch__")


def test_streaming_kernels_use_the_bulk_copy_ring(kernels):
    """grad_stream_kernel / solve_stream_kernel: cp.async.bulk (UBLKCP) completed on mbarriers (SYNCS), no
    per-thread LDGSTS staging; the persistent kernel also carries the grid barrier (RED ... .GPU) and the peer
    flags (.SYS)."""
    for pat in (r"grad_stream_kernelIdLi256ELi16ELi1", r"solve_stream_kernelIdLi256ELi16ELi1", r"solve_stream_kernelIfLi256ELi16ELi2"):
        for name, lines in _pick(kernels, pat).items():
            assert _count(lines, "UBLKCP") >= 1, name
            assert _count(lines, "SYNCS") >= 2, name
            assert _count(lines, "LDGSTS") == 0, name
            assert _count(lines, "DFMA") >= 32, name
    for name, lines in _pick(kernels, r"solve_stream_kernelIdLi256ELi16ELi1").items():
        text = "\n".join(lines)
        assert re.search(r"RED[A-Z.]*\S*GPU", text) or ".GPU" in text, name
        assert ".SYS" in text, name


def test_gram_kernels_use_tensor_maps_and_fp64_mma(kernels):
    """The default Gram build and path step: UTMALDG (cp.async.bulk.tensor) + mbarriers + DMMA, no LDGSTS; the
    A/B variants keep the cp.async ring."""
    for pat in (r"gram_syrk_tma_kernel", r"path_step_sk_kernelILi128ELb1", r"path_step_sk_kernelILi64ELb1",
                r"path_step_sk_kernelILi32ELb1", r"path_step_kernelILi128ELb1"):
        for name, lines in _pick(kernels, pat).items():
            assert _count(lines, "UTMALDG") >= 2, name
            assert _count(lines, "SYNCS") >= 3, name
            assert _count(lines, "DMMA") >= 32, name
            assert _count(lines, "LDGSTS") == 0, name
    for pat in (r"gram_syrk_kernel", r"path_step_sk_kernelILi128ELb0"):
        for name, lines in _pick(kernels, pat).items():
            assert _count(lines, "LDGSTS") >= 4 and _count(lines, "UTMALDG") == 0, name


def test_multi_rhs_kernel_uses_cluster_exchange(kernels):
    """mrhs_stream_kernel: bulk-copy ring, DMMA, and the distributed-shared-memory exchange (st.async shows up as
    ST...  with the cluster barrier instructions UCGABAR)."""
    for name, lines in _pick(kernels, r"mrhs_stream_kernel").items():
        assert _count(lines, "UBLKCP") >= 1 and _count(lines, "DMMA") >= 16, name
        assert _count(lines, "UCGABAR") >= 1, name
