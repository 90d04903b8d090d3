#!/usr/bin/env python
"""Write profiles/sass_*.txt: per-kernel mnemonic histograms of the shipped libfos_b200.so plus the
lines that prove the Blackwell paths (UBLKCP = cp.async.bulk / TMA bulk engine, SYNCS = mbarrier,
DMMA = fp64 tensor MMA, LDGSTS = cp.async, system-scope / peer LD/ST/RED of the fused exchange).

    python tools/sass_evidence.py            # needs cuobjdump on PATH; runs without a GPU
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "fastoptsolver_b200", "libfos_b200.so")
WANT = {
    "sass_grad_stream.txt": [r"solve_stream_kernelIdLi256ELi16ELi1", r"grad_stream_kernelIdLi256ELi16ELi1ELb0ELb0",
                             r"solve_stream_kernelIfLi256ELi16ELi2", r"15epilogue_kernel"],
    "sass_gram.txt": [r"gram_syrk_tma_kernel", r"gram_syrk_kernel", r"path_step_sk_kernelILi128ELb1", r"path_step_kernelILi128ELb1",
                      r"path_step_kernelILi128ELb0", r"gram_matvec_kernel", r"mrhs_stream_kernel"],
}
KEYS = ("UBLKCP", "UTMALDG", "SYNCS", "DMMA", "LDGSTS", "DFMA", "SHFL", "LDS", "BAR", "MEMBAR", "RED", "ATOM", ".SYS", "LDG", "STG",
        "UCGABAR", "ACQBULK", "CCTL", "ERRBAR", "NANOSLEEP")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    funcs = {}
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.search(r"/\*[0-9a-f]{4}\*/", line):
            funcs[name].append(line)
    for fname, pats in WANT.items():
        out = [f"# {fname}: cuobjdump -sass of fastoptsolver_b200/libfos_b200.so (sm_100a), written by tools/sass_evidence.py",
               "# per kernel: instruction count, histogram of the mnemonics that matter, first occurrences of the proof lines", ""]
        for pat in pats:
            for fn, lines in funcs.items():
                if not re.search(pat, fn):
                    continue
                demangled = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip() or fn
                ops = collections.Counter()
                for ln in lines:
                    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
                    if m:
                        ops[m.group(1).split(".")[0]] += 1
                out.append(f"== {demangled}")
                out.append(f"   {len(lines)} SASS instructions; " + ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))
                shown = collections.Counter()
                for ln in lines:
                    for k in KEYS:
                        if k in ln and shown[k] < (3 if k in ("LDS", "LDG", "STG", "DFMA", "SHFL", "BAR") else 6):
                            shown[k] += 1
                            out.append("   " + re.sub(r"\s+", " ", ln.strip())[:150])
                            break
                out.append("")
        open(os.path.join(ROOT, "profiles", fname), "w").write("\n".join(out))
        print("wrote", fname, len(out), "lines")


if __name__ == "__main__":
    sys.exit(main())
