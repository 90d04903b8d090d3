#!/usr/bin/env python
"""Summarise ncu artefacts brought back in gpurun_out/ into small tracked text files.

    python tools/summarize_ncu.py full  gpurun_out/prof.ncu-rep  profiles/name.txt
    python tools/summarize_ncu.py list  gpurun_out/launches.csv  profiles/name.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
    "dram__cycles_elapsed.avg.per_second", "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "sm__sass_inst_executed_op_shared_ld.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__ops_path_tensor_src_fp64.sum",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# source: {rep} (ncu --set full --clock-control none); one block per captured launch"]
    for r in rows[2:]:
        lines.append("")
        lines.append("kernel: " + r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"  {k:85s} {r[i]:>22s} {units[i]}")
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", ""))
            wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", ""))
            ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
            lines.append(f"  traffic (dram read + write) per launch: {rd * scale[ur] + wr * scale[uw]:.6e} bytes")
        except Exception:
            pass
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


def launch_list(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    lines = [f"# source: {path} (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised:",
             "# compare SHARES, not absolutes)", f"# total {tot / 1e6:.3f} ms over {sum(len(v) for v in agg.values())} launches", ""]
    for k, v in agg.items():
        lines.append(f"{sum(v) / tot * 100:6.2f}%  n={len(v):4d}  avg={sum(v) / len(v) / 1e3:10.1f} us  {k[:110]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    if len(sys.argv) != 4 or sys.argv[1] not in ("full", "list"):
        sys.exit(__doc__)
    {"full": full, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
