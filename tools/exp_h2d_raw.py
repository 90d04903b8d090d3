#!/usr/bin/env python
"""Ceiling of host->device copies on this box, independent of the library: pinned source -> HBM
with 1 / 2 / 4 concurrent streams and several chunk sizes, optionally with the allocating thread
pinned to the GPU's NUMA node first (first-touch decides where pinned pages live).  Under torchrun
every rank copies to its own GPU at the same time (the N-way aggregate ceiling).

    python tools/exp_h2d_raw.py [--gb 8]
    python -m torch.distributed.run --nproc-per-node 8 tools/exp_h2d_raw.py --gb 4
"""
import argparse
import glob
import json
import os
import subprocess
import time

import torch


def numa_cpus():
    nodes = {}
    for p in glob.glob("/sys/devices/system/node/node[0-9]*/cpulist"):
        node = int(p.split("node")[-1].split("/")[0])
        cpus = []
        for part in open(p).read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus += list(range(int(a), int(b) + 1))
            elif part:
                cpus.append(int(part))
        nodes[node] = cpus
    return nodes


def gpu_numa(dev):
    try:
        bus = torch.cuda.get_device_properties(dev).pci_bus_id
        dom = torch.cuda.get_device_properties(dev).pci_domain_id
        devid = torch.cuda.get_device_properties(dev).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devid:02x}.0/numa_node"
        return int(open(path).read().strip())
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=8.0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n = int(args.gb * (1 << 30)) // 8
    nodes = numa_cpus()
    gnode = gpu_numa(local)
    if rank == 0:
        print(json.dumps({"numa_nodes": {k: len(v) for k, v in nodes.items()}, "gpu_numa_node": gnode,
                          "affinity": len(os.sched_getaffinity(0))}), flush=True)
        try:
            print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout, flush=True)
        except Exception:
            pass
    dst = torch.empty(n, dtype=torch.float64, device=dev)
    results = []
    for place in ("default", "gpu_node"):
        if place == "gpu_node":
            if gnode is None or gnode < 0 or gnode not in nodes:
                continue
            allowed = set(nodes[gnode]) & os.sched_getaffinity(0)
            if not allowed:
                continue
            old = os.sched_getaffinity(0)
            os.sched_setaffinity(0, allowed)
        src = torch.empty(n, dtype=torch.float64, pin_memory=True)
        src.fill_(1.0)                       # first touch on the current CPU set
        if place == "gpu_node":
            os.sched_setaffinity(0, old)
        for nstream in (1, 2, 4):
            streams = [torch.cuda.Stream(dev) for _ in range(nstream)]
            for chunk_mb in (64, 512, 2048):
                chunk = chunk_mb * (1 << 20) // 8
                pieces = [(o, min(chunk, n - o)) for o in range(0, n, chunk)]
                best = None
                for rep in range(3):
                    if dist is not None:
                        dist.barrier()
                    torch.cuda.synchronize(dev)
                    t0 = time.perf_counter()
                    for i, (o, ln) in enumerate(pieces):
                        with torch.cuda.stream(streams[i % nstream]):
                            dst[o:o + ln].copy_(src[o:o + ln], non_blocking=True)
                    torch.cuda.synchronize(dev)
                    dt = time.perf_counter() - t0
                    best = dt if best is None else min(best, dt)
                gbps = n * 8 / best / 1e9
                if dist is not None:
                    t = torch.tensor([gbps], device=dev, dtype=torch.float64)
                    lst = [torch.empty_like(t) for _ in range(world)]
                    dist.all_gather(lst, t)
                    per = [float(v) for v in lst]
                else:
                    per = [gbps]
                rec = {"pinned_first_touch": place, "streams": nstream, "chunk_mb": chunk_mb, "GBps_per_rank_min": min(per),
                       "GBps_per_rank_max": max(per), "GBps_aggregate": sum(per), "ranks": world}
                results.append(rec)
                if rank == 0:
                    print(json.dumps(rec), flush=True)
        del src
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
