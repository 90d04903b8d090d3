#!/usr/bin/env python
"""Host->device bandwidth from pinned memory (what bounds the e2e upload)."""
import time
import torch
n = 4 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    print(f"H2D pinned 4 GiB: {n / (time.perf_counter() - t0) / 1e9:.1f} GB/s")
import subprocess
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:600])
print(subprocess.run(["bash", "-c", "nproc; numactl -H 2>/dev/null | head -5; cat /sys/devices/system/node/online"], capture_output=True, text=True).stdout)
