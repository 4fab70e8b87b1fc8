#!/usr/bin/env python
"""PCIe ceiling of the GPU box for the end-to-end path: pinned H2D alone, D2H alone, both at once, with the
pinned buffers first-touched on each NUMA node in turn (os.sched_setaffinity before the allocation)."""
import glob
import json
import os
import subprocess
import sys
import time

import torch

MB = 76 * 1024 * 1024
dev = torch.device("cuda", 0)
d_in = torch.empty(MB, dtype=torch.uint8, device=dev)
d_out = torch.empty(MB, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def cpulist(txt):
    out = []
    for part in txt.strip().split(","):
        if "-" in part:
            a, b = part.split("-")
            out.extend(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return out


def run(h_in, h_out, kind, reps=10):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        if kind in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if kind in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / reps
    return round(MB / dt / 1e9, 1)


res = {}
try:
    bdf = subprocess.check_output(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", "0"], text=True).strip().lower()
    bdf = bdf[-12:] if len(bdf) > 12 else bdf
    res["gpu_numa_node"] = open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip()
except Exception as e:  # noqa: BLE001
    res["gpu_numa_node"] = f"unknown ({e})"
all_cpus = sorted(os.sched_getaffinity(0))
nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
res["nodes"] = len(nodes)
for node in nodes + [None]:
    if node is None:
        os.sched_setaffinity(0, all_cpus)
        tag = "default"
    else:
        cpus = [c for c in cpulist(open(node + "/cpulist").read()) if c in all_cpus]
        if not cpus:
            continue
        os.sched_setaffinity(0, cpus)
        tag = os.path.basename(node)
    h_in = torch.empty(MB, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(MB, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    h_out.fill_(1)
    for kind in ("h2d", "d2h", "both"):
        run(h_in, h_out, kind, 2)
        res[f"{tag}_{kind}_GBs"] = run(h_in, h_out, kind)
    del h_in, h_out
print(json.dumps(res))
