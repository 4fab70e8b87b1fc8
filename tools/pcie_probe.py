#!/usr/bin/env python
"""PCIe ceiling of the GPU box for the end-to-end path: pinned H2D alone, D2H alone, both at once."""
import json
import sys
import time

import torch

MB = 76 * 1024 * 1024
dev = torch.device("cuda", 0)
h_in = torch.empty(MB, dtype=torch.uint8).pin_memory()
h_out = torch.empty(MB, dtype=torch.uint8).pin_memory()
d_in = torch.empty(MB, dtype=torch.uint8, device=dev)
d_out = torch.empty(MB, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(kind, chunks=1, reps=10):
    n = MB // chunks
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        for c in range(chunks):
            sl = slice(c * n, (c + 1) * n)
            if kind in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in[sl].copy_(h_in[sl], non_blocking=True)
            if kind in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out[sl].copy_(d_out[sl], non_blocking=True)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / reps
    return MB / dt / 1e9, dt * 1e3


res = {}
for kind in ("h2d", "d2h", "both"):
    for chunks in (1, 8):
        run(kind, chunks, 2)
        gbs, ms = run(kind, chunks)
        res[f"{kind}_x{chunks}"] = {"GB/s_per_direction": round(gbs, 1), "ms": round(ms, 3)}
print(json.dumps(res))
