#!/usr/bin/env python
"""Raw pinned-memory copy rates of this box (context for the e2e numbers): H2D, D2H, both at once."""
import json
import time

import torch

torch.cuda.set_device(0)
out = {}
for mb in (4, 8, 16, 64):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for name in ("h2d", "d2h", "both"):
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                if name in ("h2d", "both"):
                    with torch.cuda.stream(s1):
                        d.copy_(h, non_blocking=True)
                if name in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        h2.copy_(d2, non_blocking=True)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 10
        out[f"{name}_{mb}MB_GBps"] = round(n * (2 if name == "both" else 1) / dt / 1e9, 1)
    # one copy + sync latency
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
    out[f"h2d_sync_each_{mb}MB_us"] = round((time.perf_counter() - t0) / 20 * 1e6, 1)
# several H2D copies in flight on different streams (the library sends x, y and the sample ids that way)
for mb in (1, 2, 4, 8):
    n = mb << 20
    for k in (1, 2, 3):
        hs = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(k)]
        ds = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(k)]
        ss = [torch.cuda.Stream() for _ in range(k)]
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                for i in range(k):
                    with torch.cuda.stream(ss[i]):
                        ds[i].copy_(hs[i], non_blocking=True)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 10
        out[f"h2d_{mb}MB_x{k}streams_GBps"] = round(n * k / dt / 1e9, 1)
print(json.dumps(out))
