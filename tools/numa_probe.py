#!/usr/bin/env python
"""Does host-memory placement explain the H2D variance?  For each NUMA node: bind this process to
the node's CPUs, allocate fresh page-locked buffers (first touch lands on that node) and time
8 MB / 2 MB H2D and D2H copies."""
import glob
import json
import os
import time

import torch


def cpus_of(node):
    s = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    out = []
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-")
            out += list(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return out


torch.cuda.set_device(0)
torch.cuda.synchronize()
info = {"nodes": sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*")),
        "affinity_at_start": len(os.sched_getaffinity(0))}
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
    bdf = bdf.decode() if isinstance(bdf, bytes) else bdf
    info["gpu_bdf"] = bdf
    p = f"/sys/bus/pci/devices/{bdf[-12:].lower()}/numa_node"
    info["gpu_numa_node"] = open(p).read().strip() if os.path.exists(p) else "?"
except Exception as e:          # noqa: BLE001
    info["nvml_error"] = repr(e)
print(json.dumps(info), flush=True)
allowed = os.sched_getaffinity(0)
for node in info["nodes"]:
    cpus = [c for c in cpus_of(node) if c in allowed]
    if not cpus:
        print(json.dumps({"node": node, "skipped": "no allowed cpus"}))
        continue
    os.sched_setaffinity(0, cpus)
    out = {"node": node, "cpus": len(cpus)}
    for mb in (2, 8):
        n = mb << 20
        hbuf = torch.empty(n, dtype=torch.uint8).pin_memory()
        hbuf.fill_(1)
        d = torch.empty(n, dtype=torch.uint8, device="cuda")
        for name in ("h2d", "d2h"):
            best = 1e9
            for rep in range(5):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(10):
                    if name == "h2d":
                        d.copy_(hbuf, non_blocking=True)
                    else:
                        hbuf.copy_(d, non_blocking=True)
                torch.cuda.synchronize()
                best = min(best, (time.perf_counter() - t0) / 10)
            out[f"{name}_{mb}MB_GBps"] = round(n / best / 1e9, 1)
    print(json.dumps(out), flush=True)
    os.sched_setaffinity(0, allowed)
