#!/usr/bin/env python
"""Concurrent host<->device copy ceiling of this box: 1, 2, 4, 8 GPUs copying at the same time from one process per GPU
(page-locked buffers, H2D alone, D2H alone, both directions at once).  This is the roof over the end-to-end numbers of
`bench.py --gpus N`, whose ranks share the host's PCIe / memory path.  usage: python tools/pcie_probe_multi.py [--mb 32]"""
import argparse
import json
import multiprocessing as mp
import time


def worker(gpu, mb, start_evt, q, reps):
    import torch
    torch.cuda.set_device(gpu)
    n = mb << 20
    h, h2 = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d, d2 = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    out = {}
    for name in ("h2d", "d2h", "both"):
        for _ in range(3):
            d.copy_(h, non_blocking=True); h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize()
        q.put(("ready", gpu, name))
        start_evt[name].wait()
        t0 = time.perf_counter()
        for _ in range(reps):
            if name in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d.copy_(h, non_blocking=True)
            if name in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        out[name] = n * (2 if name == "both" else 1) / dt / 1e9
    q.put(("done", gpu, out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=32)
    ap.add_argument("--reps", type=int, default=40)
    args = ap.parse_args()
    import torch
    have = torch.cuda.device_count()
    ctx = mp.get_context("spawn")
    table = {}
    for k in [g for g in (1, 2, 4, 8) if g <= have]:
        q = ctx.Queue()
        evts = {name: ctx.Event() for name in ("h2d", "d2h", "both")}
        procs = [ctx.Process(target=worker, args=(g, args.mb, evts, q, args.reps)) for g in range(k)]
        for p in procs:
            p.start()
        results, ready = {}, {name: 0 for name in evts}
        while len(results) < k:
            kind, gpu, val = q.get()
            if kind == "ready":
                ready[val] += 1
                if ready[val] == k:
                    evts[val].set()          # all processes start this direction together
            else:
                results[gpu] = val
        for p in procs:
            p.join()
        table[str(k)] = {name: {"aggregate_GBps": round(sum(r[name] for r in results.values()), 1), "per_gpu_GBps": [round(results[g][name], 1) for g in range(k)]} for name in evts}
    print(json.dumps({"buffer_MB": args.mb, "concurrent_processes": table}))


if __name__ == "__main__":
    main()
