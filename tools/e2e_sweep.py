#!/usr/bin/env python
"""End-to-end (host buffers -> C ABI -> host results) time of one bench step as a function of the
chunk size the library cuts a call into (VSGPU_CHUNK_REGIONS, read per call).  One JSON line per
setting; the bench's workload (chr22-shaped synthetic, 1 M sorted 1 kb regions)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import bench
    from variantstore_b200 import VariantStoreIndex
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", default="0,65536,131072,262144,524288")
    ap.add_argument("--h2d-streams", default="3,1")
    ap.add_argument("--user-stream", type=int, default=0, help="1: hand torch's current stream to the index first (as bench.py does)")
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    args = argparse.Namespace(records=1_103_547, samples=2504, fmax=1100, cache_dir=os.environ.get("VSGPU_BENCH_CACHE", "/tmp/vsgpu_bench"),
                              regions=1_000_000, width=1000)
    torch.cuda.set_device(0)
    prefix, meta = bench.ensure_index(args, 0)
    x, y, s = bench.make_regions(args, meta, 0)
    n = len(x)
    idx = VariantStoreIndex(prefix, device=0)
    lib, h = idx._lib, idx._h
    if a.user_stream:
        idx.set_stream(torch.cuda.current_stream().cuda_stream)
    px = torch.from_numpy(x.astype(np.int64)).pin_memory()
    py = torch.from_numpy(y.astype(np.int64)).pin_memory()
    ps = torch.from_numpy(s.astype(np.int32)).pin_memory()
    plo, phi, pcnt = (torch.zeros(n, dtype=torch.int32).pin_memory() for _ in range(3))
    vp = C.c_void_p

    def t6():
        assert lib.vsgpu_query_t6(h, n, vp(px.data_ptr()), vp(py.data_ptr()), vp(plo.data_ptr()), vp(phi.data_ptr()), vp(pcnt.data_ptr())) == 0

    def t4():
        r = vp()
        assert lib.vsgpu_query_t4(h, n, vp(px.data_ptr()), vp(py.data_ptr()), vp(ps.data_ptr()), C.byref(r)) == 0
        total = int(lib.vsgpu_result_offsets(r)[n])
        lib.vsgpu_result_free(r)
        return total

    ref = None
    for chunk, ks in [(c, k) for k in a.h2d_streams.split(",") for c in a.chunks.split(",")]:
        os.environ["VSGPU_CHUNK_REGIONS"] = chunk
        os.environ["VSGPU_H2D_STREAMS"] = ks
        for _ in range(3):
            t6(); total = t4()
        out = {"chunk_regions": int(chunk), "h2d_streams": int(ks), "user_stream": a.user_stream}
        for name, fn in (("t6", t6), ("t4", t4)):
            ts = []
            for _ in range(a.steps):
                t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
            out[name + "_ms"] = round(float(np.median(ts)) * 1e3, 4)
        out["step_ms"] = round(out["t6_ms"] + out["t4_ms"], 4)
        out["e2e_regions_per_s"] = round(2 * n / (out["step_ms"] / 1e3))
        sig = (int(pcnt.sum()), total, int(plo[::1000].sum()))
        ref = ref or sig
        out["same_answer"] = sig == ref
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
