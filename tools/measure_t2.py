#!/usr/bin/env python
"""t2 = query_sample_from_ref (include/query.h:120-189) on the bench's chr22-shaped index: sample
sequences over batches of regions.  Per batch shape one JSON line: device time of the three stages
(count, plan, copy; CUDA events inside the library), sequence GB/s of the copy kernel against the
HBM peak (it reads and writes every byte once: algorithmic bytes = 2 x text), end-to-end time through
the C ABI with page-locked inputs, and the oracle (CPU port of the reference) on a sample of the
same regions, checked byte for byte."""
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
    import vs_testlib as T
    from variantstore_b200 import VariantStoreIndex
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="1000:1000000,100:1000000,10000:100000,100000:10000,1000000:1000")
    ap.add_argument("--oracle-sample", type=int, default=300)
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--t3", action="store_true", help="query_sample_from_sample (sample coordinates) instead of query_sample_from_ref")
    a = ap.parse_args()
    # t3 reads the per-carrier sample positions: the index is then built with fix_sample_indexes, as the reference's construct does
    args = argparse.Namespace(records=1_103_547, samples=2504, fmax=1100, cache_dir=os.environ.get("VSGPU_BENCH_CACHE", "/tmp/vsgpu_bench"),
                              regions=1_000_000, width=1000, fix_idx=a.t3)
    torch.cuda.set_device(0)
    prefix, meta = bench.ensure_index(args, 0)
    idx = VariantStoreIndex(prefix, device=0)
    oracle = T.Oracle.open(prefix) if a.oracle_sample else None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6545.3))
    lib, h = idx._lib, idx._h
    for shape in a.shapes.split(","):
        width, n = (int(v) for v in shape.split(":"))
        rng = np.random.default_rng(width)
        x = np.sort(rng.integers(max(1, meta["pos_lo"]), meta["ref_length"] - width, n)).astype(np.uint64)
        y = x + np.uint64(width)
        s = rng.integers(1, args.samples + 1, n).astype(np.uint32)
        t_first = time.perf_counter()
        off, text, st, _ = (idx.batch_sample_seq_in_sample if a.t3 else idx.batch_sample_seq_in_ref)(x, y, s)
        t_first = time.perf_counter() - t_first
        px = torch.from_numpy(x.astype(np.int64)).pin_memory()
        py = torch.from_numpy(y.astype(np.int64)).pin_memory()
        ps = torch.from_numpy(s.astype(np.int32)).pin_memory()
        ts, stages = [], []
        for _ in range(a.reps):
            t = C.c_void_p()
            t0 = time.perf_counter()
            rc = (lib.vsgpu_query_t3 if a.t3 else lib.vsgpu_query_t2)(h, n, C.c_void_p(px.data_ptr()), C.c_void_p(py.data_ptr()), C.c_void_p(ps.data_ptr()), C.byref(t))
            ts.append(time.perf_counter() - t0)
            assert rc == 0
            sm = lib.vsgpu_text_stage_ms(t)
            stages.append([sm[0], sm[1], sm[2]])
            lib.vsgpu_text_free(t)
        stages = np.median(np.array(stages), axis=0)
        nbytes = int(off[-1])
        kms = float(stages.sum())
        out = {"config": "t3 sample sequences in sample coordinates" if a.t3 else "t2 sample sequences in ref coordinates", "first_call_s": round(t_first, 3), "width": width, "regions": n, "text_bytes": nbytes, "threw": int((st == 1).sum()), "reference_hangs": int((st == 2).sum()),
               "count_ms": round(float(stages[0]), 4), "plan_ms": round(float(stages[1]), 4), "copy_ms": round(float(stages[2]), 4),
               "kernels_ms": round(kms, 4), "regions_per_s_kernels": round(n / (kms / 1e3)),
               "copy_GBps_algorithmic": round(2 * nbytes / (stages[2] / 1e3) / 1e9, 1), "copy_frac_of_hbm_peak": round(2 * nbytes / (stages[2] / 1e3) / 1e9 / hbm, 3),
               "all_kernels_GBps_algorithmic": round((2 * nbytes + 292 * n) / (kms / 1e3) / 1e9, 1),
               "e2e_ms": round(float(np.median(ts)) * 1e3, 3), "regions_per_s_e2e": round(n / np.median(ts)), "text_GBps_e2e": round(nbytes / np.median(ts) / 1e9, 2),
               "h2d_bytes": 20 * n, "d2h_bytes": nbytes + 9 * n + 8}
        if oracle is not None:
            m = min(n, a.oracle_sample)
            sub = rng.choice(n, m, replace=False)
            t0 = time.perf_counter()
            ln, dg, ost, ub, seqs = (oracle.batch_t3 if a.t3 else oracle.batch_t2)(x[sub], y[sub], s[sub], want_text=True)
            dt = time.perf_counter() - t0
            for j, i in enumerate(sub):
                assert int(ost[j]) == int(st[i]) and seqs[j].encode() == text[off[i]:off[i + 1]], (int(x[i]), int(y[i]), int(s[i]))
            out["oracle_regions_per_s"] = round(m / dt, 1)
            out["oracle_sample_regions"] = m
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
