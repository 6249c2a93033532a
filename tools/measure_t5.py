#!/usr/bin/env python
"""t5 = get_sample_var_in_sample (include/query.h:490-612) on the chr22-shaped index built with
fix_sample_indexes: a sample's variants over regions of its own coordinates.  Per batch shape one JSON
line: device time of the count + write launches (CUDA events inside the library), end-to-end time
through the C ABI with page-locked inputs, and the oracle on a sample of the regions (checked)."""
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
    ap.add_argument("--shapes", default="1000:1000000,10000:1000000,100000:100000")
    ap.add_argument("--oracle-sample", type=int, default=200)
    ap.add_argument("--reps", type=int, default=7)
    a = ap.parse_args()
    args = argparse.Namespace(records=1_103_547, samples=2504, fmax=1100, cache_dir=os.environ.get("VSGPU_BENCH_CACHE", "/tmp/vsgpu_bench"),
                              regions=1_000_000, width=1000, fix_idx=True)
    torch.cuda.set_device(0)
    prefix, meta = bench.ensure_index(args, 0)
    idx = VariantStoreIndex(prefix, device=0)
    oracle = T.Oracle.open(prefix) if a.oracle_sample else None
    lib, h = idx._lib, idx._h
    for shape in a.shapes.split(","):
        width, n = (int(v) for v in shape.split(":"))
        rng = np.random.default_rng(width)
        x = np.sort(rng.integers(max(1, meta["pos_lo"]), meta["ref_length"] - width, n)).astype(np.uint64)
        y = x + np.uint64(width)
        s = rng.integers(1, args.samples + 1, n).astype(np.uint32)
        off, hits, st, _ = idx.batch_sample_var_in_sample(x, y, s)
        px = torch.from_numpy(x.astype(np.int64)).pin_memory()
        py = torch.from_numpy(y.astype(np.int64)).pin_memory()
        ps = torch.from_numpy(s.astype(np.int32)).pin_memory()
        ts, kms = [], []
        for _ in range(a.reps):
            r = C.c_void_p()
            t0 = time.perf_counter()
            rc = lib.vsgpu_query_t5(h, n, C.c_void_p(px.data_ptr()), C.c_void_p(py.data_ptr()), C.c_void_p(ps.data_ptr()), C.byref(r))
            ts.append(time.perf_counter() - t0)
            assert rc == 0
            kms.append(float(lib.vsgpu_result_kernel_ms(r)))
            lib.vsgpu_result_free(r)
        km = float(np.median(kms))
        out = {"config": "t5 sample variants in sample coordinates", "width": width, "regions": n, "rows": int(off[-1]), "rows_per_region": round(float(off[-1]) / n, 3),
               "reference_hangs": int((st == 2).sum()), "kernels_ms": round(km, 4), "regions_per_s_kernels": round(n / (km / 1e3)),
               "e2e_ms": round(float(np.median(ts)) * 1e3, 3), "regions_per_s_e2e": round(n / np.median(ts)),
               "h2d_bytes": 20 * n, "d2h_bytes": 9 * n + 8 + 4 * int(off[-1])}
        if oracle is not None:
            m = min(n, a.oracle_sample)
            sub = rng.choice(n, m, replace=False)
            t0 = time.perf_counter()
            oc, od, ost, ub = oracle.batch_t5(x[sub], y[sub], s[sub])
            dt = time.perf_counter() - t0
            ed = idx.digest_t5(off, hits, s)
            ec = np.diff(off)
            for j, i in enumerate(sub):
                assert int(ost[j]) == int(st[i]) and (ost[j] != 0 or (int(oc[j]) == int(ec[i]) and int(od[j]) == int(ed[i]))), (int(x[i]), int(y[i]), int(s[i]))
            out["oracle_regions_per_s"] = round(m / dt, 1)
            out["oracle_sample_regions"] = m
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
