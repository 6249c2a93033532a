#!/usr/bin/env python
"""Device-side rendering of t6 rows (vsgpu_render_t6) on the bench's chr22-shaped index: rows/s and
text GB/s of the kernels, end to end through the C ABI, and the same regions through the host
materialiser (vsgpu_rows_t6, one thread) and the oracle (t6 text).  One JSON line per batch shape."""
import argparse
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
    ap.add_argument("--shapes", default="1000:20000,100:200000,10000:2000")
    ap.add_argument("--no-oracle", action="store_true")
    a = ap.parse_args()
    args = argparse.Namespace(records=1_103_547, samples=2504, fmax=1100, cache_dir=os.environ.get("VSGPU_BENCH_CACHE", "/tmp/vsgpu_bench"),
                              regions=1_000_000, width=1000)
    torch.cuda.set_device(0)
    prefix, meta = bench.ensure_index(args, 0)
    idx = VariantStoreIndex(prefix, device=0)
    oracle = None if a.no_oracle else T.Oracle.open(prefix)
    for shape in a.shapes.split(","):
        width, n = (int(v) for v in shape.split(":"))
        rng = np.random.default_rng(width)
        x = np.sort(rng.integers(max(1, meta["pos_lo"]), meta["ref_length"] - width, n)).astype(np.uint64)
        y = x + np.uint64(width)
        for ws in (True, False):
            for _ in range(2):
                off, text, rows, ms = idx.render_var_in_ref(x, y, with_samples=ws)
            # timed at the C ABI (pinned inputs, no Python-side copies of the text)
            import ctypes as C
            px = torch.from_numpy(x.astype(np.int64)).pin_memory()
            py = torch.from_numpy(y.astype(np.int64)).pin_memory()
            lib, h = idx._lib, idx._h
            ts, kms = [], []
            for _ in range(7):
                t = C.c_void_p()
                t0 = time.perf_counter()
                rc = lib.vsgpu_render_t6(h, n, C.c_void_p(px.data_ptr()), C.c_void_p(py.data_ptr()), int(ws), C.byref(t))
                ts.append(time.perf_counter() - t0)
                assert rc == 0
                kms.append(float(lib.vsgpu_text_kernel_ms(t)))
                lib.vsgpu_text_free(t)
            out = {"config": "t6 rows rendered on device", "width": width, "regions": n, "with_samples": ws, "rows": rows, "text_bytes": int(off[-1]),
                   "kernels_ms": round(float(np.median(kms)), 4), "rows_per_s_kernels": round(rows / (np.median(kms) / 1e3)),
                   "text_GBps_kernels": round(off[-1] / (np.median(kms) / 1e3) / 1e9, 1),
                   "e2e_ms": round(float(np.median(ts)) * 1e3, 3), "rows_per_s_e2e": round(rows / np.median(ts)),
                   "text_GBps_e2e": round(off[-1] / np.median(ts) / 1e9, 2)}
            if ws:
                # host materialiser on a sample of the same regions (single thread, as the CLI used it)
                lo, hi, cnt = idx.batch_var_in_ref(x, y)
                m = min(n, 300)
                t0 = time.perf_counter()
                host_rows = 0
                for i in range(m):
                    txt = idx.rows_t6_text(int(lo[i]), int(hi[i]))
                    host_rows += int(cnt[i])
                    assert txt.encode() == text[off[i]:off[i + 1]]
                dt = time.perf_counter() - t0
                out["host_rows_per_s"] = round(host_rows / dt)
                out["host_sample_regions"] = m
                if oracle is not None:
                    m2 = min(n, 30)
                    t0 = time.perf_counter()
                    orows = 0
                    for i in range(m2):
                        want = oracle.t6_text(int(x[i]), int(y[i])).split("Pos\tRef\tAlt\tSamples\n", 1)[1]
                        assert want.encode() == text[off[i]:off[i + 1]]
                        orows += int(cnt[i])
                    out["oracle_rows_per_s"] = round(orows / (time.perf_counter() - t0))
                    out["oracle_sample_regions"] = m2
            print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
