#!/usr/bin/env python
"""Differential check of the t4 kernels on the GPU box at a size where every CTA takes several tiles: k_t4p with the
32-bit-word walk (VSGPU_T4_ROW64=0), with the 64-entry-chunk walk (default), fused with t6, through the device-resident
batch API and the host-buffer API, sorted and shuffled batches, against each other and (a subsample) against the oracle."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import vs_testlib as T
from variantstore_b200 import Batch, VariantStoreIndex

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
tmp = tempfile.mkdtemp()
orc = T.Oracle.synth(os.path.join(tmp, "ser"), ref_length=12_000_000, n_records=300_000, n_samples=400, fmax=160, seed=3, cqf_log2=22)
rng = np.random.default_rng(5)
x = np.sort(rng.integers(1, 12_000_000 - 1000, n)).astype(np.uint64)
y = x + rng.choice([100, 1000, 1000, 1000, 5000], n).astype(np.uint64)
s = rng.integers(1, 401, n).astype(np.uint32)
bad = 0
for row64 in ("0", "1"):
    os.environ["VSGPU_T4_ROW64"] = row64
    with VariantStoreIndex(os.path.join(tmp, "ser"), device=0) as e:
        for name, perm in (("sorted", np.arange(n)), ("shuffled", rng.permutation(n))):
            xs, ys, ss = x[perm], y[perm], s[perm]
            off, hits = e.batch_sample_var_in_ref(xs, ys, ss)
            lo, hi, cnt = e.batch_var_in_ref(xs, ys)
            sub = rng.choice(n, 400, replace=False)
            oc4, od4, ub = orc.batch_t4(xs[sub], ys[sub], ss[sub], False)
            ed4 = e.digest_t4(off, hits, False)
            ok = np.all((oc4 == np.diff(off)[sub]) & (od4 == ed4[sub]) | (ub != 0))
            b = Batch(e, 46, xs, ys, sample_ids=ss)
            b.run(); b.run()
            flo, fhi, fcnt, foff, fhits = b.fetch()
            glo, ghi, gcnt, goff, ghits, gc4 = e.batch_var_and_sample_var_in_ref(xs.astype(np.uint32), ys.astype(np.uint32), ss)
            names = ("batch lo", "batch hi", "batch cnt6", "batch off", "batch hits", "u32 lo", "u32 hi", "u32 cnt6", "u32 off", "u32 hits")
            diff = [nm for nm, (p, q) in zip(names, ((flo, lo), (fhi, hi), (fcnt, cnt), (foff, off), (fhits, hits), (glo, lo), (ghi, hi), (gcnt, cnt), (goff, off), (ghits, hits))) if not np.array_equal(p, q)]
            same = not diff
            print(f"ROW64={row64} {name}: {len(hits)} rows, oracle subsample {'ok' if ok else 'MISMATCH'}, fused batch / fused u32 host-buffer {'ok' if same else 'MISMATCH ' + str(diff)}")
            for nm, (p, q) in zip(names, ((flo, lo), (fhi, hi), (fcnt, cnt), (foff, off), (fhits, hits), (glo, lo), (ghi, hi), (gcnt, cnt), (goff, off), (ghits, hits))):
                if nm in diff and len(p) == len(q):
                    d = np.nonzero(p != q)[0]
                    print("   ", nm, len(d), "differ, first", d[:6], p[d[:6]], q[d[:6]])
            bad += (not ok) + (not same)
print("all ok" if not bad else f"{bad} mismatch(es)")
sys.exit(1 if bad else 0)
