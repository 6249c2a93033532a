#!/usr/bin/env python
"""Long-running parity campaign on the CPU: for every seed in [lo, hi) and each of the four fuzz shapes
(overlapping deletions x explicit-id encoding), random record / sample counts, construct with the
oracle, open with the engine's loader + flattener + kernel logic compiled for the host (test-only
simulator), and compare t6, t4, closest_var, t2 / t3 (sample sequences in ref / sample coordinates) and t5 (a sample's variants in its own coordinates) on random regions.  usage: fuzz_campaign.py LO HI
Round 1: seeds 0..299 = 1 200 graphs, 480 000 t6 + t4 region queries, 360 000 closest_var: 0 mismatches;
seeds 300..1999 with t2 added (2 720 000 regions, 28 000 of which the reference throws on), 900..1999 with t3 and 1100..1999
with t5 (1 440 000 / 1 120 000 regions; the reference throws or hangs on about 5 % of them): 0 mismatches."""
import sys, os, tempfile, shutil, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import vs_testlib as T
from vs_testlib import Oracle
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = []
n_t2 = n_threw = n_odd3 = 0
t0 = time.time()
for seed in range(lo, hi):
    for overlap in (False, True):
        for sparse in (False, True):
            d = tempfile.mkdtemp(prefix="fz")
            try:
                rng = np.random.default_rng(seed)
                nrec = int(rng.choice([40, 120, 260, 500]))
                ns = int(rng.choice([3, 12, 40, 70]))
                fa, vcf, names = T.write_fuzz_inputs(d, 1000 + seed, n_records=nrec, n_samples=ns, overlap=overlap, sparse=sparse)
                o = Oracle.construct(fa, vcf, d + "/ser", force_enc=0 if sparse else -1)
                e = T.open_engine(d + "/ser", "hostsim")
                x, y, s = T.random_regions(seed + 7, 400, 4100, widths=(1, 2, 3, 7, 20, 100, 700, 5000), n_samples=len(names))
                b6, b4, ub = T.compare_all(o, e, x, y, s)
                pos = np.concatenate([rng.integers(1, 4200, 300)]).astype(np.uint64)
                b1 = T.compare_t1(o, e, pos)
                x[:4] = [0, 0, 1, 4100]
                b2, threw = T.compare_t2(o, e, x, y, s)
                n_t2 += len(x); n_threw += threw
                b3, odd3 = T.compare_t3(o, e, x, y, s)
                n_odd3 += odd3
                b5, _ = T.compare_t5(o, e, x, y, s)
                b2 = b2 + [("t3", i) for i in b3] + [("t5", i) for i in b5]
                if b6 or b4 or b1 or b2:
                    bad.append(dict(seed=seed, overlap=overlap, sparse=sparse, nrec=nrec, ns=ns, b6=b6[:5], b4=b4[:5], b1=b1[:5], b2=b2[:5]))
                    print("MISMATCH", bad[-1], flush=True)
                o.close()
            finally:
                shutil.rmtree(d, ignore_errors=True)
    if seed % 10 == 0:
        print("seed", seed, "elapsed", round(time.time() - t0), "bad", len(bad), flush=True)
print("DONE", lo, hi, "bad", len(bad), "t2 regions", n_t2, "of which the reference throws", n_threw, "; t3 on the same regions, reference throws or hangs on", n_odd3)
