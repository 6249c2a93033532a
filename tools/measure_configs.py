#!/usr/bin/env python
"""Secondary measurements on one B200 (not the bench line): BASELINE.json config [3] — batched t7
lookups on a TCGA-like sparse cohort — and config [4] — the region-width sweep for t4/t6 on the
chr22-shaped index.  Prints one JSON line per measurement; results are copied into profiles/."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timed(batch, steps=10, warmup=3):
    import torch
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        batch.run()
    torch.cuda.synchronize()
    ms = []
    for _ in range(steps):
        flush.zero_()
        batch.run()
        ms.append(sum(batch.timings_ms()))
    return float(np.mean(ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="tcga,width")
    ap.add_argument("--skip-oracle", action="store_true")
    ap.add_argument("--widths", default="", help="comma list: only these widths of the sweep (and no t7 line)")
    ap.add_argument("--tcga-records", type=int, default=3_000_000)
    ap.add_argument("--tcga-samples", type=int, default=10_000)
    ap.add_argument("--lookups", type=int, default=10_000_000)
    ap.add_argument("--t4-regions", type=int, default=1_000_000)
    args = ap.parse_args()
    import torch
    import vs_testlib as T
    from variantstore_b200 import Batch, VariantStoreIndex
    torch.cuda.set_device(0)
    if "tcga" in args.what:
        prefix = "/tmp/vsgpu_bench/tcga/ser"
        os.makedirs(os.path.dirname(prefix), exist_ok=True)
        t0 = time.time()
        o = T.Oracle.synth(prefix, chr_name="2", ref_length=243_199_373, pos_lo=10_000, n_records=args.tcga_records,
                           n_samples=args.tcga_samples, mode=1, seed=77, cqf_log2=25, gzip_level=1)
        av = o.all_variants()
        build_s = time.time() - t0
        idx = VariantStoreIndex(prefix, device=0)
        idx.set_stream(torch.cuda.current_stream().cuda_stream)
        rng = np.random.default_rng(5)
        # half of the lookups use a reported variant as is (insertions and alleles hanging off a dummy
        # vertex are findable, isolated SNPs are not — reference behaviour), half are perturbed
        pick = rng.integers(0, len(av), args.lookups)
        pos = np.array([av[i][0] for i in pick], np.uint64)
        pos[len(pos) // 2:] += 1
        refs = [av[i][1] for i in pick]
        alts = [av[i][2] for i in pick]
        b7 = Batch(idx, 7, pos, refs=refs, alts=alts)
        ms = timed(b7)
        rec = b7.fetch()
        algo, launches = b7.stats()
        # parity on a sample
        sub = rng.choice(len(pos), 3000, replace=False)
        f7, _, _ = o.batch_t7(pos[sub], [refs[i] for i in sub], [alts[i] for i in sub])
        ok = bool(np.array_equal(rec[sub] != 0xFFFFFFFF, f7 == 1))
        print(json.dumps({"config": "tcga-like sparse t7", "records": args.tcga_records, "samples": args.tcga_samples, "lookups": len(pos),
                          "k_t7_ms": ms, "lookups_per_s": len(pos) / (ms / 1000), "found_fraction": float((rec != 0xFFFFFFFF).mean()),
                          "algorithmic_GBps": algo / (ms / 1000) / 1e9, "class_mode": int(idx.info.class_mode),
                          "device_bytes": int(idx.info.device_bytes), "parity_sample_ok": ok, "build_s": build_s}), flush=True)
        idx.close()
        # ---- t4 on the same cohort: 1 M sorted 1 kb regions, a random sample each — the reference's own worst case
        # (eval_data_records/logs/query_luad.out:120: 3 319 s for 1 000 regions); per-sample carried-entry lists
        # (default for explicit-id cohorts) against the hit map (VSGPU_SPARSE_WALK=0)
        n4 = args.t4_regions
        x = np.sort(rng.integers(10_000, 243_199_373 - 1000, n4)).astype(np.uint64)
        y = x + np.uint64(1000)
        s = rng.integers(1, args.tcga_samples + 1, n4).astype(np.uint32)
        sub = rng.choice(n4, 1500, replace=False)
        oc4, od4, ub = o.batch_t4(x[sub], y[sub], s[sub], False)
        for mode in ("1", "0"):
            os.environ["VSGPU_SPARSE_WALK"] = mode
            t0 = time.time()
            idx = VariantStoreIndex(prefix, device=0)
            open_s = time.time() - t0
            idx.set_stream(torch.cuda.current_stream().cuda_stream)
            b4 = Batch(idx, 4, x, y, sample_ids=s)
            ms4 = timed(b4, steps=5, warmup=2)
            off, hits, cnt = b4.fetch()
            ed4 = idx.digest_t4(off, hits, False)
            ok = bool(np.all(((oc4 == np.diff(off)[sub]) & (od4 == ed4[sub])) | (ub != 0)))
            t0 = time.perf_counter()
            for _ in range(3):
                idx.batch_sample_var_in_ref(x.astype(np.uint32), y.astype(np.uint32), s)
            e2e_s = (time.perf_counter() - t0) / 3
            print(json.dumps({"config": "tcga-like sparse t4", "membership": "per-sample carried-entry lists" if mode == "1" else "hit map", "records": args.tcga_records,
                              "samples": args.tcga_samples, "regions": n4, "k_t4_ms": ms4, "t4_regions_per_s": n4 / (ms4 / 1000), "t4_rows_per_region": float(cnt.mean()),
                              "e2e_regions_per_s": n4 / e2e_s, "device_bytes": int(idx.info.device_bytes), "open_s": open_s, "parity_sample_ok": ok}), flush=True)
            b4.close(); idx.close()
        os.environ.pop("VSGPU_SPARSE_WALK", None)
        o.close()
    if "published" in args.what:
        # The shape of the reference's own evaluation logs: 1000 random regions of 43 185 bp on chr22, one
        # sample, types 6 and 4 (eval_data_records/logs/query_chr22_on_disk_vs_v1.log: 1946.6 s and 1237.98 s per 1000 regions, one
        # thread, on-disk vertex blocks, unknown CPU).  Here: the chr22-shaped synthetic, the same shape through the C ABI with host
        # buffers (counts + hit codes, and with the rows as -v text), beside the oracle port on one host thread.
        import bench
        class A: pass
        a = A(); a.records, a.samples, a.fmax, a.cache_dir, a.regions, a.width = 1_103_547, 2504, 1100, "/tmp/vsgpu_bench", 1_000_000, 1000
        prefix, meta = bench.ensure_index(a, 0)
        idx = VariantStoreIndex(prefix, device=0)
        rng = np.random.default_rng(43185)
        n, width = 1000, 43_185
        x = np.sort(rng.integers(meta["pos_lo"], meta["ref_length"] - width, n)).astype(np.uint64)
        y = x + np.uint64(width)
        s = np.full(n, 1234, np.uint32)
        def best(fn, reps=5):
            ts = []
            for _ in range(reps):
                t0 = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t0)
            return min(ts), r
        idx.batch_var_and_sample_var_in_ref(x, y, s)
        t_counts, r = best(lambda: idx.batch_var_and_sample_var_in_ref(x, y, s))
        lo, hi, c6, off4, hits4, _ = r
        idx.render_var_in_ref(x, y, True); idx.render_sample_var_in_ref(x, y, s, True)
        t_rows6, r6 = best(lambda: idx.render_var_in_ref(x, y, True), 3)
        t_rows4, r4 = best(lambda: idx.render_sample_var_in_ref(x, y, s, True), 3)
        # the same two calls through the C ABI alone (text left in the library's page-locked buffer)
        import ctypes as C
        lib, h = idx._lib, idx._h
        def c_call(fn, *a):
            t = C.c_void_p()
            t0 = time.perf_counter(); rc = fn(h, n, *a, 1, C.byref(t)); dt = time.perf_counter() - t0
            assert rc == 0
            lib.vsgpu_text_free(t)
            return dt
        vp = C.c_void_p
        t_c6 = min(c_call(lib.vsgpu_render_t6, vp(x.ctypes.data), vp(y.ctypes.data)) for _ in range(3))
        t_c4 = min(c_call(lib.vsgpu_render_t4, vp(x.ctypes.data), vp(y.ctypes.data), vp(s.ctypes.data)) for _ in range(3))
        o = T.Oracle.open(prefix)
        t0 = time.perf_counter(); oc6, od6 = o.batch_t6(x, y, True)[:2]; t_o6 = time.perf_counter() - t0
        t0 = time.perf_counter(); oc4, od4, ub4 = o.batch_t4(x, y, s, True); t_o4 = time.perf_counter() - t0
        ok6 = bool(np.array_equal(oc6, c6) and np.array_equal(od6, idx.digest_t6(lo, hi, True)))
        ok4 = bool(np.all(((oc4 == np.diff(off4)) & (od4 == idx.digest_t4(off4, hits4, True))) | (ub4 != 0)))
        print(json.dumps({"config": "published shape: 1000 regions x 43185 bp, one sample, chr22-shaped synthetic",
                          "t6_rows_per_region": float(c6.mean()), "t4_rows_per_region": float(np.diff(off4).mean()),
                          "vsgpu_t6_and_t4_counts_and_codes_s": t_counts, "vsgpu_t6_rows_as_text_s": t_rows6, "vsgpu_t4_rows_as_text_s": t_rows4,
                          "t6_text_bytes": int(len(r6[1])), "t4_text_bytes": int(len(r4[1])), "vsgpu_t6_rows_c_abi_s": t_c6, "vsgpu_t4_rows_c_abi_s": t_c4,
                          "t6_render_kernels_ms": r6[3], "t4_render_kernels_ms": r4[3],
                          "oracle_1_thread_t6_s": t_o6, "oracle_1_thread_t4_s": t_o4, "parity_t6": ok6, "parity_t4": ok4,
                          "reference_published_t6_s": 1946.6, "reference_published_t4_s": 1237.98,
                          "note": "oracle rows are digested, not printed; vsgpu text legs include the copy of the text into a Python bytes object"}), flush=True)
        o.close(); idx.close()
        if args.skip_oracle:
            pass
    if "width" in args.what:
        import bench
        class A: pass
        a = A(); a.records, a.samples, a.fmax, a.cache_dir, a.regions, a.width = 1_103_547, 2504, 1100, "/tmp/vsgpu_bench", 1_000_000, 1000
        prefix, meta = bench.ensure_index(a, 0)
        idx = VariantStoreIndex(prefix, device=0)
        idx.set_stream(torch.cuda.current_stream().cuda_stream)
        rng = np.random.default_rng(6)
        only = [int(w) for w in args.widths.split(",") if w]
        for width, n in [(100, 1_000_000), (1000, 1_000_000), (10_000, 1_000_000), (100_000, 200_000), (1_000_000, 50_000), (10_000_000, 10_000)]:
            if only and width not in only:
                continue
            x = np.sort(rng.integers(meta["pos_lo"], meta["ref_length"] - min(width, 30_000_000), n)).astype(np.uint64)
            y = x + np.uint64(width)
            s = rng.integers(1, 2505, n).astype(np.uint32)
            b6, b4, b46 = Batch(idx, 6, x, y), Batch(idx, 4, x, y, sample_ids=s), Batch(idx, 46, x, y, sample_ids=s)
            # one run + fetch first: the hit buffers are sized from the first answer, and a launch whose buffer is too small
            # skips its writes (so timing before the first fetch would flatter wide regions)
            b4.run(); off, hits, cnt = b4.fetch()
            b6.run(); lo, hi, c6 = b6.fetch()
            b46.run(); flo, fhi, fc6, foff, fhits = b46.fetch()
            assert np.array_equal(foff, off) and np.array_equal(fhits, hits) and np.array_equal(fc6, c6) and np.array_equal(flo, lo)
            ms6, ms4, ms46 = timed(b6), timed(b4), timed(b46)
            b46.close()
            algo4, _ = b4.stats()
            print(json.dumps({"config": "width sweep", "width": width, "regions": n, "k_t6_ms": ms6, "k_t4_ms": ms4, "k_fused_t6t4_ms": ms46, "fused_region_queries_per_s": 2 * n / (ms46 / 1000),
                              "t6_regions_per_s": n / (ms6 / 1000), "t4_regions_per_s": n / (ms4 / 1000),
                              "t6_rows_per_region": float(c6.mean()), "t4_rows_per_region": float(cnt.mean()),
                              "t4_rows_per_s": float(cnt.sum()) / (ms4 / 1000), "t4_algorithmic_GBps": algo4 / (ms4 / 1000) / 1e9}), flush=True)
            b6.close(); b4.close()
        if only:
            idx.close()
            return
        # t7 on the same index (the sweep's third operator has no width): 1 M lookups, half of them misses
        o = T.Oracle.open(prefix)
        av = o.all_variants()
        pick = rng.integers(0, len(av), 1_000_000)
        pos = np.array([av[i][0] for i in pick], np.uint64)
        pos[len(pos) // 2:] += 1
        b7 = Batch(idx, 7, pos, refs=[av[i][1] for i in pick], alts=[av[i][2] for i in pick])
        ms7 = timed(b7)
        rec = b7.fetch()
        print(json.dumps({"config": "width sweep", "operator": "t7", "lookups": len(pos), "k_t7_ms": ms7, "t7_lookups_per_s": len(pos) / (ms7 / 1000),
                          "found_fraction": float((rec != 0xFFFFFFFF).mean())}), flush=True)
        b7.close(); o.close()
        idx.close()


if __name__ == "__main__":
    main()
