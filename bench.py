#!/usr/bin/env python
"""bench.py — region queries/s (types 4/6) of the batched region path on B200.

A step = one pass of the hot path over one batch of synthetic input: t6 (get_var_in_ref) and t4
(get_sample_var_in_ref, one random sample per region) over the same regions — by default from ONE
fused launch (k_t4p<kFuse6>: the t6 slice falls out of the two index ranks t4 needs anyway;
`--unfused` runs k_t6 then k_t4p as round 1 did, and `by_kernel` always reports both).
Workload at N=1 = BASELINE.json configs[1]: a chr22-shaped synthetic index (~1.1 M records x 2,504
samples) and 1 M random 1 kb regions (sorted, as the reference's read_regions does).  With N > 1
every rank owns its own contig shard of that shape and its own regions (weak scaling, no collective
on the data path — the reference shards by contig too, eval_data_records/evaluation.txt:34).

  python bench.py --gpus 1 --steps 10 --warmup 3            this engine (libvsgpu, CUDA)
  python bench.py --impl reference ...                       the reference's CPU path (oracle port,
                                                             all host cores, bounded sample per step)
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

POS_LO, REF_LEN = 16_050_000, 51_304_566          # scripts/bm_vs_query.sh:13, scripts/run_query.sh:6
# the kernel instances a default run launches (variantstore_b200/csrc/kernels.cu: launch_t4x, 64-region tiles, 24 CTAs / SM)
T4_INSTANCE = "k_t4p<64,24,8,k32=false,fuse6=false,spill=false>"
T4_FUSED_INSTANCE = "k_t4p<64,24,8,k32=false,fuse6=true,spill=false>"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def get_oracle():
    import vs_testlib as T            # the oracle binding lives with the tests (test infrastructure)
    return T


def ensure_index(args, rank):
    """Synthetic ser/ for this rank's contig shard, built once per box with the oracle's construct
    restatement (construction is out of scope for the engine and the reference binary cannot be
    built here) and cached under --cache-dir for the other arm / later runs."""
    fix_idx = bool(getattr(args, "fix_idx", False))     # run fix_sample_indexes (variant_graph.h:1883-1997) as the reference's construct does; only t3 reads those fields
    key = hashlib.sha1((f"v3|{args.records}|{args.samples}|{args.fmax}|{rank}" + ("|fixed" if fix_idx else "")).encode()).hexdigest()[:12]
    prefix = os.path.join(args.cache_dir, f"shard_{key}", "ser")
    done = os.path.join(prefix, ".done")
    if not os.path.exists(done):
        T = get_oracle()
        os.makedirs(os.path.dirname(prefix), exist_ok=True)
        t0 = time.time()
        scale = args.records / 1_103_547
        ref_len = max(200_000, int(REF_LEN * scale)) if scale < 1 else REF_LEN
        pos_lo = int(POS_LO * min(1.0, scale)) if scale < 1 else POS_LO
        o = T.Oracle.synth(prefix, chr_name=str(22 - rank if rank < 22 else rank), ref_length=ref_len, pos_lo=max(2, pos_lo),
                           pos_hi=ref_len - 60_000 if ref_len > 200_000 else ref_len - 1000, n_records=args.records,
                           n_samples=args.samples, fmax=args.fmax, seed=2022 + rank, cqf_log2=25, fix_idx=fix_idx, gzip_level=1)
        info = getattr(o, 'construct_info', None) or o.info()
        o.close()
        with open(done, "w") as f:
            json.dump({"ref_length": ref_len, "pos_lo": pos_lo, "oracle_info": info}, f)
        log(f"[rank {rank}] built synthetic index in {time.time() - t0:.1f}s: {info}")
    meta = json.load(open(done))
    return prefix, meta


def make_regions(args, meta, rank):
    rng = np.random.default_rng(1 + 1000 * rank)
    lo = max(1, meta["pos_lo"])
    hi = meta["ref_length"] - args.width
    x = np.sort(rng.integers(lo, hi, args.regions)).astype(np.uint64)      # read_regions sorts (commands.cc:91)
    y = x + np.uint64(args.width)
    s = np.random.default_rng(2 + 1000 * rank).integers(1, args.samples + 1, args.regions).astype(np.uint32)
    return x, y, s


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons while the GPU is busy.  The timed region lasts a few milliseconds,
    so NVML is polled directly every ~2 ms (nvidia-smi -lms would get one sample); falls back to
    nvidia-smi when pynvml is missing."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.sm, self.reasons, self.sm_max, self.stop_flag = gpu, [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def run(self):
        while not self.stop_flag:
            try:
                if self.nv is not None:
                    self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                    try:
                        r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for name, bit in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(name)
                    time.sleep(0.0005)
                else:
                    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
                    out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                    self.sm.append(float(out[0]))
                    self.sm_max = float(out[1])
                    for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], out[2:]):
                        if v.strip().lower().startswith("active"):
                            self.reasons.add(name)
                    time.sleep(0.1)
            except Exception:
                time.sleep(0.05)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": "nvml" if self.nv is not None else "nvidia-smi"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_sample(args, prefix, meta, nthreads, n_sample, steps=1, warmup=0):
    """Times the oracle (CPU port of the reference's operators) on a bounded sample of the workload."""
    T = get_oracle()
    t0 = time.time()
    o = T.Oracle.open(prefix)
    load_s = time.time() - t0
    x, y, s = make_regions(args, meta, 0)
    pick = np.random.default_rng(7).choice(len(x), min(n_sample, len(x)), replace=False)
    pick.sort()
    xs, ys, ss = x[pick], y[pick], s[pick]
    times = []
    for it in range(warmup + steps):
        t0 = time.time()
        o.timed_counts(6, xs, ys, nthreads=nthreads)
        o.timed_counts(4, xs, ys, ss, nthreads=nthreads)
        dt = time.time() - t0
        if it >= warmup:
            times.append(dt)
    o.close()
    return 2 * len(xs), times, load_s


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    prefix, meta = ensure_index(args, 0)
    cores = os.cpu_count() or 1
    n_sample = args.cpu_sample
    nq, times, load_s = cpu_sample(args, prefix, meta, cores, n_sample, steps=args.steps, warmup=args.warmup)
    ms = 1000 * float(np.mean(times))
    val = nq / (ms / 1000)
    line = {
        "impl": "reference", "metric": "region queries/s (types 4/6)", "value": val, "unit": "regions/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(args, meta),
        "cpu_baseline": {"value": val, "unit": "regions/s", "cores": cores, "kind": "port",
                         "sample": f"{nq // 2} of the {args.regions} regions per type per step (t6 + t4), oracle port of include/query.h, {cores} worker threads over disjoint chunks; index load {load_s:.1f}s not timed"},
        "e2e": {"value": val, "unit": "regions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, meta):
    return {"workload": f"chr22-shaped synthetic index ({args.records} records x {args.samples} samples, ref_length {meta['ref_length']}), "
                        f"{args.regions} random {args.width} bp regions per GPU sorted by start; one step = t6 + t4 over all regions",
            "regions_per_gpu": args.regions, "region_width": args.width, "records": args.records, "samples": args.samples,
            "sharding": "one contig shard per GPU, regions routed by the host, no collective on the data path",
            "l2": "256 MiB buffer written between timed steps (L2 flush)",
            "step": "k_t6 + k_t4p (two launches)" if getattr(args, "unfused", False) else "one fused launch of k_t4p<fuse6> answers t6 and t4",
            "e2e_coordinates": "u64" if getattr(args, "e2e_u64", False) else "u32 (region bounds are parsed with std::stoi, commands.cc:76-80)"}


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json config [2]: the whole-genome-shaped synthetic (24 contigs in GRCh37 lengths, ~85 M records x 2 504
# samples, record counts per contig as in eval_data_records/logs/vs_v1.log), 10 M regions drawn in proportion to
# contig length, position-range shards of at most --genome-shard-records records (the north star's "contig /
# position shard"), placed on the GPUs by longest-processing-time and answered through the C++ router
# (variantstore_b200/csrc/router.cc) in ONE process with host threads per GPU.  Runs for N >= 2 (or --genome).
GENOME = {"1": (249_250_621, 6_468_094), "2": (243_199_373, 7_081_600), "3": (198_022_430, 5_832_276), "4": (191_154_276, 5_732_585),
          "5": (180_915_260, 5_265_763), "6": (171_115_067, 5_024_119), "7": (159_138_663, 4_716_715), "8": (146_364_022, 4_597_105),
          "9": (141_213_431, 3_560_687), "10": (135_534_747, 3_992_219), "11": (135_006_516, 4_045_628), "12": (133_851_895, 3_868_428),
          "13": (115_169_878, 2_857_916), "14": (107_349_540, 2_655_067), "15": (102_531_392, 2_424_689), "16": (90_354_753, 2_697_949),
          "17": (81_195_210, 2_329_288), "18": (78_077_248, 2_267_185), "19": (59_128_983, 1_832_506), "20": (63_025_520, 1_812_841),
          "21": (48_129_895, 1_105_538), "22": (51_304_566, 1_103_547), "X": (155_270_560, 3_468_093), "Y": (59_373_566, 62_042)}


def genome_plan(args):
    """Position-range shards of every contig: [{contig, length, lo, hi, records, seed, prefix}]."""
    out = []
    sc = args.genome_scale
    for ci, (c, (length, records)) in enumerate(GENOME.items()):
        length = max(400_000, int(length * sc)) if sc < 1 else length
        records = max(2_000, int(records * sc))
        k = max(1, -(-records // args.genome_shard_records))
        step = length // k
        for j in range(k):
            lo, hi = (1 if j == 0 else j * step), ((j + 1) * step if j + 1 < k else length + 1)
            key = hashlib.sha1(f"g2|{c}|{length}|{records}|{k}|{j}|{args.samples}|{args.fmax}".encode()).hexdigest()[:10]
            out.append({"contig": c, "length": length, "lo": lo, "hi": hi, "records": records // k, "seed": 7000 + 100 * ci + j,
                        "prefix": os.path.join(args.cache_dir, f"genome_{key}", "ser")})
    return out


def build_shard(spec, samples, fmax):
    """One shard's ser/ by the oracle's construct restatement (a child process of the bench runs this)."""
    T = get_oracle()
    prefix = spec["prefix"]
    if os.path.exists(os.path.join(prefix, ".done")):
        return
    os.makedirs(os.path.dirname(prefix), exist_ok=True)
    t0 = time.time()
    # records of this shard only, in the coordinates of the whole contig; a margin keeps records off the range ends
    pos_lo, pos_hi = max(2, spec["lo"] + 50), min(spec["hi"] - 1200, spec["length"] - 1200)
    o = T.Oracle.synth(prefix, chr_name=spec["contig"], ref_length=spec["length"], pos_lo=pos_lo, pos_hi=pos_hi, n_records=spec["records"],
                       n_samples=samples, fmax=fmax, seed=spec["seed"], cqf_log2=25 if spec["records"] > 200_000 else 20, fix_idx=False, gzip_level=1)
    info = o.construct_info
    o.close()
    with open(os.path.join(prefix, ".done"), "w") as f:
        json.dump({"oracle_info": info, "build_s": time.time() - t0}, f)


def genome_build(args, plan, rank, world):
    """Builds this rank's share of the missing shards with child processes (bounded by cores and free memory)."""
    mine = [sp for i, sp in enumerate(plan) if i % world == rank and not os.path.exists(os.path.join(sp["prefix"], ".done"))]
    if not mine:
        return 0.0
    cores = max(1, (os.cpu_count() or 1) // world)
    try:
        avail_gb = int(open("/proc/meminfo").read().split("MemAvailable:")[1].split()[0]) / 1e6
    except Exception:
        avail_gb = 64.0
    per_gb = 0.6 + 4.2 * max(sp["records"] for sp in mine) / 1_100_000 + max(sp["length"] for sp in mine) * 3e-9   # measured: 4.1 GB for a chr22-sized build
    workers = max(1, min(cores, int(avail_gb / world / per_gb), len(mine)))
    t0 = time.time()
    running, todo = [], list(mine)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    while todo or running:
        while todo and len(running) < workers:
            sp = todo.pop(0)
            running.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--build-shard", json.dumps(sp), "--samples", str(args.samples), "--fmax", str(args.fmax)], env=env))
        time.sleep(0.2)
        for pr in list(running):
            if pr.poll() is not None:
                running.remove(pr)
                if pr.returncode != 0:
                    raise RuntimeError("building a genome shard failed")
    log(f"[rank {rank}] built {len(mine)} genome shards with {workers} workers in {time.time() - t0:.0f}s")
    return time.time() - t0


def genome_regions(args, plan):
    contigs = list(GENOME)
    lengths = np.array([next(sp["length"] for sp in plan if sp["contig"] == c) for c in contigs], np.float64)
    rng = np.random.default_rng(123)
    n = max(1000, int(args.genome_regions * args.genome_scale))
    ci = rng.choice(len(contigs), n, p=lengths / lengths.sum())
    x = (rng.integers(0, 2**62, n) % (lengths[ci].astype(np.int64) - args.width - 1) + 1).astype(np.uint32)
    y = (x + args.width).astype(np.uint32)
    s = rng.integers(1, args.samples + 1, n).astype(np.uint32)
    return contigs, ci, x, y, s


def _oracle_check(job):
    """Child process: the oracle's t6 / t4 counts and row digests for a handful of regions of one shard."""
    prefix, x, y, s = job
    T = get_oracle()
    o = T.Oracle.open(prefix)
    c6, d6 = o.batch_t6(np.array(x, np.uint64), np.array(y, np.uint64), False)
    c4, d4, ub = o.batch_t4(np.array(x, np.uint64), np.array(y, np.uint64), np.array(s, np.uint32), False)
    o.close()
    return c6.tolist(), d6.tolist(), c4.tolist(), d4.tolist(), ub.tolist()


def genome_block(args, world, steps=3):
    """Rank 0, after every rank built its share: open all shards behind the router, time the routed fused call end to end
    (host regions in, host answers out, routing inside), check a subsample of every contig against the oracle."""
    from concurrent.futures import ProcessPoolExecutor
    from variantstore_b200 import Router
    plan = genome_plan(args)
    os.environ.pop("VSGPU_INDEX_CACHE", None)              # 77 flattened-index caches would not fit the box's disk
    t0 = time.time()
    r = Router([sp["prefix"] for sp in plan], ranges=[(sp["lo"], sp["hi"] if sp["hi"] <= sp["length"] else 0) for sp in plan], ndevices=world)
    open_s = time.time() - t0
    contigs, ci, x, y, s = genome_regions(args, plan)
    cid = r.contig_ids([contigs[i] for i in range(len(contigs))])[ci]
    n = len(x)
    times, stats = [], None
    for it in range(steps + 1):
        t0 = time.perf_counter()
        so, lo, c6, c4, _, _ = r.query_t6t4(cid, x, y, s, csr=False)     # hit codes stay in the shards' page-locked results
        dt = time.perf_counter() - t0
        if it:
            times.append(dt)
        stats = r.stats()
    t0 = time.perf_counter()
    so, lo, c6, c4, off, hits = r.query_t6t4(cid, x, y, s)                # the same with the hit codes gathered into one CSR in region order
    csr_s = time.perf_counter() - t0
    # ---- parity: ~90 regions of one shard per contig (>= 2 000 regions over all contigs) against the oracle, row digests included
    rng = np.random.default_rng(9)
    jobs, picks = [], []
    for c in range(len(contigs)):
        ks = np.unique(so[ci == c])
        k = int(ks[rng.integers(0, len(ks))])
        idx = np.nonzero(so == k)[0]
        idx = idx[rng.choice(len(idx), min(90, len(idx)), replace=False)]
        picks.append((k, idx))
        jobs.append((plan[k]["prefix"], x[idx].tolist(), y[idx].tolist(), s[idx].tolist()))
    bad, checked = 0, 0
    with ProcessPoolExecutor(max_workers=min(len(jobs), max(1, (os.cpu_count() or 2) // 2))) as ex:
        for (k, idx), (oc6, od6, oc4, od4, ub) in zip(picks, ex.map(_oracle_check, jobs)):
            sh = r.shard(k)
            e6 = sh.digest_t6(lo[idx], lo[idx] + c6[idx], False)
            sub_off = np.concatenate([[0], np.cumsum(c4[idx])]).astype(np.uint64)
            sub_hits = np.concatenate([hits[off[i]:off[i + 1]] for i in idx]) if len(idx) else np.zeros(0, np.uint32)
            e4 = sh.digest_t4(sub_off, sub_hits, False)
            ok = (np.array(oc6) == c6[idx]) & (np.array(od6, np.uint64) == e6) & ((np.array(ub) != 0) | ((np.array(oc4) == c4[idx]) & (np.array(od4, np.uint64) == e4)))
            bad += int((~ok).sum()); checked += len(idx)
    dev_bytes = {}
    for k in range(r.num_shards):
        dev_bytes[r.shard_device[k]] = dev_bytes.get(r.shard_device[k], 0) + int(r.shard(k).info.device_bytes)
    t = float(np.median(times))
    out = {"workload": f"{len(GENOME)} contigs in GRCh37 lengths x {args.genome_scale:g}, {sum(sp['records'] for sp in plan)} records x {args.samples} samples in {len(plan)} position-range shards "
                       f"(<= {args.genome_shard_records} records each), {n} random {args.width} bp regions drawn in proportion to contig length, unsorted, one sample each; t6 + t4 per region",
           "n_gpus": world, "shards": len(plan), "records": int(sum(sp["records"] for sp in plan)), "regions": n,
           "e2e_regions_per_s": 2 * n / t, "e2e_ms": 1000 * t, "e2e_ms_with_csr_gather": 1000 * csr_s, "route_ms": stats["route_ms"], "scatter_ms": stats["scatter_ms"],
           "per_gpu_ms": stats["device_ms"], "per_gpu_regions": stats["device_regions"],
           "imbalance_regions_max_over_mean": float(max(stats["device_regions"]) / (np.mean(stats["device_regions"]) or 1)),
           "device_bytes_per_gpu": [dev_bytes.get(d, 0) for d in sorted(dev_bytes)], "t4_rows": int(off[-1]), "open_all_shards_s": open_s,
           "router": "C++ (csrc/router.cc): routing + host threads per GPU + scatter inside the timed region; one process drives all GPUs",
           "parity": {"regions": checked, "contigs": len(contigs), "mismatches": bad, "status": "ok" if bad == 0 and checked >= 2000 * min(1.0, args.genome_scale * 10) else "FAILED",
                      "against": "oracle (t6 / t4 row counts and row digests of ~90 regions of one shard per contig)"}}
    r.close()
    return out


def run_vsgpu(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        log("bench.py: no CUDA device — libvsgpu has no CPU path")
        return 2
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from variantstore_b200 import Batch, VariantStoreIndex

    prefix, meta = ensure_index(args, rank)
    # vsgpu_open twice: decoding ser/ (and writing the flattened-index cache), then from that cache
    os.environ.setdefault("VSGPU_INDEX_CACHE", os.path.join(args.cache_dir, "flat"))
    os.makedirs(os.environ["VSGPU_INDEX_CACHE"], exist_ok=True) if os.environ["VSGPU_INDEX_CACHE"] not in ("0", "1") else None
    t0 = time.time()
    idx = VariantStoreIndex(prefix, device=local)
    open_s = time.time() - t0
    open_cached_s = None
    if idx.info.from_cache:
        open_s, open_cached_s = None, open_s
    else:
        idx.close()
        t0 = time.time()
        idx = VariantStoreIndex(prefix, device=local)
        open_cached_s = time.time() - t0 if idx.info.from_cache else None
    idx.set_stream(torch.cuda.current_stream().cuda_stream)
    x, y, s = make_regions(args, meta, rank)
    n = len(x)
    b6, b4 = Batch(idx, 6, x, y), Batch(idx, 4, x, y, sample_ids=s)
    b46 = None if args.unfused else Batch(idx, 46, x, y, sample_ids=s)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        if b46 is not None:
            b46.run()
        else:
            b6.run()
            b4.run()

    def timed(fn, batches, steps):
        ms, per = [], [[] for _ in batches]
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
            for k, b in enumerate(batches):
                per[k].append(b.timings_ms()[0])
        return ms, [float(np.mean(p)) for p in per]

    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        step()
        b6.run(); b4.run()
    torch.cuda.synchronize()
    algo6, launches6 = b6.stats()
    algo4, launches4 = b4.stats()
    algo46, launches46 = b46.stats() if b46 is not None else (algo6 + algo4, launches6 + launches4)
    # fused and unfused answers are the same arrays (checked here once, outside the timed region)
    if b46 is not None and rank == 0:
        lo_f, hi_f, cnt_f, off_f, hits_f = b46.fetch()
        lo_u, hi_u, cnt_u = b6.fetch()
        off_u, hits_u, _ = b4.fetch()
        assert np.array_equal(lo_f, lo_u) and np.array_equal(hi_f, hi_u) and np.array_equal(cnt_f, cnt_u) and np.array_equal(off_f, off_u) and np.array_equal(hits_f, hits_u), "fused launch disagrees with k_t6 + k_t4p"

    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    step_ms, kms = timed(step, [b46] if b46 is not None else [b6, b4], args.steps)
    barrier()
    # (the clock sampler keeps running through the by-kernel and end-to-end loops below: they are measured regions too)
    # the unfused kernels on their own, for by_kernel (not part of `value` unless --unfused)
    _, (t6_ms,) = timed(b6.run, [b6], max(3, args.steps // 2))
    _, (t4_ms,) = timed(b4.run, [b4], max(3, args.steps // 2))
    fused_ms = kms[0] if b46 is not None else None
    total_ms = float(np.sum(step_ms))
    if dist is not None:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = 2 * n * world / (ms_per_step / 1000)

    # ---- end to end through the C ABI with host buffers (pinned inputs), H2D + kernels + D2H each step
    lib, h = idx._lib, idx._h
    # region bounds go in as 32-bit arrays (vsgpu_query_t6_u32 / _t4_u32: the reference parses them with
    # std::stoi, commands.cc:76-80), which halves the bytes over PCIe; --e2e-u64 uses the 64-bit entry points
    cdt = np.int64 if args.e2e_u64 else np.int32
    px = torch.from_numpy(x.astype(cdt)).pin_memory()
    py = torch.from_numpy(y.astype(cdt)).pin_memory()
    ps = torch.from_numpy(s.astype(np.int32)).pin_memory()
    px64, py64 = (px, py) if args.e2e_u64 else (torch.from_numpy(x.astype(np.int64)).pin_memory(), torch.from_numpy(y.astype(np.int64)).pin_memory())
    q6 = lib.vsgpu_query_t6 if args.e2e_u64 else lib.vsgpu_query_t6_u32
    q4 = lib.vsgpu_query_t4 if args.e2e_u64 else lib.vsgpu_query_t4_u32
    q64 = lib.vsgpu_query_t6t4 if args.e2e_u64 else lib.vsgpu_query_t6t4_u32
    plo, phi, pcnt = (torch.zeros(n, dtype=torch.int32).pin_memory() for _ in range(3))      # page-locked result arrays
    vp = C.c_void_p

    def e2e_step():
        r = vp()
        if args.unfused:
            rc = q6(h, n, vp(px.data_ptr()), vp(py.data_ptr()), vp(plo.data_ptr()), vp(phi.data_ptr()), vp(pcnt.data_ptr()))
            assert rc == 0, lib.vsgpu_last_error()
            rc = q4(h, n, vp(px.data_ptr()), vp(py.data_ptr()), vp(ps.data_ptr()), C.byref(r))
        else:   # one fused call: x / y / samples up once; rec_lo, the two row counts and the hit codes back
            rc = q64(h, n, vp(px.data_ptr()), vp(py.data_ptr()), vp(ps.data_ptr()), vp(plo.data_ptr()), None, vp(pcnt.data_ptr()), C.byref(r))
        assert rc == 0, lib.vsgpu_last_error()
        total = int(lib.vsgpu_result_total(r))
        lib.vsgpu_result_free(r)
        return total

    for _ in range(2):
        hits_total = e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        hits_total = e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    sampler.stop_flag = True
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = 2 * n * world / e2e_s
    cb = 8 if args.e2e_u64 else 4
    if args.unfused:
        h2d = n * 2 * cb + n * (2 * cb + 4)              # t6 x, y; t4 x, y, sample ids
        d2h = n * 12 + n * 4 + hits_total * 4            # t6 lo, hi, counts; t4 counts + hit codes
    else:
        h2d = n * (2 * cb + 4)                           # x, y, sample ids, once
        d2h = n * 8 + n * 4 + hits_total * 4             # t6 lo + counts; t4 counts + hit codes
    # The roof over e2e on this box: every rank moves exactly these bytes (page-locked, H2D and D2H at the same time on two
    # streams, no kernels), all ranks together — what the step would cost if the GPU work and every call overhead were free.
    hb, db = torch.empty(h2d, dtype=torch.uint8).pin_memory(), torch.empty(d2h, dtype=torch.uint8).pin_memory()
    gh, gd = torch.empty(h2d, dtype=torch.uint8, device="cuda"), torch.empty(d2h, dtype=torch.uint8, device="cuda")
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    def copies():
        with torch.cuda.stream(sa):
            gh.copy_(hb, non_blocking=True)
        with torch.cuda.stream(sb):
            db.copy_(gd, non_blocking=True)
    for _ in range(3):
        copies()
    barrier()
    t0 = time.perf_counter()
    for _ in range(10):
        copies()
    torch.cuda.synchronize()
    ceil_s = (time.perf_counter() - t0) / 10
    if dist is not None:
        t = torch.tensor([ceil_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ceil_s = float(t.item())
    pcie_ceiling = {"regions_per_s": 2 * n * world / ceil_s, "ms_per_step": 1000 * ceil_s, "aggregate_GBps": (h2d + d2h) * world / ceil_s / 1e9,
                    "what": "the step's H2D + D2H bytes copied concurrently by all ranks from / to page-locked memory, nothing else"}
    del hb, db, gh, gd

    genome = None
    if (world >= 2 or args.genome) and not args.no_genome:
        # every rank builds its share of the genome shards on the host, then rank 0 alone drives all GPUs through the
        # router while the others wait on a CPU (gloo) barrier — the engine's multi-GPU front-end is one process
        for b in (b6, b4, b46):
            if b is not None:
                b.close()
        cpu_group = dist.new_group(backend="gloo") if dist is not None else None
        try:
            build_s = genome_build(args, genome_plan(args), rank, world)
            if cpu_group is not None:
                dist.barrier(group=cpu_group)
            if rank == 0:
                genome = genome_block(args, world)
                genome["build_shards_s_rank0"] = build_s
        except Exception as ex:
            genome = {"failed": repr(ex)[:400]}
        if cpu_group is not None:
            dist.barrier(group=cpu_group)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    dom_ms = fused_ms if fused_ms is not None else t4_ms
    dom_algo = algo46 if fused_ms is not None else algo4
    dom_name = T4_FUSED_INSTANCE if fused_ms is not None else T4_INSTANCE
    achieved_conv = dom_algo / (dom_ms / 1000) / 1e9
    # DRAM bytes of one launch of exactly this kernel instance on exactly this workload, from the committed ncu capture
    # (profiles/traffic.json names the instance, the region count and the capture); null when they do not match this run
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath)).get(dom_name)
            if tj and tj.get("regions") == n and tj.get("width") == args.width and tj.get("records") == args.records:
                traffic, traffic_src = int(tj["dram_bytes_per_launch"]), tj.get("capture")
        except Exception:
            traffic = None
    achieved = (traffic / (dom_ms / 1000) / 1e9) if traffic else None
    line = {
        "metric": "region queries/s (types 4/6)", "value": value, "unit": "regions/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic", "config": workload_config(args, meta),
        "by_kernel": {"fused_t6t4_ms": fused_ms, "k_t6_ms": t6_ms, "k_t4_ms": t4_ms, "t6_regions_per_s": n / (t6_ms / 1000), "t4_regions_per_s": n / (t4_ms / 1000),
                      "fused_regions_per_s": (n / (fused_ms / 1000)) if fused_ms else None, "t6_algorithmic_GBps": algo6 / (t6_ms / 1000) / 1e9,
                      "t4_hits_per_region": hits_total / n, "index_open_s": open_s, "index_open_cached_s": open_cached_s, "device_bytes": int(idx.info.device_bytes)},
        # frac = DRAM bytes the kernel really moved (ncu dram__bytes_read + write of this instance on this workload) / its live
        # CUDA-event duration / the measured HBM peak.  The SURVEY section 8(d) per-record convention is kept beside it: the
        # hit map reads one bit where that convention charges 20 bytes, so it can exceed 1 and is not a roofline fraction.
        "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                     "traffic": traffic, "traffic_source": traffic_src, "kernel_ms": dom_ms, "peak_source": peak_src,
                     "algorithmic_convention": {"bytes_per_launch": dom_algo, "GBps": achieved_conv, "frac": achieved_conv / peak,
                                                "note": "SURVEY 8(d): 288 (t6) + 292 + 20 v + 4 h (t4) per region; counts a 20-byte record read per scanned record, which the hit map replaces by one bit"}},
        "e2e": {"value": e2e_val, "unit": "regions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "pcie_ceiling": pcie_ceiling, "frac_of_pcie_ceiling": e2e_val / pcie_ceiling["regions_per_s"], "call": "vsgpu_query_t6 + vsgpu_query_t4" if args.unfused else "vsgpu_query_t6t4"},
        "gpu_launches": (launches46 if b46 is not None else launches6 + launches4) * args.steps,
        "clocks": sampler.summary(),
    }
    if genome is not None:
        line["genome"] = genome
    if world == 1 and not args.no_other_ops:
        # The same step with the answers as the text `-v` writes (print_var rows with every carrier's name and genotype, as the
        # reference's operators build them): t6 rows and t4 rows rendered on the device, host regions in -> host text out.  A
        # bounded slice of the regions: one region's t6 rows are ~46 KB of text, so this leg is the PCIe rate of the text.
        try:
            m = min(n, args.rows_regions)
            sl = slice(0, n, max(1, n // m))
            rx, ry, rs = (np.ascontiguousarray(a[sl][:m]) for a in (x, y, s))
            # through the C ABI: the text arrives in the library's page-locked buffer (vsgpu_text_bytes), from where a caller writes
            # it out; no copy into a Python object inside the timed region
            rx64, ry64, rs32 = rx.astype(np.uint64), ry.astype(np.uint64), rs.astype(np.uint32)
            vt, nbytes, rows6, rows4, ms6, ms4 = [], 0, 0, 0, 0.0, 0.0
            for rep in range(4):
                t6h, t4h = vp(), vp()
                t0 = time.perf_counter()
                rc = lib.vsgpu_render_t6(h, len(rx64), vp(rx64.ctypes.data), vp(ry64.ctypes.data), 1, C.byref(t6h))
                assert rc == 0, lib.vsgpu_last_error()
                rc = lib.vsgpu_render_t4(h, len(rx64), vp(rx64.ctypes.data), vp(ry64.ctypes.data), vp(rs32.ctypes.data), 1, C.byref(t4h))
                assert rc == 0, lib.vsgpu_last_error()
                dt = time.perf_counter() - t0
                if rep:                                                    # the first call sizes the page-locked pool
                    vt.append(dt)
                nbytes = int(lib.vsgpu_text_offsets(t6h)[len(rx64)]) + int(lib.vsgpu_text_offsets(t4h)[len(rx64)])
                rows6, rows4 = int(lib.vsgpu_text_num_rows(t6h)), int(lib.vsgpu_text_num_rows(t4h))
                ms6, ms4 = float(lib.vsgpu_text_kernel_ms(t6h)), float(lib.vsgpu_text_kernel_ms(t4h))
                lib.vsgpu_text_free(t6h); lib.vsgpu_text_free(t4h)
            line["e2e_rows"] = {"value": 2 * len(rx) / float(np.median(vt)), "unit": "regions/s", "regions": int(len(rx)), "t6_rows": rows6, "t4_rows": rows4,
                                "text_bytes": nbytes, "text_GBps": nbytes / float(np.median(vt)) / 1e9, "render_kernels_ms": ms6 + ms4,
                                "call": "vsgpu_render_t6 + vsgpu_render_t4 (rows as print_var text with carrier lists, query.h:43-50)"}
        except Exception as ex:
            line["e2e_rows"] = {"failed": repr(ex)[:300]}
    if world == 1 and not args.no_other_ops:
        # Reported beside the headline, never part of it: the widened operator t2 (a sample's sequence over the same
        # regions) through its C-ABI call — device time of its kernels and end to end.
        try:
            ms, ts, nbytes = [], [], 0
            for rep in range(4):
                t = vp()
                t0 = time.perf_counter()
                rc = lib.vsgpu_query_t2(h, n, vp(px64.data_ptr()), vp(py64.data_ptr()), vp(ps.data_ptr()), C.byref(t))
                dt = time.perf_counter() - t0
                assert rc == 0, lib.vsgpu_last_error()
                if rep:
                    ts.append(dt)
                    ms.append(float(lib.vsgpu_text_kernel_ms(t)))
                nbytes = int(lib.vsgpu_text_offsets(t)[n])
                lib.vsgpu_text_free(t)
            line["other_ops"] = {"t2_query_sample_from_ref": {"regions_per_s_kernels": n / (float(np.median(ms)) / 1e3), "kernels_ms": float(np.median(ms)),
                                                              "regions_per_s_e2e": n / float(np.median(ts)), "sequence_bytes": nbytes,
                                                              "copy_algorithmic_bytes": 2 * nbytes}}
        except Exception as ex:
            line["other_ops"] = {"failed": str(ex)}
    if world == 1 and not args.no_other_ops and not args.no_sparse:
        # BASELINE config [3] beside the headline: a TCGA-like sparse cohort (explicit sample ids, 1-3 carriers per allele) — batched
        # t7 lookups (device time, and end to end incl. hashing the query strings on the host) and t4 over 1 M regions, the
        # reference's own worst case (eval_data_records/logs/query_luad.out:120: 3 319 s for 1 000 regions).
        try:
            T = get_oracle()
            sp = os.path.join(args.cache_dir, f"tcga_{args.sparse_records}_{args.sparse_samples}", "ser")
            if not os.path.exists(os.path.join(sp, ".done")):
                os.makedirs(os.path.dirname(sp), exist_ok=True)
                so = T.Oracle.synth(sp, chr_name="2", ref_length=243_199_373, pos_lo=10_000, n_records=args.sparse_records, n_samples=args.sparse_samples,
                                    mode=1, seed=77, cqf_log2=25, gzip_level=1)
                so.close()
                open(os.path.join(sp, ".done"), "w").write("ok")
            so = T.Oracle.open(sp)
            av = so.all_variants()
            with VariantStoreIndex(sp, device=local) as sidx:
                sidx.set_stream(torch.cuda.current_stream().cuda_stream)
                rng = np.random.default_rng(5)
                m7 = args.sparse_lookups
                pick = rng.integers(0, len(av), m7)
                pos7 = np.array([av[i][0] for i in pick], np.uint64)
                pos7[m7 // 2:] += 1                           # half of the lookups miss by construction
                refs7, alts7 = [av[i][1] for i in pick], [av[i][2] for i in pick]
                b7 = Batch(sidx, 7, pos7, refs=refs7, alts=alts7)
                b7.run()
                _, (t7_ms,) = timed(b7.run, [b7], 5)
                # end to end through the C ABI: C strings in host memory -> host hashing -> H2D -> k_t7 -> D2H of the record ids
                ra7 = (C.c_char_p * m7)(*[r.encode() for r in refs7])
                aa7 = (C.c_char_p * m7)(*[a.encode() for a in alts7])
                rec7 = np.zeros(m7, np.uint32)
                t7_e2e = 1e9
                for _ in range(3):
                    t0 = time.perf_counter()
                    rc = sidx._lib.vsgpu_query_t7(sidx._h, m7, C.c_void_p(pos7.ctypes.data), ra7, aa7, C.c_void_p(rec7.ctypes.data))
                    t7_e2e = min(t7_e2e, time.perf_counter() - t0)
                    assert rc == 0, sidx._lib.vsgpu_last_error()
                sub = rng.choice(m7, 2000, replace=False)
                f7, _, _ = so.batch_t7(pos7[sub], [refs7[i] for i in sub], [alts7[i] for i in sub])
                ok7 = bool(np.array_equal(rec7[sub] != 0xFFFFFFFF, f7 == 1))
                x4 = np.sort(rng.integers(10_000, 243_199_373 - 1000, n)).astype(np.uint64)
                y4 = x4 + np.uint64(args.width)
                s4 = rng.integers(1, args.sparse_samples + 1, n).astype(np.uint32)
                b4s = Batch(sidx, 4, x4, y4, sample_ids=s4)
                b4s.run()
                _, (t4s_ms,) = timed(b4s.run, [b4s], 5)
                off4, hits4, _ = b4s.fetch()
                sub = rng.choice(n, 1000, replace=False)
                oc4, od4, ub4 = so.batch_t4(x4[sub], y4[sub], s4[sub], False)
                ok4 = bool(np.all(((oc4 == np.diff(off4)[sub]) & (od4 == sidx.digest_t4(off4, hits4, False)[sub])) | (ub4 != 0)))
                px4, py4, ps4 = (torch.from_numpy(a.astype(np.int32)).pin_memory() for a in (x4, y4, s4))
                t4s_e2e = 1e9
                for _ in range(4):
                    r4 = C.c_void_p()
                    t0 = time.perf_counter()
                    rc = sidx._lib.vsgpu_query_t4_u32(sidx._h, n, C.c_void_p(px4.data_ptr()), C.c_void_p(py4.data_ptr()), C.c_void_p(ps4.data_ptr()), C.byref(r4))
                    t4s_e2e = min(t4s_e2e, time.perf_counter() - t0)
                    assert rc == 0, sidx._lib.vsgpu_last_error()
                    sidx._lib.vsgpu_result_free(r4)
                line.setdefault("other_ops", {})["sparse_cohort"] = {
                    "workload": f"TCGA-like: {args.sparse_records} records x {args.sparse_samples} samples, explicit sample ids; membership = per-sample carried-entry lists",
                    "t7_lookups": m7, "k_t7_ms": t7_ms, "t7_lookups_per_s_kernel": m7 / (t7_ms / 1e3), "t7_lookups_per_s_e2e_c_abi_incl_host_hashing": m7 / t7_e2e,
                    "t7_found_fraction": float((rec7 != 0xFFFFFFFF).mean()), "t7_parity_sample_ok": ok7,
                    "t4_regions": n, "k_t4_ms": t4s_ms, "t4_regions_per_s_kernel": n / (t4s_ms / 1e3), "t4_regions_per_s_e2e": n / t4s_e2e, "t4_parity_sample_ok": ok4,
                    "device_bytes": int(sidx.info.device_bytes)}
                b7.close(); b4s.close()
            so.close()
        except Exception as ex:
            line.setdefault("other_ops", {})["sparse_cohort"] = {"failed": repr(ex)[:300]}
    if world == 1 and not args.no_cpu_baseline:
        try:
            nq, times, load_s = cpu_sample(args, prefix, meta, 1, args.cpu_sample_single)
            line["cpu_baseline"] = {"value": nq / times[0], "unit": "regions/s", "cores": 1, "kind": "port",
                                    "sample": f"{nq // 2} of the {n} regions per type (t6 + t4), single thread like the reference's query loop; oracle port of include/query.h; index load {load_s:.1f}s not timed"}
        except Exception as ex:           # the baseline is reported, never required
            line["cpu_baseline"] = {"value": None, "unit": "regions/s", "cores": 1, "kind": "port", "sample": f"failed: {ex}"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="vsgpu", choices=["vsgpu", "reference"])
    ap.add_argument("--regions", type=int, default=1_000_000)
    ap.add_argument("--width", type=int, default=1000)
    ap.add_argument("--records", type=int, default=1_103_547)
    ap.add_argument("--samples", type=int, default=2504)
    ap.add_argument("--fmax", type=int, default=1100)
    ap.add_argument("--cpu-sample", type=int, default=20_000, help="regions per type per step of the reference arm")
    ap.add_argument("--cpu-sample-single", type=int, default=2_000, help="regions per type of the single-thread cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-ops", action="store_true", help="skip the side measurements (t2, rows as text)")
    ap.add_argument("--no-sparse", action="store_true", help="skip the sparse-cohort side measurement (config [3])")
    ap.add_argument("--sparse-records", type=int, default=3_000_000)
    ap.add_argument("--sparse-samples", type=int, default=10_000)
    ap.add_argument("--sparse-lookups", type=int, default=10_000_000)
    ap.add_argument("--rows-regions", type=int, default=20_000, help="regions of the e2e_rows leg (answers as -v text)")
    ap.add_argument("--unfused", action="store_true", help="a step = k_t6 then k_t4p (two launches, two host-buffer calls) instead of the fused launch / call")
    ap.add_argument("--e2e-u64", action="store_true", help="end-to-end arm through the 64-bit coordinate entry points instead of the 32-bit ones")
    ap.add_argument("--cache-dir", default=os.environ.get("VSGPU_BENCH_CACHE", "/tmp/vsgpu_bench"))
    ap.add_argument("--genome", action="store_true", help="run the whole-genome block (config [2]) also at N = 1")
    ap.add_argument("--no-genome", action="store_true", help="skip the whole-genome block at N >= 2")
    ap.add_argument("--genome-scale", type=float, default=1.0, help="fraction of the 84.8 M records / 10 M regions / contig lengths")
    ap.add_argument("--genome-regions", type=int, default=10_000_000)
    ap.add_argument("--genome-shard-records", type=int, default=1_200_000)
    ap.add_argument("--build-shard", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.build_shard:
        build_shard(json.loads(args.build_shard), args.samples, args.fmax)
        return
    import __graft_entry__
    if not os.path.exists(os.path.join(ROOT, "variantstore_b200", "libvsgpu.so")) or not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        __graft_entry__.build()
    sys.exit(run_reference(args) if args.impl == "reference" else run_vsgpu(args))


if __name__ == "__main__":
    main()
