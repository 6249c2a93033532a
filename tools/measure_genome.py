#!/usr/bin/env python
"""BASELINE.json config [2] — a 1000-Genomes-shaped genome (24 contigs, GRCh37 proportions, scaled
by --scale so that it can be built on the box), contigs assigned to the GPUs by longest processing
time, regions routed by the host to the GPU owning their contig, t6 + t4 per step.  Launch with
torchrun (one rank per GPU) or plain python for one GPU.  Prints one JSON line (rank 0)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# contig -> (GRCh37 length, 1000 Genomes phase-3 records); record counts as in eval_data_records/logs/vs_v1.log (sum 84.8 M)
GENOME = {"1": (249_250_621, 6_468_094), "2": (243_199_373, 7_081_600), "3": (198_022_430, 5_832_276), "4": (191_154_276, 5_732_585),
          "5": (180_915_260, 5_265_763), "6": (171_115_067, 5_024_119), "7": (159_138_663, 4_716_715), "8": (146_364_022, 4_597_105),
          "9": (141_213_431, 3_560_687), "10": (135_534_747, 3_992_219), "11": (135_006_516, 4_045_628), "12": (133_851_895, 3_868_428),
          "13": (115_169_878, 2_857_916), "14": (107_349_540, 2_655_067), "15": (102_531_392, 2_424_689), "16": (90_354_753, 2_697_949),
          "17": (81_195_210, 2_329_288), "18": (78_077_248, 2_267_185), "19": (59_128_983, 1_832_506), "20": (63_025_520, 1_812_841),
          "21": (48_129_895, 1_105_538), "22": (51_304_566, 1_103_547), "X": (155_270_560, 3_468_093), "Y": (59_373_566, 62_042)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.1)
    ap.add_argument("--regions", type=int, default=8_000_000)
    ap.add_argument("--samples", type=int, default=2504)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    import torch
    import vs_testlib as T
    from variantstore_b200 import Batch, VariantStoreIndex
    from variantstore_b200.sharding import assign_contigs, route
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sizes = {c: max(2000, int(r * args.scale)) for c, (l, r) in GENOME.items()}
    lengths = {c: max(300_000, int(l * args.scale)) for c, (l, r) in GENOME.items()}
    owner = assign_contigs(sizes, world)
    # the host's region list: contig drawn in proportion to length, 1 kb wide, one random sample each
    rng = np.random.default_rng(123)
    names = list(GENOME)
    p = np.array([lengths[c] for c in names], float)
    p /= p.sum()
    contig_of = rng.choice(len(names), args.regions, p=p)
    parts = route([names[i] for i in contig_of], owner, world)          # what the router hands to every GPU
    mine = parts[rank]
    t0 = time.time()
    shards, batches, n_mine = {}, [], 0
    for ci, c in enumerate(names):
        if owner[c] != rank:
            continue
        prefix = f"/tmp/vsgpu_bench/genome_{args.scale}/{c}"
        os.makedirs(os.path.dirname(prefix), exist_ok=True)
        if not os.path.exists(prefix + "/.done"):
            o = T.Oracle.synth(prefix, chr_name=c, ref_length=lengths[c], pos_lo=1000, n_records=sizes[c], n_samples=args.samples,
                               fmax=1100, seed=500 + ci, cqf_log2=25, gzip_level=1)
            o.close()
            open(prefix + "/.done", "w").write("ok")
        idx = VariantStoreIndex(prefix, device=local)
        idx.set_stream(torch.cuda.current_stream().cuda_stream)
        sel = mine[contig_of[mine] == ci]
        r2 = np.random.default_rng(1000 + ci)
        x = np.sort(r2.integers(1000, lengths[c] - 2000, len(sel))).astype(np.uint64)
        y = x + np.uint64(1000)
        s = r2.integers(1, args.samples + 1, len(sel)).astype(np.uint32)
        if len(sel):
            batches += [Batch(idx, 6, x, y), Batch(idx, 4, x, y, sample_ids=s)]
        shards[c] = idx
        n_mine += len(sel)
    build_s = time.time() - t0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        for b in batches:
            b.run()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ms = []
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for b in batches:
            b.run()
        e1.record()
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    total_ms = float(np.sum(ms))
    loads = [n_mine, sum(sizes[c] for c in names if owner[c] == rank), float(np.mean(ms)), build_s,
             float(sum(int(sh.info.device_bytes) for sh in shards.values()))]
    if dist is not None:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        allloads = [None] * world
        dist.all_gather_object(allloads, loads)
    else:
        allloads = [loads]
    if rank == 0:
        step_ms = total_ms / args.steps
        print(json.dumps({"config": "genome-shaped, contig-sharded", "scale": args.scale, "n_gpus": world, "contigs": len(names),
                          "records_total": int(sum(sizes.values())), "regions_total": args.regions, "ms_per_step": step_ms,
                          "region_queries_per_s": 2 * args.regions / (step_ms / 1000),
                          "per_gpu": [{"regions": int(l[0]), "records": int(l[1]), "ms_per_step": l[2], "build_s": l[3], "device_bytes": int(l[4])} for l in allloads],
                          "contig_owner": owner}), flush=True)
    for sh in shards.values():
        sh.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
