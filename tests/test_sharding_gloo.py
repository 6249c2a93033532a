"""The N > 1 path on CPU: two processes (gloo), each owning some contigs (host simulator backend),
regions routed by rank 0, answers gathered — must equal the single-process answers and the oracle."""
import os
import sys

import numpy as np
import pytest

import vs_testlib as T
from vs_testlib import Oracle


def test_lpt_assignment_and_routing():
    from variantstore_b200.sharding import assign_contigs, route
    sizes = {"1": 65, "2": 70, "3": 58, "21": 11, "22": 11, "X": 34, "Y": 1}
    own = assign_contigs(sizes, 2)
    loads = [sum(sizes[c] for c in own if own[c] == g) for g in range(2)]
    assert set(own.values()) == {0, 1} and abs(loads[0] - loads[1]) <= 11
    assert own == assign_contigs(sizes, 2)                       # deterministic
    parts = route(["1", "22", "2", "1", "Y"], own, 2)
    assert sorted(np.concatenate(parts).tolist()) == [0, 1, 2, 3, 4]
    for g, p in enumerate(parts):
        assert all(own[["1", "22", "2", "1", "Y"][i]] == g for i in p)


def _worker(rank, world, port, prefixes, owner, queries, out_path):
    import torch.distributed as dist
    sys.path.insert(0, T.ROOT)
    sys.path.insert(0, os.path.join(T.ROOT, "tests"))
    from variantstore_b200 import VariantStoreIndex, load_library
    from variantstore_b200.sharding import ShardedIndex, distributed_sample_seq, distributed_var_in_ref
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lib = load_library(T.HOSTSIM_SO, subset=True)
    mine = {c: p for c, p in prefixes.items() if owner[c] == rank}
    sh = ShardedIndex(mine, lambda p: VariantStoreIndex(p, lib=lib))
    contigs, x, y, snames = queries
    res = distributed_var_in_ref(dist, sh, owner, contigs if rank == 0 else None, x if rank == 0 else None, y if rank == 0 else None)
    seqs = distributed_sample_seq(dist, sh, owner, contigs if rank == 0 else None, x if rank == 0 else None, y if rank == 0 else None,
                                  snames if rank == 0 else None)
    if rank == 0:
        np.save(out_path, res)
        import pickle
        pickle.dump(seqs, open(out_path + ".t2", "wb"))
    dist.barrier()
    sh.close()
    dist.destroy_process_group()


def test_two_process_sharded_queries(tmp_path):
    import torch.multiprocessing as mp
    from variantstore_b200.sharding import assign_contigs
    names = ["cA", "cB", "cC"]
    prefixes, oracles, sizes = {}, {}, {}
    for i, c in enumerate(names):
        fa, vcf, _ = T.write_fuzz_inputs(str(tmp_path / c), 40 + i, n_records=120 + 60 * i, chrom=c)
        oracles[c] = Oracle.construct(fa, vcf, str(tmp_path / c / "ser"))
        prefixes[c] = str(tmp_path / c / "ser")
        sizes[c] = 120 + 60 * i
    owner = assign_contigs(sizes, 2)
    rng = np.random.default_rng(0)
    n = 600
    contigs = [names[i] for i in rng.integers(0, 3, n)]
    x = rng.integers(1, 3900, n).astype(np.uint64)
    y = x + rng.choice([1, 5, 100, 1000], n).astype(np.uint64)
    snames = [f"S{int(i):03d}" for i in rng.integers(1, 13, n)]
    port = 29500 + os.getpid() % 400
    out_path = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(2, port, prefixes, owner, (contigs, x, y, snames), out_path), nprocs=2, join=True)
    got = np.load(out_path)
    want = np.zeros(n, np.uint64)
    for c in names:
        idx = np.nonzero(np.array(contigs) == c)[0]
        want[idx] = oracles[c].batch_t6(x[idx], y[idx])[0]
    assert np.array_equal(got.astype(np.uint64), want)
    # the routed t2 answers (sample sequences) equal the oracle's, contig by contig
    import pickle
    seqs, status = pickle.load(open(out_path + ".t2", "rb"))
    for c in names:
        idx = np.nonzero(np.array(contigs) == c)[0]
        sid = np.array([int(snames[i][1:]) for i in idx], np.uint32)
        ln, dg, st, ub, want_seqs = oracles[c].batch_t2(x[idx], y[idx], sid, want_text=True)
        for j, i in enumerate(idx):
            if not ub[j]:
                assert status[i] == st[j] and seqs[i] == want_seqs[j].encode(), (c, int(x[i]), int(y[i]), snames[i])
