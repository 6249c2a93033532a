"""Multi-GPU layout of the region path: whole contigs are assigned to GPUs, the host routes every
region to the GPU that owns its contig, results come back by plain device->host copies.  No
collective runs on the data path — the reference itself keeps one index per contig and built /
queried them as independent processes (eval_data_records/evaluation.txt:34, util.cc:93-96).

torch.distributed is only plumbing here (scatter of the routed regions, gather of the answers,
barrier + max-over-ranks for timing)."""
from typing import Dict, List, Sequence, Tuple

import numpy as np


def assign_contigs(sizes: Dict[str, int], n_gpus: int) -> Dict[str, int]:
    """Longest-processing-time assignment of contigs (weight = record count) to GPUs."""
    load = [0] * n_gpus
    out = {}
    for name, w in sorted(sizes.items(), key=lambda kv: (-kv[1], kv[0])):
        g = min(range(n_gpus), key=lambda i: (load[i], i))
        out[name] = g
        load[g] += w
    return out


def route(contigs: Sequence[str], owner: Dict[str, int], n_gpus: int) -> List[np.ndarray]:
    """Indices of the regions each GPU has to answer, grouped by GPU, original order kept."""
    ranks = np.fromiter((owner[c] for c in contigs), dtype=np.int64, count=len(contigs))
    return [np.nonzero(ranks == g)[0] for g in range(n_gpus)]


class ShardedIndex:
    """The contigs one process (one GPU) owns.  `open_fn(prefix)` returns a VariantStoreIndex."""

    def __init__(self, prefixes: Dict[str, str], open_fn):
        self.shards = {name: open_fn(p) for name, p in prefixes.items()}

    def _by_contig(self, contigs):
        contigs = np.asarray(contigs)
        for name in self.shards:
            idx = np.nonzero(contigs == name)[0]
            if len(idx):
                yield name, idx

    def var_in_ref(self, contigs, x, y) -> np.ndarray:
        """t6 counts for this process's regions (one batch per contig)."""
        x, y = np.asarray(x, np.uint64), np.asarray(y, np.uint64)
        out = np.zeros(len(x), np.uint32)
        for name, idx in self._by_contig(contigs):
            out[idx] = self.shards[name].batch_var_in_ref(x[idx], y[idx])[2]
        return out

    def sample_var_in_ref(self, contigs, x, y, sample_names) -> Tuple[np.ndarray, List[str]]:
        """t4 counts + row text per region for this process's regions."""
        x, y = np.asarray(x, np.uint64), np.asarray(y, np.uint64)
        counts = np.zeros(len(x), np.uint32)
        texts = [""] * len(x)
        for name, idx in self._by_contig(contigs):
            sh = self.shards[name]
            sid = np.array([sh.sample_id(sample_names[i]) for i in idx], np.uint32)
            off, hits = sh.batch_sample_var_in_ref(x[idx], y[idx], sid)
            counts[idx] = np.diff(off).astype(np.uint32)
            for j, i in enumerate(idx):
                texts[i] = sh.rows_t4_text(hits[off[j]:off[j + 1]])
        return counts, texts

    def sample_seq(self, contigs, x, y, sample_names, own_coordinates=False) -> Tuple[List[bytes], np.ndarray]:
        """t2 (or, with own_coordinates, t3) for this process's regions: the sequences and the status bytes."""
        x, y = np.asarray(x, np.uint64), np.asarray(y, np.uint64)
        seqs = [b""] * len(x)
        status = np.zeros(len(x), np.uint8)
        for name, idx in self._by_contig(contigs):
            sh = self.shards[name]
            sid = np.array([sh.sample_id(sample_names[i]) for i in idx], np.uint32)
            fn = sh.batch_sample_seq_in_sample if own_coordinates else sh.batch_sample_seq_in_ref
            off, text, st, _ = fn(x[idx], y[idx], sid)
            status[idx] = st
            for j, i in enumerate(idx):
                seqs[i] = text[off[j]:off[j + 1]]
        return seqs, status

    def close(self):
        for s in self.shards.values():
            s.close()


def distributed_sample_seq(dist, sharded: ShardedIndex, owner: Dict[str, int], contigs, x, y, sample_names, own_coordinates=False):
    """t2 / t3 over a routed region list: rank 0 scatters (contig, x, y, sample) by owner, every rank answers
    its part from its own shards, rank 0 gathers sequences + status bytes and restores the order."""
    rank, world = dist.get_rank(), dist.get_world_size()
    if rank == 0:
        parts = route(contigs, owner, world)
        payload = [([contigs[i] for i in p], np.asarray(x)[p], np.asarray(y)[p], [sample_names[i] for i in p]) for p in parts]
    else:
        parts, payload = None, [None] * world
    mine = [None]
    dist.scatter_object_list(mine, payload, src=0)
    c, xs, ys, names = mine[0]
    local = sharded.sample_seq(c, xs, ys, names, own_coordinates)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0)
    if rank != 0:
        return None
    seqs, status = [b""] * len(x), np.zeros(len(x), np.uint8)
    for p, (sq, st) in zip(parts, gathered):
        for j, i in enumerate(p):
            seqs[i] = sq[j]
        status[p] = st
    return seqs, status


def distributed_var_in_ref(dist, sharded: ShardedIndex, owner: Dict[str, int], contigs, x, y):
    """Rank 0 holds the region list: route, scatter, answer locally, gather, restore the order."""
    rank, world = dist.get_rank(), dist.get_world_size()
    if rank == 0:
        parts = route(contigs, owner, world)
        payload = [([contigs[i] for i in p], np.asarray(x)[p], np.asarray(y)[p]) for p in parts]
    else:
        parts, payload = None, [None] * world
    mine = [None]
    dist.scatter_object_list(mine, payload, src=0)
    c, xs, ys = mine[0]
    local = sharded.var_in_ref(c, xs, ys)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0)
    if rank != 0:
        return None
    out = np.zeros(len(x), np.uint32)
    for p, g in zip(parts, gathered):
        out[p] = g
    return out
