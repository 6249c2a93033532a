"""Position-range shards of one contig (router ranges, north star: "partitioned ... by contig / position shard").

A shard is a ser/ built by the reference's own construct from the FASTA of the whole contig and the VCF records of its
range plus a halo: records up to `W` bases past its upper end (so a region of width <= W that starts inside the range sees
every record it overlaps) and from `H` bases below its lower end (so the back-walk of get_sample_var_in_ref finds the
sample's previous variant where the whole contig would).  The router sends a region to the shard owning its start.  What a
shard answers is, by construction, what `variantstore query -p <that shard>` answers (the GPU parity tests hold that
against the oracle); this test measures how that relates to the UNSPLIT contig: t6 rows and t4 rows of routed regions
from the shard's oracle against the whole contig's oracle.
"""
import os

import numpy as np
import pytest

import vs_testlib as T
from vs_testlib import Oracle


def split_vcf(vcf, out, keep):
    with open(vcf) as f, open(out, "w") as g:
        for line in f:
            if line.startswith("#") or keep(int(line.split("\t", 2)[1])):
                g.write(line)


@pytest.mark.parametrize("seed,overlap", [(0, False), (1, False), (2, True), (3, True)])
def test_routed_shard_answers_equal_the_unsplit_contig(tmp_path, seed, overlap):
    fa, vcf, names = T.write_fuzz_inputs(str(tmp_path), 90 + seed, ref_len=6000, n_records=420, n_samples=10, overlap=overlap)
    whole = Oracle.construct(fa, vcf, str(tmp_path / "whole"))
    g, W, H = 3000, 400, 600
    va, vb = str(tmp_path / "a.vcf"), str(tmp_path / "b.vcf")
    split_vcf(vcf, va, lambda p: p < g + W)
    split_vcf(vcf, vb, lambda p: p >= g - H)
    shards = [Oracle.construct(fa, va, str(tmp_path / "sa")), Oracle.construct(fa, vb, str(tmp_path / "sb"))]
    rng = np.random.default_rng(seed)
    n = 3000
    x = rng.integers(1, 5900, n).astype(np.uint64)
    y = x + rng.choice([1, 5, 40, 200, W - 10], n).astype(np.uint64)
    s = rng.integers(1, len(names) + 1, n).astype(np.uint32)
    c6w, d6w = whole.batch_t6(x, y)
    c4w, d4w, ubw = whole.batch_t4(x, y, s)
    diff6 = diff4 = checked4 = 0
    for k, sel in ((0, x < g), (1, x >= g)):
        c6, d6 = shards[k].batch_t6(x[sel], y[sel])
        c4, d4, ub = shards[k].batch_t4(x[sel], y[sel], s[sel])
        diff6 += int(((c6 != c6w[sel]) | (d6 != d6w[sel])).sum())
        ok = (ub != 0) | (ubw[sel] != 0)
        diff4 += int((((c4 != c4w[sel]) | (d4 != d4w[sel])) & ~ok).sum())
        checked4 += int((~ok).sum())
    # t6 is position-local: with the forward halo every routed region sees its records — identical rows, identical order
    assert diff6 == 0
    # t4 additionally starts from the sample's previous variant (or the contig head): with the backward halo the rows are
    # the unsplit contig's (the walk is memoryless once it is back on the backbone)
    assert checked4 > n // 2 and diff4 == 0
    for o in shards + [whole]:
        o.close()
