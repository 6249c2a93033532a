"""GPU parity on the other BASELINE.json configurations (scaled so the oracle finishes in seconds):
[2] several contigs sharded over GPUs, [3] TCGA-like sparse cohort with t7 lookups (explicit sample
ids), [4] region widths from 100 bp to 1 Mb (search-bound to scan-bound)."""
import numpy as np
import pytest

import vs_testlib as T
from vs_testlib import Oracle

pytestmark = pytest.mark.gpu
NONE = 0xFFFFFFFF


def test_tcga_like_sparse_t7_cuda(tmp_path):
    o = Oracle.synth(str(tmp_path / "ser"), chr_name="2", ref_length=3_000_000, n_records=150_000, n_samples=10_000, mode=1, seed=9, cqf_log2=21)
    assert o.construct_info["use_bit_vector"] == 0
    with T.open_engine(str(tmp_path / "ser"), "cuda") as e:
        assert e.info.class_mode == 0 and e.info.num_samples == 10_001
        av = o.all_variants()
        rng = np.random.default_rng(1)
        pick = rng.choice(len(av), 6000, replace=False)
        pos = [av[i][0] for i in pick] + [av[i][0] + 1 for i in pick[:500]]
        refs = [av[i][1] for i in pick] + [av[i][1] for i in pick[:500]]
        alts = [av[i][2] for i in pick] + [av[i][2] for i in pick[:500]]
        f7, c7, d7 = o.batch_t7(pos, refs, alts)
        rec = e.batch_samples_has_var(pos, refs, alts)
        ec, ed = e.digest_t7(rec)
        assert np.array_equal(rec != NONE, f7 == 1)
        hit = f7 == 1
        assert 0 < hit.sum() < len(pos)                       # both outcomes are exercised
        assert np.array_equal(c7[hit], ec[hit]) and np.array_equal(d7[hit], ed[hit])
        # the same lookups through the device-resident batch
        from variantstore_b200 import Batch
        b7 = Batch(e, 7, pos, refs=refs, alts=alts)
        b7.run()
        assert np.array_equal(b7.fetch() != NONE, f7 == 1)
        # t4 / t6 in explicit-id mode at this scale (rare carriers: long back-walks)
        x = rng.integers(1, 3_000_000, 1500).astype(np.uint64)
        y = x + rng.choice([100, 1000, 50_000], 1500).astype(np.uint64)
        s = rng.integers(1, 10_001, 1500).astype(np.uint32)
        bad6, bad4, _ = T.compare_all(o, e, x, y, s)
        assert not bad6 and not bad4


@pytest.mark.parametrize("wide_entries,pool", [(None, None), ("64", None), (None, "3"), (None, "off")])
def test_width_sweep_cuda(tmp_path, monkeypatch, wide_entries, pool):
    """pool: the warp-per-region kernel keeps the rows beyond its 1 024 staged ones in chunks from a pool (default); with 3
    chunks most warps find it empty and walk their region a second time; "off": no pool at all."""
    monkeypatch.delenv("VSGPU_T4W_POOL_CHUNKS", raising=False)
    monkeypatch.delenv("VSGPU_T4W_POOL", raising=False)
    if pool == "off":
        monkeypatch.setenv("VSGPU_T4W_POOL", "0")
    elif pool:
        monkeypatch.setenv("VSGPU_T4W_POOL_CHUNKS", pool)
    if wide_entries:
        monkeypatch.setenv("VSGPU_WIDE_ENTRIES", wide_entries)     # push more regions onto the warp-cooperative path
    else:
        monkeypatch.delenv("VSGPU_WIDE_ENTRIES", raising=False)
    o = Oracle.synth(str(tmp_path / "ser"), ref_length=3_000_000, n_records=90_000, n_samples=300, fmax=120, seed=12, cqf_log2=20)
    with T.open_engine(str(tmp_path / "ser"), "cuda") as e:
        rng = np.random.default_rng(2)
        for width, n in [(100, 400), (1000, 400), (10_000, 300), (100_000, 120), (1_000_000, 30)]:
            x = rng.integers(1, 3_000_000 - width, n).astype(np.uint64)
            y = x + np.uint64(width)
            s = rng.integers(1, 301, n).astype(np.uint32)
            bad6, bad4, _ = T.compare_all(o, e, x, y, s)
            assert not bad6 and not bad4, width
            # the device-resident fused batch gives the separate calls' answers at every width (few, wide regions: two launches)
            from variantstore_b200 import Batch
            lo, hi, cnt = e.batch_var_in_ref(x, y)
            off, hits = e.batch_sample_var_in_ref(x, y, s)
            b = Batch(e, 46, x, y, sample_ids=s)
            b.run()
            assert all(t > 0 for t in b.timings_ms())
            flo, fhi, fcnt, foff, fhits = b.fetch()
            b.close()
            assert np.array_equal(flo, lo) and np.array_equal(fhi, hi) and np.array_equal(fcnt, cnt) and np.array_equal(foff, off) and np.array_equal(fhits, hits), width
        # whole contig and beyond
        x = np.array([1, 1, 2_999_000], np.uint64)
        y = np.array([3_000_001, 10_000_000, 3_100_000], np.uint64)
        bad6, bad4, _ = T.compare_all(o, e, x, y, np.array([5, 17, 250], np.uint32))
        assert not bad6 and not bad4


def test_multi_contig_sharded_cuda(tmp_path):
    """Config [2] in small: contigs as independent shards on one GPU, routed by the host."""
    from variantstore_b200 import VariantStoreIndex
    from variantstore_b200.sharding import ShardedIndex, assign_contigs, route
    names = ["20", "21", "22"]
    prefixes, oracles, sizes = {}, {}, {}
    for i, c in enumerate(names):
        oracles[c] = Oracle.synth(str(tmp_path / c), chr_name=c, ref_length=400_000 + 100_000 * i, n_records=9000 + 3000 * i,
                                  n_samples=100, fmax=60, seed=30 + i, cqf_log2=18)
        prefixes[c] = str(tmp_path / c)
        sizes[c] = 9000 + 3000 * i
    owner = assign_contigs(sizes, 2)
    assert sorted(set(owner.values())) == [0, 1]
    sh = ShardedIndex(prefixes, lambda p: VariantStoreIndex(p, device=0))
    rng = np.random.default_rng(4)
    n = 3000
    contigs = [names[i] for i in rng.integers(0, 3, n)]
    x = rng.integers(1, 400_000, n).astype(np.uint64)
    y = x + rng.choice([10, 1000, 20_000], n).astype(np.uint64)
    parts = route(contigs, owner, 2)
    assert sum(len(p) for p in parts) == n
    got6 = sh.var_in_ref(contigs, x, y)
    samples = [f"S{int(i):03d}" for i in rng.integers(1, 101, n)]
    got4, texts = sh.sample_var_in_ref(contigs, x, y, samples)
    for c in names:
        idx = np.nonzero(np.array(contigs) == c)[0]
        assert np.array_equal(got6[idx].astype(np.uint64), oracles[c].batch_t6(x[idx], y[idx])[0])
        sid = np.array([int(samples[i][1:]) for i in idx], np.uint32)
        c4, _, _ = oracles[c].batch_t4(x[idx], y[idx], sid)
        assert np.array_equal(got4[idx].astype(np.uint64), c4)
        for i in idx[:40]:
            want = "\n".join(oracles[c].t4_text(int(x[i]), int(y[i]), samples[i])[0].split("\n")[2:])
            assert texts[i] == want
    sh.close()


def test_router_contigs_and_position_shards_cuda(tmp_path):
    """Config [2] in small through the C++ router (csrc/router.cc): three contigs, one of them cut into two
    position-range shards (each built from the records of its range), regions of all of them interleaved
    in one call, answers in the caller's order and equal to the oracle of the shard that owns the start."""
    from variantstore_b200 import Router
    specs = [("20", 400_000, 9000, 1, 400_000, 30), ("21", 500_000, 12000, 1, 500_000, 31),
             ("22", 600_000, 8000, 1, 300_000, 32), ("22", 600_000, 8000, 280_000, 600_000 - 1000, 33)]   # (contig, ref_length, records, pos_lo, pos_hi, seed)
    prefixes, oracles, ranges = [], [], [(0, 0), (0, 0), (0, 300_000), (300_000, 0)]
    for k, (c, rl, nrec, plo, phi, seed) in enumerate(specs):
        p = str(tmp_path / f"s{k}")
        oracles.append(Oracle.synth(p, chr_name=c, ref_length=rl, pos_lo=max(2, plo), pos_hi=phi, n_records=nrec, n_samples=100, fmax=60, seed=seed, cqf_log2=18))
        prefixes.append(p)
    with Router(prefixes, ranges=ranges, ndevices=1) as r:
        assert r.contigs == ["20", "21", "22"] and r.num_shards == 4
        rng = np.random.default_rng(9)
        n = 6000
        contig = rng.integers(0, 3, n)
        lens = np.array([400_000, 500_000, 600_000])[contig]
        x = (rng.integers(1, 10**9, n) % (lens - 2000) + 1).astype(np.uint32)
        y = x + rng.choice([10, 1000, 20_000], n).astype(np.uint32)
        s = rng.integers(1, 101, n).astype(np.uint32)
        so, lo, c6, c4, off, hits = r.query_t6t4(r.contig_ids([["20", "21", "22"][c] for c in contig]), x, y, s)
        want_shard = np.where(contig < 2, contig, np.where(x < 300_000, 2, 3))
        assert np.array_equal(so, want_shard) and np.array_equal(np.diff(off), c4) and off[-1] == len(hits)
        for k in range(4):
            idx = np.nonzero(so == k)[0]
            assert len(idx) > 500
            o6, d6 = oracles[k].batch_t6(x[idx], y[idx])
            o4, d4, ub = oracles[k].batch_t4(x[idx], y[idx], s[idx])
            sh = r.shard(k)
            assert np.array_equal(o6, c6[idx]) and np.array_equal(d6, sh.digest_t6(lo[idx], lo[idx] + c6[idx]))
            sub_off = np.concatenate([[0], np.cumsum(c4[idx])]).astype(np.uint64)
            sub_hits = np.concatenate([hits[off[i]:off[i + 1]] for i in idx]) if len(idx) else np.zeros(0, np.uint32)
            ok = (o4 == c4[idx]) & (d4 == sh.digest_t4(sub_off, sub_hits))
            assert np.all(ok | (ub != 0))
        for i in (0, 17, 999, n - 1):                      # one region's hit codes where the copy from its GPU put them
            assert np.array_equal(r.hits_of(i), hits[off[i]:off[i + 1]])
        st = r.stats()
        assert st["devices"] == [0] and st["device_regions"] == [n] and st["device_ms"][0] > 0
        # a start no shard owns (contig 22 has no shard for nothing; an unknown contig id) is an error, not a silent zero
        from variantstore_b200 import VsgpuError
        with pytest.raises(VsgpuError):
            r.query_t6t4(np.array([7], np.uint32), x[:1], y[:1], s[:1])
    for o in oracles:
        o.close()


def test_full_size_chr22_shape_cuda(tmp_path):
    """BASELINE.json config [1] at full size (1.1 M records x 2 504 samples, 1 M regions): the oracle
    checks a 3 000-region subsample bit for bit; the full batch is checked through size-independent
    properties (split invariance = a checksum of checksums, idempotence, t6 slice algebra, CSR sanity)."""
    from variantstore_b200 import Batch
    o = Oracle.synth(str(tmp_path / "ser"), chr_name="22", ref_length=51_304_566, pos_lo=16_050_000, pos_hi=51_244_566,
                     n_records=1_103_547, n_samples=2504, fmax=1100, seed=2022, cqf_log2=25)
    ci = o.construct_info
    assert 3_000_000 < ci["vertices"] < 3_500_000 and 400_000 < ci["classes"] < 520_000        # chr22-like (vs_v1.log:3-15)
    with T.open_engine(str(tmp_path / "ser"), "cuda") as e:
        rng = np.random.default_rng(1)
        n = 1_000_000
        x = np.sort(rng.integers(16_050_000, 51_304_566 - 1000, n)).astype(np.uint64)
        y = x + np.uint64(1000)
        s = rng.integers(1, 2505, n).astype(np.uint32)
        sub = rng.choice(n, 3000, replace=False)
        bad6, bad4, _ = T.compare_all(o, e, x[sub], y[sub], s[sub])
        assert not bad6 and not bad4
        assert not T.compare_t1(o, e, x[sub[:1000]])
        # whole batch, twice (idempotence), and as two halves (split invariance)
        lo, hi, cnt = e.batch_var_in_ref(x, y)
        off, hits = e.batch_sample_var_in_ref(x, y, s)
        b4 = Batch(e, 4, x, y, sample_ids=s)
        b4.run()
        b4.run()
        off2, hits2, cnt4 = b4.fetch()
        assert np.array_equal(off, off2) and np.array_equal(hits, hits2)
        h = n // 2
        offa, hitsa = e.batch_sample_var_in_ref(x[:h], y[:h], s[:h])
        offb, hitsb = e.batch_sample_var_in_ref(x[h:], y[h:], s[h:])
        assert np.array_equal(np.concatenate([hitsa, hitsb]), hits)
        assert np.array_equal(np.concatenate([offa[:-1], offb + offa[-1]]), off)
        da = e.digest_t4(off, hits, with_samples=False)
        assert np.bitwise_xor.reduce(da) == np.bitwise_xor.reduce(np.concatenate([e.digest_t4(offa, hitsa, False), e.digest_t4(offb, hitsb, False)]))
        # t2 (query_sample_from_ref): oracle on a subsample; the whole batch (~1 GB of sequence) as two halves
        bad2, _ = T.compare_t2(o, e, x[sub[:1500]], y[sub[:1500]], s[sub[:1500]])
        assert not bad2
        soff, stext, sst, _ = e.batch_sample_seq_in_ref(x, y, s)
        sl = np.diff(soff.astype(np.int64))
        assert int(sst.sum()) < n // 100 and np.all(sl[sst == 1] == 0)
        assert np.all(np.abs(sl[sst == 0] - 1000) <= 64) and 999.0 < sl[sst == 0].mean() < 1001.0   # ref interval +- the sample's indels
        soa, sta_, ssa, _ = e.batch_sample_seq_in_ref(x[:h], y[:h], s[:h])
        sob, stb_, ssb, _ = e.batch_sample_seq_in_ref(x[h:], y[h:], s[h:])
        assert sta_ + stb_ == stext and np.array_equal(np.concatenate([soa[:-1], sob + soa[-1]]), soff)
        del stext, sta_, stb_
        # t3 (sample coordinates; the per-carrier indexes come from a second pass over ser/).  This synthetic index
        # is built without fix_sample_indexes, so a t3 walk starts from an unfixed sample position far from x and
        # the oracle needs about half a second per region: a handful of regions only
        bad3, _ = T.compare_t3(o, e, x[sub[:16]], y[sub[:16]], s[sub[:16]])
        assert not bad3
        bad5, _ = T.compare_t5(o, e, x[sub[:16]], y[sub[:16]], s[sub[:16]])
        assert not bad5
        # t6 slice algebra on sorted equal-width regions: bounds are monotone, counts add up over a split at any y
        ok = lo != NONE
        assert np.all(np.diff(lo[ok].astype(np.int64)) >= 0) and np.all(np.diff(hi[ok].astype(np.int64)) >= 0)
        assert np.array_equal(cnt[ok], (hi[ok] - lo[ok]))
        # CSR sanity: offsets non-decreasing, every hit code points at a walk entry, ~1.7 rows per region
        assert np.all(np.diff(off.astype(np.int64)) >= 0) and off[-1] == len(hits)
        assert int((hits & np.uint32(0x3FFFFFFF)).max()) < e.info.walk_entries
        assert 1.2 < len(hits) / n < 2.4
