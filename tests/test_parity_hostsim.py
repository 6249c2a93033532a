"""CPU-side parity: loader + flattener + the kernels' per-region logic (compiled for the host, test
only) + materialiser against the oracle, on the fuzz shapes of SURVEY.md §4."""
import os

import numpy as np
import pytest

import vs_testlib as T
from vs_testlib import Oracle

NONE = 0xFFFFFFFF


def t7_parity(o, e):
    av = o.all_variants()
    pos = [p for p, _, _ in av] + [p + 1 for p, _, _ in av[:40]] + [max(1, p - 1) for p, _, _ in av[:40]]
    refs = [r for _, r, _ in av] + [r for _, r, _ in av[:80]]
    alts = [a for _, _, a in av] + [a for _, _, a in av[:80]]
    f7, c7, d7 = o.batch_t7(pos, refs, alts)
    rec = e.batch_samples_has_var(pos, refs, alts)
    ec, ed = e.digest_t7(rec)
    assert np.array_equal(rec != NONE, f7 == 1)
    hit = f7 == 1
    assert np.array_equal(c7[hit], ec[hit]) and np.array_equal(d7[hit], ed[hit])
    return int(hit.sum())


@pytest.fixture(params=["hitmap", "hitmap-32bit-words", "class-bitmaps", "carried-entry-lists"])
def walk_path(request, monkeypatch):
    monkeypatch.delenv("VSGPU_DISABLE_HITMAP", raising=False)
    monkeypatch.delenv("VSGPU_T4_ROW64", raising=False)
    monkeypatch.delenv("VSGPU_SPARSE_WALK", raising=False)      # default: lists for explicit-id cohorts, hit map otherwise
    if request.param == "carried-entry-lists":
        monkeypatch.setenv("VSGPU_SPARSE_WALK", "1")              # ... here for every cohort
    elif request.param != "hitmap":
        monkeypatch.setenv("VSGPU_SPARSE_WALK", "0")              # ... here never (the explicit-id cases take the named path)
    if request.param == "class-bitmaps":
        monkeypatch.setenv("VSGPU_DISABLE_HITMAP", "1")
    elif request.param == "hitmap-32bit-words":
        monkeypatch.setenv("VSGPU_T4_ROW64", "0")          # walk_region_fast instead of walk_region_fast2
    return request.param


@pytest.mark.parametrize("overlap,sparse", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("seed", range(4))
def test_fuzz_parity(tmp_path, seed, overlap, sparse, walk_path):
    fa, vcf, names = T.write_fuzz_inputs(str(tmp_path), seed, overlap=overlap, sparse=sparse)
    o = Oracle.construct(fa, vcf, str(tmp_path / "ser"), force_enc=0 if sparse else -1)
    e = T.open_engine(str(tmp_path / "ser"), "hostsim")
    assert e.info.class_mode == (0 if sparse else 1)
    assert e.info.num_vertices_cqf == o.info()["cqf_vertices"] and e.info.num_vertices == o.info()["vertices"]
    x, y, s = T.random_regions(seed + 100, 300, 4000, n_samples=len(names))
    bad6, bad4, ub = T.compare_all(o, e, x, y, s)
    assert not bad6 and not bad4
    assert t7_parity(o, e) > 0
    # closest_var (t1): every position of the contig's head, a random spread, and past the end
    pos = np.concatenate([np.arange(1, 400), np.random.default_rng(seed).integers(1, 4000, 600), [3999, 4000, 4001, 5000]]).astype(np.uint64)
    assert not T.compare_t1(o, e, pos)
    # query_sample_from_ref (t2): the same regions, byte for byte, incl. the calls that throw
    bad2, _ = T.compare_t2(o, e, x, y, s)
    assert not bad2
    # query_sample_from_sample (t3): the sample's own coordinates; includes the regions for which the
    # reference never returns (status 2) and those whose substr throws (status 1)
    bad3, odd3 = T.compare_t3(o, e, x, y, s)
    assert not bad3 and (sparse or odd3 > 0)
    # get_sample_var_in_sample (t5): rows with var_pos in the sample's coordinates
    bad5, _ = T.compare_t5(o, e, x, y, s)
    assert not bad5


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_exhaustive_windows_reach_the_rare_walk_entries(tmp_path, seed, walk_path):
    """Every sample x a sliding window over the whole contig on overlapping-deletion graphs, so the
    out-of-step arrival markers and the rejoin vertices that carry samples themselves (the two rare
    kinds of walk entry) are hit, not just present."""
    fa, vcf, names = T.write_fuzz_inputs(str(tmp_path), 50 + seed, overlap=True, n_records=320)
    o = Oracle.construct(fa, vcf, str(tmp_path / "ser"))
    e = T.open_engine(str(tmp_path / "ser"), "hostsim")
    starts = np.arange(1, 4001, 3, dtype=np.uint64)
    x = np.tile(np.concatenate([starts, starts]), len(names))
    y = x + np.tile(np.concatenate([np.full(len(starts), 40), np.full(len(starts), 400)]).astype(np.uint64), len(names))
    s = np.repeat(np.arange(1, len(names) + 1, dtype=np.uint32), 2 * len(starts))
    bad6, bad4, ub = T.compare_all(o, e, x, y, s)
    assert not bad6 and not bad4
    off, hits = e.batch_sample_var_in_ref(x, y, s)
    assert e.info.walk_markers > 0
    # t2 on every third window: regions starting right behind a deletion whose target the neighbour
    # scan meets first make the reference's substr throw (status 1) — those must agree too
    bad2, threw = T.compare_t2(o, e, x[::3], y[::3], s[::3])
    assert not bad2 and threw > 0
    bad3, odd3 = T.compare_t3(o, e, x[1::3], y[1::3], s[1::3])
    assert not bad3 and odd3 > 0
    bad5, odd5 = T.compare_t5(o, e, x[2::3], y[2::3], s[2::3])
    assert not bad5 and odd5 > 0
    off5, hits5, st5, _ = e.batch_sample_var_in_sample(x[2::3], y[2::3], s[2::3])
    assert len(hits5) > 0
    if e.info.rejoin_carriers:
        assert int(((hits & 0x40000000) != 0).sum()) > 0


def test_many_samples_auto_sparse_detection(tmp_path):
    """40 samples with rare carriers: density <= 5 % in the first 99 records -> explicit sample ids
    (variant_graph.h:568-617), chosen by the construct restatement itself."""
    fa, vcf, names = T.write_fuzz_inputs(str(tmp_path), 21, n_samples=40, sparse=True)
    o = Oracle.construct(fa, vcf, str(tmp_path / "ser"))
    e = T.open_engine(str(tmp_path / "ser"), "hostsim")
    x, y, s = T.random_regions(4, 300, 4000, n_samples=len(names))
    bad6, bad4, _ = T.compare_all(o, e, x, y, s)
    assert not bad6 and not bad4


def test_edges_of_the_contig(tmp_path):
    fa, vcf, names = T.write_fuzz_inputs(str(tmp_path), 7)
    o = Oracle.construct(fa, vcf, str(tmp_path / "ser"))
    e = T.open_engine(str(tmp_path / "ser"), "hostsim")
    x = np.array([1, 1, 2, 3990, 3999, 4000, 4000, 4001, 5000, 1, 100, 100], np.uint64)
    y = np.array([2, 4001, 3, 4000, 4001, 4001, 9000, 4100, 6000, 100000, 100, 50], np.uint64)
    s = np.array([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12], np.uint32)
    bad6, bad4, _ = T.compare_all(o, e, x, y, s)
    assert not bad6 and not bad4
    # t2 has no gate and no pos-0 check: x = 0, empty and inverted regions, regions past the contig end
    x2 = np.concatenate([x, [0, 0, 5, 50, 4000, 4001, 7000, 1, 2**33]]).astype(np.uint64)
    y2 = np.concatenate([y, [5, 0, 5, 10, 2**40, 4002, 7001, 2**33, 2**34]]).astype(np.uint64)
    s2 = np.concatenate([s, [1, 2, 3, 4, 5, 6, 7, 8, 9]]).astype(np.uint32)
    bad2, threw = T.compare_t2(o, e, x2, y2, s2)
    assert not bad2 and threw >= 1
    bad3, odd3 = T.compare_t3(o, e, x2, y2, s2)
    assert not bad3 and odd3 >= 1
    bad5, _ = T.compare_t5(o, e, x2, y2, s2)
    assert not bad5
    rows = e.get_sample_var_in_sample(1, 4001, names[2])
    want = o.t5_text(1, 4001, names[2]).split("Pos\tRef\tAlt\tSamples\n", 1)[1]
    assert "".join(f"{v.var_pos}\t{v.ref}\t{v.alt}\t" + "".join(f"{a}({b}) " for a, b in v.samples) + "\n" for v in rows) == want and len(rows) > 10
    # 32-bit coordinate entry points: same answers as the 64-bit ones
    lo, hi, cnt = e.batch_var_in_ref(x, y)
    lo32, hi32, cnt32 = e.batch_var_in_ref(x.astype(np.uint32), y.astype(np.uint32))
    assert np.array_equal(lo, lo32) and np.array_equal(hi, hi32) and np.array_equal(cnt, cnt32)
    off, hits = e.batch_sample_var_in_ref(x, y, s)
    off32, hits32 = e.batch_sample_var_in_ref(x.astype(np.uint32), y.astype(np.uint32), s)
    assert np.array_equal(off, off32) and np.array_equal(hits, hits32)
    assert e.query_sample_from_ref(1, 4001, names[2]) == o.batch_t2([1], [4001], [3], want_text=True)[4][0]
    with pytest.raises(IndexError):
        e.query_sample_from_ref(0, 5, names[0])


def test_synthetic_generator_parity(tmp_path):
    """The bench's data path: programmatic synthetic construct (no VCF text), both modes."""
    for mode, kw in ((0, dict(n_samples=200, fmax=80)), (1, dict(n_samples=3000))):
        o = Oracle.synth(str(tmp_path / f"ser{mode}"), ref_length=300_000, n_records=12_000, mode=mode, seed=3 + mode, **kw)
        e = T.open_engine(str(tmp_path / f"ser{mode}"), "hostsim")
        assert e.info.class_mode == (1 if mode == 0 else 0)
        rng = np.random.default_rng(8)
        x = rng.integers(1, 300_000, 1200).astype(np.uint64)
        y = x + rng.choice([1, 50, 1000, 20_000], 1200).astype(np.uint64)
        s = rng.integers(1, kw["n_samples"] + 1, 1200).astype(np.uint32)
        bad6, bad4, _ = T.compare_all(o, e, x, y, s)
        assert not bad6 and not bad4
        assert t7_parity(o, e) > 0
        bad2, _ = T.compare_t2(o, e, x, y, s)
        assert not bad2
        bad3, _ = T.compare_t3(o, e, x, y, s)
        assert not bad3
        bad5, _ = T.compare_t5(o, e, x, y, s)
        assert not bad5


def test_duplicate_records_take_the_literal_path(tmp_path):
    """A VCF that repeats a record: the reference's dedup (query.h:397-414) makes the t6 answer depend
    on where the region starts; the engine detects such slices and re-counts them literally."""
    fa, vcf, names = T.write_fuzz_inputs(str(tmp_path), 2, n_records=60)
    lines = open(vcf).read().split("\n")
    body = [l for l in lines if l and not l.startswith("#")]
    snps = [l for l in body if len(l.split("\t")[3]) == 1 and len(l.split("\t")[4]) == 1]
    out = []
    for l in lines:
        out.append(l)
        if l in snps[:8]:
            out.append(l)                                   # exact duplicate record
    open(vcf, "w").write("\n".join(out))
    o = Oracle.construct(fa, vcf, str(tmp_path / "ser"))
    e = T.open_engine(str(tmp_path / "ser"), "hostsim")
    assert e.info.has_suspect_dups == 1
    x, y, s = T.random_regions(3, 400, 1200, widths=(1, 2, 5, 20, 100, 1000), n_samples=len(names))
    bad6, bad4, _ = T.compare_all(o, e, x, y, s)
    assert not bad6 and not bad4


def test_index_cache_round_trip(tmp_path, monkeypatch):
    """VSGPU_INDEX_CACHE: the loaded + flattened index is written once and read back by later opens
    (SURVEY.md section 8(f)2); a changed ser/ file, a truncated or a foreign cache file is ignored."""
    fa, vcf, names = T.write_fuzz_inputs(str(tmp_path), 31, overlap=True)
    o = Oracle.construct(fa, vcf, str(tmp_path / "ser"))
    x, y, s = T.random_regions(5, 300, 4000, n_samples=len(names))
    cache_dir = tmp_path / "cache"
    cache_dir.mkdir()
    monkeypatch.setenv("VSGPU_INDEX_CACHE", str(cache_dir))
    e1 = T.open_engine(str(tmp_path / "ser"), "hostsim")
    assert e1.info.from_cache == 0
    files = list(cache_dir.iterdir())
    assert len(files) == 1 and files[0].name.endswith(".vsgpu_cache")
    e2 = T.open_engine(str(tmp_path / "ser"), "hostsim")
    assert e2.info.from_cache == 1
    for f in ("num_vertices", "num_vertices_cqf", "seq_length", "branch_records", "walk_entries", "num_classes", "walk_markers"):
        assert getattr(e1.info, f) == getattr(e2.info, f)
    bad6, bad4, _ = T.compare_all(o, e2, x, y, s)
    assert not bad6 and not bad4
    assert t7_parity(o, e2) > 0
    assert not T.compare_t2(o, e2, x, y, s)[0]
    assert not T.compare_t3(o, e2, x, y, s)[0]        # t3 re-reads ser/ for the per-carrier indexes even when the index came from the cache
    assert e2.get_var_in_ref(1, 4001) == e1.get_var_in_ref(1, 4001)                # rows incl. sample names and phasing
    assert e2.get_sample_var_in_ref(1, 4001, names[3]) == e1.get_sample_var_in_ref(1, 4001, names[3])
    # a truncated cache is ignored (and replaced)
    blob = files[0].read_bytes()
    files[0].write_bytes(blob[: len(blob) // 2])
    e3 = T.open_engine(str(tmp_path / "ser"), "hostsim")
    assert e3.info.from_cache == 0 and files[0].stat().st_size == len(blob)
    # so is one whose ser/ changed underneath it
    p = tmp_path / "ser" / "sampleid_map.lst"
    os.utime(p, ns=(p.stat().st_atime_ns, p.stat().st_mtime_ns + 1_000_000_000))
    e4 = T.open_engine(str(tmp_path / "ser"), "hostsim")
    assert e4.info.from_cache == 0
    assert T.open_engine(str(tmp_path / "ser"), "hostsim").info.from_cache == 1
    # "1" puts it beside the data
    monkeypatch.setenv("VSGPU_INDEX_CACHE", "1")
    T.open_engine(str(tmp_path / "ser"), "hostsim")
    assert (tmp_path / "ser" / "vsgpu_flat.cache").exists()
    assert T.open_engine(str(tmp_path / "ser"), "hostsim").info.from_cache == 1
