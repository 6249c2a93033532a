"""GPU parity tests proper: the CUDA path, called through the C ABI of libvsgpu, against the oracle
on the same seeded inputs — bit-exact (integer / byte / index work)."""
import os

import numpy as np
import pytest

import vs_testlib as T
from vs_testlib import Oracle
from variantstore_b200.api import VsgpuError

pytestmark = pytest.mark.gpu
NONE = 0xFFFFFFFF


def _t7_check(o, e, limit=None):
    av = o.all_variants()
    if limit:
        av = av[:limit]
    pos = [p for p, _, _ in av] + [p + 1 for p, _, _ in av[:64]]
    refs = [r for _, r, _ in av] + [r for _, r, _ in av[:64]]
    alts = [a for _, _, a in av] + [a for _, _, a in av[:64]]
    f7, c7, d7 = o.batch_t7(pos, refs, alts)
    rec = e.batch_samples_has_var(pos, refs, alts)
    ec, ed = e.digest_t7(rec)
    assert np.array_equal(rec != NONE, f7 == 1)
    hit = f7 == 1
    assert np.array_equal(c7[hit], ec[hit]) and np.array_equal(d7[hit], ed[hit])
    return int(hit.sum()), len(pos)


@pytest.fixture(params=["hitmap", "class-bitmaps", "warp-cooperative", "cta-per-tile", "tile-256", "chunked-calls", "carried-entry-lists", "hitmap-32bit-words", "spill", "spill-scratch-full"])
def walk_path(request, monkeypatch):
    """The t4 kernel paths: the sample-major hit map with one thread per region (default), the
    per-entry class-bitmap test used when the map does not fit the memory budget, and the
    warp-cooperative scan for wide regions (forced onto every region longer than 8 walk entries);
    plus the non-persistent variant of the per-thread kernel, and host-buffer calls cut into chunks
    of 256 regions (copies of one chunk overlapping the kernels of the next; offsets chained); and the
    spilling instance of k_t4p (rows beyond the staged 8 go to a scratch), with ample and with too little scratch."""
    monkeypatch.delenv("VSGPU_CHUNK_REGIONS", raising=False)
    if request.param == "chunked-calls":
        monkeypatch.setenv("VSGPU_CHUNK_REGIONS", "128")
    monkeypatch.delenv("VSGPU_DISABLE_HITMAP", raising=False)
    monkeypatch.delenv("VSGPU_WIDE_ENTRIES", raising=False)
    monkeypatch.delenv("VSGPU_T4_PIPE", raising=False)
    monkeypatch.delenv("VSGPU_T4_TILE", raising=False)
    monkeypatch.delenv("VSGPU_T4_ROW64", raising=False)
    monkeypatch.delenv("VSGPU_T4_SPILL", raising=False)
    monkeypatch.delenv("VSGPU_T4_SPILL_WORDS", raising=False)
    if request.param.startswith("spill"):
        monkeypatch.setenv("VSGPU_T4_SPILL", "1")             # k_t4p's spilling instance for every batch (default: only where many rows per region are expected)
        if request.param == "spill-scratch-full":
            monkeypatch.setenv("VSGPU_T4_SPILL_WORDS", "64")  # 8 chunks per CTA and tile: most threads run out of scratch and take the second walk
    monkeypatch.delenv("VSGPU_SPARSE_WALK", raising=False)   # default: per-sample carried-entry lists for explicit-id cohorts, the hit map otherwise
    if request.param == "carried-entry-lists":
        monkeypatch.setenv("VSGPU_SPARSE_WALK", "1")          # the lists for every cohort
    elif request.param not in ("hitmap", "spill", "spill-scratch-full"):
        monkeypatch.setenv("VSGPU_SPARSE_WALK", "0")          # never: the explicit-id cases take the named path too
    if request.param == "hitmap-32bit-words":
        monkeypatch.setenv("VSGPU_T4_ROW64", "0")             # walk_region_fast (32 entries per step) instead of the 64-entry chunks
    if request.param == "cta-per-tile":
        monkeypatch.setenv("VSGPU_T4_PIPE", "0")        # k_t4 instead of the persistent pipelined k_t4p
    if request.param == "tile-256":
        monkeypatch.setenv("VSGPU_T4_TILE", "256")      # fewer, larger tiles than the default 64
    if request.param == "class-bitmaps":
        monkeypatch.setenv("VSGPU_DISABLE_HITMAP", "1")
    elif request.param == "warp-cooperative":
        monkeypatch.setenv("VSGPU_WIDE_ENTRIES", "8")
    return request.param


@pytest.mark.parametrize("overlap,sparse", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_fuzz_parity_cuda(tmp_path, seed, overlap, sparse, walk_path):
    fa, vcf, names = T.write_fuzz_inputs(str(tmp_path), seed, overlap=overlap, sparse=sparse)
    o = Oracle.construct(fa, vcf, str(tmp_path / "ser"), force_enc=0 if sparse else -1)
    with T.open_engine(str(tmp_path / "ser"), "cuda") as e:
        assert e.info.class_mode == (0 if sparse else 1)
        x, y, s = T.random_regions(seed + 100, 400, 4000, n_samples=len(names))
        bad6, bad4, _ = T.compare_all(o, e, x, y, s)
        assert not bad6 and not bad4
        hits, total = _t7_check(o, e)
        assert 0 < hits < total
        pos = np.concatenate([np.arange(1, 300), np.random.default_rng(seed).integers(1, 4000, 500), [3999, 4000, 4001, 5000]]).astype(np.uint64)
        assert not T.compare_t1(o, e, pos)
        x[:8] = np.arange(8)                                     # t2 has no pos-0 check: x = 0 throws inside substr
        bad2, threw = T.compare_t2(o, e, x, y, s)
        assert not bad2 and threw >= 1
        bad3, odd3 = T.compare_t3(o, e, x, y, s)                 # sample coordinates: incl. the regions the reference hangs on
        assert not bad3 and (sparse or odd3 >= 1)
        bad5, _ = T.compare_t5(o, e, x, y, s)
        assert not bad5


def test_golden_fixture_cuda():
    """README.md:93-95 of the reference: query -t 6 -r 10:105 on data/x.* -> 8 variants."""
    prefix = os.path.join(T.GOLDEN, "x_ser")
    with T.open_engine(prefix, "cuda") as e:
        assert e.info.num_vertices_cqf == 212 and e.info.seq_length == 1074 and e.info.num_classes == 2
        lo, hi, cnt = e.batch_var_in_ref([10, 14, 9, 1, 100, 466, 972], [105, 105, 105, 1001, 104, 470, 1000])
        assert list(cnt) == [8, 6, 8, 75, 1, 1, 1]
        rows = e.get_var_in_ref(10, 105)
        assert [(v.var_pos, v.ref, v.alt) for v in rows] == [(10, "C", "T"), (14, "G", "A"), (34, "T", "A"), (39, "T", "A"),
                                                             (52, "T", "G"), (58, "", "T"), (100, "T", "C"), (103, "T", "C")]
        assert [v.samples for v in rows][:2] == [[("1", "1|1")], [("1", "1|0")]]
        assert len(e.get_sample_var_in_ref(14, 105, "1")) == 7
        assert e.samples_has_var(10, "C", "T") == [("1", "1|1")]
        assert e.samples_has_var(58, "", "T") == [("1", "0|1")]
        assert e.samples_has_var(14, "G", "A") == []
        found, rows = e.closest_var(20)
        assert found and [(v.var_pos, v.ref, v.alt) for v in rows] == [(9, "G", "A")]   # first variant of the mirrored window [6, 34], not the nearest
        # t2: sample "1" over 10:20 — C>T at 10 and G>A at 14 applied to CTTGGAAATT
        assert e.query_sample_from_ref(10, 20, "1") == "TTTGAAAATT"
        exp = __import__("json").load(open(os.path.join(T.GOLDEN, "expected.json")))
        for q in exp["x"]["t2"]:
            if q["status"]:
                with pytest.raises(IndexError):
                    e.query_sample_from_ref(q["x"], q["y"], "1")
            else:
                assert e.query_sample_from_ref(q["x"], q["y"], "1") == q["seq"]


def test_synthetic_1000g_shape_cuda(tmp_path, walk_path):
    """A scaled chr22-shaped index (classes, 300 samples): batch API == host-buffer API == oracle."""
    from variantstore_b200 import Batch
    o = Oracle.synth(str(tmp_path / "ser"), ref_length=2_000_000, n_records=60_000, n_samples=300, fmax=120, seed=5, cqf_log2=20)
    with T.open_engine(str(tmp_path / "ser"), "cuda") as e:
        rng = np.random.default_rng(3)
        n = 20_000
        x = np.sort(rng.integers(1, 2_000_000, n)).astype(np.uint64)
        y = x + rng.choice([100, 1000, 10_000, 100_000], n).astype(np.uint64)
        s = rng.integers(1, 301, n).astype(np.uint32)
        # oracle on a subsample (it walks vertex by vertex), engine on everything
        sub = rng.choice(n, 1500, replace=False)
        bad6, bad4, _ = T.compare_all(o, e, x[sub], y[sub], s[sub])
        assert not bad6 and not bad4
        lo, hi, cnt = e.batch_var_in_ref(x, y)
        off, hits = e.batch_sample_var_in_ref(x, y, s)
        # the 32-bit coordinate entry points (vsgpu_query_t6_u32 / _t4_u32) give the same answers
        lo32, hi32, cnt32 = e.batch_var_in_ref(x.astype(np.uint32), y.astype(np.uint32))
        off32, hits32 = e.batch_sample_var_in_ref(x.astype(np.uint32), y.astype(np.uint32), s)
        assert np.array_equal(lo, lo32) and np.array_equal(hi, hi32) and np.array_equal(cnt, cnt32)
        assert np.array_equal(off, off32) and np.array_equal(hits, hits32)
        b6, b4 = Batch(e, 6, x, y), Batch(e, 4, x, y, sample_ids=s)
        for _ in range(2):
            b6.run()
            b4.run()
        lo2, hi2, cnt2 = b6.fetch()
        off2, hits2, cnt4 = b4.fetch()
        assert np.array_equal(lo, lo2) and np.array_equal(hi, hi2) and np.array_equal(cnt, cnt2)
        assert np.array_equal(off, off2) and np.array_equal(hits, hits2) and np.array_equal(np.diff(off), cnt4)
        assert len(b4.timings_ms()) == 1 and b4.stats()[1] == 1 and b6.stats()[0] == 288 * n
        # size-independent properties: slices are monotone in x for sorted regions of equal width,
        # a t4 answer never has more rows than twice the t6 slice (+ the start / rejoin rows)
        bad2, _ = T.compare_t2(o, e, x[sub], y[sub], s[sub])       # up to 100 kb per region
        assert not bad2
        bad3, _ = T.compare_t3(o, e, x[sub[:200]], y[sub[:200]], s[sub[:200]])
        assert not bad3
        bad5, _ = T.compare_t5(o, e, x[sub[:200]], y[sub[:200]], s[sub[:200]])
        assert not bad5
        same_w = (y - x == 1000) & (lo != NONE)
        assert np.all(np.diff(lo[same_w].astype(np.int64)) >= 0)
        assert np.all(np.diff(off).astype(np.int64) <= 2 * (hi.astype(np.int64) - lo) + 2)
        hits7, total7 = _t7_check(o, e, limit=3000)
        assert hits7 > 0


def test_chunked_calls_keep_order_and_survive_overflow_cuda(tmp_path, monkeypatch):
    """Host-buffer calls are cut into chunks whose copies overlap the kernels.  The chunks chain their
    offsets on the device; a hit buffer guessed too small (wide regions: far more than 4 rows each)
    makes the call fall back to one exact pass.  Same answer either way, and equal to the oracle."""
    o = Oracle.synth(str(tmp_path / "ser"), ref_length=1_500_000, n_records=50_000, n_samples=200, fmax=90, seed=11, cqf_log2=20)
    rng = np.random.default_rng(12)
    n = 6_000
    x = np.sort(rng.integers(1, 1_400_000, n)).astype(np.uint64)
    y = x + rng.choice([500, 20_000, 60_000], n).astype(np.uint64)
    s = rng.integers(1, 201, n).astype(np.uint32)
    answers = []
    for chunk in ("0", "256", "1000"):
        monkeypatch.setenv("VSGPU_CHUNK_REGIONS", chunk)
        with T.open_engine(str(tmp_path / "ser"), "cuda") as e:          # fresh handle: the first call guesses the hit capacity
            off, hits = e.batch_sample_var_in_ref(x, y, s)
            assert off[-1] > 4 * n                                        # the guess (4 per region) was too small
            off2, hits2 = e.batch_sample_var_in_ref(x, y, s)              # second call: capacity known, chunks stream
            assert np.array_equal(off, off2) and np.array_equal(hits, hits2)
            lo, hi, cnt = e.batch_var_in_ref(x, y)
            answers.append((off, hits, lo, hi, cnt))
            if chunk == "256":
                sub = rng.choice(n, 600, replace=False)
                bad6, bad4, _ = T.compare_all(o, e, x[sub], y[sub], s[sub])
                assert not bad6 and not bad4
    for a in answers[1:]:
        assert all(np.array_equal(p, q) for p, q in zip(answers[0], a))


def test_duplicate_records_and_contig_tail_cuda(tmp_path, monkeypatch):
    """t6 counts come from the kernel; slices holding a repeated VCF record, or regions running past
    the contig end over tail records, are flagged on the device and re-counted with the literal
    dedup rule (query.h:397-414) — also when the call is chunked."""
    fa, vcf, names = T.write_fuzz_inputs(str(tmp_path), 2, n_records=60)
    lines = open(vcf).read().split("\n")
    body = [l for l in lines if l and not l.startswith("#")]
    snps = [l for l in body if len(l.split("\t")[3]) == 1 and len(l.split("\t")[4]) == 1]
    out = []
    for l in lines:
        out.append(l)
        if l in snps[:8]:
            out.append(l)
    open(vcf, "w").write("\n".join(out))
    o = Oracle.construct(fa, vcf, str(tmp_path / "ser"))
    for chunk in ("0", "128"):
        monkeypatch.setenv("VSGPU_CHUNK_REGIONS", chunk)
        with T.open_engine(str(tmp_path / "ser"), "cuda") as e:
            assert e.info.has_suspect_dups == 1
            x, y, s = T.random_regions(3, 900, 1200, widths=(1, 2, 5, 20, 100, 1000, 5000), n_samples=len(names))
            bad6, bad4, _ = T.compare_all(o, e, x, y, s)
            assert not bad6 and not bad4
            from variantstore_b200 import Batch
            b6 = Batch(e, 6, x, y)
            b6.run(); b6.run()
            lo, hi, cnt = e.batch_var_in_ref(x, y)
            lo2, hi2, cnt2 = b6.fetch()
            assert np.array_equal(cnt, cnt2) and np.array_equal(lo, lo2) and np.array_equal(hi, hi2)


def test_index_cache_cuda(tmp_path, monkeypatch):
    """An index opened from VSGPU_INDEX_CACHE answers exactly like one decoded from ser/."""
    o = Oracle.synth(str(tmp_path / "ser"), ref_length=400_000, n_records=15_000, n_samples=120, fmax=50, seed=21, cqf_log2=18)
    monkeypatch.setenv("VSGPU_INDEX_CACHE", str(tmp_path))
    rng = np.random.default_rng(4)
    x = np.sort(rng.integers(1, 400_000, 3000)).astype(np.uint64)
    y = x + rng.choice([10, 1000, 30_000], 3000).astype(np.uint64)
    s = rng.integers(1, 121, 3000).astype(np.uint32)
    answers = []
    for expect in (0, 1):
        with T.open_engine(str(tmp_path / "ser"), "cuda") as e:
            assert e.info.from_cache == expect
            off, hits = e.batch_sample_var_in_ref(x, y, s)
            answers.append((off, hits) + e.batch_var_in_ref(x, y) + e.batch_closest_var(x))
            if expect:
                sub = rng.choice(3000, 500, replace=False)
                bad6, bad4, _ = T.compare_all(o, e, x[sub], y[sub], s[sub])
                assert not bad6 and not bad4
                assert _t7_check(o, e, limit=2000)[0] > 0
    assert all(np.array_equal(p, q) for p, q in zip(*answers))


def _check_render(o, e, x, y, oracle_regions=60):
    """device-rendered rows == host-materialised rows of the same slices == the oracle's -o bytes"""
    lo, hi, cnt = e.batch_var_in_ref(x, y)
    for ws in (True, False):
        off, text, rows, ms = e.render_var_in_ref(x, y, with_samples=ws)
        assert rows == int(cnt.sum()) and len(text) == off[-1] and np.all(np.diff(off.astype(np.int64)) >= 0)
        for i in range(len(x)):
            got = text[off[i]:off[i + 1]].decode()
            assert got.count("\n") == cnt[i]
            if cnt[i] == hi[i] - lo[i] and lo[i] != NONE:
                assert got == e.rows_t6_text(int(lo[i]), int(hi[i]), with_samples=ws), (i, x[i], y[i])
        if ws:
            for i in list(range(min(oracle_regions, len(x)))):
                want = o.t6_text(int(x[i]), int(y[i]))
                want = want.split("Pos\tRef\tAlt\tSamples\n", 1)[1]               # drop the count line and the header
                assert text[off[i]:off[i + 1]].decode() == want, (i, x[i], y[i])


@pytest.mark.parametrize("sparse", [False, True])
def test_rows_rendered_on_device_cuda(tmp_path, sparse):
    """the -v rows of t6 produced by a kernel, byte for byte."""
    fa, vcf, names = T.write_fuzz_inputs(str(tmp_path), 4, overlap=True, sparse=sparse, n_samples=40 if sparse else 12)
    o = Oracle.construct(fa, vcf, str(tmp_path / "ser"), force_enc=0 if sparse else -1)
    with T.open_engine(str(tmp_path / "ser"), "cuda") as e:
        x, y, _ = T.random_regions(9, 300, 4000, widths=(1, 5, 100, 1000, 4000), n_samples=len(names))
        x = np.concatenate([x, [1, 1, 3999, 4000, 4500]]).astype(np.uint64)
        y = np.concatenate([y, [4001, 9000, 4001, 4001, 5000]]).astype(np.uint64)
        _check_render(o, e, x, y)
        off, text, rows, ms = e.render_var_in_ref(np.zeros(0, np.uint64), np.zeros(0, np.uint64))
        assert len(off) == 1 and text == b"" and rows == 0


@pytest.mark.parametrize("sparse,overlap", [(False, True), (True, True), (False, False)])
def test_t4_rows_rendered_on_device_cuda(tmp_path, sparse, overlap):
    """get_sample_var_in_ref with print (query.h:719-726): the rows of every region's t4 answer written by a kernel
    (vsgpu_render_t4) — byte for byte the oracle's `-o` rows, and the host materialiser's (vsgpu_rows_t4) on the same hit
    codes, START / REJOIN rows included, with and without the carrier lists."""
    fa, vcf, names = T.write_fuzz_inputs(str(tmp_path), 6, overlap=overlap, sparse=sparse, n_samples=40 if sparse else 12, n_records=320)
    o = Oracle.construct(fa, vcf, str(tmp_path / "ser"), force_enc=0 if sparse else -1)
    with T.open_engine(str(tmp_path / "ser"), "cuda") as e:
        starts = np.arange(1, 4001, 7, dtype=np.uint64)
        x = np.tile(starts, len(names))
        y = x + np.tile(np.where(np.arange(len(starts)) % 3 == 0, 40, 700).astype(np.uint64), len(names))
        s = np.repeat(np.arange(1, len(names) + 1, dtype=np.uint32), len(starts))
        off4, hits = e.batch_sample_var_in_ref(x, y, s)
        assert int(((hits & 0x80000000) != 0).sum()) > 0                      # rows the walk started on (ref column empty)
        for ws in (True, False):
            off, text, rows, ms = e.render_sample_var_in_ref(x, y, s, with_samples=ws)
            assert rows == len(hits) and off[-1] == len(text) and ms > 0
            want = b"".join(e.rows_t4_text(hits[off4[i]:off4[i + 1]], ws).encode() for i in range(0, len(x), 37))
            got = b"".join(text[off[i]:off[i + 1]] for i in range(0, len(x), 37))
            assert got == want
        off, text, rows, ms = e.render_sample_var_in_ref(x, y, s, with_samples=True)
        rng = np.random.default_rng(3)
        for i in rng.choice(len(x), 150, replace=False):
            otext, ub = o.t4_text(int(x[i]), int(y[i]), names[int(s[i]) - 1])
            if ub:
                continue
            assert text[off[i]:off[i + 1]].decode() == "\n".join(otext.split("\n")[2:]), (int(x[i]), int(y[i]), int(s[i]))
        off, text, rows, ms = e.render_sample_var_in_ref(np.zeros(0, np.uint64), np.zeros(0, np.uint64), np.zeros(0, np.uint32))
        assert len(off) == 1 and text == b"" and rows == 0


@pytest.mark.parametrize("sparse,overlap", [(False, True), (True, True), (False, False)])
def test_t5_rows_rendered_on_device_cuda(tmp_path, sparse, overlap):
    """get_sample_var_in_sample with print (query.h:596-606): the rows of every region's t5 answer written by a kernel
    (vsgpu_render_t5) — byte for byte the host materialiser's (vsgpu_rows_t5) on the codes vsgpu_query_t5 returns and the
    oracle's rows; the position column is the sample's own coordinate, so it differs from sample to sample for one code."""
    fa, vcf, names = T.write_fuzz_inputs(str(tmp_path), 8, overlap=overlap, sparse=sparse, n_samples=40 if sparse else 12, n_records=320)
    o = Oracle.construct(fa, vcf, str(tmp_path / "ser"), force_enc=0 if sparse else -1)
    with T.open_engine(str(tmp_path / "ser"), "cuda") as e:
        starts = np.arange(1, 4001, 11, dtype=np.uint64)
        x = np.tile(starts, len(names))
        y = x + np.tile(np.where(np.arange(len(starts)) % 3 == 0, 40, 700).astype(np.uint64), len(names))
        s = np.repeat(np.arange(1, len(names) + 1, dtype=np.uint32), len(starts))
        off5, hits, status, _ = e.batch_sample_var_in_sample(x, y, s)
        assert len(hits) > 1000
        for ws in (True, False):
            off, text, rows, st, ms = e.render_sample_var_in_sample(x, y, s, with_samples=ws)
            assert rows == len(hits) and off[-1] == len(text) and np.array_equal(st, status) and ms[2] > 0
            pick = range(0, len(x), 23)
            want = b"".join(e.rows_t5_text(hits[off5[i]:off5[i + 1]], int(s[i]), ws).encode() for i in pick)
            got = b"".join(text[off[i]:off[i + 1]] for i in pick)
            assert got == want
        off, text, rows, st, ms = e.render_sample_var_in_sample(x, y, s, with_samples=True)
        _, _, ost, ub = o.batch_t5(x, y, s, True)
        rng = np.random.default_rng(3)
        checked = 0
        for i in rng.choice(len(x), 200, replace=False):
            if ost[i] != 0 or ub[i]:
                assert ost[i] == 0 or off[i] == off[i + 1]
                continue
            want = o.t5_text(int(x[i]), int(y[i]), names[int(s[i]) - 1]).split("Pos\tRef\tAlt\tSamples\n", 1)[1]
            assert text[off[i]:off[i + 1]].decode() == want, (int(x[i]), int(y[i]), int(s[i]))
            checked += 1
        assert checked > 50
        off, text, rows, st, ms = e.render_sample_var_in_sample(np.zeros(0, np.uint64), np.zeros(0, np.uint64), np.zeros(0, np.uint32))
        assert len(off) == 1 and text == b"" and rows == 0


def test_rows_rendered_with_long_and_mixed_names_cuda(tmp_path):
    """items of up to 15 bytes come from per-sample templates; longer names (and a mix) take the generic path"""
    for k, fmt in enumerate(["a_rather_long_sample_name_{:04d}", "n{:d}", "x{:09d}"]):          # 30-char, 2..3-char, 10-char names (items of 36, 8-9, 16 bytes)
        d = tmp_path / str(k)
        fa, vcf, names = T.write_fuzz_inputs(str(d), 6 + k, n_samples=14, name_fmt=fmt)
        if k == 1:                                                                                  # mix short and long names in one index
            txt = open(vcf).read().replace("n7\t", "the_seventh_sample_of_fourteen\t")
            open(vcf, "w").write(txt)
        o = Oracle.construct(fa, vcf, str(d / "ser"))
        with T.open_engine(str(d / "ser"), "cuda") as e:
            x, y, _ = T.random_regions(2, 120, 4000, widths=(5, 100, 1000, 4000), n_samples=len(names))
            _check_render(o, e, x, y, oracle_regions=120)


def test_rows_rendered_on_device_many_samples_cuda(tmp_path, monkeypatch):
    """chr22-shaped classes (300 samples, 5 bitmap words, long carrier lists) and a batch large enough
    for several scan CTAs; names of different lengths."""
    o = Oracle.synth(str(tmp_path / "ser"), ref_length=600_000, n_records=20_000, n_samples=300, fmax=120, seed=8, cqf_log2=19)
    with T.open_engine(str(tmp_path / "ser"), "cuda") as e:
        rng = np.random.default_rng(6)
        x = np.sort(rng.integers(1, 600_000, 5000)).astype(np.uint64)
        y = x + rng.choice([10, 300, 3000], 5000).astype(np.uint64)
        _check_render(o, e, x, y, oracle_regions=40)
        # the text leaves the device in chunks while the next chunk is rendered: same bytes
        whole = e.render_var_in_ref(x, y)
        monkeypatch.setenv("VSGPU_RENDER_CHUNK_BYTES", "65536")
        cut = e.render_var_in_ref(x, y)
        assert np.array_equal(whole[0], cut[0]) and whole[1] == cut[1] and whole[2] == cut[2] and len(whole[1]) > 16 * 65536


def test_rows_rendered_with_duplicate_records_cuda(tmp_path):
    """regions whose rows are decided by the literal dedup rule become several segments"""
    fa, vcf, names = T.write_fuzz_inputs(str(tmp_path), 2, n_records=60)
    lines = open(vcf).read().split("\n")
    body = [l for l in lines if l and not l.startswith("#")]
    snps = [l for l in body if len(l.split("\t")[3]) == 1 and len(l.split("\t")[4]) == 1]
    out = []
    for l in lines:
        out.append(l)
        if l in snps[:8]:
            out.append(l)
    open(vcf, "w").write("\n".join(out))
    o = Oracle.construct(fa, vcf, str(tmp_path / "ser"))
    with T.open_engine(str(tmp_path / "ser"), "cuda") as e:
        assert e.info.has_suspect_dups == 1
        x, y, _ = T.random_regions(3, 250, 1200, widths=(1, 5, 20, 100, 1000, 5000), n_samples=len(names))
        _check_render(o, e, x, y, oracle_regions=250)


def test_sample_sequences_cuda(tmp_path, monkeypatch):
    """t2 (query_sample_from_ref) beyond the fuzz regions: every sample over sliding windows of an
    overlapping-deletion graph (the starts right behind a deletion throw in the reference), whole-contig
    regions, split invariance of a batch, and the byte limit."""
    fa, vcf, names = T.write_fuzz_inputs(str(tmp_path), 53, overlap=True, n_records=320)
    o = Oracle.construct(fa, vcf, str(tmp_path / "ser"))
    with T.open_engine(str(tmp_path / "ser"), "cuda") as e:
        starts = np.arange(0, 4003, 2, dtype=np.uint64)
        x = np.tile(np.concatenate([starts, starts]), len(names))
        y = x + np.tile(np.concatenate([np.full(len(starts), 37), np.full(len(starts), 5000)]).astype(np.uint64), len(names))
        s = np.repeat(np.arange(1, len(names) + 1, dtype=np.uint32), 2 * len(starts))
        bad2, threw = T.compare_t2(o, e, x, y, s)
        assert not bad2 and threw > len(names)
        bad3, odd3 = T.compare_t3(o, e, x, y, s)
        assert not bad3 and odd3 > len(names)
        bad5, odd5 = T.compare_t5(o, e, x, y, s)
        assert not bad5 and odd5 > len(names)
        off5, hits5, st5, ms5 = e.batch_sample_var_in_sample(x, y, s)
        h = len(x) // 3
        o5a, h5a, s5a, _ = e.batch_sample_var_in_sample(x[:h], y[:h], s[:h])
        o5b, h5b, s5b, _ = e.batch_sample_var_in_sample(x[h:], y[h:], s[h:])
        assert np.array_equal(np.concatenate([h5a, h5b]), hits5) and np.array_equal(np.concatenate([o5a[:-1], o5b + o5a[-1]]), off5) and ms5 > 0
        st3 = e.batch_sample_seq_in_sample(x, y, s)[2]
        i2 = int(np.nonzero(st3 == 2)[0][0])
        with pytest.raises(RuntimeError):
            e.query_sample_from_sample(int(x[i2]), int(y[i2]), names[int(s[i2]) - 1])
        off, text, st, ms = e.batch_sample_seq_in_ref(x, y, s)
        h = len(x) // 3
        offa, texta, sta, _ = e.batch_sample_seq_in_ref(x[:h], y[:h], s[:h])
        offb, textb, stb, _ = e.batch_sample_seq_in_ref(x[h:], y[h:], s[h:])
        assert texta + textb == text and np.array_equal(np.concatenate([offa[:-1], offb + offa[-1]]), off)
        assert np.array_equal(np.concatenate([sta, stb]), st) and ms > 0
        assert set(text) <= set(b"ACGTN")
        off0, text0, st0, _ = e.batch_sample_seq_in_ref([], [], [])
        assert len(off0) == 1 and off0[0] == 0 and text0 == b""
        monkeypatch.setenv("VSGPU_RENDER_MAX_BYTES", "1000")
        with pytest.raises(VsgpuError) as ei:
            e.batch_sample_seq_in_ref(x, y, s)
        assert ei.value.code == -3
        monkeypatch.delenv("VSGPU_RENDER_MAX_BYTES")
        with pytest.raises(VsgpuError) as ei:
            e.batch_sample_seq_in_ref([5], [50], [len(names) + 1])
        assert ei.value.code == -1


def test_error_paths_cuda(tmp_path):
    from variantstore_b200 import VsgpuError
    prefix = os.path.join(T.GOLDEN, "x_ser")
    with T.open_engine(prefix, "cuda") as e:
        with pytest.raises(VsgpuError):
            e.batch_var_in_ref([0], [10])           # the reference aborts on pos < 1 (index.h:151-154)
        with pytest.raises(VsgpuError):
            e.sample_id("nobody")
        lo, hi, cnt = e.batch_var_in_ref([2000, 1], [3000, 1])   # beyond the contig / empty region
        assert list(cnt) == [0, 0]
        off, hits = e.batch_sample_var_in_ref(np.zeros(0, np.uint64), np.zeros(0, np.uint64), np.zeros(0, np.uint32))
        assert len(off) == 1 and len(hits) == 0
        # sample id 0 ("ref") and ids past the sample map are refused, in the plain and in the fused call
        for sid in (0, 99):
            with pytest.raises(VsgpuError):
                e.batch_sample_var_in_ref([10], [105], [sid])
            with pytest.raises(VsgpuError):
                e.batch_var_and_sample_var_in_ref([10], [105], [sid])
        # hit codes / offsets handed back for digests are checked, not trusted (stale or garbled arrays must not index past the tables)
        off, hits = e.batch_sample_var_in_ref([10], [105], [1])
        with pytest.raises(VsgpuError):
            e.digest_t4(off, hits | np.uint32(0x3FFFFFF0))
        with pytest.raises(VsgpuError):
            e.digest_t4(off[::-1].copy(), hits)
        assert len(e.digest_t4(off, hits)) == 1
        # a region whose t6 rows are decided by the literal dedup rule comes back through the render path (api.get_var_in_ref)
        assert [v.var_pos for v in e.get_var_in_ref(10, 105)] == [10, 14, 34, 39, 52, 58, 100, 103]
    with pytest.raises(VsgpuError):
        T.open_engine(str(tmp_path / "missing"), "cuda")


def test_cli_front_end_matches_reference_cli_lines(tmp_path):
    """vsgpu_query (flags of src/variantstore.cc:101-134) against the oracle's stand-in for
    `variantstore query`: same count lines on stdout, same bytes in the -o file."""
    import subprocess
    prefix = os.path.join(T.GOLDEN, "x_ser")
    cli = os.path.join(T.ROOT, "variantstore_b200", "vsgpu_query")
    ref = os.path.join(T.ORACLE_DIR, "vs_oracle")
    subprocess.run(["make", "-s", "vs_oracle"], cwd=T.ORACLE_DIR, check=True)
    subprocess.run(["make", "-s", "../vsgpu_query"], cwd=T.CSRC_DIR, check=True)
    cases = [["-t", "6", "-r", "10:105"], ["-t", "6", "-r", "30:40,10:105,2000:3000,466:470"], ["-t", "4", "-s", "1", "-r", "14:105,660:700"],
             ["-t", "7", "-r", "10", "-b", "C", "-a", "T"], ["-t", "7", "-r", "58", "-b", "", "-a", "T"],
             ["-t", "2", "-s", "1", "-r", "10:105"], ["-t", "2", "-s", "1", "-r", "1:1002,660:700,55:62"],
             ["-t", "3", "-s", "1", "-r", "1:1001,466:470,57:60"], ["-t", "5", "-s", "1", "-r", "1:1001,466:600"],
             ["-t", "1", "-r", "20"], ["-t", "1", "-r", "500,20,990"]]
    for i, c in enumerate(cases):
        outs = []
        for exe, tag in ((cli, "gpu"), (ref, "cpu")):
            of = str(tmp_path / f"{tag}{i}.txt")
            p = subprocess.run([exe, "query", "-p", prefix, "-m", "0", "-v", "-o", of] + c, capture_output=True, text=True)
            assert p.returncode == 0, p.stderr
            lines = [l for l in p.stdout.split("\n") if l.startswith(("Number of variants", "Chromosome"))]
            outs.append((lines, open(of).read() if os.path.exists(of) else None))
        assert outs[0] == outs[1], c


def test_cli_regions_file_at_batch_scale(tmp_path):
    """A 100 000-line --regions-file (a batch no command line could carry) through vsgpu_query: the reference's count line for
    every region, in read_regions' sorted order (commands.cc:64-93), equal to the oracle's counts for t6 and t4; and the same
    regions split over two ser/ directories behind --prefixes (the router front-end)."""
    import subprocess
    prefix = os.path.join(T.GOLDEN, "x_ser")
    cli = os.path.join(T.ROOT, "variantstore_b200", "vsgpu_query")
    subprocess.run(["make", "-s", "../vsgpu_query"], cwd=T.CSRC_DIR, check=True)
    o = Oracle.open(prefix)
    rng = np.random.default_rng(41)
    n = 100_000
    x = rng.integers(1, 1001, n)
    y = x + rng.choice([1, 5, 40, 300], n)
    f = tmp_path / "regions.txt"
    f.write_text("# beg:end\n" + "\n".join(f"{a}:{b}" for a, b in zip(x, y)) + "\n")
    order = np.lexsort((y, x))                                             # std::sort of (beg, end) tuples
    xs, ys = x[order].astype(np.uint64), y[order].astype(np.uint64)
    c6, _ = o.batch_t6(xs, ys, False)
    c4, _, ub = o.batch_t4(xs, ys, np.ones(n, np.uint32), False)
    for t, want in (("6", c6), ("4", c4)):
        p = subprocess.run([cli, "query", "-p", prefix, "-m", "0", "-t", t, "-s", "1", "--regions-file", str(f)], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        got = np.array([int(l.rsplit(": ", 1)[1]) for l in p.stdout.split("\n") if l.startswith("Number of variants")])
        assert len(got) == n and np.all((got == want) | ((ub != 0) if t == "4" else False))
    # router front-end: the same contig twice under two names would clash, so a second contig: the small fixture
    p2 = os.path.join(T.GOLDEN, "xsmall_ser")
    o2 = Oracle.open(p2)
    chr1, chr2 = T.open_engine(prefix, "cuda"), T.open_engine(p2, "cuda")
    names = [chr1.chr, chr2.chr]
    chr1.close(); chr2.close()
    if names[0] != names[1]:
        m = 5000
        which = rng.integers(0, 2, m)
        xx = np.where(which == 0, rng.integers(1, 1001, m), rng.integers(1, 80, m))
        yy = xx + rng.choice([1, 10, 60], m)
        f2 = tmp_path / "regions2.txt"
        f2.write_text("\n".join(f"{names[w]}\t{a}:{b}" for w, a, b in zip(which, xx, yy)) + "\n")
        p = subprocess.run([cli, "query", "--prefixes", f"{prefix},{p2}", "-m", "0", "-t", "6", "--regions-file", str(f2), "--devices", "1"], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        got = {}
        for l in p.stdout.split("\n"):
            if "Number of variants" in l:
                got.setdefault(l.split("\t")[0], []).append(int(l.rsplit(": ", 1)[1]))
        for w, oo in ((0, o), (1, o2)):
            sel = which == w
            od = np.lexsort((yy[sel], xx[sel]))
            want, _ = oo.batch_t6(xx[sel][od].astype(np.uint64), yy[sel][od].astype(np.uint64), False)
            assert np.array_equal(np.array(got[names[w]]), want)
    o.close(); o2.close()
