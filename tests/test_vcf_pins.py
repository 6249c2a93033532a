"""Oracle-independent pins for the north-star operators t6 / t4 / t7.

tests/vcf_model.py derives, from VCF + FASTA text alone, the rows the operators owe for every allele
(normalisation, carriers with phasing, order).  These tests hold BOTH the oracle and the engine
(host build of the kernel logic here; the CUDA library under `-m gpu`) to that model, on the
reference's own fixture data/x.* and on generated multi-sample VCFs with SNPs, multi-allelic SNP
records, insertions and deletions.  The reference's status table has "Correctness test: TBD" for these
operators (query_info.txt:1-58); this is the repository's own.
"""
import os
import random

import numpy as np
import pytest

import vcf_model as M
import vs_testlib as T
from vs_testlib import Oracle

HAVE_REF = os.path.exists(os.path.join(T.REF_DATA, "x.vcf.gz"))


def write_gapped_inputs(dirpath, seed, ref_len=6000, n_records=300, n_samples=9, haploid=False):
    """VCF whose records leave >= MIN_GAP reference bases between each other (the model's domain)."""
    rnd = random.Random(seed)
    ref = "".join(rnd.choice("ACGT") for _ in range(ref_len))
    names = [f"S{i:03d}" for i in range(1, n_samples + 1)]
    lines = ["##fileformat=VCFv4.1", '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">',
             "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names)]
    pos, made = rnd.randint(3, 9), 0
    while made < n_records and pos < ref_len - 12:
        u = rnd.random()
        r = ref[pos - 1]
        if u < 0.55:
            alts = [rnd.choice([b for b in "ACGT" if b != r])]
        elif u < 0.70:                                   # multi-allelic SNP record
            alts = rnd.sample([b for b in "ACGT" if b != r], rnd.choice([2, 2, 3]))
        elif u < 0.85:
            alts = [r + "".join(rnd.choice("ACGT") for _ in range(rnd.randint(1, 4)))]
        else:
            k = rnd.randint(1, 4)
            r = ref[pos - 1:pos + k]
            alts = [ref[pos - 1]]
        gts = []
        for _ in names:
            if haploid:
                gts.append(str(int(rnd.random() < 0.3)))
            else:
                a = rnd.choice([0, 0, 0, 1, len(alts)]); b = rnd.choice([0, 0, 0, 1, len(alts)])
                gts.append(f"{a}{rnd.choice('||/')}{b}")
        if all(g in ("0|0", "0/0", "0") for g in gts):
            gts[rnd.randrange(len(gts))] = "1" if haploid else "1|0"
        lines.append(f"g\t{pos}\t.\t{r}\t{','.join(alts)}\t99\t.\t.\tGT\t" + "\t".join(gts))
        made += 1
        pos += len(r) + M.MIN_GAP + rnd.choice([0, 0, 1, 2, 3, 5, 9, 17, 30])
    os.makedirs(dirpath, exist_ok=True)
    fa, vcf = os.path.join(dirpath, "ref.fa"), os.path.join(dirpath, "in.vcf")
    with open(fa, "w") as f:
        f.write(">g\n" + "\n".join(ref[i:i + 80] for i in range(0, ref_len, 80)) + "\n")
    with open(vcf, "w") as f:
        f.write("\n".join(lines) + "\n")
    return fa, vcf


class OracleSide:
    def __init__(self, o):
        self.o = o

    def t6(self, x, y):
        return M.parse_rows(self.o.t6_text(x, y))

    def t4(self, x, y, name):
        return M.parse_rows(self.o.t4_text(x, y, name)[0])

    def t7(self, pos, ref, alt):
        t = self.o.t7_text(pos, ref, alt)
        return None if t.startswith("There is no such variant") else t.rstrip("\n")


class EngineSide:
    def __init__(self, e):
        self.e = e

    def t6(self, x, y):
        lo, hi, cnt = self.e.batch_var_in_ref([x], [y])
        assert cnt[0] == hi[0] - lo[0]
        return M.parse_rows(self.e.rows_t6_text(lo[0], hi[0]))

    def t4(self, x, y, name):
        off, hits = self.e.batch_sample_var_in_ref([x], [y], [self.e.sample_id(name)])
        return M.parse_rows(self.e.rows_t4_text(hits))

    def t7(self, pos, ref, alt):
        rec = self.e.batch_samples_has_var([pos], [ref], [alt])
        return None if rec[0] == 0xFFFFFFFF else self.e.rows_t7_text(rec[0])


def hold_to_model(side, fa, vcf, strict=True):
    names, ref, rows = M.alleles(fa, vcf, strict)
    n6 = M.check_t6_whole_contig(side.t6(1, len(ref) + 1), rows)
    sites = exact = 0
    for name in names:
        s, ex = M.check_t4_whole_contig(side.t4(1, len(ref) + 1, name), rows, name)
        sites += s; exact += ex
    # t4 region rules on substitutions the sample carries: a region starting ON the site prints ref "" when the walk
    # starts on that very vertex (cur_ref is still empty: query.h:650, :706 — every site but the first few of a contig,
    # where the back-walk runs into vertex 0); a region starting one base earlier prints the model row
    started = empty_ref = 0
    for m in rows:
        if m["kind"] != "sub" or m["nsite"] != 1:
            continue
        name = m["carriers"][0][0]
        on = side.t4(m["pos"], m["pos"] + 1, name)
        before = side.t4(m["pos"] - 1, m["pos"] + 1, name)
        want = M.row_text(m).split("\t")
        assert [(r[0], r[2], r[3]) for r in on] == [(m["pos"], m["alt"], want[3])] and on[0][1] in ("", m["ref"])
        empty_ref += on[0][1] == ""
        assert before and "\t".join(map(str, before[-1])) == M.row_text(m)      # (an abutting site at pos - 1 comes first)
        started += 1
        if started >= 25:
            break
    # t7: what the operator finds, it must report with exactly the VCF's carriers; an insertion is always found by
    # (pos, "", inserted bases) — its branch hangs off the vertex containing pos
    found = 0
    for m in rows:
        got = side.t7(m["pos"], m["ref"], m["alt"])
        if m["kind"] == "ins" and m["nsite"] == 1:
            assert got is not None, ("insertion not found", m["pos"], m["alt"])
        if got is not None:
            assert got == M.t7_expected(m), (m["pos"], m["ref"], m["alt"])
            found += 1
    assert found > 0 and started > 0 and empty_ref >= started - 3 and exact > sites // 2
    return n6, sites, exact, found


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not present")
def test_reference_fixture_held_to_the_vcf_model():
    """data/x.fa + data/x.vcf.gz (75 records, one sample): oracle and engine vs the text-derived model."""
    fa, vcf = T.REF_DATA + "/x.fa", T.REF_DATA + "/x.vcf.gz"
    o = Oracle.open(os.path.join(T.GOLDEN, "x_ser"))
    n6, sites, exact, found = hold_to_model(OracleSide(o), fa, vcf, strict=False)
    assert n6 == 75 and sites == 75
    e = T.open_engine(os.path.join(T.GOLDEN, "x_ser"), "hostsim")
    assert hold_to_model(EngineSide(e), fa, vcf, strict=False) == (n6, sites, exact, found)


@pytest.mark.parametrize("seed,haploid", [(0, False), (1, False), (2, False), (3, True)])
def test_generated_vcfs_held_to_the_vcf_model(tmp_path, seed, haploid):
    fa, vcf = write_gapped_inputs(str(tmp_path), seed, haploid=haploid)
    o = Oracle.construct(fa, vcf, str(tmp_path / "ser"))
    a = hold_to_model(OracleSide(o), fa, vcf)
    e = T.open_engine(str(tmp_path / "ser"), "hostsim")
    assert hold_to_model(EngineSide(e), fa, vcf) == a


@pytest.mark.gpu
@pytest.mark.parametrize("seed,haploid", [(0, False), (4, False), (5, True)])
def test_cuda_path_held_to_the_vcf_model(tmp_path, seed, haploid):
    """The CUDA library through the C ABI against the text-derived model (the oracle only builds the ser/)."""
    fa, vcf = write_gapped_inputs(str(tmp_path), seed, n_records=500, ref_len=12000, n_samples=17, haploid=haploid)
    Oracle.construct(fa, vcf, str(tmp_path / "ser")).close()
    e = T.open_engine(str(tmp_path / "ser"), "cuda")
    n6, sites, exact, found = hold_to_model(EngineSide(e), fa, vcf)
    assert n6 >= 500 and sites > 500


@pytest.mark.gpu
def test_cuda_path_on_the_committed_reference_fixture_rows():
    """tests/golden/x_ser through the CUDA library: the 75 rows of data/x.vcf.gz as committed in expected.json's
    model-checked form (positions / alleles / phasing derived from the VCF when the reference is present)."""
    e = T.open_engine(os.path.join(T.GOLDEN, "x_ser"), "cuda")
    rows = EngineSide(e).t6(1, 1002)
    assert len(rows) == 75 and rows[1][:3] == (10, "C", "T") and rows[1][3] == "1(1|1) "
    assert [r[0] for r in rows] == sorted(r[0] for r in rows)
    t4 = EngineSide(e).t4(1, 1002, "1")
    assert len(t4) == 75
