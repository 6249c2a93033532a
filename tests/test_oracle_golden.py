"""Pins the oracle: the reference's own golden numbers (README.md:51-62, :84-96 — the only
machine-checkable outputs the reference publishes for this path) and the committed fixtures."""
import json
import os

import numpy as np
import pytest

import vs_testlib as T
from vs_testlib import Oracle

EXPECTED = json.load(open(os.path.join(T.GOLDEN, "expected.json")))
HAVE_REF = os.path.exists(os.path.join(T.REF_DATA, "x.vcf.gz"))


def rows(text):
    return [tuple(line.split("\t")[:3]) for line in text.split("\n")[2:] if line]


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not present")
@pytest.mark.parametrize("use_ref_gqf", [False, True])
def test_readme_construct_goldens(tmp_path, use_ref_gqf):
    if use_ref_gqf and not os.path.exists(os.path.join(T.ORACLE_DIR, "_ref", "libgqf_ref.so")):
        pytest.skip("oracle/_ref not built")
    o = Oracle.construct(T.REF_DATA + "/x.fa", T.REF_DATA + "/x.vcf.gz", str(tmp_path / "ser"), use_ref_gqf=use_ref_gqf)
    ci = o.construct_info
    # README.md:55-60: "Num mutations: 75 num mutations-sample: 75", "Num vars: 75",
    # "#Vertices: 212 #Edges: 287 Seq length: 1074", "Number of sample vector classes: 2"
    assert (ci["num_mutations"], ci["num_mutations_samples"], ci["num_vars"]) == (75, 75, 75)
    assert (ci["cqf_vertices"], ci["edges"], ci["seq_length"], ci["classes"]) == (212, 287, 1074, 2)
    assert ci["use_bit_vector"] == 1 and ci["vertices"] == 213 and ci["index_ones"] == 138
    # README.md:93-95: after reload "#Vertices: 212 #Edges: 0", "Number of variants get_var_in_ref: 8"
    li = o.info()
    assert (li["cqf_vertices"], li["edges"], li["seq_length"]) == (212, 0, 1074)
    assert o.t6_text(10, 105).startswith("Number of variants get_var_in_ref: 8\n")
    assert ci == EXPECTED["x"]["construct_info"]


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not present")
def test_small_fixture_construct(tmp_path):
    o = Oracle.construct(T.REF_DATA + "/x.small.fa", T.REF_DATA + "/x.small.vcf", str(tmp_path / "ser"), cqf_log2=10)
    ci = o.construct_info
    assert (ci["cqf_vertices"], ci["edges"], ci["seq_length"], ci["vertices"], ci["classes"]) == (18, 27, 88, 19, 2)
    # tri-allelic site at 10 plus the insertion there (reported as a substitution hanging off a dummy vertex)
    assert rows(o.t6_text(1, 80)) == [("9", "G", "A"), ("10", "C", "T"), ("10", "C", "A"), ("10", "C", "AAA"), ("25", "", "T"),
                                      ("26", "", "A"), ("39", "T", ""), ("41", "C", ""), ("55", "C", "")]
    t4 = rows(o.t4_text(1, 80, "1")[0])
    assert len(t4) == 7 and ("10", "C", "T") in t4 and ("10", "C", "A") not in t4


def test_committed_fixture_matches_expected():
    for name in ("x", "xsmall"):
        o = Oracle.open(os.path.join(T.GOLDEN, name + "_ser"))
        ex = EXPECTED[name]
        for key, text in ex["t6"].items():
            x, y = map(int, key.split(":"))
            assert o.t6_text(x, y) == text, (name, key)
        for key, text in ex["t4"].items():
            x, y = map(int, key.split(":"))
            assert o.t4_text(x, y, "1")[0] == text, (name, key)
        for key, text in ex["t7"].items():
            p, r, a = key.split("|")
            assert o.t7_text(int(p), r, a) == text, (name, key)
        for q in ex["t2"]:
            ln, dg, st, ub, seqs = o.batch_t2([q["x"]], [q["y"]], [1], want_text=True)
            assert (int(st[0]), seqs[0]) == (q["status"], q["seq"]), (name, q["x"], q["y"])
        o.close()


def test_sample_sequence_equals_vcf_consensus():
    """query_sample_from_ref over the whole contig of data/x.* (one sample, no overlapping records) =
    the FASTA with the sample's alt alleles applied — computed without the oracle by
    tests/golden/make_golden.py::consensus and committed as a SHA-1; recomputed here when the
    reference's data directory is present."""
    import hashlib
    o = Oracle.open(os.path.join(T.GOLDEN, "x_ser"))
    seq = o.batch_t2([1], [o.info()["ref_length"] + 1], [1], want_text=True)[4][0]
    assert hashlib.sha1(seq.encode()).hexdigest() == EXPECTED["x"]["t2_consensus_sha1"]
    assert len(seq) == 1005 and seq[9:19] == "TTTGAAAATT"      # C>T at 10, G>A at 14 on CTTGGAAATT
    if os.path.isdir(T.REF_DATA):
        src = open(os.path.join(T.GOLDEN, "make_golden.py")).read()
        ns = {}
        exec(src[src.index("def consensus"):src.index("out = {}")], ns)
        ref, cons = ns["consensus"](os.path.join(T.REF_DATA, "x.fa"), os.path.join(T.REF_DATA, "x.vcf.gz"))
        assert cons == seq and len(ref) == 1001
    o.close()


def test_sample_coordinate_operators_agree_with_the_consensus():
    """Independent pins for t3 and t5 on data/x.*: a region of the sample's own coordinates is a slice of the
    consensus sequence (FASTA + the sample's alt alleles, computed without the oracle), and a substitution row
    of t5 carries the position at which its alt allele sits in that sequence.  Regions the reference hangs or
    throws on are excluded (they have no answer to compare)."""
    cons = EXPECTED["x"]["t2_consensus"]
    o = Oracle.open(os.path.join(T.GOLDEN, "x_ser"))
    rng = np.random.default_rng(0)
    n = 4000
    x = rng.integers(1, 1000, n).astype(np.uint64)
    y = x + rng.integers(1, 200, n).astype(np.uint64)
    ln, dg, st, ub, seqs = o.batch_t3(x, y, np.ones(n, np.uint32), want_text=True)
    checked = 0
    for i in range(n):
        if st[i] == 0 and not ub[i]:
            assert seqs[i] == cons[int(x[i]) - 1:int(y[i]) - 1], (int(x[i]), int(y[i]))
            checked += 1
    assert checked > n // 2
    rows = o.t5_text(1, len(cons) + 1, "1").split("Pos\tRef\tAlt\tSamples\n", 1)[1].strip().split("\n")
    subs = 0
    for line in rows:
        pos, ref, alt, _ = line.split("\t")
        if ref and alt:
            assert cons[int(pos) - 1:int(pos) - 1 + len(alt)] == alt, line
            subs += 1
    assert subs > 40 and len(rows) == 75          # every record of data/x.vcf.gz
    o.close()


def test_known_answers_x():
    """Hand-checkable answers on data/x.* (README golden + the survey's behavioural-model vectors)."""
    o = Oracle.open(os.path.join(T.GOLDEN, "x_ser"))
    assert rows(o.t6_text(10, 105)) == [("10", "C", "T"), ("14", "G", "A"), ("34", "T", "A"), ("39", "T", "A"), ("52", "T", "G"),
                                        ("58", "", "T"), ("100", "T", "C"), ("103", "T", "C")]
    counts = {(14, 105): 6, (9, 105): 8, (1, 1001): 75, (100, 104): 1, (466, 470): 1, (972, 1000): 1}
    for (x, y), c in counts.items():
        assert o.t6_text(x, y).startswith(f"Number of variants get_var_in_ref: {c}\n")
    assert rows(o.t6_text(466, 470)) == [("467", "C", "")] and rows(o.t6_text(972, 1000)) == [("973", "GG", "")]
    # the is_empty gate prints the t4 label even for t6 (query.h:746)
    assert o.t6_text(2000, 3000).startswith("Number of variants get_sample_var_in_ref: 0\n")
    t4 = rows(o.t4_text(14, 105, "1")[0])
    assert len(t4) == 7 and t4[0] == ("14", "", "A")          # walk starts on the SNP: ref column empty
    assert rows(o.t4_text(660, 700, "1")[0]) == [("668", "", "A"), ("670", "G", ""), ("681", "", "T"), ("688", "T", "C"), ("698", "T", "A")]
    assert len(rows(o.t4_text(1, 1001, "1")[0])) == 75 and len(rows(o.t4_text(10, 105, "1")[0])) == 8
    assert o.t7_text(10, "C", "T") == "1 1|1\n" and o.t7_text(58, "", "T") == "1 0|1\n"
    for q in [(14, "G", "A"), (58, "G", "GT"), (467, "C", ""), (100, "T", "C")]:
        assert o.t7_text(*q) == "There is no such variant!\n"
    o.close()


def test_read_regions_matches_reference_parsing():
    from variantstore_b200 import read_regions, read_sequences
    assert read_regions("30:40,10:105,7") == [(7, 0), (10, 105), (30, 40)]      # sorted (commands.cc:91)
    assert read_sequences("A,,GT") == ["A", "", "GT"]
