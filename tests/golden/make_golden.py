"""Regenerates tests/golden/: run here (where /root/reference exists).

x_ser/, xsmall_ser/ = `ser/` directories the oracle's construct writes for the reference's own
fixtures data/x.fa + data/x.vcf.gz and data/x.small.fa + data/x.small.vcf (CQF table 2^12 / 2^10
slots instead of the reference's 2^25 so the files stay small).  expected.json = the oracle's
answers on them; the first entries are pinned by the reference's README (README.md:58-60, :93-95).
"""
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import vs_testlib as T  # noqa: E402
from vs_testlib import Oracle  # noqa: E402



def consensus(fa, vcf, sample_col=9):
    """Independent of the oracle: the FASTA with every record applied whose genotype names an alt allele
    (what query_sample_from_ref returns over the whole contig when no two records overlap)."""
    import gzip
    ref = open(fa).read().split("\n", 1)[1].replace("\n", "")
    opener = gzip.open if vcf.endswith(".gz") else open
    pieces, cur = [], 0
    for line in opener(vcf, "rt"):
        if line.startswith("#"):
            continue
        r = line.rstrip("\n").split("\t")
        pos, alleles = int(r[1]) - 1, [int(g) for g in r[sample_col].split(":")[0].replace("/", "|").split("|") if g != "." and int(g) > 0]
        if not alleles:
            continue
        assert ref[pos:pos + len(r[3])] == r[3] and pos >= cur
        pieces += [ref[cur:pos], r[4].split(",")[alleles[0] - 1]]
        cur = pos + len(r[3])
    return ref, "".join(pieces) + ref[cur:]


out = {}
for name, fa, vcf, log2 in [("x", "x.fa", "x.vcf.gz", 12), ("xsmall", "x.small.fa", "x.small.vcf", 10)]:
    prefix = os.path.join(HERE, name + "_ser")
    shutil.rmtree(prefix, ignore_errors=True)
    o = Oracle.construct(os.path.join(T.REF_DATA, fa), os.path.join(T.REF_DATA, vcf), prefix, cqf_log2=log2)
    ref_len = o.info()["ref_length"]
    ex = {"construct_info": o.construct_info, "t6": {}, "t4": {}, "t7": {}}
    regions = [(10, 105), (14, 105), (9, 105), (1, ref_len + 1), (100, 104), (466, 470), (972, 1000), (660, 700), (1, 80), (25, 27), (500, 400)]
    for x, y in regions:
        ex["t6"][f"{x}:{y}"] = o.t6_text(x, y)
        ex["t4"][f"{x}:{y}"] = o.t4_text(x, y, "1")[0]
    # t2 (query_sample_from_ref) for sample "1": status 1 = the reference's substr throws
    ex["t2"] = []
    for x, y in regions + [(0, 5), (58, 60), (57, 60), (ref_len, ref_len + 50), (ref_len + 10, ref_len + 20)]:
        ln, dg, st, ub, seqs = o.batch_t2([x], [y], [1], want_text=True)
        ex["t2"].append({"x": x, "y": y, "status": int(st[0]), "seq": seqs[0]})
    if name == "x":       # single-sample fixture without overlapping records: pinned by an independent consensus
        ref, cons = consensus(os.path.join(T.REF_DATA, fa), os.path.join(T.REF_DATA, vcf))
        whole = [q for q in ex["t2"] if (q["x"], q["y"]) == (1, ref_len + 1)][0]
        assert whole["status"] == 0 and whole["seq"] == cons, "oracle t2 != VCF consensus"
        ex["t2_consensus_sha1"] = __import__("hashlib").sha1(cons.encode()).hexdigest()
        ex["t2_consensus"] = cons        # 1005 characters: lets the t3 / t5 pins run where /root/reference is absent
    for p, r, a in o.all_variants() + [(58, "G", "GT"), (11, "C", "T")]:
        ex["t7"][f"{p}|{r}|{a}"] = o.t7_text(p, r, a)
    out[name] = ex
    o.close()
json.dump(out, open(os.path.join(HERE, "expected.json"), "w"), indent=1, sort_keys=True)
print("golden fixtures written")
