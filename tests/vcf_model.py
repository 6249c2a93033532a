"""An oracle-independent model of what t6 / t4 / t7 must return, derived from the VCF + FASTA text alone.

Nothing in here imports the oracle, the engine or vs_testlib: it restates, from the reference's
construct rules, what each VCF allele becomes (variant_graph.h:1509-1532 normalisation; GT parsing
:655-700; phasing strings :882-900) and which rows the three operators owe for it:

  * t6 `get_var_in_ref` over the whole contig lists every allele once, by position
    (query.h:336-392: a deletion as (first deleted base, deleted bases, ""), an insertion as
    (base before the inserted ones, "", inserted bases), a substitution as (pos, ref, alt)),
    each with all its carriers as name(gt1|gt2);
  * t4 `get_sample_var_in_ref` over the whole contig lists the alleles the sample carries
    (query.h:677-710), same row text except where the reference itself is order-dependent
    (documented below and asserted as such, not skipped);
  * t7 `samples_has_var` of an allele the operator can find returns that allele's carriers
    as "name gt" pairs (query.h:807-816).

The model only speaks for VCFs whose records leave at least `MIN_GAP` reference bases between one
record's REF span and the next record's POS (no abutting or overlapping sites, which is where the
reference's dummy-vertex and sibling-order quirks live — SURVEY.md section 3.2) and whose header is
name-sorted.  Multi-allelic records may list several SNP alts; two construct rules are then part of the
model: "any non-zero GT digit adds the sample to every alt of the record" (variant_graph.h:655-691), and
the first alt of such a record is printed with an empty ref (see alleles()).
"""
import gzip

MIN_GAP = 2


def read_fasta(path):
    return open(path).read().split("\n", 1)[1].replace("\n", "")


def read_vcf(path):
    names, recs = [], []
    opener = gzip.open if path.endswith(".gz") else open
    for line in opener(path, "rt"):
        if line.startswith("##"):
            continue
        f = line.rstrip("\n").split("\t")
        if line.startswith("#"):
            names = f[9:]
            continue
        recs.append((int(f[1]), f[3], f[4].split(","), [g.split(":")[0] for g in f[9:]]))
    return names, recs


def carriers_of(names, gts):
    """(name, phasing) of every sample with a non-zero digit on either side of a 3-character GT, or a
    non-zero single-character GT (haploid: gt1 = 1, gt2 = 0, unphased), in header order."""
    out = []
    for name, g in zip(names, gts):
        if len(g) == 3 and g[0].isdigit() and g[2].isdigit():
            a, b = int(g[0]), int(g[2])
            if a > 0 or b > 0:
                out.append((name, f"{int(a > 0)}{g[1]}{int(b > 0)}"))
        elif len(g) == 1 and g.isdigit() and int(g):
            out.append((name, "1/0"))
    return out


def alleles(fa, vcf, strict=True):
    """Every (record, alt) as the row the operators print: dict(pos, ref, alt, kind, carriers, site, nsite).
    strict=False admits abutting records (the reference's own fixture has some); the checks then still hold
    for it, as test_vcf_pins shows, but the model does not claim them in general."""
    ref = read_fasta(fa)
    names, recs = read_vcf(vcf)
    assert names == sorted(names), "the model needs a name-sorted header (vcflib iterates samples by name)"
    out, prev_end = [], 0
    for site, (pos, r, alts, gts) in enumerate(recs):
        assert ref[pos - 1:pos - 1 + len(r)] == r, "REF does not match the FASTA"
        assert not strict or pos >= prev_end + MIN_GAP, "records abut or overlap: outside the model"
        assert pos >= prev_end, "records overlap: outside the model"
        prev_end = pos + len(r)
        car = carriers_of(names, gts)
        if not car:
            continue
        for a in alts:
            if len(r) == len(a):
                # Quirk of the reference, part of the model: the second alt of a record splits the site's reference
                # vertex again and leaves a zero-length dummy vertex in front of it (variant_graph.h:1564-1567,
                # :1584-1601); the FIRST alt then hangs off the vertex before that dummy, and both operators print
                # "the next backbone vertex" as its ref — the dummy, i.e. "" (query.h:378-392, :660-674).
                row = dict(pos=pos, ref=r if (len(alts) == 1 or a != alts[0]) else "", alt=a, kind="sub")
            elif len(r) > len(a):                      # deletion: pos += |alt|, ref = ref[|alt|:]
                assert r.startswith(a)
                row = dict(pos=pos + len(a), ref=r[len(a):], alt="", kind="del")
            else:                                      # insertion: pos += |ref|, alt = alt[|ref|:]; printed at the base before
                assert a.startswith(r)
                row = dict(pos=pos + len(r) - 1, ref="", alt=a[len(r):], kind="ins")
            row.update(carriers=car, site=site, nsite=len(alts))
            out.append(row)
    return names, ref, out


def row_text(row, with_samples=True):
    s = f"{row['pos']}\t{row['ref']}\t{row['alt']}\t"
    if with_samples:
        s += "".join(f"{n}({g}) " for n, g in row["carriers"])
    return s


def parse_rows(text):
    """Rows of an operator's `-o` text (header line skipped when present) as (pos, ref, alt, carriers-string)."""
    out = []
    for line in text.split("\n"):
        if not line or line.startswith("Pos\t") or line.startswith("Number of variants"):
            continue
        p, r, a, c = line.split("\t")
        out.append((int(p), r, a, c))
    return out


def check_t6_whole_contig(rows, model_rows):
    """t6 over [1, len + 1): every allele exactly once, ascending position; alleles of one multi-allelic record
    may come in either order (std::unordered_set iteration order)."""
    want = sorted((m["pos"], m["site"], row_text(m)) for m in model_rows)
    got_pos = [r[0] for r in rows]
    assert got_pos == sorted(got_pos), "t6 rows are not in position order"
    assert len(rows) == len(want), (len(rows), len(want))
    site_of = {}
    for m in model_rows:
        site_of.setdefault(m["pos"], m["site"])
    got = sorted((r[0], site_of.get(r[0], -1), f"{r[0]}\t{r[1]}\t{r[2]}\t{r[3]}") for r in rows)
    assert got == want
    return len(want)


def check_t4_whole_contig(rows, model_rows, sample):
    """t4 of `sample` over [1, len + 1): one row per SITE the sample carries, ascending.  At a multi-allelic record
    the sample carries every alt (construct rule) and the walk takes the first carrying sibling in
    unordered_set order: exactly one of the record's alleles must be reported.  Substitutions and insertions
    print exactly the model row.  A deletion row is (first deleted base, deleted bases, "") when the walk's
    ref_pos is in step; when the deletion target is listed last among the origin's neighbours the reference
    reports the preceding backbone vertex instead (query.h:660-674, :689-697) — then alt is still "" and the
    row's position lies before the deleted span."""
    mine = [m for m in model_rows if any(n == sample for n, _ in m["carriers"])]
    sites = []
    for m in mine:
        if not sites or sites[-1][0]["site"] != m["site"]:
            sites.append([])
        sites[-1].append(m)
    assert len(rows) == len(sites), (sample, len(rows), len(sites))
    exact = 0
    for r, cands in zip(rows, sites):
        got = f"{r[0]}\t{r[1]}\t{r[2]}\t{r[3]}"
        if cands[0]["kind"] == "del":
            m = cands[0]
            assert r[2] == "" and r[3] == row_text(m).split("\t")[3]
            if r[0] == m["pos"]:
                assert r[1] == m["ref"]
                exact += 1
            else:
                assert r[0] < m["pos"]
        else:
            assert got in [row_text(m) for m in cands], (sample, got)
            exact += 1
    return len(sites), exact


def t7_expected(model_row):
    """samples_has_var output for an allele the operator finds: "name gt" pairs, no separator (query.h:810-813)."""
    return "".join(f"{n} {g}" for n, g in model_row["carriers"])
