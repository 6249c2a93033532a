"""Shared test helpers: builds, the oracle binding (tests may use the oracle; the product may not),
synthetic FASTA/VCF generators and parity comparison utilities."""
import ctypes as C
import os
import random
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
HOSTSIM_DIR = os.path.join(ROOT, "tests", "hostsim")
CSRC_DIR = os.path.join(ROOT, "variantstore_b200", "csrc")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
HOSTSIM_SO = os.path.join(HOSTSIM_DIR, "libvsgpu_hostsim.so")
VSGPU_SO = os.path.join(ROOT, "variantstore_b200", "libvsgpu.so")
REF_DATA = "/root/reference/data"
GOLDEN = os.path.join(ROOT, "tests", "golden")

_built = False


def _make(cwd, target=None):
    cmd = ["make", "-s"] + ([target] if target else [])
    subprocess.run(cmd, cwd=cwd, check=True, stdout=subprocess.DEVNULL)


def ensure_built():
    global _built
    if _built:
        return
    _make(ORACLE_DIR, "liboracle.so")
    if os.path.isdir("/root/reference/src/gqf"):
        _make(ORACLE_DIR, "ref")
    _make(HOSTSIM_DIR)
    if os.path.exists("/usr/local/cuda/bin/nvcc"):
        _make(CSRC_DIR, "../libvsgpu.so")
    _built = True


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# ----------------------------------------------------------------------------- oracle binding
class Oracle:
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            ensure_built()
            L = C.CDLL(ORACLE_SO)
            vp, u64 = C.c_void_p, C.c_uint64
            L.vso_last_error.restype = C.c_char_p
            L.vso_construct.restype = vp
            L.vso_construct.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
            L.vso_open.restype = vp
            L.vso_open.argtypes = [C.c_char_p, C.c_int]
            L.vso_close.argtypes = [vp]
            L.vso_info.argtypes = [vp, vp]
            L.vso_free.argtypes = [vp]
            for f in ("vso_query_t6_text", "vso_query_t4_text", "vso_query_t7_text", "vso_all_variants_text", "vso_sample_name"):
                getattr(L, f).restype = vp
            L.vso_query_t6_text.argtypes = [vp, u64, u64]
            L.vso_query_t4_text.argtypes = [vp, u64, u64, C.c_char_p, C.POINTER(C.c_int)]
            L.vso_query_t7_text.argtypes = [vp, u64, C.c_char_p, C.c_char_p]
            L.vso_all_variants_text.argtypes = [vp]
            L.vso_sample_name.argtypes = [vp, C.c_uint32]
            L.vso_batch_t6.argtypes = [vp, u64, vp, vp, vp, vp, C.c_int]
            L.vso_batch_t4.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp, C.c_int]
            L.vso_batch_t1.argtypes = [vp, u64, vp, vp, vp, vp, C.c_int]
            L.vso_batch_t2.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp, vp, C.POINTER(vp)]
            L.vso_batch_t5.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp, vp, C.c_int]
            L.vso_query_t5_text.restype = vp
            L.vso_query_t5_text.argtypes = [vp, u64, u64, C.c_char_p]
            L.vso_batch_t3.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp, vp, C.POINTER(vp)]
            L.vso_batch_t2_mt.argtypes = [vp, u64, vp, vp, vp, vp, C.c_int]
            L.vso_batch_t6_mt.argtypes = [vp, u64, vp, vp, vp, C.c_int]
            L.vso_batch_t4_mt.argtypes = [vp, u64, vp, vp, vp, vp, C.c_int]
            L.vso_batch_t7.argtypes = [vp, u64, vp, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), vp, vp, vp]
            L.vso_synth.restype = vp
            L.vso_synth.argtypes = [C.c_char_p, C.c_char_p, u64, u64, u64, u64, C.c_uint32, C.c_uint32, C.c_double, C.c_double,
                                    C.c_int, C.c_int, u64, C.c_int, C.c_int, C.c_int, C.c_double]
            L.vso_rrr_roundtrip.argtypes = [vp, u64, C.c_char_p]
            L.vso_encode_vertices.restype = vp
            L.vso_encode_vertices.argtypes = [u64, vp, vp, vp, vp, vp, vp, C.POINTER(u64)]
            L.vso_cqf_differential.argtypes = [u64, u64, C.c_uint32, C.c_int, C.c_char_p]
            cls._lib = L
        return cls._lib

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle: " + self.lib().vso_last_error().decode())
        self.h = C.c_void_p(handle)

    @classmethod
    def construct(cls, fasta, vcf, prefix, cqf_log2=12, use_ref_gqf=False, fix_idx=True, force_enc=-1, reopen=True):
        """`variantstore construct`.  With reopen (default) the returned oracle is a fresh load of the
        serialised directory, as `variantstore query` would see it: neighbour sets are rebuilt by
        insertion at load time (graph.h:162-171), so their iteration order differs from construct time."""
        o = cls(cls.lib().vso_construct(fasta.encode(), vcf.encode(), prefix.encode(), cqf_log2, int(use_ref_gqf), int(fix_idx), force_enc))
        if not reopen:
            return o
        o.construct_info = o.info()
        ci = o.construct_info
        o.close()
        r = cls.open(prefix, use_ref_gqf)
        r.construct_info = ci
        return r

    @classmethod
    def open(cls, prefix, use_ref_gqf=False):
        return cls(cls.lib().vso_open(prefix.encode(), int(use_ref_gqf)))

    @classmethod
    def synth(cls, prefix, chr_name="22", ref_length=200000, pos_lo=1000, pos_hi=None, n_records=5000, n_samples=64,
              fmax=40, frac_multi=0.0025, frac_indel=0.035, mode=0, overlap=0, seed=1, cqf_log2=18, fix_idx=False, gzip_level=1,
              reuse_prob=0.55):
        pos_hi = pos_hi or ref_length - 1000
        o = cls(cls.lib().vso_synth(prefix.encode(), chr_name.encode(), ref_length, pos_lo, pos_hi, n_records, n_samples, fmax,
                                    frac_multi, frac_indel, mode, overlap, seed, cqf_log2, int(fix_idx), gzip_level, reuse_prob))
        ci = o.info()
        o.close()
        r = cls.open(prefix)
        r.construct_info = ci
        return r

    def close(self):
        if self.h:
            self.lib().vso_close(self.h)
            self.h = None

    def info(self):
        a = np.zeros(12, np.uint64)
        self.lib().vso_info(self.h, a.ctypes.data_as(C.c_void_p))
        keys = ["cqf_vertices", "edges", "seq_length", "ref_length", "num_samples", "classes", "vertices", "index_ones",
                "num_vars", "num_mutations", "num_mutations_samples", "use_bit_vector"]
        return dict(zip(keys, (int(v) for v in a)))

    def _text(self, p):
        if not p:
            raise RuntimeError("oracle: " + self.lib().vso_last_error().decode())
        try:
            return C.string_at(p).decode()
        finally:
            self.lib().vso_free(p)

    def t6_text(self, x, y):
        return self._text(self.lib().vso_query_t6_text(self.h, x, y))

    def t4_text(self, x, y, sample):
        ub = C.c_int(0)
        return self._text(self.lib().vso_query_t4_text(self.h, x, y, sample.encode(), C.byref(ub))), bool(ub.value)

    def t7_text(self, pos, ref, alt):
        return self._text(self.lib().vso_query_t7_text(self.h, pos, ref.encode(), alt.encode()))

    def all_variants(self):
        rows = []
        for line in self._text(self.lib().vso_all_variants_text(self.h)).split("\n"):
            if line:
                pos, ref, alt, _ = line.split("\t")
                rows.append((int(pos), ref, alt))
        return rows

    def sample_name(self, sid):
        return self._text(self.lib().vso_sample_name(self.h, sid))

    def batch_t6(self, x, y, with_samples=True):
        x, y = np.ascontiguousarray(x, np.uint64), np.ascontiguousarray(y, np.uint64)
        n = len(x)
        cnt, dig = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        rc = self.lib().vso_batch_t6(self.h, n, x.ctypes.data, y.ctypes.data, cnt.ctypes.data, dig.ctypes.data, int(with_samples))
        if rc != 0:
            raise RuntimeError("oracle: " + self.lib().vso_last_error().decode())
        return cnt, dig

    def batch_t4(self, x, y, sample_ids, with_samples=True):
        x, y = np.ascontiguousarray(x, np.uint64), np.ascontiguousarray(y, np.uint64)
        s = np.ascontiguousarray(sample_ids, np.uint32)
        n = len(x)
        cnt, dig, ub = np.zeros(n, np.uint64), np.zeros(n, np.uint64), np.zeros(n, np.uint8)
        rc = self.lib().vso_batch_t4(self.h, n, x.ctypes.data, y.ctypes.data, s.ctypes.data, cnt.ctypes.data, dig.ctypes.data, ub.ctypes.data, int(with_samples))
        if rc != 0:
            raise RuntimeError("oracle: " + self.lib().vso_last_error().decode())
        return cnt, dig, ub

    def t5_text(self, x, y, sample):
        return self._text(self.lib().vso_query_t5_text(self.h, x, y, sample.encode()))

    def batch_t5(self, x, y, sample_ids, with_samples=True):
        """get_sample_var_in_sample: (counts, digests, status, ub); status 2 = the reference never returns."""
        x, y = np.ascontiguousarray(x, np.uint64), np.ascontiguousarray(y, np.uint64)
        s = np.ascontiguousarray(sample_ids, np.uint32)
        n = len(x)
        cnt, dig, st, ub = np.zeros(n, np.uint64), np.zeros(n, np.uint64), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        rc = self.lib().vso_batch_t5(self.h, n, x.ctypes.data, y.ctypes.data, s.ctypes.data, cnt.ctypes.data, dig.ctypes.data, st.ctypes.data, ub.ctypes.data, int(with_samples))
        if rc != 0:
            raise RuntimeError("oracle: " + self.lib().vso_last_error().decode())
        return cnt, dig, st, ub

    def batch_t1(self, pos, with_samples=True):
        pos = np.ascontiguousarray(pos, np.uint64)
        n = len(pos)
        found, cnt, dig = np.zeros(n, np.uint8), np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        rc = self.lib().vso_batch_t1(self.h, n, pos.ctypes.data, found.ctypes.data, cnt.ctypes.data, dig.ctypes.data, int(with_samples))
        if rc != 0:
            raise RuntimeError("oracle: " + self.lib().vso_last_error().decode())
        return found, cnt, dig

    def batch_t3(self, x, y, sample_ids, want_text=False):
        """query_sample_from_sample: like batch_t2; status 2 = the reference never leaves its loop."""
        return self.batch_t2(x, y, sample_ids, want_text, fn="vso_batch_t3")

    def batch_t2(self, x, y, sample_ids, want_text=False, fn="vso_batch_t2"):
        """query_sample_from_ref: (lengths, digests, status, ub[, list of sequences]); status 1 = the
        reference call ends in std::out_of_range."""
        x, y = np.ascontiguousarray(x, np.uint64), np.ascontiguousarray(y, np.uint64)
        s = np.ascontiguousarray(sample_ids, np.uint32)
        n = len(x)
        ln, dig, st, ub = np.zeros(n, np.uint64), np.zeros(n, np.uint64), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        tp = C.c_void_p(0)
        rc = getattr(self.lib(), fn)(self.h, n, x.ctypes.data, y.ctypes.data, s.ctypes.data, ln.ctypes.data, dig.ctypes.data,
                                     st.ctypes.data, ub.ctypes.data, C.byref(tp) if want_text else None)
        if rc != 0:
            raise RuntimeError("oracle: " + self.lib().vso_last_error().decode())
        if want_text:
            seqs = self._text(tp.value).split("\n")[:n]
            return ln, dig, st, ub, seqs
        return ln, dig, st, ub

    def timed_counts(self, qtype, x, y, sample_ids=None, nthreads=1):
        """Counts only, optionally over several worker threads (bench timing arms)."""
        x, y = np.ascontiguousarray(x, np.uint64), np.ascontiguousarray(y, np.uint64)
        cnt = np.zeros(len(x), np.uint64)
        if qtype == 6:
            rc = self.lib().vso_batch_t6_mt(self.h, len(x), x.ctypes.data, y.ctypes.data, cnt.ctypes.data, nthreads)
        elif qtype == 2:
            s = np.ascontiguousarray(sample_ids, np.uint32)
            rc = self.lib().vso_batch_t2_mt(self.h, len(x), x.ctypes.data, y.ctypes.data, s.ctypes.data, cnt.ctypes.data, nthreads)
        else:
            s = np.ascontiguousarray(sample_ids, np.uint32)
            rc = self.lib().vso_batch_t4_mt(self.h, len(x), x.ctypes.data, y.ctypes.data, s.ctypes.data, cnt.ctypes.data, nthreads)
        if rc != 0:
            raise RuntimeError("oracle batch failed")
        return cnt

    def batch_t7(self, pos, refs, alts):
        pos = np.ascontiguousarray(pos, np.uint64)
        n = len(pos)
        ra = (C.c_char_p * n)(*[r.encode() for r in refs])
        aa = (C.c_char_p * n)(*[a.encode() for a in alts])
        found, cnt, dig = np.zeros(n, np.uint8), np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        rc = self.lib().vso_batch_t7(self.h, n, pos.ctypes.data, ra, aa, found.ctypes.data, cnt.ctypes.data, dig.ctypes.data)
        if rc != 0:
            raise RuntimeError("oracle: " + self.lib().vso_last_error().decode())
        return found, cnt, dig


# ----------------------------------------------------------------------------- synthetic VCF text
def write_fuzz_inputs(dirpath, seed, ref_len=4000, n_records=260, n_samples=12, overlap=False, sparse=False, chrom="f", name_fmt="S{:03d}"):
    """The fuzz shapes of SURVEY.md §4: SNPs (some bi-allelic), small insertions and deletions with
    abutting / adjacent sites; `overlap` lets a record start inside the previous record's REF span."""
    rnd = random.Random(seed)
    ref = "".join(rnd.choice("ACGT") for _ in range(ref_len))
    names = [name_fmt.format(i) for i in range(1, n_samples + 1)]
    lines = ["##fileformat=VCFv4.1", '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">',
             "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names)]
    pos = rnd.randint(2, 12)
    made = 0
    while made < n_records and pos < ref_len - 10:
        u = rnd.random()
        if u < 0.70:
            r = ref[pos - 1]
            alts = [rnd.choice([b for b in "ACGT" if b != r])]
            if rnd.random() < 0.15:
                alts.append(rnd.choice([b for b in "ACGT" if b != r and b != alts[0]]))
        elif u < 0.85:
            r = ref[pos - 1]
            alts = [r + "".join(rnd.choice("ACGT") for _ in range(rnd.randint(1, 3)))]
        else:
            k = rnd.randint(1, 3)
            r = ref[pos - 1:pos + k]
            alts = [ref[pos - 1]]
        gts = []
        any_carrier = False
        for _ in names:
            if sparse:
                g = "0/1" if rnd.random() < 0.04 else "0/0"
            else:
                a, b = int(rnd.random() < 0.3), int(rnd.random() < 0.3)
                g = f"{a}|{b}"
            any_carrier |= g not in ("0|0", "0/0")
            gts.append(g)
        if not any_carrier:
            gts[rnd.randrange(len(gts))] = "0/1" if sparse else "1|0"
        lines.append(f"{chrom}\t{pos}\t.\t{r}\t{','.join(alts)}\t99\t.\t.\tGT\t" + "\t".join(gts))
        made += 1
        if overlap:
            pos += rnd.choice([1, 1, 2, 3, 5, 9, 17, 30, len(r), len(r) + 1])
        else:
            pos += len(r) + rnd.choice([0, 0, 1, 1, 2, 3, 5, 9, 17, 30])
    os.makedirs(dirpath, exist_ok=True)
    fa, vcf = os.path.join(dirpath, "ref.fa"), os.path.join(dirpath, "in.vcf")
    with open(fa, "w") as f:
        f.write(f">{chrom}\n")
        for i in range(0, ref_len, 80):
            f.write(ref[i:i + 80] + "\n")
    with open(vcf, "w") as f:
        f.write("\n".join(lines) + "\n")
    return fa, vcf, names


def random_regions(seed, n, ref_len, widths=(1, 2, 5, 20, 100, 1000), n_samples=12):
    rnd = random.Random(seed)
    x = np.array([rnd.randint(1, ref_len) for _ in range(n)], np.uint64)
    w = np.array([rnd.choice(widths) for _ in range(n)], np.uint64)
    s = np.array([rnd.randint(1, n_samples) for _ in range(n)], np.uint32)
    return x, x + w, s


def open_engine(prefix, backend):
    """backend 'hostsim' (CPU, test-only) or 'cuda' (libvsgpu.so through the C ABI)."""
    from variantstore_b200 import VariantStoreIndex, load_library
    ensure_built()
    if backend == "hostsim":
        return VariantStoreIndex(prefix, lib=load_library(HOSTSIM_SO, subset=True))
    return VariantStoreIndex(prefix, device=0)


def compare_t1(oracle, eng, pos, with_samples=True):
    """closest_var on both sides; returns mismatching indices."""
    f, c, d = oracle.batch_t1(pos, with_samples)
    lo, hi = eng.batch_closest_var(pos)
    ec, ed = eng.digest_t1(lo, hi, with_samples)
    efound = lo != 0xFFFFFFFF
    bad = (efound != (f == 1)) | ((f == 1) & ((c != ec) | (d != ed)))
    return [int(i) for i in np.nonzero(bad)[0]]


def compare_t5(oracle, eng, x, y, s, with_samples=True, skip_ub=True):
    """get_sample_var_in_sample on both sides (row counts, row digests, "never returns" flags); returns
    (mismatching indices, regions the reference hangs on)."""
    oc, od, ost, ub = oracle.batch_t5(x, y, s, with_samples)
    off, hits, est, _ = eng.batch_sample_var_in_sample(x, y, s)
    ed = eng.digest_t5(off, hits, s, with_samples)
    ec = np.diff(off)
    ok = ost == 0
    mism = (ost != est) | (ok & ((oc != ec) | (od != ed))) | (~ok & (ec != 0))
    if skip_ub:
        mism &= ub == 0
    return [int(i) for i in np.nonzero(mism)[0]], int((ost != 0).sum())


def compare_t3(oracle, eng, x, y, s, skip_ub=True):
    """query_sample_from_sample on both sides; returns (mismatching indices, regions where the reference
    throws or never returns)."""
    return compare_t2(oracle, eng, x, y, s, skip_ub, t3=True)


def compare_t2(oracle, eng, x, y, s, skip_ub=True, t3=False):
    """query_sample_from_ref on both sides, byte for byte; returns (mismatching indices, regions where
    the reference throws std::out_of_range)."""
    if t3:
        ln, dg, st, ub, seqs = oracle.batch_t3(x, y, s, want_text=True)
        off, text, est, _ = eng.batch_sample_seq_in_sample(x, y, s)
    else:
        ln, dg, st, ub, seqs = oracle.batch_t2(x, y, s, want_text=True)
        off, text, est, _ = eng.batch_sample_seq_in_ref(x, y, s)
    bad = []
    for i in range(len(seqs)):
        if skip_ub and ub[i]:
            continue
        if est[i] != st[i] or int(off[i + 1] - off[i]) != int(ln[i]) or text[off[i]:off[i + 1]] != seqs[i].encode():
            bad.append(i)
    return bad, int((st != 0).sum())


def compare_all(oracle, eng, x, y, s, with_samples=True, skip_ub=True):
    """Run t6 and t4 on both sides; return lists of mismatching query indices."""
    oc6, od6 = oracle.batch_t6(x, y, with_samples)
    lo, hi, cnt = eng.batch_var_in_ref(x, y)
    ed6 = eng.digest_t6(lo, hi, with_samples)
    bad6 = [int(i) for i in np.nonzero((oc6 != cnt) | (od6 != ed6))[0]]
    oc4, od4, ub = oracle.batch_t4(x, y, s, with_samples)
    off, hits = eng.batch_sample_var_in_ref(x, y, s)
    ec4 = np.diff(off)
    ed4 = eng.digest_t4(off, hits, with_samples)
    mism = (oc4 != ec4) | (od4 != ed4)
    if skip_ub:
        mism &= ub == 0
    bad4 = [int(i) for i in np.nonzero(mism)[0]]
    # the fused pass (t6 slice from the two ranks of the t4 walk, one kernel) must give the very same arrays,
    # through the 64-bit and the 32-bit coordinate entry points
    for xs, ys in ((x, y), (np.minimum(x, 0xFFFFFFFF).astype(np.uint32), np.minimum(y, 0xFFFFFFFF).astype(np.uint32))):
        if xs.dtype == np.uint32 and (np.any(x > 0xFFFFFFFF) or np.any(y > 0xFFFFFFFF)):
            continue
        flo, fhi, fcnt, foff, fhits, fc4 = eng.batch_var_and_sample_var_in_ref(xs, ys, s)
        assert np.array_equal(flo, lo) and np.array_equal(fhi, hi) and np.array_equal(fcnt, cnt), "fused t6 differs from vsgpu_query_t6"
        assert np.array_equal(foff, off) and np.array_equal(fhits, hits) and np.array_equal(fc4, np.diff(off).astype(np.uint32)), "fused t4 differs from vsgpu_query_t4"
    return bad6, bad4, int(ub.sum())
