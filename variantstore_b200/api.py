"""Host-side mirror of the reference's query operators (include/query.h) over libvsgpu.

Names, argument meaning and error behaviour follow the reference:
  get_var_in_ref(pos_x, pos_y)                 query.h:736-784   (t6)
  get_sample_var_in_ref(pos_x, pos_y, sample)  query.h:618-729   (t4)
  samples_has_var(pos, ref, alt)               query.h:792-823   (t7)
  query_sample_from_ref(pos_x, pos_y, sample)  query.h:120-189   (t2)
  query_sample_from_sample(...)                query.h:195-261   (t3)
  get_sample_var_in_sample(...)                query.h:490-612   (t5)
  closest_var(pos)                             query.h:441-483   (t1)
plus batched forms that take whole arrays of regions — the reason this engine exists.
Positions are 1-based, regions are [pos_x, pos_y).
"""
import ctypes as C
import weakref
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib

NONE = 0xFFFFFFFF


class VsgpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"vsgpu error {code}: {msg}")
        self.code = code


@dataclass
class Variant:                      # struct Variant, query.h:30-36
    var_pos: int
    ref: str
    alt: str
    samples: List[Tuple[str, str]] = field(default_factory=list)


def load_library(path: Optional[str] = None, subset: bool = False):
    return _lib.load(path, subset)


def read_regions(region: str) -> List[Tuple[int, int]]:
    """src/commands.cc:64-93 — comma list of beg[:end]; the result is sorted like the reference's."""
    out = []
    for tok in region.split(","):
        if ":" in tok:
            b, e = tok.split(":", 1)
            out.append((int(b), int(e)))
        else:
            out.append((int(tok), 0))
    out.sort()
    return out


def read_sequences(s: str) -> List[str]:
    """src/commands.cc:96-111"""
    return s.split(",")


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _is_u32(a):
    return isinstance(a, np.ndarray) and a.dtype == np.uint32


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _parse_rows(text: str) -> List[Variant]:
    rows = []
    for line in text.split("\n"):
        if not line:
            continue
        pos, ref, alt, samples = line.split("\t")
        carriers = []
        for tok in samples.split(" "):
            if tok:
                name, gt = tok[:-1].split("(")
                carriers.append((name, gt))
        rows.append(Variant(int(pos), ref, alt, carriers))
    return rows


class VariantStoreIndex:
    """`Index idx(prefix); VariantGraph vg(prefix, mode)` of query_main (commands.cc:116-132) in one
    object: loads ser/, flattens it and keeps it resident on one GPU."""

    def __init__(self, prefix: str, device: int = 0, lib=None, _borrowed=None):
        self._lib = lib or _lib.load()
        self._owned = _borrowed is None
        if _borrowed is None:
            h = C.c_void_p()
            rc = self._lib.vsgpu_open(prefix.encode(), device, C.byref(h))
            if rc != 0:
                raise VsgpuError(rc, self._lib.vsgpu_last_error().decode())
        else:
            h = C.c_void_p(_borrowed)          # a shard of a Router: the router closes it
        self._h = h
        self._batches = weakref.WeakSet()     # device-resident batches must be freed before the index they point into
        info = _lib.InfoT()
        self._lib.vsgpu_info(self._h, C.byref(info))
        self.info = info
        self.chr = info.chr.decode()

    def close(self):
        if getattr(self, "_h", None):
            for b in list(getattr(self, "_batches", ())):
                b.close()
            if self._owned:
                self._lib.vsgpu_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise VsgpuError(rc, self._lib.vsgpu_last_error().decode())

    def set_stream(self, cuda_stream: int):
        self._check(self._lib.vsgpu_set_stream(self._h, C.c_void_p(cuda_stream)))

    # ---- sample ids
    def sample_id(self, name: str) -> int:
        v = C.c_uint32()
        self._check(self._lib.vsgpu_sample_id(self._h, name.encode(), C.byref(v)))
        return v.value

    def sample_name(self, sid: int) -> str:
        return self._lib.vsgpu_sample_name(self._h, sid).decode()

    # ---- batched operators
    def batch_var_in_ref(self, x, y):
        """t6 over arrays: returns (rec_lo, rec_hi, counts).  uint32 arrays go through the 32-bit entry
        point (half the bytes over PCIe), anything else as 64-bit."""
        n = len(x)
        lo, hi, cnt = (np.zeros(n, np.uint32) for _ in range(3))
        if _is_u32(x) and _is_u32(y):
            x, y = np.ascontiguousarray(x), np.ascontiguousarray(y)
            self._check(self._lib.vsgpu_query_t6_u32(self._h, n, _ptr(x), _ptr(y), _ptr(lo), _ptr(hi), _ptr(cnt)))
            return lo, hi, cnt
        x, y = _u64(x), _u64(y)
        self._check(self._lib.vsgpu_query_t6(self._h, n, _ptr(x), _ptr(y), _ptr(lo), _ptr(hi), _ptr(cnt)))
        return lo, hi, cnt

    def batch_sample_var_in_ref(self, x, y, sample_ids):
        """t4 over arrays: returns (offsets[n+1], hit codes).  uint32 coordinate arrays use the 32-bit entry point."""
        s = np.ascontiguousarray(sample_ids, dtype=np.uint32)
        n = len(x)
        r = C.c_void_p()
        if _is_u32(x) and _is_u32(y):
            x, y = np.ascontiguousarray(x), np.ascontiguousarray(y)
            self._check(self._lib.vsgpu_query_t4_u32(self._h, n, _ptr(x), _ptr(y), _ptr(s), C.byref(r)))
        else:
            x, y = _u64(x), _u64(y)
            self._check(self._lib.vsgpu_query_t4(self._h, n, _ptr(x), _ptr(y), _ptr(s), C.byref(r)))
        try:
            off = np.ctypeslib.as_array(self._lib.vsgpu_result_offsets(r), shape=(n + 1,)).copy()
            total = int(off[-1])
            hits = np.ctypeslib.as_array(self._lib.vsgpu_result_hits(r), shape=(total,)).copy() if total else np.zeros(0, np.uint32)
        finally:
            self._lib.vsgpu_result_free(r)
        return off, hits

    def batch_var_and_sample_var_in_ref(self, x, y, sample_ids, want_hi=True):
        """t6 and t4 over the same regions from one fused pass (vsgpu_query_t6t4*): returns
        (rec_lo, rec_hi or None, counts6, offsets[n+1], hit codes, counts4)."""
        s = np.ascontiguousarray(sample_ids, dtype=np.uint32)
        n = len(x)
        lo, cnt = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        hi = np.zeros(n, np.uint32) if want_hi else None
        r = C.c_void_p()
        if _is_u32(x) and _is_u32(y):
            x, y = np.ascontiguousarray(x), np.ascontiguousarray(y)
            fn = self._lib.vsgpu_query_t6t4_u32
        else:
            x, y = _u64(x), _u64(y)
            fn = self._lib.vsgpu_query_t6t4
        self._check(fn(self._h, n, _ptr(x), _ptr(y), _ptr(s), _ptr(lo), _ptr(hi) if want_hi else None, _ptr(cnt), C.byref(r)))
        try:
            c4 = np.ctypeslib.as_array(self._lib.vsgpu_result_counts(r), shape=(n,)).copy() if n else np.zeros(0, np.uint32)
            off = np.ctypeslib.as_array(self._lib.vsgpu_result_offsets(r), shape=(n + 1,)).copy()
            total = int(self._lib.vsgpu_result_total(r))
            assert total == int(off[-1])
            hits = np.ctypeslib.as_array(self._lib.vsgpu_result_hits(r), shape=(total,)).copy() if total else np.zeros(0, np.uint32)
        finally:
            self._lib.vsgpu_result_free(r)
        return lo, hi, cnt, off, hits, c4

    def render_var_in_ref(self, x, y, with_samples=True):
        """t6 over arrays with the rows rendered on the device: (offsets[n+1], text bytes, rows,
        kernel ms).  Region i's rows (print_var lines, query.h:43-50) are text[offsets[i]:offsets[i+1]]."""
        x, y = _u64(x), _u64(y)
        n = len(x)
        t = C.c_void_p()
        self._check(self._lib.vsgpu_render_t6(self._h, n, _ptr(x), _ptr(y), int(with_samples), C.byref(t)))
        try:
            off = np.ctypeslib.as_array(self._lib.vsgpu_text_offsets(t), shape=(n + 1,)).copy()
            text = C.string_at(self._lib.vsgpu_text_bytes(t), int(off[-1]))
            rows = int(self._lib.vsgpu_text_num_rows(t))
            ms = float(self._lib.vsgpu_text_kernel_ms(t))
        finally:
            self._lib.vsgpu_text_free(t)
        return off, text, rows, ms

    def render_sample_var_in_ref(self, x, y, sample_ids, with_samples=True):
        """t4 over arrays with the rows rendered on the device (vsgpu_render_t4): (offsets[n+1], text bytes, rows, kernel ms)."""
        x, y = _u64(x), _u64(y)
        s = np.ascontiguousarray(sample_ids, dtype=np.uint32)
        n = len(x)
        t = C.c_void_p()
        self._check(self._lib.vsgpu_render_t4(self._h, n, _ptr(x), _ptr(y), _ptr(s), int(with_samples), C.byref(t)))
        try:
            off = np.ctypeslib.as_array(self._lib.vsgpu_text_offsets(t), shape=(n + 1,)).copy()
            text = C.string_at(self._lib.vsgpu_text_bytes(t), int(off[-1]))
            rows = int(self._lib.vsgpu_text_num_rows(t))
            ms = float(self._lib.vsgpu_text_kernel_ms(t))
        finally:
            self._lib.vsgpu_text_free(t)
        return off, text, rows, ms

    def render_sample_var_in_sample(self, x, y, sample_ids, with_samples=True):
        """t5 over arrays with the rows rendered on the device (vsgpu_render_t5): (offsets[n+1], text bytes, rows, status[n],
        (count, write, rows) kernel ms); status 2 = the reference never returns (no rows)."""
        x, y = _u64(x), _u64(y)
        s = np.ascontiguousarray(sample_ids, dtype=np.uint32)
        n = len(x)
        t = C.c_void_p()
        self._check(self._lib.vsgpu_render_t5(self._h, n, _ptr(x), _ptr(y), _ptr(s), int(with_samples), C.byref(t)))
        try:
            off = np.ctypeslib.as_array(self._lib.vsgpu_text_offsets(t), shape=(n + 1,)).copy()
            text = C.string_at(self._lib.vsgpu_text_bytes(t), int(off[-1]))
            rows = int(self._lib.vsgpu_text_num_rows(t))
            status = np.frombuffer(C.string_at(self._lib.vsgpu_text_status(t), n), np.uint8).copy() if n else np.zeros(0, np.uint8)
            ms = tuple(float(v) for v in np.ctypeslib.as_array(self._lib.vsgpu_text_stage_ms(t), shape=(3,)))
        finally:
            self._lib.vsgpu_text_free(t)
        return off, text, rows, status, ms

    def batch_sample_var_in_sample(self, x, y, sample_ids):
        """t5 over arrays (get_sample_var_in_sample, query.h:490-612): (offsets[n+1], hit codes, status, kernel ms);
        status 2 = the reference never returns."""
        x, y = _u64(x), _u64(y)
        s = np.ascontiguousarray(sample_ids, dtype=np.uint32)
        n = len(x)
        r = C.c_void_p()
        self._check(self._lib.vsgpu_query_t5(self._h, n, _ptr(x), _ptr(y), _ptr(s), C.byref(r)))
        try:
            off = np.ctypeslib.as_array(self._lib.vsgpu_result_offsets(r), shape=(n + 1,)).copy()
            total = int(off[-1])
            hits = np.ctypeslib.as_array(self._lib.vsgpu_result_hits(r), shape=(total,)).copy() if total else np.zeros(0, np.uint32)
            status = np.frombuffer(C.string_at(self._lib.vsgpu_result_status(r), n), np.uint8).copy() if n else np.zeros(0, np.uint8)
            ms = float(self._lib.vsgpu_result_kernel_ms(r))
        finally:
            self._lib.vsgpu_result_free(r)
        return off, hits, status, ms

    def digest_t5(self, offsets, hits, sample_ids, with_samples=True):
        n = len(offsets) - 1
        d = np.zeros(n, np.uint64)
        hits = np.ascontiguousarray(hits, np.uint32)
        s = np.ascontiguousarray(sample_ids, np.uint32)
        self._check(self._lib.vsgpu_digest_t5(self._h, n, _ptr(offsets), _ptr(hits), _ptr(s), int(with_samples), _ptr(d)))
        return d

    def get_sample_var_in_sample(self, pos_x: int, pos_y: int, sample_id: str) -> List[Variant]:
        """query.h:490-612.  RuntimeError where the reference would spin forever."""
        sid = self.sample_id(sample_id)
        off, hits, status, _ = self.batch_sample_var_in_sample([pos_x], [pos_y], [sid])
        if status[0] == 2:
            raise RuntimeError("the reference does not terminate for this region (query.h:505-510)")
        p = C.c_void_p()
        self._check(self._lib.vsgpu_rows_t5(self._h, _ptr(hits), len(hits), sid, 1, C.byref(p)))
        return _parse_rows(self._take_text(p))

    def batch_sample_seq_in_sample(self, x, y, sample_ids):
        """t3 over arrays (query_sample_from_sample, query.h:195-261): like batch_sample_seq_in_ref with the
        regions in the sample's own coordinates; status 2 = the reference never returns."""
        return self.batch_sample_seq_in_ref(x, y, sample_ids, _fn="vsgpu_query_t3")

    def query_sample_from_sample(self, pos_x: int, pos_y: int, sample_id: str) -> str:
        """query.h:195-261.  IndexError where substr throws; RuntimeError where the reference would spin forever."""
        off, text, status, _ = self.batch_sample_seq_in_sample([pos_x], [pos_y], [self.sample_id(sample_id)])
        if status[0] == 1:
            raise IndexError("basic_string::substr: __pos > this->size() (query.h:235,239)")
        if status[0] == 2:
            raise RuntimeError("the reference does not terminate for this region (query.h:209-214)")
        return text.decode()

    def batch_sample_seq_in_ref(self, x, y, sample_ids, _fn="vsgpu_query_t2"):
        """t2 over arrays (query_sample_from_ref, query.h:120-189): (offsets[n+1], bytes, status, kernel ms).
        Region i's sequence is bytes[offsets[i]:offsets[i+1]]; status[i] = 1 where the reference call
        ends in std::out_of_range."""
        x, y = _u64(x), _u64(y)
        s = np.ascontiguousarray(sample_ids, dtype=np.uint32)
        n = len(x)
        t = C.c_void_p()
        self._check(getattr(self._lib, _fn)(self._h, n, _ptr(x), _ptr(y), _ptr(s), C.byref(t)))
        try:
            off = np.ctypeslib.as_array(self._lib.vsgpu_text_offsets(t), shape=(n + 1,)).copy()
            text = C.string_at(self._lib.vsgpu_text_bytes(t), int(off[-1]))
            status = np.frombuffer(C.string_at(self._lib.vsgpu_text_status(t), n), np.uint8).copy() if n else np.zeros(0, np.uint8)
            ms = float(self._lib.vsgpu_text_kernel_ms(t))
        finally:
            self._lib.vsgpu_text_free(t)
        return off, text, status, ms

    def query_sample_from_ref(self, pos_x: int, pos_y: int, sample_id: str) -> str:
        """query.h:120-189.  Raises IndexError where the reference's substr throws std::out_of_range."""
        off, text, status, _ = self.batch_sample_seq_in_ref([pos_x], [pos_y], [self.sample_id(sample_id)])
        if status[0]:
            raise IndexError("basic_string::substr: __pos > this->size() (query.h:163,167)")
        return text.decode()

    def batch_closest_var(self, pos):
        """t1 over an array of positions: (rec_lo, rec_hi); both NONE where the operator returns false."""
        pos = _u64(pos)
        n = len(pos)
        lo, hi = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        self._check(self._lib.vsgpu_query_t1(self._h, n, _ptr(pos), _ptr(lo), _ptr(hi)))
        return lo, hi

    def digest_t1(self, lo, hi, with_samples=True):
        n = len(lo)
        c, d = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        self._check(self._lib.vsgpu_digest_t1(self._h, n, _ptr(lo), _ptr(hi), int(with_samples), _ptr(c), _ptr(d)))
        return c, d

    def closest_var(self, pos: int):
        """closest_var(vg, idx, pos, vars) of query.h:441-483: (found, rows)."""
        lo, hi = self.batch_closest_var([pos])
        if lo[0] == NONE:
            return False, []
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self._lib.vsgpu_rows_t1(self._h, int(lo[0]), int(hi[0]), 1, C.byref(p), C.byref(n)))
        return True, _parse_rows(self._take_text(p))

    def batch_samples_has_var(self, pos, refs: Sequence[str], alts: Sequence[str]):
        """t7 over arrays: returns record ids (NONE where the reference says "There is no such variant!")."""
        pos = _u64(pos)
        n = len(pos)
        ra = (C.c_char_p * n)(*[r.encode() for r in refs])
        aa = (C.c_char_p * n)(*[a.encode() for a in alts])
        rec = np.zeros(n, np.uint32)
        self._check(self._lib.vsgpu_query_t7(self._h, n, _ptr(pos), ra, aa, _ptr(rec)))
        return rec

    # ---- digests / text (parity tests, -v output)
    def digest_t6(self, lo, hi, with_samples=True):
        d = np.zeros(len(lo), np.uint64)
        self._check(self._lib.vsgpu_digest_t6(self._h, len(lo), _ptr(lo), _ptr(hi), int(with_samples), _ptr(d)))
        return d

    def digest_t4(self, offsets, hits, with_samples=True):
        n = len(offsets) - 1
        d = np.zeros(n, np.uint64)
        hits = np.ascontiguousarray(hits, np.uint32)
        self._check(self._lib.vsgpu_digest_t4(self._h, n, _ptr(offsets), _ptr(hits), int(with_samples), _ptr(d)))
        return d

    def digest_t7(self, rec):
        n = len(rec)
        d, c = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        self._check(self._lib.vsgpu_digest_t7(self._h, n, _ptr(rec), _ptr(c), _ptr(d)))
        return c, d

    def _take_text(self, p):
        try:
            return C.string_at(p).decode()
        finally:
            self._lib.vsgpu_free(p)

    def rows_t6_text(self, lo, hi, with_samples=True):
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self._lib.vsgpu_rows_t6(self._h, int(lo), int(hi), int(with_samples), C.byref(p), C.byref(n)))
        return self._take_text(p)

    def rows_t4_text(self, hits, with_samples=True):
        hits = np.ascontiguousarray(hits, np.uint32)
        p = C.c_void_p()
        self._check(self._lib.vsgpu_rows_t4(self._h, _ptr(hits), len(hits), int(with_samples), C.byref(p)))
        return self._take_text(p)

    def rows_t5_text(self, hits, sample_id: int, with_samples=True):
        hits = np.ascontiguousarray(hits, np.uint32)
        p = C.c_void_p()
        self._check(self._lib.vsgpu_rows_t5(self._h, _ptr(hits), len(hits), int(sample_id), int(with_samples), C.byref(p)))
        return self._take_text(p)

    def rows_t7_text(self, rec):
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self._lib.vsgpu_rows_t7(self._h, int(rec), C.byref(p), C.byref(n)))
        return self._take_text(p)

    # ---- single-region operators with the reference's signatures
    def get_var_in_ref(self, pos_x: int, pos_y: int) -> List[Variant]:
        lo, hi, cnt = self.batch_var_in_ref([pos_x], [pos_y])
        if cnt[0] == hi[0] - lo[0]:
            return _parse_rows(self.rows_t6_text(lo[0], hi[0]))
        # the literal dedup rule decided this region's rows (a repeated record in the slice, or a region past the contig end
        # over tail records: DESIGN.md section 9): they are not a prefix of the slice — the render path builds them exactly
        off, text, rows, _ = self.render_var_in_ref([pos_x], [pos_y])
        out = _parse_rows(text.decode())
        assert len(out) == int(cnt[0])
        return out

    def get_sample_var_in_ref(self, pos_x: int, pos_y: int, sample_id: str) -> List[Variant]:
        off, hits = self.batch_sample_var_in_ref([pos_x], [pos_y], [self.sample_id(sample_id)])
        return _parse_rows(self.rows_t4_text(hits))

    def samples_has_var(self, pos: int, ref: str, alt: str) -> List[Tuple[str, str]]:
        rec = self.batch_samples_has_var([pos], [ref], [alt])
        if rec[0] == NONE:
            return []                 # the reference logs "There is no such variant!" (query.h:820)
        c = self._lib
        ids = []
        text = self.rows_t7_text(rec[0])
        # "name gt" pairs are concatenated without a separator; gt is always 3 characters
        i = 0
        while i < len(text):
            sp = text.index(" ", i)
            ids.append((text[i:sp], text[sp + 1:sp + 4]))
            i = sp + 4
        return ids


class Router:
    """Several ser/ directories (whole contigs or position ranges of one) on the GPUs of this node behind one
    handle (vsgpu_router_*): the multi-contig front-end the reference does not have — it runs one process per
    contig (eval_data_records/evaluation.txt:34).  Routing, the per-GPU host threads and the scatter of the
    answers are C++ (csrc/router.cc); this class only marshals arrays."""

    def __init__(self, prefixes: Sequence[str], ranges: Optional[Sequence[Tuple[int, int]]] = None, devices: Optional[Sequence[int]] = None,
                 ndevices: int = 0, lib=None):
        self._lib = lib or _lib.load()
        n = len(prefixes)
        arr = (C.c_char_p * n)(*[p.encode() for p in prefixes])
        lo = np.array([r[0] for r in ranges], np.uint64) if ranges is not None else None
        hi = np.array([r[1] for r in ranges], np.uint64) if ranges is not None else None
        dev = np.array(devices, np.int32) if devices is not None else None
        h = C.c_void_p()
        rc = self._lib.vsgpu_router_open(n, arr, _ptr(lo) if lo is not None else None, _ptr(hi) if hi is not None else None,
                                         _ptr(dev) if dev is not None else None, int(ndevices), C.byref(h))
        if rc != 0:
            raise VsgpuError(rc, self._lib.vsgpu_router_last_error().decode())
        self._h = h
        self.num_shards = n
        self.contigs = [self._lib.vsgpu_router_contig_name(h, i).decode() for i in range(self._lib.vsgpu_router_num_contigs(h))]
        self.shard_device = [self._lib.vsgpu_router_shard_device(h, k) for k in range(n)]
        self._shards = {}

    def contig_ids(self, names) -> np.ndarray:
        table = {c: i for i, c in enumerate(self.contigs)}
        try:
            return np.fromiter((table[str(c)] for c in names), np.uint32, count=len(names))
        except KeyError as e:
            raise VsgpuError(-1, f"no shard holds contig {e.args[0]}")

    def shard(self, k: int) -> VariantStoreIndex:
        """Shard k as an index object (rows / digests of its record ids and hit codes); owned by the router."""
        if k not in self._shards:
            self._shards[k] = VariantStoreIndex("", lib=self._lib, _borrowed=self._lib.vsgpu_router_shard_index(self._h, k))
        return self._shards[k]

    def query_t6t4(self, contig_ids, x, y, sample_ids, csr=True):
        """(shard_of, rec_lo, counts6, counts4, offsets[n+1], hit codes) in the caller's region order; csr=False leaves the
        hit codes in the shards' results (offsets / hits come back None; hits_of(i) reads one region's)."""
        c = np.ascontiguousarray(contig_ids, np.uint32); x = np.ascontiguousarray(x, np.uint32); y = np.ascontiguousarray(y, np.uint32)
        s = np.ascontiguousarray(sample_ids, np.uint32)
        n = len(x)
        so, lo, c6, c4 = (np.empty(n, np.uint32) for _ in range(4))
        rc = self._lib.vsgpu_router_query_t6t4(self._h, n, _ptr(c), _ptr(x), _ptr(y), _ptr(s), _ptr(so), _ptr(lo), _ptr(c6), _ptr(c4))
        if rc != 0:
            raise VsgpuError(rc, self._lib.vsgpu_router_last_error().decode())
        if not csr:
            return so, lo, c6, c4, None, None
        off = np.ctypeslib.as_array(self._lib.vsgpu_router_offsets(self._h), shape=(n + 1,)).copy()
        total = int(off[-1])
        hits = np.ctypeslib.as_array(self._lib.vsgpu_router_hits(self._h), shape=(total,)).copy() if total else np.zeros(0, np.uint32)
        return so, lo, c6, c4, off, hits

    def hits_of(self, i: int) -> np.ndarray:
        p, c = _lib.u32p(), C.c_uint32()
        rc = self._lib.vsgpu_router_region_hits(self._h, int(i), C.byref(p), C.byref(c))
        if rc != 0:
            raise VsgpuError(rc, self._lib.vsgpu_router_last_error().decode())
        return np.ctypeslib.as_array(p, shape=(c.value,)).copy() if c.value else np.zeros(0, np.uint32)

    def stats(self):
        dev = np.zeros(16, np.int32); ms = np.zeros(16, np.float64); reg = np.zeros(16, np.uint64)
        nd = C.c_uint32(); r_ms, s_ms = C.c_double(), C.c_double()
        self._lib.vsgpu_router_stats(self._h, 16, _ptr(dev), _ptr(ms), _ptr(reg), C.byref(nd), C.byref(r_ms), C.byref(s_ms))
        k = nd.value
        return {"devices": dev[:k].tolist(), "device_ms": ms[:k].tolist(), "device_regions": reg[:k].astype(int).tolist(), "route_ms": r_ms.value, "scatter_ms": s_ms.value}

    def close(self):
        if getattr(self, "_h", None):
            for sh in self._shards.values():
                sh.close()
            self._lib.vsgpu_router_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    """Device-resident batch of regions (vsgpu_batch_*): the bench harness that replaces the timing
    loop of src/bm_query.cc:74-135.  Inputs and results stay in HBM; run() only enqueues kernels."""

    def __init__(self, index: VariantStoreIndex, qtype: int, x, y=None, sample_ids=None, refs=None, alts=None):
        self._ix, self._lib, self.type = index, index._lib, qtype
        x = _u64(x)
        self.n = len(x)
        y = _u64(y) if y is not None else None
        s = np.ascontiguousarray(sample_ids, np.uint32) if sample_ids is not None else None
        ra = (C.c_char_p * self.n)(*[r.encode() for r in refs]) if refs is not None else None
        aa = (C.c_char_p * self.n)(*[a.encode() for a in alts]) if alts is not None else None
        h = C.c_void_p()
        index._check(self._lib.vsgpu_batch_create(index._h, qtype, self.n, _ptr(x), _ptr(y) if y is not None else None,
                                                  _ptr(s) if s is not None else None, ra, aa, C.byref(h)))
        self._h = h
        index._batches.add(self)

    def run(self):
        self._ix._check(self._lib.vsgpu_batch_run(self._h))

    def timings_ms(self):
        ms = (C.c_float * 4)()
        n = C.c_uint32()
        self._ix._check(self._lib.vsgpu_batch_timings(self._h, ms, 4, C.byref(n)))
        return [ms[i] for i in range(n.value)]

    def stats(self):
        b, k = C.c_uint64(), C.c_uint32()
        self._ix._check(self._lib.vsgpu_batch_stats(self._h, C.byref(b), C.byref(k)))
        return b.value, k.value

    def fetch(self):
        n = self.n
        if self.type == 6:
            lo, hi, cnt = (np.zeros(n, np.uint32) for _ in range(3))
            self._ix._check(self._lib.vsgpu_batch_fetch(self._h, _ptr(lo), _ptr(hi), _ptr(cnt), None))
            return lo, hi, cnt
        if self.type == 7:
            rec = np.zeros(n, np.uint32)
            self._ix._check(self._lib.vsgpu_batch_fetch(self._h, _ptr(rec), None, None, None))
            return rec
        if self.type == 46:               # fused: (rec_lo, rec_hi, counts6, offsets, hits)
            lo, hi, cnt = (np.zeros(n, np.uint32) for _ in range(3))
            r = C.c_void_p()
            self._ix._check(self._lib.vsgpu_batch_fetch(self._h, _ptr(lo), _ptr(hi), _ptr(cnt), C.byref(r)))
            try:
                off = np.ctypeslib.as_array(self._lib.vsgpu_result_offsets(r), shape=(n + 1,)).copy()
                total = int(off[-1])
                hits = np.ctypeslib.as_array(self._lib.vsgpu_result_hits(r), shape=(total,)).copy() if total else np.zeros(0, np.uint32)
            finally:
                self._lib.vsgpu_result_free(r)
            return lo, hi, cnt, off, hits
        cnt = np.zeros(n, np.uint32)
        r = C.c_void_p()
        self._ix._check(self._lib.vsgpu_batch_fetch(self._h, None, None, _ptr(cnt), C.byref(r)))
        try:
            off = np.ctypeslib.as_array(self._lib.vsgpu_result_offsets(r), shape=(n + 1,)).copy()
            total = int(off[-1])
            hits = np.ctypeslib.as_array(self._lib.vsgpu_result_hits(r), shape=(total,)).copy() if total else np.zeros(0, np.uint32)
        finally:
            self._lib.vsgpu_result_free(r)
        return off, hits, cnt

    def close(self):
        if getattr(self, "_h", None):
            self._lib.vsgpu_batch_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
