"""ctypes binding of include/vsgpu.h (one prototype per exported symbol)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libvsgpu.so")


class InfoT(C.Structure):
    _fields_ = [
        ("ref_length", C.c_uint64), ("seq_length", C.c_uint64), ("num_vertices_cqf", C.c_uint64),
        ("num_vertices", C.c_uint64), ("num_samples", C.c_uint32), ("num_classes", C.c_uint32),
        ("class_mode", C.c_uint32), ("backbone_vertices", C.c_uint32), ("distinct_starts", C.c_uint32),
        ("branch_records", C.c_uint32), ("walk_entries", C.c_uint32), ("has_suspect_dups", C.c_uint32),
        ("device_bytes", C.c_uint64), ("chr", C.c_char * 64), ("walk_markers", C.c_uint32), ("rejoin_carriers", C.c_uint32), ("from_cache", C.c_uint32),
    ]


u64p, u32p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
vp, cpp = C.c_void_p, C.POINTER(C.c_char_p)

# symbol -> (restype, argtypes); every symbol include/vsgpu.h declares
PROTOTYPES = {
    "vsgpu_open": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(vp)]),
    "vsgpu_close": (None, [vp]),
    "vsgpu_last_error": (C.c_char_p, []),
    "vsgpu_info": (C.c_int, [vp, C.POINTER(InfoT)]),
    "vsgpu_set_stream": (C.c_int, [vp, vp]),
    "vsgpu_sample_id": (C.c_int, [vp, C.c_char_p, u32p]),
    "vsgpu_sample_name": (C.c_char_p, [vp, C.c_uint32]),
    "vsgpu_query_t6": (C.c_int, [vp, C.c_uint64, vp, vp, vp, vp, vp]),
    "vsgpu_query_t4": (C.c_int, [vp, C.c_uint64, vp, vp, vp, C.POINTER(vp)]),
    "vsgpu_query_t6_u32": (C.c_int, [vp, C.c_uint64, vp, vp, vp, vp, vp]),
    "vsgpu_query_t4_u32": (C.c_int, [vp, C.c_uint64, vp, vp, vp, C.POINTER(vp)]),
    "vsgpu_result_num_queries": (C.c_uint64, [vp]),
    "vsgpu_result_offsets": (u64p, [vp]),
    "vsgpu_result_counts": (u32p, [vp]),
    "vsgpu_result_total": (C.c_uint64, [vp]),
    "vsgpu_query_t6t4": (C.c_int, [vp, C.c_uint64, vp, vp, vp, vp, vp, vp, C.POINTER(vp)]),
    "vsgpu_query_t6t4_u32": (C.c_int, [vp, C.c_uint64, vp, vp, vp, vp, vp, vp, C.POINTER(vp)]),
    "vsgpu_result_hits": (u32p, [vp]),
    "vsgpu_result_free": (None, [vp]),
    "vsgpu_query_t1": (C.c_int, [vp, C.c_uint64, vp, vp, vp]),
    "vsgpu_rows_t1": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(vp), u64p]),
    "vsgpu_digest_t1": (C.c_int, [vp, C.c_uint64, vp, vp, C.c_int, vp, vp]),
    "vsgpu_query_t7": (C.c_int, [vp, C.c_uint64, vp, cpp, cpp, vp]),
    "vsgpu_rows_t6": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(vp), u64p]),
    "vsgpu_rows_t4": (C.c_int, [vp, vp, C.c_uint64, C.c_int, C.POINTER(vp)]),
    "vsgpu_rows_t7": (C.c_int, [vp, C.c_uint32, C.POINTER(vp), u64p]),
    "vsgpu_free": (None, [vp]),
    "vsgpu_digest_t6": (C.c_int, [vp, C.c_uint64, vp, vp, C.c_int, vp]),
    "vsgpu_digest_t4": (C.c_int, [vp, C.c_uint64, vp, vp, C.c_int, vp]),
    "vsgpu_digest_t7": (C.c_int, [vp, C.c_uint64, vp, vp, vp]),
    "vsgpu_render_t6": (C.c_int, [vp, C.c_uint64, vp, vp, C.c_int, C.POINTER(vp)]),
    "vsgpu_render_t4": (C.c_int, [vp, C.c_uint64, vp, vp, vp, C.c_int, C.POINTER(vp)]),
    "vsgpu_render_t5": (C.c_int, [vp, C.c_uint64, vp, vp, vp, C.c_int, C.POINTER(vp)]),
    "vsgpu_text_bytes": (vp, [vp]),
    "vsgpu_text_offsets": (u64p, [vp]),
    "vsgpu_text_num_rows": (C.c_uint64, [vp]),
    "vsgpu_text_kernel_ms": (C.c_float, [vp]),
    "vsgpu_text_free": (None, [vp]),
    "vsgpu_query_t2": (C.c_int, [vp, C.c_uint64, vp, vp, vp, C.POINTER(vp)]),
    "vsgpu_query_t3": (C.c_int, [vp, C.c_uint64, vp, vp, vp, C.POINTER(vp)]),
    "vsgpu_query_t5": (C.c_int, [vp, C.c_uint64, vp, vp, vp, C.POINTER(vp)]),
    "vsgpu_result_status": (vp, [vp]),
    "vsgpu_result_kernel_ms": (C.c_float, [vp]),
    "vsgpu_rows_t5": (C.c_int, [vp, vp, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(vp)]),
    "vsgpu_digest_t5": (C.c_int, [vp, C.c_uint64, vp, vp, vp, C.c_int, vp]),
    "vsgpu_text_status": (vp, [vp]),
    "vsgpu_text_stage_ms": (C.POINTER(C.c_float), [vp]),
    "vsgpu_batch_create": (C.c_int, [vp, C.c_int, C.c_uint64, vp, vp, vp, cpp, cpp, C.POINTER(vp)]),
    "vsgpu_batch_run": (C.c_int, [vp]),
    "vsgpu_batch_fetch": (C.c_int, [vp, vp, vp, vp, C.POINTER(vp)]),
    "vsgpu_batch_stats": (C.c_int, [vp, u64p, u32p]),
    "vsgpu_batch_timings": (C.c_int, [vp, vp, C.c_uint32, u32p]),
    "vsgpu_batch_free": (None, [vp]),
    "vsgpu_router_open": (C.c_int, [C.c_uint32, cpp, vp, vp, vp, C.c_int, C.POINTER(vp)]),
    "vsgpu_router_close": (None, [vp]),
    "vsgpu_router_last_error": (C.c_char_p, []),
    "vsgpu_router_num_shards": (C.c_uint32, [vp]),
    "vsgpu_router_num_contigs": (C.c_uint32, [vp]),
    "vsgpu_router_contig_name": (C.c_char_p, [vp, C.c_uint32]),
    "vsgpu_router_contig_id": (C.c_int, [vp, C.c_char_p, u32p]),
    "vsgpu_router_shard_index": (vp, [vp, C.c_uint32]),
    "vsgpu_router_shard_device": (C.c_int, [vp, C.c_uint32]),
    "vsgpu_router_query_t6t4": (C.c_int, [vp, C.c_uint64, vp, vp, vp, vp, vp, vp, vp, vp]),
    "vsgpu_router_offsets": (u64p, [vp]),
    "vsgpu_router_hits": (u32p, [vp]),
    "vsgpu_router_region_hits": (C.c_int, [vp, C.c_uint64, C.POINTER(u32p), u32p]),
    "vsgpu_router_stats": (C.c_int, [vp, C.c_uint32, vp, vp, vp, u32p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}
# subset a test-only host simulator has to provide
QUERY_SUBSET = [s for s in PROTOTYPES if not s.startswith(("vsgpu_batch", "vsgpu_render", "vsgpu_router")) and s != "vsgpu_set_stream"]


def load(path=None, subset=False):
    """dlopen libvsgpu (or a library exporting the same symbols) and attach prototypes."""
    path = path or DEFAULT_LIB
    if not os.path.exists(path):
        raise OSError(f"{path} not found — build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(path)
    for name in (QUERY_SUBSET if subset else PROTOTYPES):
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype, fn.argtypes = PROTOTYPES[name]
    return lib
