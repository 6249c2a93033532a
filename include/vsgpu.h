/*
 * libvsgpu — B200-native batched region-query engine for VariantStore indexes.
 *
 * C ABI boundary.  The reference (Kingsford-Group/variantstore) has no plugin/FFI layer: its
 * query path is a set of C++ free functions over (VariantGraph*, Index*) called from the switch
 * in query_main (src/commands.cc:150-193).  Each entry point below names the reference interface
 * it replaces.  Plain pointers and sizes only; the caller owns every input array; result objects
 * are owned by the library until the matching *_free call.  Every function returns 0 on success
 * or a negative VSGPU_E* code; vsgpu_last_error() gives the message for the calling thread.
 * Nothing here writes to stdout — the front-end prints the reference's count lines.
 *
 * There is no CPU fallback: every query entry point fails with VSGPU_ENODEVICE when no CUDA
 * device is usable.
 */
#ifndef VSGPU_H_
#define VSGPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSGPU_OK 0
#define VSGPU_EINVAL (-1)      /* bad argument (e.g. region start < 1: the reference aborts, index.h:151-154) */
#define VSGPU_EIO (-2)         /* ser/ directory unreadable or malformed */
#define VSGPU_ESHAPE (-3)      /* graph shape the flattened form cannot represent */
#define VSGPU_ENODEVICE (-4)   /* no usable CUDA device / CUDA runtime error */
#define VSGPU_ENOMEM (-5)

#define VSGPU_NONE 0xFFFFFFFFu

/* t4 hit code layout (vsgpu_result_hits) */
#define VSGPU_HIT_START 0x80000000u   /* the walk started on this vertex: a substitution prints ref "" (query.h:650,706) */
#define VSGPU_HIT_REJOIN 0x40000000u  /* the row is the backbone vertex the alt edge rejoins */
#define VSGPU_HIT_ENTRY_MASK 0x3FFFFFFFu

typedef struct vsgpu_index vsgpu_index;
typedef struct vsgpu_result vsgpu_result;
typedef struct vsgpu_batch vsgpu_batch;

typedef struct vsgpu_info_t {
	uint64_t ref_length;        /* VariantGraph::get_ref_length  variant_graph.h:1215 */
	uint64_t seq_length;        /* VariantGraph::get_seq_length  variant_graph.h:1207 */
	uint64_t num_vertices_cqf;  /* Graph::get_num_vertices = "#Vertices" printed by query_main (commands.cc:134-136) */
	uint64_t num_vertices;      /* vertices in the protobuf blocks */
	uint32_t num_samples;       /* including "ref" */
	uint32_t num_classes;       /* sample-vector classes (0 in explicit-id mode) */
	uint32_t class_mode;        /* 1 = bit-vector encoding, 0 = explicit sample ids */
	uint32_t backbone_vertices; /* M */
	uint32_t distinct_starts;   /* D = set bits of index.sdsl */
	uint32_t branch_records;    /* R */
	uint32_t walk_entries;      /* compact t4 entries */
	uint32_t has_suspect_dups;
	uint64_t device_bytes;      /* HBM held by the flattened index */
	char chr[64];               /* VariantGraph::get_chr */
	uint32_t walk_markers;      /* walk entries that are out-of-step arrival markers (deletion target listed last) */
	uint32_t rejoin_carriers;   /* alt entries whose rejoin vertex carries samples itself (rows with VSGPU_HIT_REJOIN) */
	uint32_t from_cache;        /* 1: the flattened index came from VSGPU_INDEX_CACHE, ser/ was not decoded */
} vsgpu_info_t;

/* ---- lifecycle -------------------------------------------------------------------------------
 * Replaces `Index idx(prefix); VariantGraph vg(prefix, mode);` (src/commands.cc:116-132,
 * include/index.h:108-117, include/variant_graph.h:366-446, include/graph.h:149-172): loads the
 * serialised directory once, flattens it and uploads it to `device`. */
int vsgpu_open(const char* ser_prefix, int device, vsgpu_index** out);
void vsgpu_close(vsgpu_index* idx);
const char* vsgpu_last_error(void);
int vsgpu_info(const vsgpu_index* idx, vsgpu_info_t* out);
/* Launch kernels and copies on this CUDA stream (a cudaStream_t) instead of the library's own. */
int vsgpu_set_stream(vsgpu_index* idx, void* cuda_stream);

/* sampleid_map lookups (variant_graph.h:1230-1236, :1333-1338) */
int vsgpu_sample_id(const vsgpu_index* idx, const char* name, uint32_t* id);
const char* vsgpu_sample_name(const vsgpu_index* idx, uint32_t id);

/* ---- t6: get_var_in_ref(vg, idx, pos_x, pos_y) — include/query.h:736-784 -----------------------
 * For every region [x[i], y[i]) writes the slice [rec_lo[i], rec_hi[i]) of the branch-record table
 * (records in the order the reference pushes them) and counts[i] = rows the reference returns
 * ("Number of variants get_var_in_ref: N").  Where the reference's Index::is_empty gate fires
 * (index.h:150-166; it then prints the t4 label, query.h:746) rec_lo = rec_hi = VSGPU_NONE and the
 * count is 0.  rec_lo / rec_hi may be NULL.  Host buffers. */
int vsgpu_query_t6(vsgpu_index* idx, uint64_t n, const uint64_t* x, const uint64_t* y,
                   uint32_t* rec_lo, uint32_t* rec_hi, uint32_t* counts);

/* The same with 32-bit coordinates: region bounds are parsed with std::stoi (src/commands.cc:76-80), so
 * they fit, and half as many bytes cross PCIe; a small kernel widens them in HBM. */
int vsgpu_query_t6_u32(vsgpu_index* idx, uint64_t n, const uint32_t* x, const uint32_t* y,
                       uint32_t* rec_lo, uint32_t* rec_hi, uint32_t* counts);

/* ---- t4: get_sample_var_in_ref(vg, idx, pos_x, pos_y, sample) — include/query.h:618-729 ---------
 * Result = CSR: offsets[n+1] into hits[]; each hit is a walk-entry code (VSGPU_HIT_*), in the
 * order the reference pushes the rows.  sample_ids are sampleid_map ids (1..num_samples-1).  Id 0 ("ref") is refused
 * with VSGPU_EINVAL, like a region start < 1: the reference accepts the name but then reports every backbone vertex of the
 * region as a "deletion" row (the ref path carries "ref" everywhere, query.h:677-700) — an answer no caller can want; a
 * batch with such an entry fails as a whole, as the reference's process would on its first abort(). */
int vsgpu_query_t4(vsgpu_index* idx, uint64_t n, const uint64_t* x, const uint64_t* y,
                   const uint32_t* sample_ids, vsgpu_result** out);
int vsgpu_query_t4_u32(vsgpu_index* idx, uint64_t n, const uint32_t* x, const uint32_t* y,
                       const uint32_t* sample_ids, vsgpu_result** out);   /* 32-bit coordinates, see vsgpu_query_t6_u32 */
uint64_t vsgpu_result_num_queries(const vsgpu_result* r);
const uint64_t* vsgpu_result_offsets(const vsgpu_result* r);   /* n + 1 */
const uint32_t* vsgpu_result_counts(const vsgpu_result* r);    /* n: rows per region = offsets[i + 1] - offsets[i].  A host-buffer call brings back
                                                                  the counts (4 bytes a region over PCIe); whichever of offsets / counts did not travel is
                                                                  built on the host the first time it is asked for */
uint64_t vsgpu_result_total(const vsgpu_result* r);            /* rows over all regions = offsets[n] */
const uint32_t* vsgpu_result_hits(const vsgpu_result* r);
void vsgpu_result_free(vsgpu_result* r);                       /* before vsgpu_close of its index: the buffers go back to the index's page-locked pool */

/* ---- t6 + t4 over the same regions in one pass — the loop body of query_main (src/commands.cc:150-193) run for
 * `-t 6` and `-t 4` on one region list.  get_var_in_ref's slice comes from the two index ranks that
 * get_sample_var_in_ref needs anyway (Index::find / is_empty, index.h:119-166), so one kernel (k_t4p<kFuse6>)
 * answers both operators: x / y / sample_ids cross PCIe once, and per region rec_lo, counts6 (rec_hi optional: NULL)
 * and the t4 row count come back with the hit codes.  Same answers as vsgpu_query_t6 + vsgpu_query_t4. */
int vsgpu_query_t6t4(vsgpu_index* idx, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample_ids,
                     uint32_t* rec_lo, uint32_t* rec_hi, uint32_t* counts6, vsgpu_result** out);
int vsgpu_query_t6t4_u32(vsgpu_index* idx, uint64_t n, const uint32_t* x, const uint32_t* y, const uint32_t* sample_ids,
                         uint32_t* rec_lo, uint32_t* rec_hi, uint32_t* counts6, vsgpu_result** out);

/* ---- t1: closest_var(vg, idx, pos, vars) — include/query.h:441-483 (SURVEY.md §8f "next" row 1) -----
 * rec_lo[i] = rec_hi[i] = VSGPU_NONE where the operator returns false (no variant on the contig);
 * otherwise the rows are the records of [rec_lo, rec_hi) that a fresh next_variant_in_ref call keeps
 * (vsgpu_rows_t1 / vsgpu_digest_t1 apply that), possibly none. */
int vsgpu_query_t1(vsgpu_index* idx, uint64_t n, const uint64_t* pos, uint32_t* rec_lo, uint32_t* rec_hi);

/* ---- t7: samples_has_var(vg, idx, pos, ref, alt) — include/query.h:792-823 ----------------------
 * rec[i] = branch record whose (ref, pos, alt) equals the query among the records found by one
 * next_variant_in_ref(pos) call, or VSGPU_NONE ("There is no such variant!").  refs/alts are n
 * NUL-terminated strings ("" for an empty ref/alt). */
int vsgpu_query_t7(vsgpu_index* idx, uint64_t n, const uint64_t* pos, const char* const* refs,
                   const char* const* alts, uint32_t* rec);

/* ---- materialisation (host side) ---------------------------------------------------------------
 * Build the rows the reference's operators return (struct Variant, query.h:30-36) from record ids
 * / hit codes.  Text form = what print_var writes (query.h:43-50): "pos\tref\talt\tname(gt) ...\n".
 * The returned buffer is malloc'd; free it with vsgpu_free.  with_samples = 0 leaves the carrier
 * list empty (rows end "\t\n"). */
int vsgpu_rows_t6(const vsgpu_index* idx, uint32_t rec_lo, uint32_t rec_hi, int with_samples, char** text, uint64_t* nrows);
int vsgpu_rows_t1(const vsgpu_index* idx, uint32_t rec_lo, uint32_t rec_hi, int with_samples, char** text, uint64_t* nrows);
int vsgpu_rows_t4(const vsgpu_index* idx, const uint32_t* hits, uint64_t nhits, int with_samples, char** text);
/* "name gt" pairs concatenated as samples_has_var writes them (query.h:807-816) */
int vsgpu_rows_t7(const vsgpu_index* idx, uint32_t rec, char** text, uint64_t* ncarriers);
void vsgpu_free(void* p);
/* FNV-1a 64 digests of the row text per query (multi-threaded); used by the parity tests. */
int vsgpu_digest_t6(const vsgpu_index* idx, uint64_t n, const uint32_t* rec_lo, const uint32_t* rec_hi, int with_samples, uint64_t* digests);
int vsgpu_digest_t1(const vsgpu_index* idx, uint64_t n, const uint32_t* rec_lo, const uint32_t* rec_hi, int with_samples, uint64_t* counts, uint64_t* digests);
int vsgpu_digest_t4(const vsgpu_index* idx, uint64_t n, const uint64_t* offsets, const uint32_t* hits, int with_samples, uint64_t* digests);
int vsgpu_digest_t7(const vsgpu_index* idx, uint64_t n, const uint32_t* rec, uint64_t* ncarriers, uint64_t* digests);

/* ---- t6 rows rendered on the device (SURVEY.md §8f "next" row 3) -----------------------------------
 * get_var_in_ref(vg, idx, x, y, print = true, outfile) — include/query.h:736-784 with print_var
 * (:43-50) and get_samples (:268-285): the rows of every region as the bytes the reference writes
 * to `-o` after its header line, produced by a kernel (one warp per row expands the class bitmap /
 * id list of the row's vertex into "name(gt) " items) instead of the host loop of vsgpu_rows_t6.
 * Region i owns bytes [offsets[i], offsets[i+1]) of vsgpu_text_bytes; byte-identical to
 * vsgpu_rows_t6 on the slice vsgpu_query_t6 returns for it.  Returns VSGPU_ESHAPE when the text of
 * the batch exceeds VSGPU_RENDER_MAX_BYTES (default 2 GiB): split the batch. */
typedef struct vsgpu_text vsgpu_text;
int vsgpu_render_t6(vsgpu_index* idx, uint64_t n, const uint64_t* x, const uint64_t* y, int with_samples, vsgpu_text** out);
/* The same for t4: get_sample_var_in_ref(vg, idx, x, y, sample, print = true, outfile) — include/query.h:618-729 with print_var
 * (:43-50): t4 runs on the device and the rows of its hit codes (position / ref / alt by the emission rules of :677-710,
 * carriers of the row's vertex) are rendered there too.  Region i owns bytes [offsets[i], offsets[i+1]); byte-identical to
 * vsgpu_rows_t4 on the codes vsgpu_query_t4 returns for it. */
int vsgpu_render_t4(vsgpu_index* idx, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample_ids, int with_samples, vsgpu_text** out);
const char* vsgpu_text_bytes(const vsgpu_text* t);          /* NUL-terminated after the last region */
const uint64_t* vsgpu_text_offsets(const vsgpu_text* t);    /* n + 1 byte offsets */
uint64_t vsgpu_text_num_rows(const vsgpu_text* t);          /* rows rendered over all regions */
float vsgpu_text_kernel_ms(const vsgpu_text* t);            /* device time of the offset + render kernels (CUDA events) */
void vsgpu_text_free(vsgpu_text* t);

/* ---- t2: query_sample_from_ref(vg, idx, pos_x, pos_y, sample) — include/query.h:120-189 (SURVEY.md §8f "next" row 4) ---
 * The sample's sequence over the reference interval [x[i], y[i]): region i owns bytes
 * [offsets[i], offsets[i+1]) of vsgpu_text_bytes (what the reference writes to `-o` before its
 * newline).  vsgpu_text_status()[i] = 1 where the reference call ends in an uncaught
 * std::out_of_range from std::string::substr (query.h:163,167 — e.g. x = 0, or x right behind a
 * deletion whose target the neighbour scan meets first); the region's sequence is then empty.
 * sample_ids are sampleid_map ids (1..num_samples-1).  Returns VSGPU_ESHAPE when the index cannot
 * serve t2 (backbone sequences not contiguous in seq_buffer.sdsl) or the batch exceeds
 * VSGPU_RENDER_MAX_BYTES. */
int vsgpu_query_t2(vsgpu_index* idx, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample_ids, vsgpu_text** out);
/* ---- t3: query_sample_from_sample(vg, idx, pos_x, pos_y, sample) — include/query.h:195-261 ------------------
 * The sample's sequence over [x[i], y[i]) of the sample's OWN coordinates (the `index` fields
 * fix_sample_indexes writes, variant_graph.h:1883-1997).  Same result object as t2; status 1 = substr
 * throws (query.h:235,239), status 2 = the reference never returns: its loop at :209-214 repeats
 * get_prev_vertex_with_sample from a position that maps to itself (any x at or just behind one of
 * the sample's variants).  The first call reads the vertex blocks of ser/ a second time (vsgpu_open
 * does not keep the per-carrier indexes) and uploads them: 4 bytes per genotype entry. */
int vsgpu_query_t3(vsgpu_index* idx, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample_ids, vsgpu_text** out);
/* ---- t5: get_sample_var_in_sample(vg, idx, pos_x, pos_y, sample) — include/query.h:490-612 -----------------
 * The sample's variants over [x[i], y[i]) of its own coordinates.  Result = the CSR of t4 (same hit
 * codes, in the reference's push order) plus vsgpu_result_status()[i] = 2 where the reference never
 * returns (the loop at :505-510 is the one of t3).  Rows: vsgpu_rows_t5 / vsgpu_digest_t5 — the t4 row of
 * the same code with var_pos = ref_pos for an insertion and the sample's own position in the vertex
 * otherwise (:556-579).  Like t3, the first call reads the vertex blocks of ser/ a second time. */
int vsgpu_query_t5(vsgpu_index* idx, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample_ids, vsgpu_result** out);
const uint8_t* vsgpu_result_status(const vsgpu_result* r);      /* n bytes (t5 results; NULL for t4) */
float vsgpu_result_kernel_ms(const vsgpu_result* r);            /* t5: device time of the count and write launches (CUDA events) */
int vsgpu_rows_t5(const vsgpu_index* idx, const uint32_t* hits, uint64_t nhits, uint32_t sample_id, int with_samples, char** text);
int vsgpu_digest_t5(const vsgpu_index* idx, uint64_t n, const uint64_t* offsets, const uint32_t* hits, const uint32_t* sample_ids, int with_samples, uint64_t* digests);
/* t5 with print (get_sample_var_in_sample(..., print = true, outfile), query.h:596-606) for a batch, rows written on the device: region i's
 * rows are vsgpu_text_bytes()[offsets[i] .. offsets[i+1]) — byte for byte vsgpu_rows_t5 on the codes vsgpu_query_t5 returns for it;
 * vsgpu_text_status()[i] = 2 where the reference never returns (no rows).  vsgpu_text_stage_ms: count, write, rows.  The first call
 * uploads sample_info.index of every genotype entry (4 bytes each) beside the t3 tables. */
int vsgpu_render_t5(vsgpu_index* idx, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample_ids, int with_samples, vsgpu_text** out);
const uint8_t* vsgpu_text_status(const vsgpu_text* t);      /* n bytes (t2 / t3 results and rendered t5 rows; NULL for rendered t6 / t4 rows) */
const float* vsgpu_text_stage_ms(const vsgpu_text* t);      /* t2: device time of the count, plan and copy launches (CUDA events); their sum = vsgpu_text_kernel_ms */

/* ---- device-resident batches (bench harness; replaces the timing loop of src/bm_query.cc:74-135)
 * A batch keeps its regions and results in HBM so a run times the kernels alone.
 * type = 4, 6, 7 or 46 (t6 + t4 fused: one launch answers both, as vsgpu_query_t6t4).  For type 7 pass refs/alts;
 * for types 4 and 46 pass sample_ids. */
int vsgpu_batch_create(vsgpu_index* idx, int type, uint64_t n, const uint64_t* x, const uint64_t* y,
                       const uint32_t* sample_ids, const char* const* refs, const char* const* alts,
                       vsgpu_batch** out);
/* Enqueue one pass of the hot path over the batch on the index's stream (no host sync). */
int vsgpu_batch_run(vsgpu_batch* b);
/* Synchronise, then copy results out.  Any pointer may be NULL.  t6: rec_lo/rec_hi/counts;
 * t4: counts (per query) and *out (CSR); t7: rec_lo receives the record ids; 46: rec_lo/rec_hi/counts of t6
 * and *out = the t4 CSR. */
int vsgpu_batch_fetch(vsgpu_batch* b, uint32_t* rec_lo, uint32_t* rec_hi, uint32_t* counts, vsgpu_result** out);
/* Algorithmic bytes of the last run (SURVEY.md §8d formulas) and the number of kernels it launched. */
int vsgpu_batch_stats(vsgpu_batch* b, uint64_t* algorithmic_bytes, uint32_t* kernel_launches);
/* Device time (ms) of each kernel of the last run, from CUDA events on the launch stream
 * (t6/t7: 1 kernel; t4: walk, scan, gather).  Synchronises. */
int vsgpu_batch_timings(vsgpu_batch* b, float* ms, uint32_t cap, uint32_t* n);
void vsgpu_batch_free(vsgpu_batch* b);                         /* before vsgpu_close of its index: a batch points into it */

/* ---- several indexes on several GPUs behind one handle ---------------------------------------------
 * The reference keeps one ser/ directory per contig and runs one process per contig (util.cc:93-96,
 * eval_data_records/evaluation.txt:34); query_main (src/commands.cc:113-215) serves one.  A router opens
 * nshards directories — whole contigs, or position ranges [range_lo[k], range_hi[k]) of a contig built
 * from the records of that range (range_hi[k] = 0: to the contig's end; both arrays may be NULL: whole
 * contigs) — on the GPUs of this node: devices[k] names the GPU of shard k, or devices = NULL spreads the
 * shards over the first ndevices GPUs (0: all) by longest-processing-time on their size.  A query names a
 * contig per region (vsgpu_router_contig_id; the contig of a shard is the one its sampleid_map.lst names)
 * and is answered by the shard owning (contig, x): one host thread per GPU, the fused host-buffer call per
 * shard, no traffic between GPUs.  Record ids and hit codes are local to shard_of[i]
 * (vsgpu_router_shard_index gives that shard's index for vsgpu_rows_* / vsgpu_digest_*). */
typedef struct vsgpu_router vsgpu_router;
int vsgpu_router_open(uint32_t nshards, const char* const* ser_prefixes, const uint64_t* range_lo, const uint64_t* range_hi,
                      const int* devices, int ndevices, vsgpu_router** out);
void vsgpu_router_close(vsgpu_router* r);
const char* vsgpu_router_last_error(void);
uint32_t vsgpu_router_num_shards(const vsgpu_router* r);
uint32_t vsgpu_router_num_contigs(const vsgpu_router* r);
const char* vsgpu_router_contig_name(const vsgpu_router* r, uint32_t contig);
int vsgpu_router_contig_id(const vsgpu_router* r, const char* name, uint32_t* contig);
vsgpu_index* vsgpu_router_shard_index(const vsgpu_router* r, uint32_t shard);
int vsgpu_router_shard_device(const vsgpu_router* r, uint32_t shard);
/* t6 + t4 for n regions (contig[i], [x[i], y[i]), sample_ids[i]) in the caller's order: shard_of[i], rec_lo[i],
 * counts6[i] (t6 slice = [rec_lo, rec_lo + counts6) except where the literal dedup rule lowered the count) and
 * counts4[i]; the t4 hit codes per region through vsgpu_router_region_hits, or as one CSR in the caller's order through
 * vsgpu_router_offsets / _hits (all valid until the next call). */
int vsgpu_router_query_t6t4(vsgpu_router* r, uint64_t n, const uint32_t* contig, const uint32_t* x, const uint32_t* y, const uint32_t* sample_ids,
                            uint32_t* shard_of, uint32_t* rec_lo, uint32_t* counts6, uint32_t* counts4);
const uint64_t* vsgpu_router_offsets(const vsgpu_router* r);   /* n + 1; the first call after a query gathers the CSR (host threads) */
const uint32_t* vsgpu_router_hits(const vsgpu_router* r);
/* The hit codes of region i of the last call where the device->host copy put them (page-locked memory of its shard's result): no gather. */
int vsgpu_router_region_hits(const vsgpu_router* r, uint64_t i, const uint32_t** hits, uint32_t* count);
/* Last call: wall-clock milliseconds each GPU's host thread spent on its shards and the regions it answered (up to cap
 * entries; *ndev = GPUs in use), the routing time before and the scatter time after. */
int vsgpu_router_stats(const vsgpu_router* r, uint32_t cap, int* devices, double* device_ms, uint64_t* device_regions, uint32_t* ndev,
                       double* route_ms, double* scatter_ms);

#ifdef __cplusplus
}
#endif
#endif /* VSGPU_H_ */
