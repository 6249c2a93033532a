// libvsgpu device side — shared declarations between kernels.cu and the C-ABI host code.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vsgpu {

constexpr uint32_t kScratchHits = 8;    // t4 hits kept inline per region before the overflow path
constexpr uint32_t kBucketTarget = 12;  // average number of distinct starts per rank bucket

struct DevIndex {
	uint32_t D, M, R, num_cent, words_per_set, num_samples, class_mode;
	uint64_t index_bits;
	uint32_t last_end;                    // vstart[M-1] + vlen[M-1]
	const uint32_t* dstart;               // D distinct backbone starts, ascending
	const uint32_t* bucket;               // nbuckets + 1: bucket[b] = number of starts < (b << bucket_shift)
	uint32_t nbuckets, bucket_shift;
	uint32_t t1_fallback_pos;             // largest position from which next_variant_in_ref still finds a variant (0: no variants)
	const uint4* dlev;                    // D + 1: {k, rec_lo, rec_hi_prev, cent_begin}
	const uint64_t* dinfo;                // D
	const uint2* t7rng;                   // D
	const uint4* cent;                    // num_cent: {src, tgt|flags, set_id, arrival}
	const uint32_t* bb_set;               // M
	const uint32_t* vstart;               // M
	const uint64_t* bitmap;
	const uint64_t* list_begin;
	const uint32_t* list_ids;
	const uint32_t* rec_pos;              // R
	const uint64_t* rec_hash;             // R
	const uint8_t* rec_flags;             // R
	const uint32_t* rec_dup_prefix;       // R + 1 prefix count of suspect duplicate records; nullptr when the index has none
	uint32_t tail_records;                // 1: the last backbone vertex has branch records (regions past the contig end need the literal rule)
	// sample-major hit map over the walk entries (nullptr: not built, kernels use the class bitmaps)
	const uint32_t* hitmap;               // num_samples rows x row_words
	uint32_t row_words;                   // 32-bit words per row, multiple of 32 (128-byte rows)
	const uint32_t* marker_bits;          // row_words words: bit c = walk entry c is a marker
	const uint32_t* cent_begin_k;         // M + 1: first walk entry with src >= k
	const uint32_t* dtin;                 // D + 1: pre-order time of back-walk state c in the back-walk forest
	const uint2* cent_anc;                // per walk entry: pre-order interval of the state examining its source
	const uint4* d4;                      // D + 1: {dlev.k, dlev.cent_begin, entries below the back-walk start of state d + 1, dtin[d]} (built at upload)
	// Sparse cohorts (explicit-id encoding, or few carriers per sample): instead of the hit map, per sample the sorted list of
	// the walk entries it carries — 4 bytes per genotype entry, a predecessor search for the back-walk, a range scan forward.
	const uint64_t* car_begin;            // num_samples + 1 (nullptr: not built; then the hit map or the class bitmaps answer)
	const uint32_t* car;                  // walk-entry ids, ascending per sample (markers excluded)
	const uint32_t* marker_list;          // the marker entries, ascending; num_markers of them
	uint32_t num_markers;
	uint32_t marker_span;                 // a marker can only end walks whose scan bound lies at most this many entries behind it (max over markers)
	// The walk of a sample from the start of the contig (what a region gets when nothing the sample carries lies on the
	// back-walk's chain) does not depend on the region: can_entry lists the entries that walk takes, can_pmax the running
	// maximum of their arrival positions, so a region joins the walk at the last entry arriving before x instead of at entry 0.
	const uint64_t* can_begin;            // num_samples + 1
	const uint32_t* can_entry;
	const uint32_t* can_pmax;
	uint32_t walk2;                       // 1: the per-thread hit-map walk is walk_region_fast2 (64-entry row chunks fetched together); VSGPU_T4_ROW64=0 clears it
};

// Tables for rendering t6 rows as text on the device; uploaded on first use.
// A row is print_var's line (query.h:43-50): "pos\tref\talt\tname(gt) name(gt) ...\n"; its length
// depends on the record alone, so text offsets are differences of two prefix arrays.
struct RenderTables {
	const uint4* rec_seq;        // R: {ref_off, ref_len, alt_off, alt_len} into `seq`
	const uint4* rec_car;        // R: {carrier set id, s_info count | row flags << 28, s_info begin lo, hi}
	const uint64_t* text_prefix[2];   // R + 1 each: bytes of rows [0, r) without / with the carrier list
	const uint8_t* seq;          // 3-bit base codes, one per byte (seq_buffer.sdsl)
	const uint8_t* s_flags;      // per s_info: bit0 phase, bit1 gt_1, bit2 gt_2
	const uint32_t* s_sample_id; // explicit-id mode: per s_info sample id
	const uint32_t* name_off;    // num_samples + 1 offsets into name_chars
	const char* name_chars;
	const uint4* item16;         // num_samples: "name(0|0) " + its length in byte 15 when that fits 15 bytes, else zeros
};

// Tables for rendering t4 rows (get_sample_var_in_ref with print, query.h:677-726) on the device: per walk entry and variant
// (0: plain, 1: the walk started on it — a substitution prints ref "", 2: the row is the backbone vertex the alt edge rejoins)
// what the row prints.  Built on the host with the rules of materialize.cc (hit_row); uploaded on first use.
struct HitTables {
	const uint32_t* pos[3];      // num_cent each: var_pos
	const uint4* seq[3];         // {ref_off, ref_len, alt_off, alt_len} into RenderTables::seq
	const uint4* car[3];         // {carrier set id, s_info count | row flags << 28, s_info begin lo, hi} of the vertex whose carriers are printed
	const uint32_t* len[2][3];   // bytes of the row without / with the carrier list
};
// byte_off[nh + 1] = exclusive text offsets of the rows of hits[0 .. nh); region_off[i] = byte_off[offsets[i]] for i <= n; scratch as launch_render_offsets
cudaError_t launch_hit_offsets(const HitTables& ht, const uint32_t* hits, uint64_t nh, int with_samples, uint64_t n, const uint64_t* offsets, uint64_t* byte_off, uint64_t* region_off,
                               uint64_t* scratch, cudaStream_t stream, const uint32_t* pos5 = nullptr);
cudaError_t launch_render_hits(const DevIndex& ix, const RenderTables& rt, const HitTables& ht, const uint32_t* hits, int with_samples, const uint64_t* byte_off,
                               uint64_t row_begin, uint64_t row_end, char* text, cudaStream_t stream, const uint32_t* pos5 = nullptr);
// t5 rows: pos5[h] = var_pos of hit h of a t5 answer (offsets / samples: the batch's CSR and sample ids; gsidx: sample_info.index per s_info entry);
// passed to the two launchers above, the rows come out as get_sample_var_in_sample prints them (query.h:553-590)
cudaError_t launch_t5_row_pos(const DevIndex& ix, const RenderTables& rt, const HitTables& ht, const uint32_t* gsidx, uint64_t n, const uint64_t* offsets, const uint32_t* samples,
                              const uint32_t* hits, uint64_t nh, uint32_t* pos5, cudaStream_t stream);

// Tables of t2 = query_sample_from_ref (include/query.h:120-189); uploaded on first use.
struct T2Tables {
	const uint32_t* bbs;          // M + 1: backbone starts, bbs[M] = end of the last backbone vertex
	const uint32_t* nrp1;         // M: next_ref_pos computed at P[k] (first ref-carrying neighbour, query.h:143-151)
	const uint32_t* first_reach;  // D + 1: first backbone index j with nrp1[j] >= dstart[d]
	const uint2* cent_seq;        // per walk entry: {seq offset, length} of its alt target
	const char* seq_ascii;        // seq_buffer.sdsl as ASCII (A C T G N, util.cc:32-41); 64 readable bytes before and after
};
// Extra tables of t3 = query_sample_from_sample (include/query.h:195-261): sample_info.index of every carrier of
// every walk entry's target vertex, in s_info order (uploaded on first use; read back from ser/ by a second pass).
struct T3Tables {
	const uint64_t* sidx_begin;   // num_cent + 1
	const uint32_t* sidx;         // sample_info.index per carrier
	const uint32_t* sid;          // explicit-id mode: sample id per carrier (nullptr in class mode)
	uint32_t first_index;         // REF's index in node_list[0] (where a walk from the contig start begins)
};
constexpr uint32_t kT3Hang = 2;           // per-region status: the loop of query.h:209-214 never ends (the reference hangs)
constexpr uint32_t kT2Tile = 512;         // bytes of output one warp of the copy kernel writes per step (16 per lane)
constexpr uint32_t kT2Keep = 16;          // copy records per region the count pass keeps for the plan pass (more: the plan pass walks again)
constexpr uint32_t kT2Throw = 1;          // per-region status: the reference call ends in std::out_of_range (substr, query.h:163,167)
// t2 runs as five launches over n regions (count and plan: CTA b owns regions [256 b, 256 b + 256)):
//   count : walk every region, write cnt[i] = {copy records, bytes}, status[i] and the first kT2Keep pieces {src, len}
//           of the region into keep[k * n + i]; per-CTA sums into cta_sums[2 b]
//   bases : exclusive scan of the per-CTA sums in place (entry nctas = totals); launched by launch_t2_count
//   plan  : offsets[i] (bytes) from the scans, walk again and write the copy records {src, len, dst lo, dst hi}
//           and tile_first[t] = the record covering output byte t * kT2Tile
//   copy  : one warp per kT2Tile bytes of `text`; a lane finds the record of its 16 bytes starting from
//           tile_first and stores them with one 128-bit store (k_t2_copy); the 16-byte chunks in which a
//           record starts are built by one thread per record (k_t2_seams)
uint64_t t2_ctas(uint64_t n);
cudaError_t launch_t2_count(const DevIndex& ix, const T2Tables& t2, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample,
                            uint2* cnt, uint2* keep, uint8_t* status, uint64_t* cta_sums, uint32_t* gstatus, cudaStream_t stream, const T3Tables* t3 = nullptr);
cudaError_t launch_t2_plan(const DevIndex& ix, const T2Tables& t2, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample,
                           const uint2* cnt, const uint2* keep, const uint64_t* cta_sums, uint64_t* offsets, uint4* recs, uint32_t* tile_first, cudaStream_t stream,
                           const T3Tables* t3 = nullptr);   // t3 != nullptr: the same launches walk in the sample's own coordinates (t3)
// `totals` = {records, bytes} on the device (entry nctas of the scanned CTA sums); text must have room for bytes rounded up to kT2Tile
cudaError_t launch_t2_copy(const T2Tables& t2, const uint4* recs, const uint32_t* tile_first, const uint64_t* totals, uint64_t recs_hint, uint64_t bytes_hint, char* text, cudaStream_t stream);

// t5 = get_sample_var_in_sample (query.h:490-612): count (+ scan of the CTA sums, as t2), then write.  cnt[n], status[n];
// offsets[n+1]; hits sized from the total in cta_sums[2 * ceil(n / 256)].
// keep: 8 hit codes of room per region (t5_keep_words(n) 32-bit words), filled by the count launch and read by the write launch
uint64_t t5_keep_words(uint64_t n);
cudaError_t launch_t5_count(const DevIndex& ix, const T2Tables& t2, const T3Tables& t3, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample,
                            uint32_t* cnt, uint8_t* status, uint64_t* cta_sums, uint32_t* gstatus, uint32_t* keep, cudaStream_t stream);
cudaError_t launch_t5_write(const DevIndex& ix, const T2Tables& t2, const T3Tables& t3, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample,
                            const uint32_t* cnt, const uint64_t* cta_sums, uint64_t* offsets, uint32_t* hits, const uint32_t* keep, cudaStream_t stream);

// Segment s of a render call = records [seg_lo[s], seg_hi[s]) ((NONE, NONE) = empty).  Three small
// launches turn the per-segment row and byte counts into exclusive offsets: row_off / byte_off get
// nseg + 1 entries (the last = totals); `scratch` needs 2 * (ceil(nseg / 1024) + 1) words.
cudaError_t launch_render_offsets(const DevIndex& ix, const RenderTables& rt, uint64_t nseg, const uint32_t* seg_lo, const uint32_t* seg_hi, int with_samples,
                                  uint64_t* row_off, uint64_t* byte_off, uint64_t* scratch, cudaStream_t stream);
// One warp per row of [row_begin, row_end) writes its text at its final position in `text`.
cudaError_t launch_render(const DevIndex& ix, const RenderTables& rt, uint64_t nseg, const uint32_t* seg_lo, int with_samples,
                          const uint64_t* row_off, const uint64_t* byte_off, uint64_t row_begin, uint64_t row_end, char* text, cudaStream_t stream);

// fills `hitmap` (zeroed, num_samples x row_words) from the walk entries and their carrier sets
cudaError_t launch_build_hitmap(const DevIndex& ix, uint32_t* hitmap, cudaStream_t stream);

// 32-bit host coordinates widened to the 64-bit arrays the query kernels read
cudaError_t launch_widen(uint64_t n, const uint32_t* x32, const uint32_t* y32, uint64_t* x, uint64_t* y, cudaStream_t stream);

// All launchers enqueue on `stream` and return the CUDA error of the launch.
// `status` is two words: [0] status bits, [1] number of entries appended to a t6 `flagged` list.
// t6 writes the record slice of every region as two arrays lo[n], hi[n]; optionally counts[n] (slice
// lengths) and, into flagged[], flag_base + i for every region whose count needs the literal rule.
cudaError_t launch_t6(const DevIndex& ix, uint64_t n, const uint64_t* x, const uint64_t* y, uint32_t* lo, uint32_t* hi, uint32_t* counts,
                      uint32_t* flagged, uint32_t flag_base, uint32_t* status, cudaStream_t stream);
cudaError_t launch_t1(const DevIndex& ix, uint64_t n, const uint64_t* pos, uint32_t* lo, uint32_t* hi, uint32_t* status, cudaStream_t stream);
cudaError_t launch_t7(const DevIndex& ix, uint64_t n, const uint64_t* pos, const uint64_t* qhash, uint32_t* rec, uint32_t* status, cudaStream_t stream);
// t4: one launch — walk, CTA scan, decoupled look-back, ordered write of the hits.
// offsets[n+1] (exclusive, offsets[n] = total); hits has room for `cap` codes, kStatusOverflow is
// raised (and nothing past cap written) when the total exceeds it.  `tile_state` needs
// t4_state_words(n) zeroed 64-bit words before every launch.  `base_ptr` (nullable) holds the offset
// the launch starts from: a batch cut into chunks passes the previous chunk's offsets[n] slot.
uint64_t t4_state_words(uint64_t n);
cudaError_t launch_t4(const DevIndex& ix, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample,
                      uint64_t* offsets, uint32_t* hits, uint64_t cap, uint64_t* tile_state, uint32_t* status, bool wide_regions,
                      cudaStream_t stream, const uint64_t* base_ptr = nullptr, uint32_t* spill = nullptr);
uint32_t t4_wide_entries();

// t6 outputs of a fused launch: as launch_t6's; hi / counts / flagged nullable
struct T6Out { uint32_t* lo; uint32_t* hi; uint32_t* counts; uint32_t* flagged; uint32_t flag_base; };
// launch_t4 with everything k_t4p can do: x / y as 64-bit or 32-bit arrays (coords32), per-region hit counts
// beside the offsets (counts, nullable), and the t6 slice of every region from the same two ranks (fuse6, nullable).
// Batches of few, wide regions (wide_regions) and the one-CTA-per-tile kernel take 64-bit coordinates only and
// cannot fuse; t4x_supported() says whether a launch can take these options.
struct T4Launch {
	uint64_t n; const void* x; const void* y; bool coords32; const uint32_t* sample;
	uint64_t* offsets; uint32_t* counts; uint32_t* hits; uint64_t cap; uint64_t* tile_state; uint32_t* status; const uint64_t* base_ptr;
	const T6Out* fuse6;
	uint32_t* spill = nullptr;     // t4x_spill_bytes() of scratch: regions with more rows than the staging holds spill there instead of walking twice (64-region tiles only)
};
uint64_t t4x_spill_bytes();
uint64_t t4w_pool_bytes();       // launch_t4 with wide_regions: `spill` = this many bytes (a pool of hit-code chunks), or nullptr
bool t4x_supported(bool wide_regions);
cudaError_t launch_t4x(const DevIndex& ix, const T4Launch& a, cudaStream_t stream);

// status word bits
constexpr uint32_t kStatusBadRegion = 1;   // some region had x < 1
constexpr uint32_t kStatusOverflow = 2;    // hits capacity too small

}  // namespace vsgpu
