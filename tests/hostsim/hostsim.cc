// TEST-ONLY host simulator of libvsgpu's kernels.
//
// Exports the same C symbols as include/vsgpu.h (the query subset), but runs the per-region logic
// of variantstore_b200/csrc/device_logic.cuh compiled for the host, over the same flattened tables.
// Purpose: let the `-m "not gpu"` suite check loader + flattener + walk rules + materialiser against
// the oracle on a machine without a GPU.  It is never linked into, loaded by, or shipped with
// libvsgpu.so — the product has no CPU path.
#include "../../include/vsgpu.h"

#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>

#include "../../variantstore_b200/csrc/device_logic.cuh"
#include "../../variantstore_b200/csrc/host_index.h"

using namespace vsgpu;

struct vsgpu_result { std::vector<uint64_t> offsets; std::vector<uint32_t> hits; std::vector<uint8_t> status; std::vector<uint32_t> counts; };
struct vsgpu_text { std::string bytes; std::vector<uint64_t> offsets; std::vector<uint8_t> status; };
struct vsgpu_index : HostIndex {
	DevIndex dev;
	T2Tables t2{};
	T3Tables t3{};
	std::vector<uint64_t> sidx_begin; std::vector<uint32_t> sidx, sid;
	std::vector<uint32_t> bbs;
	std::string seq_ascii;
	std::vector<uint32_t> bucket, d4, car, marker_list, can_entry, can_pmax; std::vector<uint64_t> car_begin, can_begin;
	std::vector<uint2> t7;
	std::vector<uint32_t> hitmap;
};

namespace {
thread_local std::string g_err;
int set_err(int code, const std::string& m) { g_err = m; return code; }
char* dup_text(const std::string& s) { char* p = (char*)malloc(s.size() + 1); memcpy(p, s.data(), s.size()); p[s.size()] = 0; return p; }
struct VecSink { std::vector<uint32_t>* v; void emit(uint32_t c) { v->push_back(c); } };
}

extern "C" {
const char* vsgpu_last_error(void) { return g_err.c_str(); }
void vsgpu_free(void* p) { free(p); }

int vsgpu_open(const char* prefix, int, vsgpu_index** out) {
	*out = nullptr;
	std::unique_ptr<vsgpu_index> ix(new vsgpu_index);
	int stage = 0;
	try { build_host_index(prefix, *ix, &stage); }
	catch (const std::exception& e) { return set_err(stage == 0 ? VSGPU_EIO : VSGPU_ESHAPE, e.what()); }
	FlatIndex& f = ix->flat; DevIndex& d = ix->dev;
	memset(&d, 0, sizeof d);
	d.D = f.D; d.M = f.M; d.R = f.R; d.num_cent = (uint32_t)f.cent.size(); d.words_per_set = f.words_per_set;
	d.num_samples = f.num_samples; d.class_mode = f.class_mode; d.index_bits = f.index_bits; d.last_end = ix->last_end; d.t1_fallback_pos = ix->t1_fallback_pos;
	build_buckets(f, ix->bucket, d.bucket_shift);
	build_d4(f, ix->d4);
	d.nbuckets = (uint32_t)ix->bucket.size() - 1; d.bucket = ix->bucket.data(); d.dstart = f.dstart.data();
	d.d4 = (const uint4*)ix->d4.data();
	{ const char* e = getenv("VSGPU_T4_ROW64"); d.walk2 = (!e || atoi(e) != 0) ? 1 : 0; }
	ix->t7.resize(f.D);
	for (uint32_t i = 0; i < f.D; i++) ix->t7[i] = make_uint2(f.t7_lo[i], f.t7_hi[i]);
	d.dlev = (const uint4*)f.dlev.data(); d.dinfo = f.dinfo.data(); d.t7rng = ix->t7.data(); d.cent = (const uint4*)f.cent.data();
	d.bb_set = f.bb_set.data(); d.vstart = f.vstart.data(); d.bitmap = f.bitmap.data(); d.list_begin = f.list_begin.data();
	d.list_ids = f.list_ids.data(); d.rec_pos = f.rec_pos.data(); d.rec_hash = f.rec_hash.data(); d.rec_flags = f.rec_flags.data();
	d.marker_bits = f.marker_bits.data(); d.cent_begin_k = f.cent_begin.data(); d.dtin = f.dtin.data(); d.cent_anc = (const uint2*)f.cent_anc.data(); d.row_words = f.row_words; d.hitmap = nullptr;
	if (want_sparse_walk(f)) {
		build_sparse_walk(f, ix->car_begin, ix->car, ix->marker_list);
		d.car_begin = ix->car_begin.data(); d.car = ix->car.data(); d.marker_list = ix->marker_list.data(); d.num_markers = (uint32_t)ix->marker_list.size(); d.marker_span = marker_span(f, ix->marker_list);
		build_canonical_walks(f, ix->car_begin, ix->car, ix->marker_list, ix->can_begin, ix->can_entry, ix->can_pmax);
		d.can_begin = ix->can_begin.data(); d.can_entry = ix->can_entry.data(); d.can_pmax = ix->can_pmax.data();
	} else if (!getenv("VSGPU_DISABLE_HITMAP")) {   // host copy of k_build_hitmap
		ix->hitmap.assign((size_t)f.num_samples * f.row_words, 0);
		for (size_t c = 0; c < f.cent.size(); c++) {
			const CEntry& e = f.cent[c];
			if (e.tgt & kEntMarker) continue;
			if (f.class_mode) {
				for (uint32_t w = 0; w < f.words_per_set; w++) {
					uint64_t bits = f.bitmap[(uint64_t)e.set_id * f.words_per_set + w];
					while (bits) { uint32_t s = w * 64 + (uint32_t)__builtin_ctzll(bits); bits &= bits - 1; if (s) ix->hitmap[(size_t)s * f.row_words + (c >> 5)] |= 1u << (c & 31); }
				}
			} else for (uint64_t i = f.list_begin[e.set_id]; i < f.list_begin[e.set_id + 1]; i++) ix->hitmap[(size_t)f.list_ids[i] * f.row_words + (c >> 5)] |= 1u << (c & 31);
		}
		d.hitmap = ix->hitmap.data();
	}
	ix->bbs.assign(f.vstart.begin(), f.vstart.end()); ix->bbs.push_back(ix->last_end);
	static const char kBase[8] = {'A', 'C', 'T', 'G', 'N', 5, 5, 5};
	ix->seq_ascii.resize(ix->ser.seq.size() + 16, 0);
	for (size_t i = 0; i < ix->ser.seq.size(); i++) ix->seq_ascii[i] = kBase[ix->ser.seq[i] & 7];
	ix->t2.bbs = ix->bbs.data(); ix->t2.nrp1 = f.nrp1.data(); ix->t2.first_reach = f.first_reach.data();
	ix->t2.cent_seq = (const uint2*)f.cent_seq.data(); ix->t2.seq_ascii = ix->seq_ascii.data();
	*out = ix.release();
	return VSGPU_OK;
}
void vsgpu_close(vsgpu_index* ix) { delete ix; }

// t2 through the same two passes as the kernels: count, then copy records, then the copy itself
static int query_seq(vsgpu_index* ix, bool t3, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* s, vsgpu_text** out);
int vsgpu_query_t2(vsgpu_index* ix, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* s, vsgpu_text** out) { return query_seq(ix, false, n, x, y, s, out); }
static int ensure_t3(vsgpu_index* ix) {
	if (ix->sidx_begin.empty()) {     // same construction as ensure_t3_tables of libvsgpu
		const FlatIndex& f = ix->flat; const SerData& sd = ix->ser;
		std::vector<uint32_t>& sindex = ix->sindex;
		try { load_sample_indexes(ix->prefix, sd.v_sinfo_begin.back(), sindex); } catch (const std::exception& e) { return set_err(VSGPU_ESHAPE, e.what()); }
		const size_t E = f.cent.size();
		ix->sidx_begin.assign(E + 1, 0);
		for (size_t c = 0; c < E; c++) { const uint32_t v = f.cent_vertex[c]; ix->sidx_begin[c + 1] = ix->sidx_begin[c] + (v == kNone ? 0 : sd.v_sinfo_begin[v + 1] - sd.v_sinfo_begin[v]); }
		ix->sidx.resize(ix->sidx_begin[E]); if (!f.class_mode) ix->sid.resize(ix->sidx_begin[E]);
		for (size_t c = 0; c < E; c++) {
			const uint32_t v = f.cent_vertex[c];
			if (v == kNone) continue;
			const uint64_t s0 = sd.v_sinfo_begin[v], cnt = sd.v_sinfo_begin[v + 1] - s0;
			memcpy(ix->sidx.data() + ix->sidx_begin[c], sindex.data() + s0, cnt * 4);
			if (!f.class_mode) memcpy(ix->sid.data() + ix->sidx_begin[c], sd.s_sample_id.data() + s0, cnt * 4);
		}
		ix->t3.sidx_begin = ix->sidx_begin.data(); ix->t3.sidx = ix->sidx.data(); ix->t3.sid = f.class_mode ? nullptr : ix->sid.data();
		ix->t3.first_index = f.vstart[f.dlev[0].k];
	}
	return VSGPU_OK;
}
int vsgpu_query_t3(vsgpu_index* ix, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* s, vsgpu_text** out) {
	if (int rc = ensure_t3(ix)) return rc;
	return query_seq(ix, true, n, x, y, s, out);
}
int vsgpu_query_t5(vsgpu_index* ix, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* s, vsgpu_result** out) {
	if (!ix->flat.t2_ok) return set_err(VSGPU_ESHAPE, "vsgpu_query_t5: " + ix->flat.t2_why);
	if (int rc = ensure_t3(ix)) return rc;
	std::unique_ptr<vsgpu_result> r(new vsgpu_result);
	r->offsets.assign(n + 1, 0); r->status.assign(n + 1, 0);
	VecSink sink{&r->hits};
	for (uint64_t i = 0; i < n; i++) {
		if (s[i] == 0 || s[i] >= ix->dev.num_samples) return set_err(VSGPU_EINVAL, "vsgpu_query_t5: sample id out of range");
		const size_t before = r->hits.size();
		const uint32_t st = logic::t5_walk(ix->dev, ix->t2, ix->t3, x[i], y[i], s[i], sink);
		if (st) r->hits.resize(before);
		r->status[i] = (uint8_t)st;
		r->offsets[i + 1] = r->hits.size();
	}
	*out = r.release();
	return VSGPU_OK;
}
const uint8_t* vsgpu_result_status(const vsgpu_result* r) { return r->status.empty() ? nullptr : r->status.data(); }
float vsgpu_result_kernel_ms(const vsgpu_result*) { return 0.f; }
int vsgpu_rows_t5(const vsgpu_index* ix, const uint32_t* hits, uint64_t nhits, uint32_t sample, int ws, char** text) {
	std::string s;
	for (uint64_t i = 0; i < nhits; i++) t5_row(ix, hits[i], sample, ws != 0, s);
	*text = dup_text(s); return VSGPU_OK;
}
int vsgpu_digest_t5(const vsgpu_index* ix, uint64_t n, const uint64_t* off, const uint32_t* hits, const uint32_t* samples, int ws, uint64_t* d) { digests_t5(ix, n, off, hits, samples, ws != 0, d); return VSGPU_OK; }
static int query_seq(vsgpu_index* ix, bool t3, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* s, vsgpu_text** out) {
	if (!ix->flat.t2_ok) return set_err(VSGPU_ESHAPE, "vsgpu_query_t2: " + ix->flat.t2_why);
	std::unique_ptr<vsgpu_text> t(new vsgpu_text);
	t->offsets.assign(n + 1, 0); t->status.assign(n + 1, 0);
	for (uint64_t i = 0; i < n; i++) {
		if (s[i] == 0 || s[i] >= ix->dev.num_samples) return set_err(VSGPU_EINVAL, "vsgpu_query_t2: sample id out of range");
		logic::T2CountSink cs{0, 0, 0, 0, nullptr, 0};
		uint32_t st = t3 ? logic::t3_walk(ix->dev, ix->t2, ix->t3, x[i], y[i], s[i], cs) : logic::t2_walk(ix->dev, ix->t2, x[i], y[i], s[i], cs); cs.flush();
		t->status[i] = (uint8_t)st;
		if (!st && cs.nrec) {
			std::vector<uint4> recs(cs.nrec);
			std::vector<uint32_t> tile_first((t->bytes.size() + cs.bytes) / kT2Tile + 2, 0);
			logic::T2WriteSink ws{0, 0, recs.data(), t->bytes.size(), recs.data(), tile_first.data()};
			if (t3) logic::t3_walk(ix->dev, ix->t2, ix->t3, x[i], y[i], s[i], ws); else logic::t2_walk(ix->dev, ix->t2, x[i], y[i], s[i], ws);
			ws.flush();
			if (ws.out != recs.data() + cs.nrec || ws.dst != t->bytes.size() + cs.bytes) return set_err(VSGPU_EINVAL, "hostsim: t2 count and write passes disagree");
			t->bytes.resize(t->bytes.size() + cs.bytes);
			for (const uint4& r : recs) memcpy(&t->bytes[r.z | ((uint64_t)r.w << 32)], ix->seq_ascii.data() + r.x, r.y);
		}
		t->offsets[i + 1] = t->bytes.size();
	}
	*out = t.release();
	return VSGPU_OK;
}
const char* vsgpu_text_bytes(const vsgpu_text* t) { return t->bytes.c_str(); }
const uint64_t* vsgpu_text_offsets(const vsgpu_text* t) { return t->offsets.data(); }
const uint8_t* vsgpu_text_status(const vsgpu_text* t) { return t->status.data(); }
uint64_t vsgpu_text_num_rows(const vsgpu_text* t) { return t->offsets.size() - 1; }
float vsgpu_text_kernel_ms(const vsgpu_text*) { return 0.f; }
const float* vsgpu_text_stage_ms(const vsgpu_text*) { static const float z[3] = {0, 0, 0}; return z; }
void vsgpu_text_free(vsgpu_text* t) { delete t; }

int vsgpu_info(const vsgpu_index* ix, vsgpu_info_t* o) {
	memset(o, 0, sizeof *o);
	o->ref_length = ix->ser.ref_length; o->seq_length = ix->ser.seq.size(); o->num_vertices_cqf = ix->ser.cqf_distinct;
	o->num_vertices = ix->ser.num_vertices; o->num_samples = ix->ser.num_samples;
	o->num_classes = ix->flat.class_mode ? ix->flat.num_sets - 1 : 0; o->class_mode = ix->flat.class_mode;
	o->backbone_vertices = ix->flat.M; o->distinct_starts = ix->flat.D; o->branch_records = ix->flat.R;
	o->walk_entries = (uint32_t)ix->flat.cent.size(); o->has_suspect_dups = ix->flat.has_suspect_dups;
	strncpy(o->chr, ix->ser.chr.c_str(), sizeof o->chr - 1);
	o->from_cache = ix->from_cache ? 1 : 0;
	for (const auto& e : ix->flat.cent) { if (e.tgt & kEntMarker) o->walk_markers++; else if ((e.tgt & kEntAlt) && (e.tgt & kEntTgtCarriers)) o->rejoin_carriers++; }
	return VSGPU_OK;
}
int vsgpu_sample_id(const vsgpu_index* ix, const char* name, uint32_t* id) {
	auto it = ix->name2id.find(name);
	if (it == ix->name2id.end()) return set_err(VSGPU_EINVAL, std::string("Sample not found: ") + name);
	*id = it->second; return VSGPU_OK;
}
const char* vsgpu_sample_name(const vsgpu_index* ix, uint32_t id) { return id < ix->ser.num_samples ? ix->ser.sample_names[id].c_str() : nullptr; }

int vsgpu_query_t6(vsgpu_index* ix, uint64_t n, const uint64_t* x, const uint64_t* y, uint32_t* lo, uint32_t* hi, uint32_t* counts) {
	std::vector<uint32_t> tmp;
	for (uint64_t i = 0; i < n; i++) {
		bool bad = false;
		uint2 r = logic::t6_bounds(ix->dev, x[i], y[i], &bad);
		if (bad) return set_err(VSGPU_EINVAL, "Can't find node corresponding to pos 0");
		if (lo) lo[i] = r.x;
		if (hi) hi[i] = r.y;
		if (counts) { if (t6_needs_literal(ix, y[i], r.x, r.y)) { t6_literal(ix, x[i], y[i], tmp); counts[i] = (uint32_t)tmp.size(); } else counts[i] = r.y - r.x; }
	}
	return VSGPU_OK;
}

int vsgpu_query_t4(vsgpu_index* ix, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* s, vsgpu_result** out) {
	std::unique_ptr<vsgpu_result> r(new vsgpu_result);
	r->offsets.assign(n + 1, 0);
	VecSink sink{&r->hits};
	for (uint64_t i = 0; i < n; i++) {
		if (x[i] < 1 || s[i] == 0 || s[i] >= ix->dev.num_samples) return set_err(VSGPU_EINVAL, "region start < 1 or sample id out of range");
		uint2 fused;                                       // the t6 slice a fused launch takes from the walk's two ranks
		logic::walk_any(ix->dev, x[i], y[i], s[i], sink, &fused);
		bool bad = false;
		const uint2 b = logic::t6_bounds(ix->dev, x[i], y[i], &bad);
		if (fused.x != b.x || fused.y != b.y) return set_err(VSGPU_EINVAL, "hostsim: fused t6 slice differs from t6_bounds");
		r->offsets[i + 1] = r->hits.size();
	}
	*out = r.release();
	return VSGPU_OK;
}
int vsgpu_query_t6_u32(vsgpu_index* ix, uint64_t n, const uint32_t* x, const uint32_t* y, uint32_t* lo, uint32_t* hi, uint32_t* counts) {
	std::vector<uint64_t> x64(x, x + n), y64(y, y + n);
	return vsgpu_query_t6(ix, n, x64.data(), y64.data(), lo, hi, counts);
}
int vsgpu_query_t4_u32(vsgpu_index* ix, uint64_t n, const uint32_t* x, const uint32_t* y, const uint32_t* s, vsgpu_result** out) {
	std::vector<uint64_t> x64(x, x + n), y64(y, y + n);
	return vsgpu_query_t4(ix, n, x64.data(), y64.data(), s, out);
}
int vsgpu_query_t6t4(vsgpu_index* ix, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* s, uint32_t* lo, uint32_t* hi, uint32_t* c6, vsgpu_result** out) {
	if (int rc = vsgpu_query_t6(ix, n, x, y, lo, hi, c6)) return rc;
	return vsgpu_query_t4(ix, n, x, y, s, out);
}
int vsgpu_query_t6t4_u32(vsgpu_index* ix, uint64_t n, const uint32_t* x, const uint32_t* y, const uint32_t* s, uint32_t* lo, uint32_t* hi, uint32_t* c6, vsgpu_result** out) {
	std::vector<uint64_t> x64(x, x + n), y64(y, y + n);
	return vsgpu_query_t6t4(ix, n, x64.data(), y64.data(), s, lo, hi, c6, out);
}
const uint32_t* vsgpu_result_counts(const vsgpu_result* cr) {
	vsgpu_result* r = const_cast<vsgpu_result*>(cr);
	if (r->counts.size() + 1 != r->offsets.size()) { r->counts.resize(r->offsets.size() - 1); for (size_t i = 0; i + 1 < r->offsets.size(); i++) r->counts[i] = (uint32_t)(r->offsets[i + 1] - r->offsets[i]); }
	return r->counts.data();
}
uint64_t vsgpu_result_total(const vsgpu_result* r) { return r->offsets.back(); }
uint64_t vsgpu_result_num_queries(const vsgpu_result* r) { return r->offsets.size() - 1; }
const uint64_t* vsgpu_result_offsets(const vsgpu_result* r) { return r->offsets.data(); }
const uint32_t* vsgpu_result_hits(const vsgpu_result* r) { return r->hits.data(); }
void vsgpu_result_free(vsgpu_result* r) { delete r; }

int vsgpu_query_t1(vsgpu_index* ix, uint64_t n, const uint64_t* pos, uint32_t* lo, uint32_t* hi) {
	for (uint64_t i = 0; i < n; i++) {
		bool bad = false;
		uint2 r = logic::t1_lookup(ix->dev, pos[i], &bad);
		if (bad) return set_err(VSGPU_EINVAL, "Can't find node corresponding to pos 0");
		lo[i] = r.x; hi[i] = r.y;
	}
	return VSGPU_OK;
}
int vsgpu_rows_t1(const vsgpu_index* ix, uint32_t lo, uint32_t hi, int ws, char** text, uint64_t* nrows) {
	std::string s; uint64_t cnt = 0; rows_t1(ix, lo, hi, ws != 0, s, cnt);
	if (nrows) *nrows = cnt;
	*text = dup_text(s); return VSGPU_OK;
}
int vsgpu_digest_t1(const vsgpu_index* ix, uint64_t n, const uint32_t* lo, const uint32_t* hi, int ws, uint64_t* c, uint64_t* d) { digests_t1(ix, n, lo, hi, ws != 0, c, d); return VSGPU_OK; }

int vsgpu_query_t7(vsgpu_index* ix, uint64_t n, const uint64_t* pos, const char* const* refs, const char* const* alts, uint32_t* rec) {
	for (uint64_t i = 0; i < n; i++) {
		bool bad = false;
		uint32_t r = logic::t7_lookup(ix->dev, pos[i], hash_query(refs[i], alts[i]), &bad);
		if (bad) return set_err(VSGPU_EINVAL, "Can't find node corresponding to pos 0");
		rec[i] = t7_confirm(ix, pos[i], refs[i], alts[i], r);
	}
	return VSGPU_OK;
}

int vsgpu_rows_t6(const vsgpu_index* ix, uint32_t lo, uint32_t hi, int ws, char** text, uint64_t* nrows) {
	if ((lo != VSGPU_NONE || hi != VSGPU_NONE) && (lo > hi || hi > ix->flat.R)) return set_err(VSGPU_EINVAL, "bad record slice");
	std::string s; uint64_t cnt = 0; rows_t6(ix, lo, hi, ws != 0, s, cnt);
	if (nrows) *nrows = cnt;
	*text = dup_text(s); return VSGPU_OK;
}
int vsgpu_rows_t4(const vsgpu_index* ix, const uint32_t* hits, uint64_t nhits, int ws, char** text) {
	std::string s;
	for (uint64_t i = 0; i < nhits; i++) t4_row(ix, hits[i], ws != 0, s);
	*text = dup_text(s); return VSGPU_OK;
}
int vsgpu_rows_t7(const vsgpu_index* ix, uint32_t rec, char** text, uint64_t* nc) {
	std::string s; uint64_t c = t7_carriers(ix, rec, &s, nullptr);
	if (nc) *nc = c;
	*text = dup_text(s); return VSGPU_OK;
}
int vsgpu_digest_t6(const vsgpu_index* ix, uint64_t n, const uint32_t* lo, const uint32_t* hi, int ws, uint64_t* d) { bool bad = false; digests_t6(ix, n, lo, hi, ws != 0, d, &bad); return bad ? VSGPU_EINVAL : VSGPU_OK; }
int vsgpu_digest_t4(const vsgpu_index* ix, uint64_t n, const uint64_t* off, const uint32_t* hits, int ws, uint64_t* d) { digests_t4(ix, n, off, hits, ws != 0, d); return VSGPU_OK; }
int vsgpu_digest_t7(const vsgpu_index* ix, uint64_t n, const uint32_t* rec, uint64_t* nc, uint64_t* d) { digests_t7(ix, n, rec, nc, d); return VSGPU_OK; }
}  // extern "C"
