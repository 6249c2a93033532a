// libvsgpu host side — see host_index.h.
#include "host_index.h"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "../../include/vsgpu.h"
#include "kernels.cuh"

namespace vsgpu {

void build_buckets(const FlatIndex& f, std::vector<uint32_t>& bucket, uint32_t& shift) {
	const uint64_t span = std::max<uint64_t>(f.dstart.empty() ? 1 : f.dstart.back(), 1);
	uint64_t target = kBucketTarget;                       // tuning knob, read per open
	if (const char* e = getenv("VSGPU_BUCKET_TARGET")) { const long v = atol(e); if (v >= 1 && v <= 4096) target = (uint64_t)v; }
	shift = 0;
	while (shift < 31 && (span >> (shift + 1)) * target >= f.D) shift++;    // ~target starts per bucket
	while ((span >> shift) + 2 > (1u << 28)) shift++;
	const uint32_t nb = (uint32_t)(span >> shift) + 1;
	bucket.assign(nb + 1, 0);
	size_t j = 0;
	for (uint32_t b = 0; b <= nb; b++) {
		const uint64_t lim = (uint64_t)b << shift;
		while (j < f.dstart.size() && f.dstart[j] < lim) j++;
		bucket[b] = (uint32_t)j;
	}
}

void build_d4(const FlatIndex& f, std::vector<uint32_t>& d4) {
	d4.assign(4 * ((size_t)f.D + 1), 0);
	for (uint32_t d = 0; d <= f.D; d++) {
		d4[4 * (size_t)d] = f.dlev[d].k; d4[4 * (size_t)d + 1] = f.dlev[d].cent_begin;
		if (d < f.D) d4[4 * (size_t)d + 2] = (uint32_t)f.dinfo[d] + ((uint32_t)(f.dinfo[d] >> 32) & 0xFFFF);   // entries below it are back-walk candidates
		d4[4 * (size_t)d + 3] = f.dtin[d];
	}
}

// Sparse cohorts: per sample the ascending list of the walk entries whose target it carries, and the marker entries.
void build_sparse_walk(const FlatIndex& f, std::vector<uint64_t>& car_begin, std::vector<uint32_t>& car, std::vector<uint32_t>& marker_list) {
	const size_t E = f.cent.size();
	car_begin.assign((size_t)f.num_samples + 1, 0);
	marker_list.clear();
	auto for_carriers = [&](uint32_t set_id, auto&& fn) {
		if (f.class_mode) {
			for (uint32_t w = 0; w < f.words_per_set; w++)
				for (uint64_t m = f.bitmap[(uint64_t)set_id * f.words_per_set + w]; m; m &= m - 1) { const uint32_t s = w * 64 + (uint32_t)__builtin_ctzll(m); if (s && s < f.num_samples) fn(s); }
		} else for (uint64_t i = f.list_begin[set_id]; i < f.list_begin[set_id + 1]; i++) { const uint32_t s = f.list_ids[i]; if (s && s < f.num_samples) fn(s); }
	};
	for (size_t c = 0; c < E; c++) {
		if (f.cent[c].tgt & kEntMarker) { marker_list.push_back((uint32_t)c); continue; }
		for_carriers(f.cent[c].set_id, [&](uint32_t s) { car_begin[s + 1]++; });
	}
	for (uint32_t s = 0; s < f.num_samples; s++) car_begin[s + 1] += car_begin[s];
	car.assign(car_begin[f.num_samples], 0);
	std::vector<uint64_t> at(car_begin.begin(), car_begin.end() - 1);
	for (size_t c = 0; c < E; c++) {
		if (f.cent[c].tgt & kEntMarker) continue;
		for_carriers(f.cent[c].set_id, [&](uint32_t s) { car[at[s]++] = (uint32_t)c; });
	}
}
// The walk of every sample from the head of the contig under the rules of get_sample_var_in_ref (query.h:649-716) as the
// kernels apply them (device_logic.cuh: fwd_step), with no region bounds: first carrying entry of the current vertex wins,
// a taken entry hides everything before the vertex it rejoins, an alt vertex without an out-edge ends the path.
void build_canonical_walks(const FlatIndex& f, const std::vector<uint64_t>& car_begin, const std::vector<uint32_t>& car, const std::vector<uint32_t>& marker_list,
                           std::vector<uint64_t>& can_begin, std::vector<uint32_t>& can_entry, std::vector<uint32_t>& can_pmax) {
	(void)marker_list;                                      // markers are not edges: with no region end they never stop the walk
	const uint32_t S = f.num_samples;
	std::vector<std::vector<uint32_t>> taken(S);
	parallel_for(S, [&](uint64_t a, uint64_t b) {
		for (uint64_t s = a; s < b; s++) {
			uint32_t cur_k = 0;
			for (uint64_t i = car_begin[s]; i < car_begin[s + 1]; i++) {
				const CEntry& e = f.cent[car[i]];
				if (e.src < cur_k) continue;                        // hidden behind a taken detour / an earlier sibling
				taken[s].push_back(car[i]);
				const uint32_t tk = e.tgt & kEntTgtMask;
				if ((e.tgt & kEntAlt) && tk == kEntTgtNone) break;  // the path ends on this vertex
				cur_k = tk;
			}
		}
	});
	can_begin.assign((size_t)S + 1, 0);
	for (uint32_t s = 0; s < S; s++) can_begin[s + 1] = can_begin[s] + taken[s].size();
	can_entry.resize(can_begin[S]); can_pmax.resize(can_begin[S]);
	parallel_for(S, [&](uint64_t a, uint64_t b) {
		for (uint64_t s = a; s < b; s++) {
			uint32_t pm = 0; uint64_t at = can_begin[s];
			for (uint32_t c : taken[s]) { pm = std::max(pm, f.cent[c].arrival); can_entry[at] = c; can_pmax[at] = pm; at++; }
		}
	});
}
// marker_span (DevIndex): a marker ends a walk only if its arrival is >= the region's y; the region's scan bound is then at
// most the first entry of the backbone vertex starting at or after that arrival.  The largest distance, in entries, from a
// marker to that bound says how far below a scan bound a marker can still matter.
uint32_t marker_span(const FlatIndex& f, const std::vector<uint32_t>& marker_list) {
	uint32_t span = 0;
	for (uint32_t c : marker_list) {
		const uint32_t a = f.cent[c].arrival;
		const uint32_t k = (uint32_t)(std::lower_bound(f.vstart.begin(), f.vstart.end(), a) - f.vstart.begin());
		const uint32_t reach = f.cent_begin[std::min<uint32_t>(k, f.M)];
		if (reach > c) span = std::max(span, reach - c);
	}
	return span + 1;
}
// Which membership structure the t4 walk gets: the per-sample lists when the cohort is sparse (explicit-id encoding, or on
// average fewer than one carried entry per 1024), else the hit map.  VSGPU_SPARSE_WALK=0/1 overrides.
bool want_sparse_walk(const FlatIndex& f) {
	if (const char* e = getenv("VSGPU_SPARSE_WALK")) return atoi(e) != 0;
	if (!f.class_mode) return true;
	return false;
}

void build_host_index(const std::string& prefix, HostIndex& h, int* stage) {
	h.prefix = prefix;
	if (stage) *stage = 0;
	PhaseClock pc;
	h.from_cache = load_index_cache(prefix, h);          // VSGPU_INDEX_CACHE (index_cache.cc); false when unset, stale or unreadable
	if (h.from_cache) pc.lap("flattened index read from cache");
	else {
		load_ser(prefix, h.ser);
		if (stage) *stage = 1;
		flatten(h.ser, h.flat);
		uint64_t le = (uint64_t)h.flat.vstart[h.flat.M - 1] + h.flat.vlen[h.flat.M - 1];
		h.last_end = le > 0xFFFFFFFEull ? 0xFFFFFFFEu : (uint32_t)le;
		{   // largest d whose first-branchy range is non-empty; the position just before the next start
			const FlatIndex& f = h.flat;
			h.t1_fallback_pos = 0;
			for (uint32_t d = f.D; d-- > 0;) if (f.t7_hi[d] > f.t7_lo[d]) { h.t1_fallback_pos = d + 1 < f.D ? f.dstart[d + 1] - 1 : 0xFFFFFFFEu; break; }
		}
		// host copies nothing reads after flatten()
		SerData& s = h.ser;
		std::vector<uint64_t>().swap(s.sample_vector); std::vector<uint32_t>().swap(s.index_ones); std::vector<uint32_t>().swap(s.node_list);
		std::vector<uint32_t>().swap(s.v_first_index); std::vector<uint32_t>().swap(s.v_ref0_index);
		std::vector<uint64_t>().swap(s.adj_begin); std::vector<uint32_t>().swap(s.adj);
		pc = PhaseClock();
		save_index_cache(prefix, h);
		if (getenv("VSGPU_INDEX_CACHE")) pc.lap("flattened index written to cache");
	}
	for (uint32_t i = 0; i < h.ser.num_samples; i++) h.name2id[h.ser.sample_names[i]] = i;
	if (stage) *stage = 2;
}

// ------------------------------------------------------------------ row materialisation
static const char kBase[8] = {'A', 'C', 'T', 'G', 'N', 5, 5, 5};   // map_int, src/util.cc:32-41

void append_seq(const HostIndex* ix, uint32_t v, std::string& out) {   // VariantGraph::get_sequence, variant_graph.h:1261-1268
	if (v == kNone) return;
	const uint8_t* p = &ix->ser.seq[ix->ser.v_offset[v]];
	for (uint32_t i = 0, n = ix->ser.v_length[v]; i < n; i++) out += kBase[p[i] & 7];
}

// get_samples (query.h:268-285): every non-ref carrier of vertex v as name(phasing), s_info order
template <class F>
void for_each_carrier(const HostIndex* ix, uint32_t v, F&& fn) {
	const SerData& s = ix->ser; const FlatIndex& f = ix->flat;
	const uint64_t b = s.v_sinfo_begin[v], e = s.v_sinfo_begin[v + 1];
	if (s.class_mode) {
		const uint64_t* row = &f.bitmap[(uint64_t)s.v_class[v] * f.words_per_set];
		uint64_t i = b;
		for (uint32_t w = 0; w < f.words_per_set && i < e; w++) {
			uint64_t bits = row[w];
			while (bits && i < e) {
				uint32_t id = w * 64 + (uint32_t)__builtin_ctzll(bits); bits &= bits - 1;
				if (id != 0) fn(id, s.s_flags[i]);
				i++;
			}
		}
	} else {
		for (uint64_t i = b; i < e; i++) if (s.s_sample_id[i] != 0) fn(s.s_sample_id[i], s.s_flags[i]);
	}
}
void append_carriers(const HostIndex* ix, uint32_t v, std::string& out) {
	for_each_carrier(ix, v, [&](uint32_t id, uint8_t fl) {
		out += ix->ser.sample_names[id]; out += '(';
		out += (fl & 2) ? '1' : '0'; out += (fl & 1) ? '|' : '/'; out += (fl & 4) ? '1' : '0';   // get_sample_phasing, variant_graph.h:882-900
		out += ") ";
	});
}

void t6_row(const HostIndex* ix, uint32_t r, bool with_samples, std::string& out) {
	const FlatIndex& f = ix->flat;
	out += std::to_string(f.rec_pos[r]); out += '\t';
	append_seq(ix, f.rec_refv[r], out); out += '\t';
	append_seq(ix, f.rec_altv[r], out); out += '\t';
	if (with_samples && !(f.rec_flags[r] & 4)) append_carriers(ix, f.rec_vertex[r], out);
	out += '\n';
}

bool rec_same_pos_alt(const HostIndex* ix, uint32_t a, uint32_t b) {
	const FlatIndex& f = ix->flat; const SerData& s = ix->ser;
	if (f.rec_pos[a] != f.rec_pos[b]) return false;
	uint32_t va = f.rec_altv[a], vb = f.rec_altv[b];
	uint32_t la = va == kNone ? 0 : s.v_length[va], lb = vb == kNone ? 0 : s.v_length[vb];
	if (la != lb) return false;
	return la == 0 || memcmp(&s.seq[s.v_offset[va]], &s.seq[s.v_offset[vb]], la) == 0;
}

// the "only add var if not seen before" rule of next_variant_in_ref (query.h:397-414)
bool push_rule(const HostIndex* ix, const std::vector<uint32_t>& vars, uint32_t r) {
	const FlatIndex& f = ix->flat;
	if (vars.empty()) return true;
	if (rec_same_pos_alt(ix, vars.back(), r)) return false;
	if (vars.size() > 1 && f.rec_pos[vars.back()] == f.rec_pos[r]) {
		for (size_t q = vars.size(); q-- > 0;) {
			if (f.rec_pos[vars[q]] < f.rec_pos[r]) break;
			if (rec_same_pos_alt(ix, vars[q], r)) return false;
		}
	}
	return true;
}

uint32_t host_rank(const FlatIndex& f, uint64_t pos) { return (uint32_t)(std::upper_bound(f.dstart.begin(), f.dstart.end(), pos > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)pos) - f.dstart.begin()); }

// Literal get_var_in_ref loop (query.h:758-771) over the flattened tables.  Only used for regions
// whose record slice contains suspect duplicates or that run past the end of the contig, where the
// reference's re-find / dedup interplay is not a plain slice.
void t6_literal(const HostIndex* ix, uint64_t x, uint64_t y, std::vector<uint32_t>& vars) {
	const FlatIndex& f = ix->flat;
	vars.clear();
	uint64_t cur = x; uint64_t guard = 0;
	while (cur < y && guard++ < (uint64_t)f.M + 8) {
		bool found = false;
		uint32_t rk = cur >= f.index_bits ? f.D : host_rank(f, cur);
		if (rk < 1) rk = 1;
		uint32_t k = f.dlev[rk - 1].k;
		for (; k < f.M; k++) {
			if ((uint64_t)f.vstart[k] + f.vlen[k] >= y) break;
			for (uint32_t r = f.rec_begin[k]; r < f.rec_begin[k + 1]; r++) if (push_rule(ix, vars, r)) { vars.push_back(r); found = true; }
			if (found) break;
		}
		if (!found) break;
		cur = k + 1 < f.M ? f.vstart[k + 1] : 1;   // *next_it is vertex 0 (index 1) once the ref path is exhausted
		if (cur >= y) break;
	}
}

bool t6_needs_literal(const HostIndex* ix, uint64_t y, uint32_t lo, uint32_t hi) {
	const FlatIndex& f = ix->flat;
	if (hi > lo && f.has_suspect_dups && f.rec_dup_prefix[hi] - f.rec_dup_prefix[lo] > 0) return true;
	if (y > ix->last_end && hi > lo && f.rec_begin[f.M] > f.rec_begin[f.M - 1]) return true;
	return false;
}

// sample_info.index of `sample` in `vertex` (HostIndex::sindex must be loaded): s_info[i] belongs to the
// i-th set bit of the vertex's class, or to the i-th listed id (variant_graph.h:1296-1326)
uint32_t host_sample_index(const HostIndex* ix, uint32_t vertex, uint32_t sample) {
	const FlatIndex& f = ix->flat; const SerData& s = ix->ser;
	const uint64_t b = s.v_sinfo_begin[vertex], e = s.v_sinfo_begin[vertex + 1];
	if (f.class_mode) {
		const uint64_t* row = &f.bitmap[(uint64_t)s.v_class[vertex] * f.words_per_set];
		uint64_t r = 0;
		for (uint32_t w = 0; w < (sample >> 6); w++) r += (uint64_t)__builtin_popcountll(row[w]);
		r += (uint64_t)__builtin_popcountll(row[sample >> 6] & (((uint64_t)1 << (sample & 63)) - 1));
		return b + r < e ? ix->sindex[b + r] : 0;
	}
	for (uint64_t i = b; i < e; i++) if (s.s_sample_id[i] == sample) return ix->sindex[i];
	return 0;
}

// Row of a t4 / t5 hit code: emission rules of get_sample_var_in_ref (query.h:677-710) and, with
// sample != kNone, of get_sample_var_in_sample (:553-590) — the same three cases with var_pos = ref_pos
// for an insertion and the sample's own position in the vertex otherwise.
// What the row of a hit code prints, as vertex ids: var_pos, the vertex whose sequence is the ref column (kNone: ""), the one
// of the alt column, and the vertex u whose carriers are listed.  sample == kNone: t4; else t5 (positions in the sample's coordinates).
void hit_row_parts(const HostIndex* ix, uint32_t code, uint32_t sample, uint64_t& pos, uint32_t& refv, uint32_t& altv, uint32_t& u, bool* insertion) {
	const FlatIndex& f = ix->flat; const SerData& s = ix->ser;
	const bool t5 = sample != kNone;
	const uint32_t c = code & VSGPU_HIT_ENTRY_MASK;
	const CEntry& e = f.cent[c];
	uint32_t ref_pos, cur_ref_v; bool cur_ref_empty = false, u_is_bb; uint32_t u_k = kNone;
	if (code & VSGPU_HIT_REJOIN) {
		u_k = e.tgt & kEntTgtMask; u = f.bb_vertex[u_k]; u_is_bb = true;
		ref_pos = f.vstart[u_k]; cur_ref_v = u;                         // state left by the alt vertex: ref_pos = index(N), cur_ref = seq(N)
	} else {
		u = f.cent_vertex[c]; ref_pos = e.arrival; cur_ref_v = f.bb_nref[e.src];
		u_is_bb = !(e.tgt & kEntAlt);
		if (u_is_bb) u_k = e.tgt & kEntTgtMask;
		if (code & VSGPU_HIT_START) cur_ref_empty = true;
	}
	// next_ref_pos at u (query.h:660-674)
	uint64_t next_ref_pos;
	if (u_is_bb) next_ref_pos = f.bb_nrp[u_k];
	else {
		uint32_t tk = e.tgt & kEntTgtMask;
		next_ref_pos = tk == kEntTgtNone ? (uint64_t)ref_pos + s.v_length[u] : f.vstart[tk];
	}
	refv = kNone; altv = kNone;
	if (insertion) *insertion = ref_pos == next_ref_pos;
	if (ref_pos == next_ref_pos) { pos = t5 ? (uint64_t)ref_pos : (uint64_t)ref_pos - 1; altv = u; }   // insertion
	else if (u_is_bb) {                                                                                   // deletion: cur_ref = seq(find(ref_pos - 1))
		uint64_t p = ref_pos > 1 ? ref_pos - 1 : 1;
		uint32_t rk = p >= f.index_bits ? f.D : host_rank(f, p);
		if (rk < 1) rk = 1;
		uint32_t wk = f.dlev[rk - 1].k;
		refv = f.bb_vertex[wk]; pos = t5 ? host_sample_index(ix, u, sample) : f.vstart[wk];
	} else { pos = t5 ? host_sample_index(ix, u, sample) : ref_pos; if (!cur_ref_empty) refv = cur_ref_v; altv = u; } // substitution
}

// Row of a t4 / t5 hit code: emission rules of get_sample_var_in_ref (query.h:677-710) and, with
// sample != kNone, of get_sample_var_in_sample (:553-590) — the same three cases with var_pos = ref_pos
// for an insertion and the sample's own position in the vertex otherwise.
static void hit_row(const HostIndex* ix, uint32_t code, uint32_t sample, bool with_samples, std::string& out) {
	uint64_t pos; uint32_t refv, altv, u;
	hit_row_parts(ix, code, sample, pos, refv, altv, u);
	std::string ref, alt;
	if (refv != kNone) append_seq(ix, refv, ref);
	if (altv != kNone) append_seq(ix, altv, alt);
	out += std::to_string(pos); out += '\t'; out += ref; out += '\t'; out += alt; out += '\t';
	if (with_samples) append_carriers(ix, u, out);
	out += '\n';
}
void t4_row(const HostIndex* ix, uint32_t code, bool with_samples, std::string& out) { hit_row(ix, code, kNone, with_samples, out); }
void t5_row(const HostIndex* ix, uint32_t code, uint32_t sample, bool with_samples, std::string& out) { hit_row(ix, code, sample, with_samples, out); }

uint64_t fnv1a(uint64_t h, const void* data, size_t n) {
	const unsigned char* p = (const unsigned char*)data;
	for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ULL; }
	return h;
}

uint64_t hash_query(const char* ref, const char* alt) { return hash_ref_alt(ref, strlen(ref), alt, strlen(alt)); }

// t7 post-check on the host: a device hit is a 64-bit hash match; confirm the strings (a colliding
// hash would otherwise be a false positive) and, if they differ, finish the compare on the strings.
uint32_t t7_confirm(const HostIndex* ix, uint64_t pos, const char* ref, const char* alt, uint32_t r) {
	if (r == VSGPU_NONE) return r;
	const FlatIndex& f = ix->flat;
	std::string a, b;
	append_seq(ix, f.rec_refv[r], a); append_seq(ix, f.rec_altv[r], b);
	if (a == ref && b == alt) return r;
	uint32_t k = f.rec_k[r];
	for (uint32_t q = f.rec_begin[k]; q < f.rec_begin[k + 1]; q++) {
		if (!(f.rec_flags[q] & 1) || f.rec_pos[q] != pos) continue;
		a.clear(); b.clear(); append_seq(ix, f.rec_refv[q], a); append_seq(ix, f.rec_altv[q], b);
		if (a == ref && b == alt) return q;
	}
	return VSGPU_NONE;
}


uint64_t t7_carriers(const HostIndex* ix, uint32_t rec, std::string* text, uint64_t* digest) {
	uint64_t h = kFnvInit, cnt = 0;
	if (rec != kNone && rec < ix->flat.R && !(ix->flat.rec_flags[rec] & 4))
		for_each_carrier(ix, ix->flat.rec_vertex[rec], [&](uint32_t id, uint8_t fl) {
			const std::string& nm = ix->ser.sample_names[id];
			char gt[3] = {(fl & 2) ? '1' : '0', (fl & 1) ? '|' : '/', (fl & 4) ? '1' : '0'};
			if (text) { *text += nm; *text += ' '; text->append(gt, 3); }
			h = fnv1a(h, nm.data(), nm.size()); h = fnv1a(h, " ", 1); h = fnv1a(h, gt, 3); cnt++;
		});
	if (digest) *digest = h;
	return cnt;
}

void rows_t1(const HostIndex* ix, uint32_t lo, uint32_t hi, bool with_samples, std::string& s, uint64_t& cnt) {
	cnt = 0;
	if (lo == kNone || hi > ix->flat.R) return;
	for (uint32_t r = lo; r < hi; r++) if (ix->flat.rec_flags[r] & 1) { t6_row(ix, r, with_samples, s); cnt++; }
}

void digests_t1(const HostIndex* ix, uint64_t n, const uint32_t* lo, const uint32_t* hi, bool with_samples, uint64_t* counts, uint64_t* digests) {
	parallel_for(n, [&](uint64_t a, uint64_t b) {
		std::string row;
		for (uint64_t i = a; i < b; i++) {
			uint64_t h = kFnvInit, c = 0;
			if (lo[i] != kNone && hi[i] <= ix->flat.R)
				for (uint32_t r = lo[i]; r < hi[i]; r++) if (ix->flat.rec_flags[r] & 1) { row.clear(); t6_row(ix, r, with_samples, row); h = fnv1a(h, row.data(), row.size()); c++; }
			if (digests) digests[i] = h;
			if (counts) counts[i] = c;
		}
	});
}

void rows_t6(const HostIndex* ix, uint32_t lo, uint32_t hi, bool with_samples, std::string& s, uint64_t& cnt) {
	cnt = 0;
	if (lo == kNone && hi == kNone) return;        // the is_empty gate fired: no rows
	std::vector<uint32_t> vars;
	if (ix->flat.has_suspect_dups && ix->flat.rec_dup_prefix[hi] - ix->flat.rec_dup_prefix[lo] > 0) {
		for (uint32_t r = lo; r < hi; r++) if (push_rule(ix, vars, r)) vars.push_back(r);
		for (uint32_t r : vars) { t6_row(ix, r, with_samples, s); cnt++; }
	} else for (uint32_t r = lo; r < hi; r++) { t6_row(ix, r, with_samples, s); cnt++; }
}

void digests_t6(const HostIndex* ix, uint64_t n, const uint32_t* lo, const uint32_t* hi, bool with_samples, uint64_t* digests, bool* bad_out) {
	std::atomic<bool> bad{false};
	parallel_for(n, [&](uint64_t a, uint64_t b) {
		std::string row; std::vector<uint32_t> vars;
		for (uint64_t i = a; i < b; i++) {
			uint64_t h = kFnvInit;
			if (lo[i] == kNone && hi[i] == kNone) { digests[i] = h; continue; }
			if (lo[i] > hi[i] || hi[i] > ix->flat.R) { bad = true; continue; }
			bool dd = ix->flat.has_suspect_dups && ix->flat.rec_dup_prefix[hi[i]] - ix->flat.rec_dup_prefix[lo[i]] > 0;
			vars.clear();
			for (uint32_t r = lo[i]; r < hi[i]; r++) {
				if (dd) { if (!push_rule(ix, vars, r)) continue; vars.push_back(r); }
				row.clear(); t6_row(ix, r, with_samples, row); h = fnv1a(h, row.data(), row.size());
			}
			digests[i] = h;
		}
	});
	if (bad_out) *bad_out = bad;
}

void digests_t4(const HostIndex* ix, uint64_t n, const uint64_t* offsets, const uint32_t* hits, bool with_samples, uint64_t* digests) {
	parallel_for(n, [&](uint64_t a, uint64_t b) {
		std::string row;
		for (uint64_t i = a; i < b; i++) {
			uint64_t h = kFnvInit;
			for (uint64_t j = offsets[i]; j < offsets[i + 1]; j++) { row.clear(); t4_row(ix, hits[j], with_samples, row); h = fnv1a(h, row.data(), row.size()); }
			digests[i] = h;
		}
	});
}

void digests_t5(const HostIndex* ix, uint64_t n, const uint64_t* offsets, const uint32_t* hits, const uint32_t* samples, bool with_samples, uint64_t* digests) {
	parallel_for(n, [&](uint64_t a, uint64_t b) {
		std::string row;
		for (uint64_t i = a; i < b; i++) {
			uint64_t h = kFnvInit;
			for (uint64_t j = offsets[i]; j < offsets[i + 1]; j++) { row.clear(); t5_row(ix, hits[j], samples[i], with_samples, row); h = fnv1a(h, row.data(), row.size()); }
			digests[i] = h;
		}
	});
}

void digests_t7(const HostIndex* ix, uint64_t n, const uint32_t* rec, uint64_t* ncarriers, uint64_t* digests) {
	parallel_for(n, [&](uint64_t a, uint64_t b) {
		for (uint64_t i = a; i < b; i++) { uint64_t d; uint64_t c = t7_carriers(ix, rec[i], nullptr, &d); if (digests) digests[i] = d; if (ncarriers) ncarriers[i] = c; }
	});
}

}  // namespace vsgpu
