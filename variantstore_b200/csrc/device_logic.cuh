// libvsgpu — per-region logic of the query kernels, written once as __host__ __device__ inline
// functions.  kernels.cu instantiates it inside the sm_100a kernels; tests/hostsim compiles the
// same text for the host so the flattened tables and the walk rules can be checked against the
// oracle on a machine without a GPU (test-only: libvsgpu.so never runs this code on the CPU).
#pragma once
#include "kernels.cuh"

#if defined(__CUDACC__)
#define VSGPU_HD __host__ __device__ __forceinline__
#else
#define VSGPU_HD inline
#endif

namespace vsgpu {
namespace logic {

template <class T> VSGPU_HD T ldg(const T* p) {
#if defined(__CUDA_ARCH__)
	return __ldg(p);
#else
	return *p;
#endif
}
VSGPU_HD uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }

constexpr uint32_t kEntAlt = 1u << 31, kEntTgtCarriers = 1u << 30, kEntMarker = 1u << 29, kEntTgtMask = (1u << 29) - 1;
constexpr uint32_t kHitStart = 0x80000000u, kHitRejoin = 0x40000000u;
constexpr uint32_t kNoneU32 = 0xFFFFFFFFu;

VSGPU_HD uint32_t clamp_pos(uint64_t v) { return v > 0xFFFFFFFEull ? 0xFFFFFFFEu : (uint32_t)v; }

// rank(x) = number of distinct backbone starts <= x  (rank_rrrb(pos) of index.h:128,142,158).
// A direct-mapped bucket table over positions gives the slice of `dstart` that can hold the answer
// (about kBucketTarget keys, one or two 128-byte lines); a short binary search finishes it.
VSGPU_HD uint32_t rank_le(const DevIndex& ix, uint32_t x) {
	uint32_t b = x >> ix.bucket_shift;
	if (b >= ix.nbuckets) b = ix.nbuckets - 1;
	const uint2 r = make_uint2(ldg(ix.bucket + b), ldg(ix.bucket + b + 1));
	uint32_t lo = r.x, hi = b + 1 == ix.nbuckets ? ix.D : r.y;
	while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (ldg(ix.dstart + mid) <= x) lo = mid + 1; else hi = mid; }
	return lo;
}

// Two independent ranks at once: both bucket lookups are issued together and the two binary
// searches advance in lock step, so every step has two loads in flight instead of one.
VSGPU_HD void rank_le2(const DevIndex& ix, uint32_t xa, uint32_t xb, uint32_t& ra, uint32_t& rb) {
	uint32_t ba = xa >> ix.bucket_shift, bb = xb >> ix.bucket_shift;
	if (ba >= ix.nbuckets) ba = ix.nbuckets - 1;
	if (bb >= ix.nbuckets) bb = ix.nbuckets - 1;
	uint32_t lo1 = ldg(ix.bucket + ba), hi1 = ldg(ix.bucket + ba + 1), lo2 = ldg(ix.bucket + bb), hi2 = ldg(ix.bucket + bb + 1);
	if (ba + 1 == ix.nbuckets) hi1 = ix.D;
	if (bb + 1 == ix.nbuckets) hi2 = ix.D;
	while (lo1 < hi1 || lo2 < hi2) {
		const uint32_t m1 = (lo1 + hi1) >> 1, m2 = (lo2 + hi2) >> 1;
		const bool a1 = lo1 < hi1, a2 = lo2 < hi2;
		const uint32_t k1 = a1 ? ldg(ix.dstart + m1) : 0, k2 = a2 ? ldg(ix.dstart + m2) : 0;
		if (a1) { if (k1 <= xa) lo1 = m1 + 1; else hi1 = m1; }
		if (a2) { if (k2 <= xb) lo2 = m2 + 1; else hi2 = m2; }
	}
	ra = lo1; rb = lo2;
}

VSGPU_HD bool member(const DevIndex& ix, uint32_t s, uint32_t set_id) {
	if (ix.class_mode) return (ldg(ix.bitmap + (uint64_t)set_id * ix.words_per_set + (s >> 6)) >> (s & 63)) & 1;
	for (uint64_t i = ldg(ix.list_begin + set_id), e = ldg(ix.list_begin + set_id + 1); i < e; i++)
		if (ldg(ix.list_ids + i) == s) return true;
	return false;
}

// ------------------------------------------------------------------ t6 slice bounds (query.h:736-784)
VSGPU_HD uint2 t6_bounds(const DevIndex& ix, uint64_t x64, uint64_t y64, bool* bad) {
	uint2 r = make_uint2(kNoneU32, kNoneU32);                                // is_empty gate fired: (NONE, NONE)
	if (x64 < 1) { *bad = true; return r; }
	if (x64 > ix.index_bits) return r;                                     // is_empty: pos_x > size -> empty
	const uint32_t x = clamp_pos(x64), y = clamp_pos(y64);
	uint32_t rk, e;                                                        // rank(x); first start >= y
	rank_le2(ix, x, y ? y - 1 : 0, rk, e);
	if (!y) e = 0;
	// gate (index.h:158-165): next distinct start s' must satisfy s' - 1 <= y
	if (rk >= 1 && rk < ix.D && (uint64_t)ldg(ix.dstart + rk) <= (uint64_t)y + 1) {
		const uint32_t lo = ldg(&ix.dlev[rk - 1].y);
		uint32_t hi;
		if (e < ix.D) hi = ldg(&ix.dlev[e].z);                              // rec_begin[k(e) - 1]
		else hi = ((uint64_t)ix.last_end >= y) ? ldg(&ix.dlev[ix.D].z) : ix.R;
		r = make_uint2(lo, hi > lo ? hi : lo);
	}
	return r;
}

// ------------------------------------------------------------------ t7 lookup (query.h:792-823)
VSGPU_HD uint32_t t7_lookup(const DevIndex& ix, uint64_t p64, uint64_t h, bool* bad) {
	if (p64 < 1) { *bad = true; return kNoneU32; }
	const uint32_t p = clamp_pos(p64);
	uint32_t rk = p64 >= ix.index_bits ? ix.D : rank_le(ix, p);      // Index::find (index.h:125-132)
	if (rk < 1) rk = 1;
	const uint2 rng = ldg(ix.t7rng + (rk - 1));
	for (uint32_t r = rng.x; r < rng.y; r++)
		if ((ldg(ix.rec_flags + r) & 1) && p64 == ldg(ix.rec_pos + r) && ldg(ix.rec_hash + r) == h) return r;
	return kNoneU32;
}

// ------------------------------------------------------------------ t1 closest_var (query.h:441-483)
// Records of the first backbone vertex at/after `pos` that has any (one next_variant_in_ref call).
VSGPU_HD uint2 first_branchy(const DevIndex& ix, uint64_t p64) {
	uint32_t rk = p64 >= ix.index_bits ? ix.D : rank_le(ix, clamp_pos(p64));
	if (rk < 1) rk = 1;
	return ldg(ix.t7rng + (rk - 1));
}
// (NONE, NONE): the operator returns false; lo == hi: true with no rows; else the record range whose
// kept records (rec_flags bit 0) are the rows.
VSGPU_HD uint2 t1_lookup(const DevIndex& ix, uint64_t p64, bool* bad) {
	if (p64 < 1) { *bad = true; return make_uint2(kNoneU32, kNoneU32); }
	const uint2 a = first_branchy(ix, p64);
	if (a.x < a.y) {
		const uint64_t p1 = ldg(ix.rec_pos + a.x);                            // next_var[0].var_pos
		const int cur = (int)(uint32_t)(p64 - (p1 - p64));                    // int cur_pos = pos-(next_var_pos-pos)  (:451)
		if (cur > 0) {
			const uint2 b = first_branchy(ix, (uint64_t)cur);
			if (b.x < b.y && ldg(ix.rec_pos + b.x) != p1) return b;             // an earlier variant sits inside the mirrored window
		}
		return a;
	}
	// nothing at or after pos: step back until a call finds something (:464-470)
	if (p64 < 2) return make_uint2(0, 0);
	if (ix.t1_fallback_pos == 0) return make_uint2(kNoneU32, kNoneU32);      // no variant anywhere: returns false at cur_pos == 1
	uint64_t cur = p64 - 1;
	if (cur > ix.t1_fallback_pos) cur = ix.t1_fallback_pos;                  // largest position whose call succeeds
	return first_branchy(ix, cur);
}

// ------------------------------------------------------------------ t4 walk (one thread per region)
struct CountSink {
	uint32_t* dst; uint32_t n;
	VSGPU_HD void emit(uint32_t code) { if (n < kScratchHits) dst[n] = code; n++; }
};
struct DirectSink {
	uint32_t* dst; uint32_t n;
	VSGPU_HD void emit(uint32_t code) { dst[n++] = code; }
};

template <class Sink>
VSGPU_HD void walk_region(const DevIndex& ix, uint64_t x64, uint64_t y64, uint32_t s, Sink& sink) {
	if (x64 > ix.index_bits) return;
	const uint32_t x = clamp_pos(x64), y = clamp_pos(y64);
	const uint32_t rk = rank_le(ix, x);
	if (rk < 1 || rk >= ix.D) return;
	if ((uint64_t)ldg(ix.dstart + rk) > (uint64_t)y + 1) return;                // is_empty gate
	// ---- get_prev_vertex_with_sample (query.h:57-113)
	uint64_t cur = (x64 >= ix.index_bits) ? ix.D - 1 : rk - 1;                   // ref_node_rank (index.h:135-148)
	uint32_t c_found = kNoneU32;
	for (;;) {
		if (cur > ix.D) cur = 0;                                                   // reference: unsigned wrap + OOB read; defined here as vertex 0
		if (cur <= 1) break;
		const uint64_t info = ldg(ix.dinfo + (cur - 1));
		const uint32_t cb = (uint32_t)info, ncar = (uint32_t)(info >> 32) & 0xFFFF, deg = (uint32_t)(info >> 48);
		for (uint32_t c = cb; c < cb + ncar; c++) if (member(ix, s, ldg(&ix.cent[c].z))) c_found = c;   // last carrier wins
		cur -= deg;                                                                // once per neighbour (:103)
		if (c_found != kNoneU32) break;
	}
	// ---- forward walk (query.h:649-716)
	const uint32_t e_y = y ? rank_le(ix, y - 1) : 0;
	const uint32_t k_end = ldg(&ix.dlev[e_y].x);                                // first backbone vertex whose start >= y
	uint32_t cur_k = 0, c = 0;
	if (c_found != kNoneU32) {
		const uint4 e = ldg(ix.cent + c_found);
		if (e.w >= y) return;
		if (e.w >= x) sink.emit(c_found | kHitStart);
		if (e.y & kEntAlt) {
			const uint32_t tk = e.y & kEntTgtMask;
			if (tk == kEntTgtMask || tk >= k_end) return;
			if ((e.y & kEntTgtCarriers) && member(ix, s, ldg(ix.bb_set + tk)) && ldg(ix.vstart + tk) >= x) sink.emit(c_found | kHitRejoin);
			cur_k = tk;
		} else cur_k = e.y & kEntTgtMask;
		c = c_found + 1;
	} else {
		if (1 >= y) return;
	}
	for (; c < ix.num_cent; c++) {
		const uint4 e = ldg(ix.cent + c);
		if (e.x < cur_k) continue;                                                 // hidden behind a taken detour / later sibling
		if (e.x > cur_k && e.x >= k_end) break;                                    // the walk stopped before reaching P[e.x]
		if (e.y & kEntMarker) { if (e.w >= y) break; continue; }
		if (!member(ix, s, e.z)) continue;
		if (e.w >= y) break;
		if (e.w >= x) sink.emit(c);
		if (e.y & kEntAlt) {
			const uint32_t tk = e.y & kEntTgtMask;
			if (tk == kEntTgtMask || tk >= k_end) break;
			if ((e.y & kEntTgtCarriers) && member(ix, s, ldg(ix.bb_set + tk)) && ldg(ix.vstart + tk) >= x) sink.emit(c | kHitRejoin);
			cur_k = tk;
		} else cur_k = e.y & kEntTgtMask;
	}
}

// ------------------------------------------------------------------ t4 walk over the sample-major hit map
// hitmap row s, bit c = "sample s is a carrier of the target of walk entry c".  One 32-bit load
// answers the membership question for 32 consecutive entries, so the scan only touches the entries
// the sample actually carries (plus the rare markers).  Same rules, same order, same output as
// walk_region above; used whenever the map fits the memory budget.
VSGPU_HD uint32_t ctz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
	return (uint32_t)__ffs((int)v) - 1;
#else
	return (uint32_t)__builtin_ctz(v);
#endif
}
VSGPU_HD uint32_t clz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
	return (uint32_t)__clz((int)v);
#else
	return (uint32_t)__builtin_clz(v);
#endif
}
// bits [pos, pos+len) of a row, len <= 32
VSGPU_HD uint32_t row_bits(const uint32_t* row, uint32_t pos, uint32_t len) {
	const uint32_t w = pos >> 5, off = pos & 31;
	uint64_t v = ldg(row + w);
	if (off + len > 32) v |= (uint64_t)ldg(row + w + 1) << 32;
	return (uint32_t)(v >> off) & (len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1));
}

// State of the forward walk: scan cursor `c` over the walk entries, the backbone index the walk
// has reached (`cur_k`: entries with a smaller source are hidden), the scan bound `limit`.
struct FwdState { const uint32_t* row; uint32_t c, cur_k, limit, k_end, x, y; };

// One step of the forward walk on a carried entry (or marker) `ci`.  Returns true when the walk ends.
template <class Sink>
VSGPU_HD bool fwd_step(const DevIndex& ix, FwdState& st, uint32_t s, uint32_t ci, const uint4& e, Sink& sink) {
	if (ci >= st.limit) return true;
	if (e.x < st.cur_k) return false;                                          // hidden behind a taken detour / later sibling
	if (e.x > st.cur_k && e.x >= st.k_end) return true;                        // the walk stopped before reaching P[e.x]
	if (e.y & kEntMarker) return e.w >= st.y;
	if (e.w >= st.y) return true;
	if (e.w >= st.x) sink.emit(ci);
	if (e.y & kEntAlt) {
		const uint32_t tk = e.y & kEntTgtMask;
		if (tk == kEntTgtMask || tk >= st.k_end) return true;
		if ((e.y & kEntTgtCarriers) && member(ix, s, ldg(ix.bb_set + tk)) && ldg(ix.vstart + tk) >= st.x) sink.emit(ci | kHitRejoin);
		st.cur_k = tk;
	} else {
		st.cur_k = e.y & kEntTgtMask;
		if (st.cur_k >= st.k_end) { const uint32_t nl = ldg(ix.cent_begin_k + st.cur_k + 1); if (nl > st.limit) st.limit = nl; }   // its own entries still count
	}
	return false;
}

// Ranks, gate, back-walk and the vertex the walk starts on.  Returns true when a forward scan from
// st.c remains to be done.
template <class Sink>
VSGPU_HD bool fast_setup(const DevIndex& ix, uint64_t x64, uint64_t y64, uint32_t s, Sink& sink, FwdState& st) {
	if (x64 > ix.index_bits) return false;
	const uint32_t x = clamp_pos(x64), y = clamp_pos(y64);
	uint32_t rk, e_y;                                                           // rank(x); first start >= y
	rank_le2(ix, x, y ? y - 1 : 0, rk, e_y);
	if (!y) e_y = 0;
	if (rk < 1 || rk >= ix.D) return false;
	// everything the next steps need from the index, issued together
	const uint32_t next_start = ldg(ix.dstart + rk);
	const uint64_t info = ldg(ix.dinfo + (rk >= 2 ? rk - 2 : 0));
	const uint32_t t = ldg(ix.dtin + (rk - 1));
	const uint4 dl = ldg(ix.dlev + e_y);
	if ((uint64_t)next_start > (uint64_t)y + 1) return false;                   // is_empty gate
	const uint32_t* row = ix.hitmap + (uint64_t)s * ix.row_words;
	// ---- get_prev_vertex_with_sample (query.h:57-113)
	// The reference steps back through node_list by out-degree until a neighbour carries the sample.
	// Those steps are the ancestors of the start state in the back-walk forest, so instead of
	// stepping: take the sample's carried entries below the start, highest first (one row word covers
	// 32 entries), and stop at the first whose source is examined by an ancestor state.
	const uint32_t cur = rk - 1;                                                // ref_node_rank (index.h:135-148); x < index_bits here since rk < D
	uint32_t c_found = kNoneU32;
	if (cur >= 2) {
		const uint32_t pos = (uint32_t)info + ((uint32_t)(info >> 32) & 0xFFFF);    // entries below pos are candidates
		if (pos > 0) {
			uint32_t w = (pos - 1) >> 5;
			uint32_t m = ldg(row + w) & (0xFFFFFFFFu >> (31 - ((pos - 1) & 31)));
			for (;;) {
				while (m == 0 && w > 0) { w--; m = ldg(row + w); }
				if (m == 0) break;
				const uint32_t b = 31 - clz32(m);
				m &= ~(1u << b);
				const uint32_t p = (w << 5) + b;
				const uint2 a = ldg(ix.cent_anc + p);
				if (a.x <= t && t <= a.y) { c_found = p; break; }                      // last carrier of the nearest examined vertex
			}
		}
	}
	// ---- the vertex the walk starts on (query.h:649-654)
	st.row = row; st.c = 0; st.cur_k = 0; st.limit = dl.w; st.k_end = dl.x; st.x = x; st.y = y;   // entries >= limit have src >= k_end
	if (c_found != kNoneU32) {
		const uint4 e = ldg(ix.cent + c_found);
		if (e.w >= y) return false;
		if (e.w >= x) sink.emit(c_found | kHitStart);
		if (e.y & kEntAlt) {
			const uint32_t tk = e.y & kEntTgtMask;
			if (tk == kEntTgtMask || tk >= st.k_end) return false;
			if ((e.y & kEntTgtCarriers) && member(ix, s, ldg(ix.bb_set + tk)) && ldg(ix.vstart + tk) >= x) sink.emit(c_found | kHitRejoin);
			st.cur_k = tk;
		} else {
			st.cur_k = e.y & kEntTgtMask;
			if (st.cur_k >= st.k_end) { const uint32_t nl = ldg(ix.cent_begin_k + st.cur_k + 1); if (nl > st.limit) st.limit = nl; }   // its own entries still count
		}
		st.c = c_found + 1;
	} else {
		if (1 >= y) return false;
	}
	return st.c < st.limit;
}

// ---- forward walk (query.h:649-716), one thread
template <class Sink>
VSGPU_HD void fast_forward(const DevIndex& ix, FwdState& st, uint32_t s, Sink& sink) {
	uint32_t w = st.c >> 5;
	uint32_t m = (ldg(st.row + w) | ldg(ix.marker_bits + w)) & (0xFFFFFFFFu << (st.c & 31));
	for (;;) {
		while (m == 0) {
			w++;
			if ((w << 5) >= st.limit) return;
			m = ldg(st.row + w) | ldg(ix.marker_bits + w);
		}
		const uint32_t ci = (w << 5) + ctz32(m);
		m &= m - 1;
		if (ci >= st.limit) return;
		const uint4 e = ldg(ix.cent + ci);
		if (fwd_step(ix, st, s, ci, e, sink)) return;
	}
}

template <class Sink>
VSGPU_HD void walk_region_fast(const DevIndex& ix, uint64_t x64, uint64_t y64, uint32_t s, Sink& sink) {
	FwdState st;
	if (fast_setup(ix, x64, y64, s, sink, st)) fast_forward(ix, st, s, sink);
}


// ------------------------------------------------------------------ t4, second form of the hit-map walk (k_t4p's default)
// Same rules and output as fast_setup / fast_forward above, restructured around what the ncu records of
// those said (profiles/README.md): the walk is a chain of dependent loads, so the sample's hit-map row is
// fetched as three 64-bit chunks at once, right after the ranks, and both the back-walk and the forward
// scan work on those registers, a 64-entry chunk per step; the three per-start lookups read one packed
// table (d4) instead of three.  The two ranks are handed back so that a fused launch can answer t6
// (get_var_in_ref) for the same region without searching again.
VSGPU_HD uint64_t ld64(const uint32_t* p) {          // two consecutive 32-bit words as one little-endian chunk; p is 8-byte aligned
#if defined(__CUDA_ARCH__)
	return __ldg((const unsigned long long*)p);
#else
	uint64_t v; __builtin_memcpy(&v, p, 8); return v;
#endif
}
VSGPU_HD uint32_t ctz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
	return (uint32_t)__ffsll((long long)v) - 1;
#else
	return (uint32_t)__builtin_ctzll(v);
#endif
}
VSGPU_HD uint32_t clz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
	return (uint32_t)__clzll((long long)v);
#else
	return (uint32_t)__builtin_clzll(v);
#endif
}

// three consecutive 64-bit chunks of the sample's hit-map row held in registers; other chunks come from global memory
struct RowChunks {
	const uint32_t* row; uint32_t q0; uint64_t v0, v1, v2;
	VSGPU_HD uint64_t get(uint32_t q) const { const uint32_t j = q - q0; return j == 0 ? v0 : j == 1 ? v1 : j == 2 ? v2 : ld64(row + 2 * q); }
};
struct FwdState2 { RowChunks rc; uint32_t c, cur_k, limit, k_end, x, y; };

template <class Sink>
VSGPU_HD bool fwd_step2(const DevIndex& ix, FwdState2& st, uint32_t s, uint32_t ci, const uint4& e, Sink& sink) {
	if (ci >= st.limit) return true;
	if (e.x < st.cur_k) return false;                                          // hidden behind a taken detour / later sibling
	if (e.x > st.cur_k && e.x >= st.k_end) return true;                        // the walk stopped before reaching P[e.x]
	if (e.y & kEntMarker) return e.w >= st.y;
	if (e.w >= st.y) return true;
	if (e.w >= st.x) sink.emit(ci);
	if (e.y & kEntAlt) {
		const uint32_t tk = e.y & kEntTgtMask;
		if (tk == kEntTgtMask || tk >= st.k_end) return true;
		if ((e.y & kEntTgtCarriers) && member(ix, s, ldg(ix.bb_set + tk)) && ldg(ix.vstart + tk) >= st.x) sink.emit(ci | kHitRejoin);
		st.cur_k = tk;
	} else {
		st.cur_k = e.y & kEntTgtMask;
		if (st.cur_k >= st.k_end) { const uint32_t nl = ldg(ix.cent_begin_k + st.cur_k + 1); if (nl > st.limit) st.limit = nl; }
	}
	return false;
}

template <class Sink>
VSGPU_HD bool fast2_setup(const DevIndex& ix, uint64_t x64, uint64_t y64, uint32_t s, Sink& sink, FwdState2& st, uint32_t& rk, uint32_t& e_y) {
	rk = 0; e_y = 0;
	if (x64 > ix.index_bits) return false;
	const uint32_t x = clamp_pos(x64), y = clamp_pos(y64);
	rank_le2(ix, x, y ? y - 1 : 0, rk, e_y);
	if (!y) e_y = 0;
	if (rk < 1 || rk >= ix.D) return false;
	const uint32_t* row = ix.hitmap + (uint64_t)s * ix.row_words;
	// entries below `pos` are the back-walk's candidates: the chunk holding pos - 1, the one below it and the one
	// above it are fetched now, together; everything up to the forward scan works on them
	const uint32_t pos = ldg(&ix.d4[rk >= 2 ? rk - 2 : 0].z);
	RowChunks& rc = st.rc;
	rc.row = row;
	const uint32_t qb = pos ? (pos - 1) >> 6 : 0, nq = ix.row_words >> 1;
	rc.q0 = qb ? qb - 1 : 0;
	rc.v0 = ld64(row + 2 * rc.q0);
	rc.v1 = rc.q0 + 1 < nq ? ld64(row + 2 * rc.q0 + 2) : 0;
	rc.v2 = rc.q0 + 2 < nq ? ld64(row + 2 * rc.q0 + 4) : 0;
	const uint32_t next_start = ldg(ix.dstart + rk);
	const uint32_t t = ldg(&ix.d4[rk - 1].w);
	const uint4 dl = ldg(ix.d4 + e_y);
	if ((uint64_t)next_start > (uint64_t)y + 1) return false;                   // is_empty gate
	// ---- get_prev_vertex_with_sample (query.h:57-113): highest carried entry below pos whose source an ancestor state examines
	const uint32_t cur = rk - 1;
	uint32_t c_found = kNoneU32;
	if (cur >= 2 && pos > 0) {
		uint32_t q = (pos - 1) >> 6;
		uint64_t m = rc.get(q) & (~(uint64_t)0 >> (63 - ((pos - 1) & 63)));
		for (;;) {
			while (m == 0 && q > 0) { q--; m = rc.get(q); }
			if (m == 0) break;
			const uint32_t b = 63 - clz64(m);
			m &= ~((uint64_t)1 << b);
			const uint32_t p = (q << 6) + b;
			const uint2 a = ldg(ix.cent_anc + p);
			if (a.x <= t && t <= a.y) { c_found = p; break; }
		}
	}
	st.c = 0; st.cur_k = 0; st.limit = dl.y; st.k_end = dl.x; st.x = x; st.y = y;
	if (c_found != kNoneU32) {
		const uint4 e = ldg(ix.cent + c_found);
		if (e.w >= y) return false;
		if (e.w >= x) sink.emit(c_found | kHitStart);
		if (e.y & kEntAlt) {
			const uint32_t tk = e.y & kEntTgtMask;
			if (tk == kEntTgtMask || tk >= st.k_end) return false;
			if ((e.y & kEntTgtCarriers) && member(ix, s, ldg(ix.bb_set + tk)) && ldg(ix.vstart + tk) >= x) sink.emit(c_found | kHitRejoin);
			st.cur_k = tk;
		} else {
			st.cur_k = e.y & kEntTgtMask;
			if (st.cur_k >= st.k_end) { const uint32_t nl = ldg(ix.cent_begin_k + st.cur_k + 1); if (nl > st.limit) st.limit = nl; }
		}
		st.c = c_found + 1;
	} else {
		if (1 >= y) return false;
	}
	return st.c < st.limit;
}

template <class Sink>
VSGPU_HD void fast2_forward(const DevIndex& ix, FwdState2& st, uint32_t s, Sink& sink) {
	uint32_t q = st.c >> 6;
	uint64_t m = (st.rc.get(q) | ld64(ix.marker_bits + 2 * q)) & (~(uint64_t)0 << (st.c & 63));
	for (;;) {
		while (m == 0) {
			q++;
			if ((q << 6) >= st.limit) return;
			m = st.rc.get(q) | ld64(ix.marker_bits + 2 * q);
		}
		const uint32_t ci = (q << 6) + ctz64(m);
		m &= m - 1;
		if (ci >= st.limit) return;
		const uint4 e = ldg(ix.cent + ci);
		if (fwd_step2(ix, st, s, ci, e, sink)) return;
	}
}

// t6 slice bounds from the two ranks of a t4 setup (t6_bounds above, query.h:736-784)
VSGPU_HD uint2 t6_from_ranks(const DevIndex& ix, uint64_t x64, uint64_t y64, uint32_t rk, uint32_t e) {
	uint2 r = make_uint2(kNoneU32, kNoneU32);
	if (x64 > ix.index_bits) return r;
	const uint32_t y = clamp_pos(y64);
	if (rk >= 1 && rk < ix.D && (uint64_t)ldg(ix.dstart + rk) <= (uint64_t)y + 1) {
		const uint32_t lo = ldg(&ix.dlev[rk - 1].y);
		uint32_t hi;
		if (e < ix.D) hi = ldg(&ix.dlev[e].z);
		else hi = ((uint64_t)ix.last_end >= y) ? ldg(&ix.dlev[ix.D].z) : ix.R;
		r = make_uint2(lo, hi > lo ? hi : lo);
	}
	return r;
}

// ------------------------------------------------------------------ t4 on a sparse cohort: per-sample carried-entry lists
// car[car_begin[s] .. car_begin[s+1]) = the walk entries whose target sample s carries, ascending.  The hit map of such a
// cohort is gigabytes of zeros and a back-walk over it scans words until it meets a set bit — tens of thousands of
// dependent loads for a sample with a few hundred variants on the contig.  Here the back-walk is a predecessor search
// and the forward walk a merge of the sample's list with the (global, short) list of marker entries.  Same rules, same
// order, same output as the other walks.
VSGPU_HD uint64_t lower_bound_u32(const uint32_t* a, uint64_t lo, uint64_t hi, uint32_t v) {      // first index in [lo, hi) with a[i] >= v
	while (lo < hi) { const uint64_t m = (lo + hi) >> 1; if (ldg(a + m) < v) lo = m + 1; else hi = m; }
	return lo;
}
template <class Sink>
VSGPU_HD void walk_region_sparse(const DevIndex& ix, uint64_t x64, uint64_t y64, uint32_t s, Sink& sink, uint32_t& rk, uint32_t& e_y) {
	rk = 0; e_y = 0;
	if (x64 > ix.index_bits) return;
	const uint32_t x = clamp_pos(x64), y = clamp_pos(y64);
	rank_le2(ix, x, y ? y - 1 : 0, rk, e_y);
	if (!y) e_y = 0;
	if (rk < 1 || rk >= ix.D) return;
	const uint32_t pos = ldg(&ix.d4[rk >= 2 ? rk - 2 : 0].z);
	const uint32_t next_start = ldg(ix.dstart + rk);
	const uint32_t t = ldg(&ix.d4[rk - 1].w);
	const uint4 dl = ldg(ix.d4 + e_y);
	const uint64_t cb = ldg(ix.car_begin + s), ce = ldg(ix.car_begin + s + 1);
	if ((uint64_t)next_start > (uint64_t)y + 1) return;                         // is_empty gate
	// ---- get_prev_vertex_with_sample: the sample's carried entries below pos, highest first, until one whose source an ancestor state examines
	uint64_t i_next = cb;                                                       // where the forward scan starts in the sample's list
	uint32_t c_found = kNoneU32;
	if (rk - 1 >= 2 && pos > 0) {
		for (uint64_t i = lower_bound_u32(ix.car, cb, ce, pos); i > cb;) {
			i--;
			const uint32_t p = ldg(ix.car + i);
			const uint2 a = ldg(ix.cent_anc + p);
			if (a.x <= t && t <= a.y) { c_found = p; i_next = i + 1; break; }
		}
	}
	FwdState st;
	st.row = nullptr; st.c = 0; st.cur_k = 0; st.limit = dl.y; st.k_end = dl.x; st.x = x; st.y = y;
	if (c_found != kNoneU32) {
		const uint4 e = ldg(ix.cent + c_found);
		if (e.w >= y) return;
		if (e.w >= x) sink.emit(c_found | kHitStart);
		if (e.y & kEntAlt) {
			const uint32_t tk = e.y & kEntTgtMask;
			if (tk == kEntTgtMask || tk >= st.k_end) return;
			if ((e.y & kEntTgtCarriers) && member(ix, s, ldg(ix.bb_set + tk)) && ldg(ix.vstart + tk) >= x) sink.emit(c_found | kHitRejoin);
			st.cur_k = tk;
		} else {
			st.cur_k = e.y & kEntTgtMask;
			if (st.cur_k >= st.k_end) { const uint32_t nl = ldg(ix.cent_begin_k + st.cur_k + 1); if (nl > st.limit) st.limit = nl; }
		}
		st.c = c_found + 1;
	} else {
		if (1 >= y) return;
		// The walk starts at the head of the contig.  Which entries it takes up to x is the sample's business alone
		// (can_entry), and none of them can be reported while the running maximum of their arrivals stays below x: join
		// the walk at the last such entry — it is stepped through fwd_step like any other, which also leaves the walk
		// position it implies — instead of stepping through every entry the sample carries before x.
		uint64_t lo = ldg(ix.can_begin + s), hi = ldg(ix.can_begin + s + 1);
		const uint64_t lo0 = lo;
		while (lo < hi) { const uint64_t m = (lo + hi) >> 1; if (ldg(ix.can_pmax + m) < x) lo = m + 1; else hi = m; }
		if (lo > lo0) {
			const uint32_t c0 = ldg(ix.can_entry + lo - 1);
			if (c0 < st.limit) {
				const uint4 e0 = ldg(ix.cent + c0);
				st.cur_k = e0.x;
				if (fwd_step(ix, st, s, c0, e0, sink)) return;
				st.c = c0 + 1;
				i_next = lower_bound_u32(ix.car, cb, ce, st.c);
			}
		}
	}
	if (st.c >= st.limit) return;
	// ---- forward: merge of the sample's entries and the markers from st.c on, in entry order, until the scan bound
	// A marker only matters when its arrival is >= y (it then ends the walk), and such a marker lies at most marker_span
	// entries below the scan bound — the markers of the long gap between two entries a sparse sample carries are skipped.
	const uint32_t m_lo = st.limit > ix.marker_span ? st.limit - ix.marker_span : 0;
	uint64_t im = lower_bound_u32(ix.marker_list, 0, ix.num_markers, st.c > m_lo ? st.c : m_lo);
	uint64_t ic = i_next;
	for (;;) {
		const uint32_t a = ic < ce ? ldg(ix.car + ic) : kNoneU32, b = im < ix.num_markers ? ldg(ix.marker_list + im) : kNoneU32;
		const uint32_t ci = a < b ? a : b;
		if (ci >= st.limit) return;                                               // (kNoneU32 >= any limit)
		if (a < b) ic++; else im++;
		const uint4 e = ldg(ix.cent + ci);
		if (fwd_step(ix, st, s, ci, e, sink)) return;
	}
}

// One region of t4 by whichever walk the index supports; *t6 (nullable) receives the region's t6 slice.
template <class Sink>
VSGPU_HD void walk_any(const DevIndex& ix, uint64_t x64, uint64_t y64, uint32_t s, Sink& sink, uint2* t6 = nullptr) {
	if (ix.car_begin) {
		uint32_t rk, e_y;
		walk_region_sparse(ix, x64, y64, s, sink, rk, e_y);
		if (t6) *t6 = t6_from_ranks(ix, x64, y64, rk, e_y);
		return;
	}
	if (ix.hitmap && ix.walk2) {
		FwdState2 st; uint32_t rk, e_y;
		if (fast2_setup(ix, x64, y64, s, sink, st, rk, e_y)) fast2_forward(ix, st, s, sink);
		if (t6) *t6 = t6_from_ranks(ix, x64, y64, rk, e_y);
		return;
	}
	if (ix.hitmap) walk_region_fast(ix, x64, y64, s, sink);
	else walk_region(ix, x64, y64, s, sink);
	if (t6) { bool bad = false; *t6 = t6_bounds(ix, x64, y64, &bad); }
}

// ------------------------------------------------------------------ t2: query_sample_from_ref (query.h:120-189)
// The reference walks the sample's path vertex by vertex from the vertex get_prev_vertex_with_sample
// returns, keeping (ref_pos, next_ref_pos) and cutting the vertex sequences with four rules
// (:157-173).  Here the path is taken stretch by stretch: between two walk entries the sample takes,
// it follows the backbone, whose sequences are contiguous in seq_buffer, so a whole stretch is at
// most three pieces (a clipped first vertex, whole vertices, a clipped last one) and the vertices
// where recording starts / ends come from `first_reach` instead of a per-vertex loop.

// get_prev_vertex_with_sample (query.h:57-113) for back-walk state `cur`: the walk entry it stops on,
// or kNoneU32 when it reaches the start of the contig.
VSGPU_HD uint32_t back_walk(const DevIndex& ix, uint32_t s, uint64_t cur) {
	if (ix.car_begin) {
		if (cur < 2 || cur > ix.D) return kNoneU32;
		const uint64_t info = ldg(ix.dinfo + (cur - 1));
		const uint32_t t = ldg(ix.dtin + cur);
		const uint32_t pos = (uint32_t)info + ((uint32_t)(info >> 32) & 0xFFFF);
		const uint64_t cb = ldg(ix.car_begin + s), ce = ldg(ix.car_begin + s + 1);
		for (uint64_t i = lower_bound_u32(ix.car, cb, ce, pos); i > cb;) {
			i--;
			const uint32_t p = ldg(ix.car + i);
			const uint2 a = ldg(ix.cent_anc + p);
			if (a.x <= t && t <= a.y) return p;
		}
		return kNoneU32;
	}
	if (ix.hitmap) {
		if (cur < 2 || cur > ix.D) return kNoneU32;
		const uint32_t* row = ix.hitmap + (uint64_t)s * ix.row_words;
		const uint64_t info = ldg(ix.dinfo + (cur - 1));
		const uint32_t t = ldg(ix.dtin + cur);
		const uint32_t pos = (uint32_t)info + ((uint32_t)(info >> 32) & 0xFFFF);     // entries below pos are candidates
		if (pos == 0) return kNoneU32;
		uint32_t w = (pos - 1) >> 5;
		uint32_t m = ldg(row + w) & (0xFFFFFFFFu >> (31 - ((pos - 1) & 31)));
		for (;;) {
			while (m == 0 && w > 0) { w--; m = ldg(row + w); }
			if (m == 0) return kNoneU32;
			const uint32_t b = 31 - clz32(m);
			m &= ~(1u << b);
			const uint32_t p = (w << 5) + b;
			const uint2 a = ldg(ix.cent_anc + p);
			if (a.x <= t && t <= a.y) return p;                                      // last carrier of the nearest examined vertex
		}
	}
	uint32_t c_found = kNoneU32;
	for (;;) {
		if (cur > ix.D) cur = 0;
		if (cur <= 1) break;
		const uint64_t info = ldg(ix.dinfo + (cur - 1));
		const uint32_t cb = (uint32_t)info, ncar = (uint32_t)(info >> 32) & 0xFFFF, deg = (uint32_t)(info >> 48);
		for (uint32_t c = cb; c < cb + ncar; c++) if (member(ix, s, ldg(&ix.cent[c].z))) c_found = c;
		cur -= deg;
		if (c_found != kNoneU32) break;
	}
	return c_found;
}

// first walk entry in [c, limit) whose target the sample carries (markers are not edges), or kNoneU32
VSGPU_HD uint32_t next_carried(const DevIndex& ix, uint32_t s, uint32_t c, uint32_t limit) {
	if (c >= limit) return kNoneU32;
	if (ix.car_begin) {
		const uint64_t cb = ldg(ix.car_begin + s), ce = ldg(ix.car_begin + s + 1);
		const uint64_t i = lower_bound_u32(ix.car, cb, ce, c);
		if (i >= ce) return kNoneU32;
		const uint32_t ci = ldg(ix.car + i);
		return ci < limit ? ci : kNoneU32;
	}
	if (ix.hitmap) {
		const uint32_t* row = ix.hitmap + (uint64_t)s * ix.row_words;
		uint32_t w = c >> 5;
		uint32_t m = ldg(row + w) & (0xFFFFFFFFu << (c & 31));
		while (m == 0) { w++; if ((w << 5) >= limit) return kNoneU32; m = ldg(row + w); }
		const uint32_t ci = (w << 5) + ctz32(m);
		return ci < limit ? ci : kNoneU32;
	}
	for (; c < limit; c++) { const uint4 e = ldg(ix.cent + c); if (!(e.y & kEntMarker) && member(ix, s, e.z)) return c; }
	return kNoneU32;
}

struct T2State { bool rec; uint64_t x, y; uint32_t e_x, e_y; };

// The four rules of query.h:157-173 on one vertex: `off`/`l` = its sequence in seq_buffer, rp = ref_pos
// on arrival, nrp = next_ref_pos.  0: go on, 1: the walk ends, 2: substr throws std::out_of_range.
template <class Sink>
VSGPU_HD int t2_apply(T2State& st, uint32_t off, uint32_t l, uint64_t rp, uint64_t nrp, Sink& sink) {
	if (st.rec) {
		if (nrp < st.y) { sink.seg(off, l); return 0; }                             // :157-159
		const uint64_t cnt = st.y - rp;                                            // :160-163 substr(0, pos_y - ref_pos)
		sink.seg(off, cnt < l ? (uint32_t)cnt : l);
		return 1;
	}
	if (nrp < st.x) return 0;
	const uint64_t start = st.x - rp;                                            // substr(pos_x - ref_pos ...): throws past the end
	if (start > l) return 2;
	const uint32_t avail = l - (uint32_t)start;
	if (nrp < st.y) { st.rec = true; sink.seg(off + (uint32_t)start, avail); return 0; }   // :164-167
	const uint64_t cnt = st.y - st.x;                                            // :168-172 substr(pos_x - ref_pos, pos_y - pos_x)
	sink.seg(off + (uint32_t)start, cnt < avail ? (uint32_t)cnt : avail);
	return 1;
}

// first backbone index j >= k with nrp1[j] >= v, where e = number of distinct starts < v; M: none.
// first_reach answers it unless the vertex it names was bypassed by a detour (then the answer lies a
// few vertices ahead: nrp1[j] >= start of P[j+1]).
VSGPU_HD uint32_t t2_first_ge(const DevIndex& ix, const T2Tables& t2, uint32_t k, uint32_t e, uint64_t v) {
	const uint32_t fr = ldg(t2.first_reach + e);
	if (fr >= k && e < ix.D) return fr;
	for (uint32_t j = fr > k ? fr : k; j < ix.M; j++) if ((uint64_t)ldg(t2.nrp1 + j) >= v) return j;
	return ix.M;
}

// backbone vertices [k_a, k_b] visited in order, arriving at P[k_a] with ref_pos
template <class Sink>
VSGPU_HD int t2_stretch(const DevIndex& ix, const T2Tables& t2, T2State& st, uint32_t k_a, uint32_t k_b, uint64_t ref_pos, Sink& sink) {
	uint32_t k = k_a;
	if (!st.rec) {
		const uint32_t j = t2_first_ge(ix, t2, k_a, st.e_x, st.x);
		if (j > k_b) return 0;
		const uint64_t rp = j == k_a ? ref_pos : (uint64_t)ldg(t2.nrp1 + j - 1);
		const uint32_t b0 = ldg(t2.bbs + j), b1 = ldg(t2.bbs + j + 1);
		const int r = t2_apply(st, b0 - 1, b1 - b0, rp, ldg(t2.nrp1 + j), sink);
		if (r) return r;
		k = j + 1;
		if (k > k_b) return 0;
	}
	const uint32_t j = t2_first_ge(ix, t2, k, st.e_y, st.y);
	const uint32_t b0 = ldg(t2.bbs + k);
	if (j > k_b) { sink.seg(b0 - 1, ldg(t2.bbs + k_b + 1) - b0); return 0; }        // whole vertices, one piece
	const uint32_t bj = ldg(t2.bbs + j), bj1 = ldg(t2.bbs + j + 1);
	if (j > k) sink.seg(b0 - 1, bj - b0);
	const uint64_t rp = j == k_a ? ref_pos : (uint64_t)ldg(t2.nrp1 + j - 1);
	return t2_apply(st, bj - 1, bj1 - bj, rp, ldg(t2.nrp1 + j), sink);
}

// returns 0, or kT2Throw when the reference call ends in std::out_of_range (nothing is emitted then)
template <class Sink>
VSGPU_HD uint32_t t2_walk(const DevIndex& ix, const T2Tables& t2, uint64_t x64, uint64_t y64, uint32_t s, Sink& sink) {
	const uint32_t xc = clamp_pos(x64), yc = clamp_pos(y64);
	uint32_t rk, e_y;                                                           // rank(x); number of starts < y
	rank_le2(ix, xc, yc ? yc - 1 : 0, rk, e_y);
	if (!yc) e_y = 0;
	uint32_t e_x = rk;                                                          // number of starts < x
	if (rk >= 1 && ldg(ix.dstart + rk - 1) == xc) e_x = rk - 1;
	T2State st{false, x64, y64, e_x, e_y};
	// entries whose source starts at or after max(x, y) are never reached: the vertex before ends the walk
	const uint4 dl = ldg(ix.dlev + (e_x > e_y ? e_x : e_y));
	const uint32_t limit = dl.w, k_last = dl.x;
	// ---- get_prev_vertex_with_sample (query.h:57-113); Index::find(pos, rank) index.h:135-148
	const uint64_t cur = x64 >= ix.index_bits ? ix.D - 1 : (rk ? rk - 1 : 0);
	const uint32_t c_found = back_walk(ix, s, cur);
	uint32_t cur_k, c;
	uint64_t ref_pos;
	if (c_found != kNoneU32) {
		const uint4 e = ldg(ix.cent + c_found);
		ref_pos = e.w;                                                            // index of the last ref-carrying neighbour seen (:93-95)
		if (e.y & kEntAlt) {
			const uint2 sq = ldg(t2.cent_seq + c_found);
			const uint32_t tk = e.y & kEntTgtMask;
			const uint64_t nrp = tk == kEntTgtMask ? ref_pos + sq.y : (uint64_t)ldg(t2.bbs + tk);
			const int r = t2_apply(st, sq.x, sq.y, ref_pos, nrp, sink);
			if (r) return r == 2 ? kT2Throw : 0;
			if (tk == kEntTgtMask) return 0;                                        // the path ends on this vertex
			cur_k = tk; ref_pos = nrp;
		} else cur_k = e.y & kEntTgtMask;
		c = c_found + 1;
	} else { cur_k = ldg(&ix.dlev[0].x); ref_pos = 1; c = 0; }
	// ---- the path: backbone stretches between the walk entries the sample takes
	for (;;) {
		const uint32_t ci = next_carried(ix, s, c, limit);
		if (ci == kNoneU32) break;
		c = ci + 1;
		const uint4 e = ldg(ix.cent + ci);
		if (e.x < cur_k) continue;                                                // hidden behind a taken detour / an earlier sibling
		int r = t2_stretch(ix, t2, st, cur_k, e.x, ref_pos, sink);
		if (r) return r == 2 ? kT2Throw : 0;
		ref_pos = ldg(t2.nrp1 + e.x);
		if (e.y & kEntAlt) {
			const uint2 sq = ldg(t2.cent_seq + ci);
			const uint32_t tk = e.y & kEntTgtMask;
			const uint64_t nrp = tk == kEntTgtMask ? ref_pos + sq.y : (uint64_t)ldg(t2.bbs + tk);
			r = t2_apply(st, sq.x, sq.y, ref_pos, nrp, sink);
			if (r) return r == 2 ? kT2Throw : 0;
			if (tk == kEntTgtMask) return 0;
			cur_k = tk; ref_pos = nrp;
		} else cur_k = e.y & kEntTgtMask;
	}
	uint32_t k_b = k_last ? k_last - 1 : 0;
	if (k_b < cur_k) k_b = cur_k;
	if (k_b > ix.M - 1) k_b = ix.M - 1;
	if (cur_k > k_b) return 0;
	const int r = t2_stretch(ix, t2, st, cur_k, k_b, ref_pos, sink);
	return r == 2 ? kT2Throw : 0;
}

// ------------------------------------------------------------------ t3: query_sample_from_sample (query.h:195-261)
// Same answer shape as t2, but the cutting rules run on the sample's own coordinate: it starts from
// the sample's `index` in the vertex get_prev_vertex_with_sample returns and grows by the length of
// every vertex on the path, so along a backbone stretch it is monotone and the vertices where
// recording starts / ends come from one rank each.

// sample_info.index of sample s in the target vertex of walk entry c (s is a carrier of it)
VSGPU_HD uint32_t sample_index_at(const DevIndex& ix, const T3Tables& t3, uint32_t c, uint32_t s) {
	const uint64_t b = ldg(t3.sidx_begin + c);
	if (ix.class_mode) {                                   // s_info[i] belongs to the i-th set bit of the class (variant_graph.h:1302-1315)
		const uint64_t* row = ix.bitmap + (uint64_t)ldg(&ix.cent[c].z) * ix.words_per_set;
		uint32_t r = 0;
		for (uint32_t w = 0; w < (s >> 6); w++) r += (uint32_t)
#if defined(__CUDA_ARCH__)
			__popcll(ldg(row + w));
#else
			__builtin_popcountll(row[w]);
#endif
		const uint64_t last = ldg(row + (s >> 6)) & (((uint64_t)1 << (s & 63)) - 1);
#if defined(__CUDA_ARCH__)
		r += (uint32_t)__popcll(last);
#else
		r += (uint32_t)__builtin_popcountll(last);
#endif
		return ldg(t3.sidx + b + r);
	}
	for (uint64_t i = b, e = ldg(t3.sidx_begin + c + 1); i < e; i++) if (ldg(t3.sid + i) == s) return ldg(t3.sidx + i);
	return 0;
}

struct PrevHit { uint32_t c; uint64_t ref_pos, sample_pos; };
// get_prev_vertex_with_sample (query.h:57-113) with all three outputs
VSGPU_HD PrevHit prev_with_sample(const DevIndex& ix, const T3Tables& t3, uint64_t pos, uint32_t s) {
	const uint32_t rk = rank_le(ix, clamp_pos(pos));
	const uint64_t cur = pos >= ix.index_bits ? ix.D - 1 : (rk ? rk - 1 : 0);
	PrevHit h;
	h.c = back_walk(ix, s, cur);
	if (h.c == kNoneU32) { h.ref_pos = 1; h.sample_pos = t3.first_index; }
	else { h.ref_pos = ldg(&ix.cent[h.c].w); h.sample_pos = sample_index_at(ix, t3, h.c, s); }
	return h;
}

// first backbone index k >= k_a with sp + (start of P[k+1] - start of P[k_a]) >= v; M: none
VSGPU_HD uint32_t t3_first_ge(const DevIndex& ix, const T2Tables& t2, uint32_t k_a, uint64_t sp, uint64_t v) {
	const uint64_t base = ldg(t2.bbs + k_a);
	if (v <= sp || v - sp + base <= (uint64_t)ldg(t2.bbs + k_a + 1)) return k_a;
	const uint64_t V = v - sp + base;                      // first k with start of P[k+1] >= V
	if (V > (uint64_t)ix.last_end) return ix.M;
	const uint32_t e = rank_le(ix, (uint32_t)V - 1);       // number of distinct starts < V
	if (e >= ix.D) return ix.M - 1;                        // only the end of the last vertex reaches V
	const uint32_t j = ldg(&ix.dlev[e].x);                 // first backbone vertex starting at or after V
	return j ? j - 1 : 0;
}

template <class Sink>
VSGPU_HD int t3_stretch(const DevIndex& ix, const T2Tables& t2, T2State& st, uint32_t k_a, uint32_t k_b, uint64_t sp, Sink& sink) {
	const uint32_t base = ldg(t2.bbs + k_a);
	uint32_t k = k_a;
	if (!st.rec) {
		uint32_t j = t3_first_ge(ix, t2, k_a, sp, st.x);
		if (j < k_a) j = k_a;
		if (j > k_b) return 0;
		const uint32_t b0 = ldg(t2.bbs + j), b1 = ldg(t2.bbs + j + 1);
		const uint64_t spj = sp + (b0 - base);
		const int r = t2_apply(st, b0 - 1, b1 - b0, spj, spj + (b1 - b0), sink);
		if (r) return r;
		k = j + 1;
		if (k > k_b) return 0;
	}
	uint32_t j = t3_first_ge(ix, t2, k_a, sp, st.y);
	if (j < k) j = k;
	const uint32_t b0 = ldg(t2.bbs + k);
	if (j > k_b) { sink.seg(b0 - 1, ldg(t2.bbs + k_b + 1) - b0); return 0; }
	const uint32_t bj = ldg(t2.bbs + j), bj1 = ldg(t2.bbs + j + 1);
	if (j > k) sink.seg(b0 - 1, bj - b0);
	const uint64_t spj = sp + (bj - base);
	return t2_apply(st, bj - 1, bj1 - bj, spj, spj + (bj1 - bj), sink);
}

// returns 0, kT2Throw (substr throws) or kT3Hang (the reference never leaves the loop at :209-214)
template <class Sink>
VSGPU_HD uint32_t t3_walk(const DevIndex& ix, const T2Tables& t2, const T3Tables& t3, uint64_t x64, uint64_t y64, uint32_t s, Sink& sink) {
	PrevHit h = prev_with_sample(ix, t3, x64, s);
	// :209-214: step back from the ref position of the vertex found while its sample position is >= x.
	// The chain of positions is deterministic; Brent's test finds the cycle the reference would spin in.
	uint64_t saved = ~(uint64_t)0;
	for (uint32_t steps = 0, power = 1; h.sample_pos >= x64 && h.c != kNoneU32;) {
		const uint64_t pos = h.ref_pos;
		if (pos == saved) return kT3Hang;
		if (++steps == power) { saved = pos; power <<= 1; steps = 0; }
		h = prev_with_sample(ix, t3, pos, s);
	}
	T2State st{false, x64, y64, 0, 0};
	const uint64_t m = x64 > y64 ? x64 : y64;
	uint32_t cur_k, c;
	uint64_t sp = h.sample_pos;
	if (h.c != kNoneU32) {
		const uint4 e = ldg(ix.cent + h.c);
		if (e.y & kEntAlt) {
			const uint2 sq = ldg(t2.cent_seq + h.c);
			const uint32_t tk = e.y & kEntTgtMask;
			const int r = t2_apply(st, sq.x, sq.y, sp, sp + sq.y, sink);
			if (r) return r == 2 ? kT2Throw : 0;
			if (tk == kEntTgtMask) return 0;
			cur_k = tk; sp += sq.y;
		} else cur_k = e.y & kEntTgtMask;
		c = h.c + 1;
	} else { cur_k = ldg(&ix.dlev[0].x); c = 0; }
	for (;;) {
		// without another entry taken, the backbone from here ends the walk at k_end
		const uint32_t k_end = t3_first_ge(ix, t2, cur_k, sp, m);
		const uint32_t limit = k_end < ix.M ? ldg(ix.cent_begin_k + k_end + 1) : ix.num_cent;
		uint32_t ci;
		uint4 e = make_uint4(0, 0, 0, 0);
		for (;;) {                                             // next entry the sample takes that is not hidden behind cur_k
			ci = next_carried(ix, s, c, limit);
			if (ci == kNoneU32) break;
			c = ci + 1;
			e = ldg(ix.cent + ci);
			if (e.x >= cur_k) break;
		}
		if (ci == kNoneU32) {
			const uint32_t k_b = k_end < ix.M ? k_end : ix.M - 1;
			const int r = t3_stretch(ix, t2, st, cur_k, k_b, sp, sink);
			return r == 2 ? kT2Throw : 0;
		}
		int r = t3_stretch(ix, t2, st, cur_k, e.x, sp, sink);
		if (r) return r == 2 ? kT2Throw : 0;
		sp += ldg(t2.bbs + e.x + 1) - ldg(t2.bbs + cur_k);
		if (e.y & kEntAlt) {
			const uint2 sq = ldg(t2.cent_seq + ci);
			const uint32_t tk = e.y & kEntTgtMask;
			r = t2_apply(st, sq.x, sq.y, sp, sp + sq.y, sink);
			if (r) return r == 2 ? kT2Throw : 0;
			if (tk == kEntTgtMask) return 0;
			cur_k = tk; sp += sq.y;
		} else cur_k = e.y & kEntTgtMask;
	}
}

// ------------------------------------------------------------------ t5: get_sample_var_in_sample (query.h:490-612)
// Same start as t3 (incl. the loop that may never end), then the walk of t4 — but from the backbone
// vertex holding ref_pos (:513), gated on the sample's own coordinate: a vertex carrying the sample is
// reported when x < sample_pos < y on arrival (:531, :553; sample_pos only grows along the path).
// Emits the same hit codes as t4; the row rules differ only in var_pos (host_index.h: t5_row).
template <class Sink>
VSGPU_HD uint32_t t5_walk(const DevIndex& ix, const T2Tables& t2, const T3Tables& t3, uint64_t x64, uint64_t y64, uint32_t s, Sink& sink) {
	PrevHit h = prev_with_sample(ix, t3, x64, s);
	uint64_t saved = ~(uint64_t)0;
	for (uint32_t steps = 0, power = 1; h.sample_pos >= x64 && h.c != kNoneU32;) {
		const uint64_t pos = h.ref_pos;
		if (pos == saved) return kT3Hang;
		if (++steps == power) { saved = pos; power <<= 1; steps = 0; }
		h = prev_with_sample(ix, t3, pos, s);
	}
	// closest_v = idx->find(ref_pos) (:513, index.h:119-133); seq_len = ref_pos - its start (:517-520)
	uint32_t rk = h.ref_pos >= ix.index_bits ? ix.D : rank_le(ix, clamp_pos(h.ref_pos));
	if (rk < 1) rk = 1;
	uint32_t cur_k = ldg(&ix.dlev[rk - 1].x);
	uint64_t sp = h.sample_pos - (h.ref_pos - (uint64_t)ldg(t2.bbs + cur_k));
	uint32_t c = ldg(ix.cent_begin_k + cur_k);
	for (;;) {
		if (sp >= y64) return 0;                                 // :531
		// the first backbone vertex from here at which sample_pos has reached y: entries from it on are never taken
		const uint32_t k_end = t3_first_ge(ix, t2, cur_k, sp, y64);
		const uint32_t limit = k_end < ix.M ? ldg(ix.cent_begin_k + k_end + 1) : ix.num_cent;
		uint32_t ci;
		uint4 e = make_uint4(0, 0, 0, 0);
		for (;;) {
			ci = next_carried(ix, s, c, limit);
			if (ci == kNoneU32) return 0;
			c = ci + 1;
			e = ldg(ix.cent + ci);
			if (e.x >= cur_k) break;                               // else hidden behind a taken detour / an earlier sibling
		}
		sp += ldg(t2.bbs + e.x + 1) - ldg(t2.bbs + cur_k);       // sample_pos on arrival at the entry's target
		if (sp >= y64) return 0;
		if (sp > x64) sink.emit(ci);
		if (e.y & kEntAlt) {
			const uint32_t tk = e.y & kEntTgtMask;
			if (tk == kEntTgtMask) return 0;
			sp += ldg(&t2.cent_seq[ci].y);
			if ((e.y & kEntTgtCarriers) && member(ix, s, ldg(ix.bb_set + tk)) && sp > x64 && sp < y64) sink.emit(ci | kHitRejoin);
			cur_k = tk;
		} else cur_k = e.y & kEntTgtMask;
	}
}

// Pieces of one region's answer, merged while they are contiguous in seq_buffer: one copy record
// {src, len, dst lo, dst hi} per maximal run.  Count: how many records / bytes.  Write: the records
// themselves plus, for every kT2Tile-byte boundary of the output a record covers, its index in
// tile_first (the copy kernel starts each tile from there; every boundary below the total is covered
// by exactly one record because the records tile the output without gaps).
struct T2CountSink {
	uint32_t off, len; uint32_t nrec; uint64_t bytes; uint2* keep; uint64_t stride;   // keep: nullable
	VSGPU_HD void flush() { if (len) { if (keep && nrec < kT2Keep) keep[nrec * stride] = make_uint2(off, len); nrec++; bytes += len; len = 0; } }
	VSGPU_HD void seg(uint32_t o, uint32_t l) { if (!l) return; if (len && off + len == o) { len += l; return; } flush(); off = o; len = l; }
};
struct T2WriteSink {
	uint32_t off, len; uint4* out; uint64_t dst; const uint4* base; uint32_t* tile_first;
	VSGPU_HD void flush() {
		if (!len) return;
		*out = make_uint4(off, len, (uint32_t)dst, (uint32_t)(dst >> 32));
		const uint32_t r = (uint32_t)(out - base);
		for (uint64_t b = (dst + kT2Tile - 1) / kT2Tile, e = (dst + len - 1) / kT2Tile; b <= e; b++) tile_first[b] = r;
		out++; dst += len; len = 0;
	}
	VSGPU_HD void seg(uint32_t o, uint32_t l) { if (!l) return; if (len && off + len == o) { len += l; return; } flush(); off = o; len = l; }
};

}  // namespace logic
}  // namespace vsgpu
