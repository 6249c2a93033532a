// libvsgpu host side — flattening of the loaded variation graph + position index into the
// structure-of-arrays the kernels scan.  One pass, done at vsgpu_open().
//
// Every order- or quirk-dependent decision of the reference's operators is evaluated here once
// per backbone vertex with the reference's own rules, so that the kernels only do searches,
// membership tests and an ordered chain resolution:
//   * backbone = the "ref" path: repeated VariantGraph::get_neighbor_vertex(v, 0) from vertex 0
//     (include/variant_graph.h:1402-1451, :2025-2032)
//   * per backbone vertex, the branch records next_variant_in_ref builds (include/query.h:316-415)
//   * per backbone vertex, what get_sample_var_in_ref needs of its out-neighbours
//     (include/query.h:660-674: next_ref_pos = index of the LAST ref-carrying neighbour in
//     unordered_set iteration order; :677-710 emission; get_neighbor_vertex: FIRST carrier wins)
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "ser_reader.h"

namespace vsgpu {

constexpr uint32_t kNone = 0xFFFFFFFFu;

// bits 29..31 of CEntry::tgt
constexpr uint32_t kEntAlt = 1u << 31;          // target is an alt vertex (else: a backbone vertex that carries samples)
constexpr uint32_t kEntTgtCarriers = 1u << 30;  // alt only: the backbone vertex it rejoins carries samples itself
constexpr uint32_t kEntMarker = 1u << 29;       // not an edge: "arrival at src+1 along the backbone has ref_pos = arrival"
constexpr uint32_t kEntTgtMask = (1u << 29) - 1;
constexpr uint32_t kEntTgtNone = kEntTgtMask;   // alt vertex without an out-edge

struct CEntry {            // 16 bytes, scanned with one 128-bit load per lane
	uint32_t src;            // backbone index k of the source vertex
	uint32_t tgt;            // flags | backbone index the walk continues at (alt: its rejoin vertex)
	uint32_t set_id;         // carrier-set id of the target (class id, or list id in explicit-id mode)
	uint32_t arrival;        // ref_pos on arrival at the target = next_ref_pos computed at P[src]
};

struct DLevel {            // per distinct backbone start (one per set bit of index.sdsl)
	uint32_t k;              // backbone index of node_list[d]
	uint32_t rec_lo;         // rec_begin[k]
	uint32_t rec_hi_prev;    // rec_begin[k-1] (0 for k == 0): t6 upper bound when y lands on this start
	uint32_t cent_begin;     // first compact entry with src >= k
};

struct FlatIndex {
	// sizes
	uint64_t ref_length = 0, index_bits = 0;
	uint32_t num_samples = 0, M = 0, D = 0, R = 0, num_sets = 0, words_per_set = 0;
	bool class_mode = true;
	bool has_suspect_dups = false;

	// backbone (path order)
	std::vector<uint32_t> bb_vertex, vstart, vlen, rec_begin /*M+1*/, cent_begin /*M+1*/, bb_set /*carrier set of P[k] or 0*/;
	std::vector<uint32_t> vertex_bb;      // vertex id -> backbone index or kNone
	std::vector<uint32_t> bb_nrp, bb_nref;  // host: next_ref_pos computed at P[k] and the neighbour vertex that set it (kNone: none)

	// distinct-start level
	std::vector<uint32_t> dstart;         // D, ascending
	std::vector<DLevel> dlev;             // D + 1 (sentinel k = M)
	std::vector<uint64_t> dinfo;          // D: cent_begin(32) | carrier entries(16) | out-degree(16) of P[dlev[d].k]
	std::vector<uint32_t> t7_lo, t7_hi;   // D: record range of the first backbone vertex >= dlev[d].k that has records

	// compact t4 entries
	std::vector<CEntry> cent;
	std::vector<uint32_t> cent_vertex;    // target vertex id (kNone for markers)
	uint32_t row_words = 0;               // hit-map row length in 32-bit words (multiple of 32)
	std::vector<uint32_t> marker_bits;    // row_words: bit c = walk entry c is a marker
	// get_prev_vertex_with_sample steps back through node_list by out-degree (query.h:103); as a forest
	// parent(c) = c - outdeg(node_list[c-1]) that walk visits exactly the ancestors of its start.
	std::vector<uint32_t> dtin;           // D + 1: Euler-tour entry time of back-walk state c
	std::vector<uint32_t> cent_anc;       // 2 per walk entry: [tin, last] of the state that examines its source vertex (1,0: never examined)

	// t2 (query_sample_from_ref, query.h:120-189): next_ref_pos computed at P[k] from the FIRST ref-carrying
	// neighbour (:143-151; t4 takes the last), the sequence of every walk entry's alt target, and per
	// distinct start d the first backbone index j with nrp1[j] >= dstart[d] (D + 1 entries; the last =
	// first j with nrp1[j] > dstart[D-1]).  t2_ok: backbone sequences are contiguous in seq_buffer
	// (offset = start - 1) and nrp1[k] >= vstart[k+1] everywhere; otherwise t2 is refused (t2_why).
	std::vector<uint32_t> nrp1;           // M
	std::vector<uint32_t> first_reach;    // D + 1
	std::vector<uint32_t> cent_seq;       // 2 per walk entry: {seq offset, length} of an alt target (0,0 otherwise)
	bool t2_ok = true;
	std::string t2_why;

	// t6/t7 branch records, (backbone index, out-order) order
	std::vector<uint32_t> rec_k, rec_vertex, rec_pos, rec_refv, rec_altv;   // *_v: vertex whose sequence is the string, kNone = ""
	std::vector<uint8_t> rec_flags;       // bit0: kept by a fresh next_variant_in_ref call (t7); bit1: suspect duplicate (t6)
	std::vector<uint64_t> rec_hash;       // hash of (ref, alt) strings for the t7 compare
	std::vector<uint32_t> rec_dup_prefix; // R+1 prefix count of suspect duplicates

	// carrier sets: class bitmaps (class mode) or sorted id lists (explicit-id mode)
	std::vector<uint64_t> bitmap;         // (num_sets) x words_per_set, row 0 = {ref}
	std::vector<uint64_t> list_begin;     // explicit-id mode: num_sets + 1
	std::vector<uint32_t> list_ids;

	bool member(uint32_t sample, uint32_t set_id) const {
		if (class_mode) return (bitmap[(uint64_t)set_id * words_per_set + (sample >> 6)] >> (sample & 63)) & 1;
		for (uint64_t i = list_begin[set_id]; i < list_begin[set_id + 1]; i++) if (list_ids[i] == sample) return true;
		return false;
	}
};

uint64_t hash_ref_alt(const char* ref, size_t nref, const char* alt, size_t nalt);

// Throws std::runtime_error if the graph has a shape the flattened form cannot represent
// (documented in DESIGN.md: non-backbone vertices with more than one out-edge, backbone gaps).
void flatten(const SerData& d, FlatIndex& f);

}  // namespace vsgpu
