// libvsgpu host side — the loaded + flattened index as the host sees it, and the materialisation
// of result rows (struct Variant of include/query.h:30-36 as print_var text, query.h:43-50) from
// record ids / walk-entry hit codes.  No CUDA in here.
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

#include "flatten.h"
#include "ser_reader.h"

namespace vsgpu {

struct HostIndex {
	SerData ser;
	FlatIndex flat;
	uint32_t last_end = 0;                                  // start + length of the last backbone vertex
	uint32_t t1_fallback_pos = 0;                           // see DevIndex::t1_fallback_pos
	std::unordered_map<std::string, uint32_t> name2id;
	std::string prefix;                                     // the ser/ directory (operators that need a second pass over it: t3)
	std::vector<uint32_t> sindex;                           // sample_info.index per s_info (second pass over ser/, on the first t3 / t5 call)
	bool from_cache = false;                                // read from VSGPU_INDEX_CACHE instead of decoding ser/
};

// index_cache.cc: opt-in on-disk cache of everything build_host_index computes
bool load_index_cache(const std::string& prefix, HostIndex& h);
void save_index_cache(const std::string& prefix, HostIndex& h);

// load_ser + flatten + derived fields; throws std::runtime_error
void build_host_index(const std::string& prefix, HostIndex& h, int* stage = nullptr);
// direct-mapped rank buckets over positions: bucket[b] = number of distinct starts < (b << shift)
void build_buckets(const FlatIndex& f, std::vector<uint32_t>& bucket, uint32_t& shift);
// DevIndex::d4 (kernels.cuh), 4 words per distinct start: what the t4 walk looks up per start, in one row
void build_d4(const FlatIndex& f, std::vector<uint32_t>& d4);

// per-sample carried walk entries + marker entries (DevIndex::car_begin / car / marker_list) and when to use them
void build_sparse_walk(const FlatIndex& f, std::vector<uint64_t>& car_begin, std::vector<uint32_t>& car, std::vector<uint32_t>& marker_list);
bool want_sparse_walk(const FlatIndex& f);
uint32_t marker_span(const FlatIndex& f, const std::vector<uint32_t>& marker_list);
// per sample: the entries its walk from the head of the contig takes, with the running maximum of their arrivals (DevIndex::can_*)
void build_canonical_walks(const FlatIndex& f, const std::vector<uint64_t>& car_begin, const std::vector<uint32_t>& car, const std::vector<uint32_t>& marker_list,
                           std::vector<uint64_t>& can_begin, std::vector<uint32_t>& can_entry, std::vector<uint32_t>& can_pmax);

void append_seq(const HostIndex* ix, uint32_t v, std::string& out);
void append_carriers(const HostIndex* ix, uint32_t v, std::string& out);
void t6_row(const HostIndex* ix, uint32_t r, bool with_samples, std::string& out);
void t4_row(const HostIndex* ix, uint32_t code, bool with_samples, std::string& out);
// the parts of that row as vertex ids (what the device-side renderer's tables are built from); sample == kNone (0xFFFFFFFF): t4
void hit_row_parts(const HostIndex* ix, uint32_t code, uint32_t sample, uint64_t& pos, uint32_t& refv, uint32_t& altv, uint32_t& u, bool* insertion = nullptr);
// t5 row (get_sample_var_in_sample, query.h:553-590): the t4 row of the same hit code with var_pos in the sample's coordinates
void t5_row(const HostIndex* ix, uint32_t code, uint32_t sample, bool with_samples, std::string& out);
void digests_t5(const HostIndex* ix, uint64_t n, const uint64_t* offsets, const uint32_t* hits, const uint32_t* samples, bool with_samples, uint64_t* digests);
uint32_t host_sample_index(const HostIndex* ix, uint32_t vertex, uint32_t sample);
bool push_rule(const HostIndex* ix, const std::vector<uint32_t>& vars, uint32_t r);
uint32_t host_rank(const FlatIndex& f, uint64_t pos);
void t6_literal(const HostIndex* ix, uint64_t x, uint64_t y, std::vector<uint32_t>& vars);
bool t6_needs_literal(const HostIndex* ix, uint64_t y, uint32_t lo, uint32_t hi);
uint64_t hash_query(const char* ref, const char* alt);
uint32_t t7_confirm(const HostIndex* ix, uint64_t pos, const char* ref, const char* alt, uint32_t r);
uint64_t t7_carriers(const HostIndex* ix, uint32_t rec, std::string* text, uint64_t* digest);

uint64_t fnv1a(uint64_t h, const void* data, size_t n);
constexpr uint64_t kFnvInit = 14695981039346656037ULL;

// rows / digests shared by the C ABI and the test-only host simulator
// rows of a closest_var answer: the records of [lo, hi) a fresh next_variant_in_ref call keeps
void rows_t1(const HostIndex* ix, uint32_t lo, uint32_t hi, bool with_samples, std::string& text, uint64_t& nrows);
void digests_t1(const HostIndex* ix, uint64_t n, const uint32_t* lo, const uint32_t* hi, bool with_samples, uint64_t* counts, uint64_t* digests);
void rows_t6(const HostIndex* ix, uint32_t lo, uint32_t hi, bool with_samples, std::string& text, uint64_t& nrows);
void digests_t6(const HostIndex* ix, uint64_t n, const uint32_t* lo, const uint32_t* hi, bool with_samples, uint64_t* digests, bool* bad);
void digests_t4(const HostIndex* ix, uint64_t n, const uint64_t* offsets, const uint32_t* hits, bool with_samples, uint64_t* digests);
void digests_t7(const HostIndex* ix, uint64_t n, const uint32_t* rec, uint64_t* ncarriers, uint64_t* digests);

template <class F> void parallel_for(uint64_t n, F&& fn);

}  // namespace vsgpu

#include <algorithm>
#include <thread>
namespace vsgpu {
template <class F>
void parallel_for(uint64_t n, F&& fn) {
	unsigned nt = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), (unsigned)((n + 4095) / 4096)));
	if (nt <= 1) { fn((uint64_t)0, n); return; }
	std::vector<std::thread> th;
	uint64_t chunk = (n + nt - 1) / nt;
	for (unsigned t = 0; t < nt; t++) { uint64_t a = t * chunk, b = std::min<uint64_t>(n, a + chunk); if (a < b) th.emplace_back([=, &fn]() { fn(a, b); }); }
	for (auto& t : th) t.join();
}
}  // namespace vsgpu
