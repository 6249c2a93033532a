// ORACLE — test infrastructure only (see vso.h).  C entry points for tests/ and bench.py
// (ctypes): construct from FASTA+VCF, synthetic construct, load, and the three operators with
// text / digest outputs that the engine's results are compared against.
#include "vso.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <sys/stat.h>
#include <thread>

using namespace vso;

namespace {
struct Handle {
	std::unique_ptr<VariantGraph> vg;
	std::unique_ptr<Index> idx;
};
thread_local std::string g_err;
char* dup_str(const std::string& s) { char* p = (char*)malloc(s.size() + 1); memcpy(p, s.data(), s.size()); p[s.size()] = 0; return p; }

inline uint64_t fnv1a(uint64_t h, const void* data, size_t n) {
	const unsigned char* p = (const unsigned char*)data;
	for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ULL; }
	return h;
}
constexpr uint64_t kFnvInit = 14695981039346656037ULL;

// one output row as print_var (query.h:43-50) writes it
void row_text(const Variant& v, bool with_samples, std::string& out) {
	out += std::to_string(v.var_pos); out += '\t'; out += v.ref; out += '\t'; out += v.alt; out += '\t';
	if (with_samples) for (const auto& s : v.samples) { out += s.first; out += '('; out += s.second; out += ") "; }
	out += '\n';
}
uint64_t rows_digest(const std::vector<Variant>& vars, bool with_samples) {
	uint64_t h = kFnvInit; std::string t;
	for (const auto& v : vars) { t.clear(); row_text(v, with_samples, t); h = fnv1a(h, t.data(), t.size()); }
	return h;
}

struct Rng {   // splitmix64
	uint64_t s;
	explicit Rng(uint64_t seed) : s(seed) {}
	uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ULL); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }
	uint64_t below(uint64_t n) { return (uint64_t)(((unsigned __int128)next() * n) >> 64); }
	double unit() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
};
}  // namespace

extern "C" {

const char* vso_last_error() { return g_err.c_str(); }
void vso_free(void* p) { free(p); }

// variantstore construct (commands.cc:32-60): build, serialize graph, build + serialize index.
void* vso_construct(const char* fasta, const char* vcf, const char* prefix, int cqf_log2, int use_ref_gqf,
                    int fix_idx, int force_enc) {
	try {
		ConstructOpts o; o.cqf_log2_slots = cqf_log2 > 0 ? cqf_log2 : 25; o.use_ref_gqf = use_ref_gqf != 0;
		o.fix_sample_indexes = fix_idx != 0; o.force_encoding = force_enc;
		mkdir(prefix, 0755);
		auto h = new Handle;
		h->vg.reset(new VariantGraph(fasta, vcf, prefix, o));
		h->vg->serialize();
		h->idx.reset(new Index(h->vg.get()));
		h->idx->serialize(prefix);
		return h;
	} catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}

// variantstore query: Index idx(prefix); VariantGraph vg(prefix, mode)  (commands.cc:116-132)
void* vso_open(const char* prefix, int use_ref_gqf) {
	try {
		auto h = new Handle;
		h->idx.reset(new Index(std::string(prefix)));
		h->vg.reset(new VariantGraph(std::string(prefix), use_ref_gqf != 0));
		return h;
	} catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}

void vso_close(void* hp) { delete (Handle*)hp; }

// out[0..10]: cqf distinct keys (#Vertices), #Edges, seq length, ref length, num_samples, classes,
// total vertices, index ones, num_vars, num_mutations, num_mutations_samples, use_bit_vector
int vso_info(void* hp, uint64_t* out) {
	Handle* h = (Handle*)hp;
	out[0] = h->vg->get_num_vertices(); out[1] = h->vg->get_num_edges(); out[2] = h->vg->get_seq_length();
	out[3] = h->vg->get_ref_length(); out[4] = h->vg->num_samples; out[5] = h->vg->get_num_sample_classes();
	out[6] = h->vg->vertices.size(); out[7] = h->idx->ones.size(); out[8] = h->vg->num_vars;
	out[9] = h->vg->num_mutations; out[10] = h->vg->num_mutations_samples; out[11] = h->vg->use_bit_vector ? 1 : 0;
	return 0;
}

char* vso_sample_name(void* hp, uint32_t id) {
	Handle* h = (Handle*)hp;
	try { return dup_str(h->vg->get_sample_name(id)); } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}

// text = the count line the operator prints + header + rows (what -v writes to the -o file)
char* vso_query_t6_text(void* hp, uint64_t x, uint64_t y) {
	Handle* h = (Handle*)hp;
	try {
		QueryLog log;
		auto vars = get_var_in_ref(h->vg.get(), h->idx.get(), x, y, false, "", &log);
		std::string t = log.out + "Pos\tRef\tAlt\tSamples\n";
		for (auto& v : vars) row_text(v, true, t);
		return dup_str(t);
	} catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
char* vso_query_t4_text(void* hp, uint64_t x, uint64_t y, const char* sample, int* ub) {
	Handle* h = (Handle*)hp;
	try {
		QueryLog log; bool u = false;
		auto vars = get_sample_var_in_ref(h->vg.get(), h->idx.get(), x, y, sample, false, "", &log, &u);
		if (ub) *ub = u;
		std::string t = log.out + "Pos\tRef\tAlt\tSamples\n";
		for (auto& v : vars) row_text(v, true, t);
		return dup_str(t);
	} catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
// "name phasing" pairs concatenated exactly as the -o file of samples_has_var (query.h:807-816),
// or "There is no such variant!" when nothing matches.
char* vso_query_t7_text(void* hp, uint64_t pos, const char* ref, const char* alt) {
	Handle* h = (Handle*)hp;
	try {
		QueryLog log;
		auto s = samples_has_var(h->vg.get(), h->idx.get(), pos, ref, alt, false, "", &log);
		if (!log.err.empty()) return dup_str(log.err);
		std::string t;
		for (auto& p : s) { t += p.first; t += ' '; t += p.second; }
		t += '\n';
		return dup_str(t);
	} catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}

// batched forms: counts[i] = rows, digests[i] = FNV-1a over the row text (with or without carriers)
int vso_batch_t6(void* hp, uint64_t n, const uint64_t* x, const uint64_t* y, uint64_t* counts, uint64_t* digests, int with_samples) {
	Handle* h = (Handle*)hp;
	try {
		QueryLog log;
		for (uint64_t i = 0; i < n; i++) {
			log.out.clear();
			auto vars = get_var_in_ref(h->vg.get(), h->idx.get(), x[i], y[i], false, "", &log);
			counts[i] = vars.size();
			if (digests) digests[i] = rows_digest(vars, with_samples != 0);
		}
		return 0;
	} catch (const std::exception& e) { g_err = e.what(); return -1; }
}
int vso_batch_t4(void* hp, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample_ids,
                 uint64_t* counts, uint64_t* digests, uint8_t* ub, int with_samples) {
	Handle* h = (Handle*)hp;
	try {
		QueryLog log;
		for (uint64_t i = 0; i < n; i++) {
			log.out.clear();
			bool u = false;
			std::string name = h->vg->get_sample_name(sample_ids[i]);
			auto vars = get_sample_var_in_ref(h->vg.get(), h->idx.get(), x[i], y[i], name, false, "", &log, &u);
			counts[i] = vars.size();
			if (digests) digests[i] = rows_digest(vars, with_samples != 0);
			if (ub) ub[i] = u;
		}
		return 0;
	} catch (const std::exception& e) { g_err = e.what(); return -1; }
}
// get_sample_var_in_sample (query.h:490-612): status 2 = the reference never returns (no rows then)
int vso_batch_t5(void* hp, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample_ids,
                 uint64_t* counts, uint64_t* digests, uint8_t* status, uint8_t* ub, int with_samples) {
	Handle* h = (Handle*)hp;
	try {
		QueryLog log;
		for (uint64_t i = 0; i < n; i++) {
			log.out.clear();
			bool u = false, hang = false;
			std::string name = h->vg->get_sample_name(sample_ids[i]);
			auto vars = get_sample_var_in_sample(h->vg.get(), h->idx.get(), x[i], y[i], name, false, "", &log, &u, &hang);
			counts[i] = vars.size();
			if (digests) digests[i] = rows_digest(vars, with_samples != 0);
			if (status) status[i] = hang ? 2 : 0;
			if (ub) ub[i] = u;
		}
		return 0;
	} catch (const std::exception& e) { g_err = e.what(); return -1; }
}
char* vso_query_t5_text(void* hp, uint64_t x, uint64_t y, const char* sample) {
	Handle* h = (Handle*)hp;
	try {
		QueryLog log; bool hang = false;
		auto vars = get_sample_var_in_sample(h->vg.get(), h->idx.get(), x, y, sample, false, "", &log, nullptr, &hang);
		if (hang) return dup_str("HANG\n");
		std::string t = log.out + "Pos\tRef\tAlt\tSamples\n";
		for (auto& v : vars) row_text(v, true, t);
		return dup_str(t);
	} catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
// refs/alts: n NUL-terminated strings each.  found[i] = 1 when a record matched; counts = carriers.
int vso_batch_t7(void* hp, uint64_t n, const uint64_t* pos, const char* const* refs, const char* const* alts,
                 uint8_t* found, uint64_t* counts, uint64_t* digests) {
	Handle* h = (Handle*)hp;
	try {
		for (uint64_t i = 0; i < n; i++) {
			QueryLog log;
			auto s = samples_has_var(h->vg.get(), h->idx.get(), pos[i], refs[i], alts[i], false, "", &log);
			found[i] = log.err.empty() ? 1 : 0;
			counts[i] = s.size();
			if (digests) { uint64_t d = kFnvInit; for (auto& p : s) { d = fnv1a(d, p.first.data(), p.first.size()); d = fnv1a(d, " ", 1); d = fnv1a(d, p.second.data(), p.second.size()); } digests[i] = d; }
		}
		return 0;
	} catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// query_sample_from_ref (query.h:120-189).  status[i]: 0 = a sequence came back, 1 = the call ended in
// std::out_of_range (the reference process would terminate).  lengths / digests describe the sequence;
// `text` (nullable) receives the sequences joined by '\n' (malloc'd).
int vso_batch_t2(void* hp, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample_ids,
                 uint64_t* lengths, uint64_t* digests, uint8_t* status, uint8_t* ub, char** text) {
	Handle* h = (Handle*)hp;
	try {
		std::string all;
		for (uint64_t i = 0; i < n; i++) {
			bool u = false;
			std::string name = h->vg->get_sample_name(sample_ids[i]);
			std::string seq; uint8_t st = 0;
			try { seq = query_sample_from_ref(h->vg.get(), h->idx.get(), x[i], y[i], name, false, "", &u); }
			catch (const std::out_of_range&) { seq.clear(); st = 1; }
			lengths[i] = seq.size();
			if (digests) digests[i] = fnv1a(kFnvInit, seq.data(), seq.size());
			if (status) status[i] = st;
			if (ub) ub[i] = u;
			if (text) { all += seq; all += '\n'; }
		}
		if (text) *text = dup_str(all);
		return 0;
	} catch (const std::exception& e) { g_err = e.what(); return -1; }
}
// query_sample_from_sample (query.h:195-261).  status: 0 = a sequence came back, 1 = std::out_of_range,
// 2 = the loop at :209-214 never ends (the reference hangs).
int vso_batch_t3(void* hp, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample_ids,
                 uint64_t* lengths, uint64_t* digests, uint8_t* status, uint8_t* ub, char** text) {
	Handle* h = (Handle*)hp;
	try {
		std::string all;
		for (uint64_t i = 0; i < n; i++) {
			bool u = false, hang = false;
			std::string name = h->vg->get_sample_name(sample_ids[i]);
			std::string seq; uint8_t st = 0;
			try { seq = query_sample_from_sample(h->vg.get(), h->idx.get(), x[i], y[i], name, false, "", &u, &hang); }
			catch (const std::out_of_range&) { seq.clear(); st = 1; }
			if (hang) { seq.clear(); st = 2; }
			lengths[i] = seq.size();
			if (digests) digests[i] = fnv1a(kFnvInit, seq.data(), seq.size());
			if (status) status[i] = st;
			if (ub) ub[i] = u;
			if (text) { all += seq; all += '\n'; }
		}
		if (text) *text = dup_str(all);
		return 0;
	} catch (const std::exception& e) { g_err = e.what(); return -1; }
}
int vso_batch_t2_mt(void* hp, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample_ids, uint64_t* lengths, int nthreads) {
	Handle* h = (Handle*)hp;
	if (nthreads < 1) nthreads = 1;
	std::vector<std::thread> th; std::vector<int> rc(nthreads, 0);
	for (int t = 0; t < nthreads; t++) th.emplace_back([&, t]() {
		try {
			for (uint64_t i = t; i < n; i += nthreads) {
				std::string name = h->vg->get_sample_name(sample_ids[i]);
				try { lengths[i] = query_sample_from_ref(h->vg.get(), h->idx.get(), x[i], y[i], name).size(); }
				catch (const std::out_of_range&) { lengths[i] = 0; }
			}
		} catch (const std::exception&) { rc[t] = -1; }
	});
	for (auto& t : th) t.join();
	for (int r : rc) if (r) return r;
	return 0;
}

// closest_var (query.h:441-483): found flag, row count, digest of the rows
int vso_batch_t1(void* hp, uint64_t n, const uint64_t* pos, uint8_t* found, uint64_t* counts, uint64_t* digests, int with_samples) {
	Handle* h = (Handle*)hp;
	try {
		for (uint64_t i = 0; i < n; i++) {
			std::vector<Variant> vars;
			found[i] = closest_var(h->vg.get(), h->idx.get(), pos[i], vars) ? 1 : 0;
			counts[i] = vars.size();
			if (digests) digests[i] = rows_digest(vars, with_samples != 0);
		}
		return 0;
	} catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// Multi-threaded timing arms for bench.py (`--impl reference`, cpu_baseline): the reference query
// path is single-threaded; "all host cores" = independent workers over disjoint region chunks, which
// is how its evaluation ran contigs side by side (eval_data_records/evaluation.txt:34).
int vso_batch_t6_mt(void* hp, uint64_t n, const uint64_t* x, const uint64_t* y, uint64_t* counts, int nthreads) {
	Handle* h = (Handle*)hp;
	if (nthreads < 1) nthreads = 1;
	std::vector<std::thread> th; std::vector<int> rc(nthreads, 0);
	for (int t = 0; t < nthreads; t++) th.emplace_back([&, t]() {
		try {
			QueryLog log;
			for (uint64_t i = t; i < n; i += nthreads) { log.out.clear(); counts[i] = get_var_in_ref(h->vg.get(), h->idx.get(), x[i], y[i], false, "", &log).size(); }
		} catch (const std::exception&) { rc[t] = -1; }
	});
	for (auto& t : th) t.join();
	for (int r : rc) if (r) return r;
	return 0;
}
int vso_batch_t4_mt(void* hp, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample_ids, uint64_t* counts, int nthreads) {
	Handle* h = (Handle*)hp;
	if (nthreads < 1) nthreads = 1;
	std::vector<std::thread> th; std::vector<int> rc(nthreads, 0);
	for (int t = 0; t < nthreads; t++) th.emplace_back([&, t]() {
		try {
			QueryLog log;
			for (uint64_t i = t; i < n; i += nthreads) {
				log.out.clear();
				std::string name = h->vg->get_sample_name(sample_ids[i]);
				counts[i] = get_sample_var_in_ref(h->vg.get(), h->idx.get(), x[i], y[i], name, false, "", &log).size();
			}
		} catch (const std::exception&) { rc[t] = -1; }
	});
	for (auto& t : th) t.join();
	for (int r : rc) if (r) return r;
	return 0;
}

// Every distinct record next_variant_in_ref can report, in backbone order: used by tests to build
// t7 lookups that hit.  Returns a malloc'd text "pos\tref\talt\n"... (one row per record).
char* vso_all_variants_text(void* hp) {
	Handle* h = (Handle*)hp;
	try {
		QueryLog log;
		auto vars = get_var_in_ref(h->vg.get(), h->idx.get(), 1, h->vg->get_ref_length() + 1, false, "", &log);
		std::string t;
		for (auto& v : vars) row_text(v, false, t);
		return dup_str(t);
	} catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}

// ---------------------------------------------------------------- synthetic construct
// No VCF text: records are drawn here and handed to add_allele with name-sorted carrier lists
// (sample names are zero-padded, so name order == id order == column order).
//   mode 0: 1000-Genomes-like (phased a|b, carrier count ~ 1/f up to fmax haplotypes) -> classes
//   mode 1: TCGA-like sparse (1-3 carriers, unphased 0/1)                              -> explicit ids
// overlap != 0 lets a record start inside the previous record's REF span.
void* vso_synth(const char* prefix, const char* chr, uint64_t ref_length, uint64_t pos_lo, uint64_t pos_hi,
                uint64_t n_records, uint32_t n_samples, uint32_t fmax, double frac_multi, double frac_indel,
                int mode, int overlap, uint64_t seed, int cqf_log2, int fix_idx, int gzip_level, double reuse_prob) {
	try {
		Rng rng(seed);
		std::string ref(ref_length, 'A');
		static const char B[4] = {'A', 'C', 'G', 'T'};
		for (uint64_t i = 0; i < ref_length; i += 32) { uint64_t r = rng.next(); for (unsigned k = 0; k < 32 && i + k < ref_length; k++) ref[i + k] = B[(r >> (2 * k)) & 3]; }
		ConstructOpts o; o.cqf_log2_slots = cqf_log2 > 0 ? cqf_log2 : 25; o.fix_sample_indexes = fix_idx != 0; o.gzip_level = gzip_level;
		mkdir(prefix, 0755);
		auto h = new Handle;
		h->vg.reset(new VariantGraph(chr, ref, prefix, o, mode == 0));
		std::vector<std::string> names;
		unsigned digits = 1; for (uint32_t t = n_samples; t >= 10; t /= 10) digits++;
		char buf[32];
		for (uint32_t i = 1; i <= n_samples; i++) { snprintf(buf, sizeof buf, "S%0*u", (int)digits, i); names.push_back(buf); }
		h->vg->set_sample_names(names);
		// sorted distinct positions
		if (pos_hi > ref_length - 8) pos_hi = ref_length - 8;
		if (pos_lo < 2) pos_lo = 2;
		std::vector<uint64_t> pos(n_records);
		for (auto& p : pos) p = pos_lo + rng.below(pos_hi - pos_lo + 1);
		std::sort(pos.begin(), pos.end());
		pos.erase(std::unique(pos.begin(), pos.end()), pos.end());
		// harmonic CDF for the 1/f spectrum
		std::vector<double> cdf;
		if (mode == 0) { cdf.resize(fmax); double s = 0; for (uint32_t f = 1; f <= fmax; f++) { s += 1.0 / f; cdf[f - 1] = s; } for (auto& c : cdf) c /= s; }
		std::vector<uint32_t> hap_mark(2 * (size_t)n_samples, 0); uint32_t stamp = 0;
		std::vector<uint32_t> haps;
		auto draw_carriers = [&](std::vector<SampleStruct>& out) {
			out.clear();
			if (mode == 0) {
				uint32_t f = (uint32_t)(std::lower_bound(cdf.begin(), cdf.end(), rng.unit()) - cdf.begin()) + 1;
				f = std::min<uint32_t>(f, 2 * n_samples);
				stamp++; haps.clear();
				while (haps.size() < f) { uint32_t hpl = (uint32_t)rng.below(2 * (uint64_t)n_samples); if (hap_mark[hpl] != stamp) { hap_mark[hpl] = stamp; haps.push_back(hpl); } }
				std::sort(haps.begin(), haps.end());
				for (size_t i = 0; i < haps.size();) {
					uint32_t s = haps[i] / 2; bool g1 = false, g2 = false;
					while (i < haps.size() && haps[i] / 2 == s) { if (haps[i] & 1) g2 = true; else g1 = true; i++; }
					out.push_back(SampleStruct{s + 1, true, g1, g2});
				}
			} else {
				uint32_t k = 1 + (uint32_t)rng.below(3);
				stamp++; haps.clear();
				while (haps.size() < k) { uint32_t s = (uint32_t)rng.below(n_samples); if (hap_mark[s] != stamp) { hap_mark[s] = stamp; haps.push_back(s); } }
				std::sort(haps.begin(), haps.end());
				for (auto s : haps) out.push_back(SampleStruct{s + 1, false, false, true});
			}
		};
		std::vector<SampleStruct> carriers;
		// linkage stand-in: with probability reuse_prob an allele repeats the carrier set of one of the
		// last 64 alleles, so that ~half of the alleles share a sample class as in 1000 Genomes
		// (eval_data_records/logs/vs_v1.log: 1 106 184 alleles -> 490 775 classes on chr22)
		std::vector<std::vector<SampleStruct>> recent; size_t recent_next = 0;
		auto next_carriers = [&]() {
			if (!recent.empty() && rng.unit() < reuse_prob) { carriers = recent[rng.below(recent.size())]; return; }
			draw_carriers(carriers);
			if (recent.size() < 64) recent.push_back(carriers); else { recent[recent_next] = carriers; recent_next = (recent_next + 1) % 64; }
		};
		uint64_t prev_end = 0;   // last reference base covered by the previous record's REF
		for (size_t i = 0; i < pos.size(); i++) {
			uint64_t p = pos[i];
			if (!overlap && p <= prev_end) continue;
			uint64_t room = (i + 1 < pos.size() ? pos[i + 1] : ref_length) - p;   // bases before the next record
			double u = rng.unit();
			std::string r(1, ref[p - 1]);
			std::vector<std::string> alts;
			auto other_base = [&](char c) { char a; do { a = B[rng.below(4)]; } while (a == c); return a; };
			if (u < frac_indel / 2) {            // insertion R -> R + 1..3 bases
				unsigned len = 1; while (len < 3 && rng.unit() < 0.3) len++;
				std::string a = r; for (unsigned k = 0; k < len; k++) a += B[rng.below(4)];
				alts.push_back(a);
			} else if (u < frac_indel) {         // deletion of 1..3 bases
				unsigned len = 1; while (len < 3 && rng.unit() < 0.3) len++;
				if (!overlap && len + 1 > room) len = (unsigned)std::max<uint64_t>(1, room) - (room > 1 ? 1 : 0);
				if (len >= 1 && p + len <= ref_length && (overlap || len + 1 <= room)) { r = ref.substr(p - 1, len + 1); alts.push_back(std::string(1, ref[p - 1])); }
				else alts.push_back(std::string(1, other_base(ref[p - 1])));
			} else {
				char a = other_base(ref[p - 1]);
				alts.push_back(std::string(1, a));
				if (rng.unit() < frac_multi) { char b2; do { b2 = other_base(ref[p - 1]); } while (b2 == a); alts.push_back(std::string(1, b2)); }
			}
			h->vg->count_record();
			for (auto& a : alts) { next_carriers(); h->vg->add_allele(r, a, p, carriers); }
			prev_end = p + r.size() - 1;
		}
		h->vg->finish_construct();
		h->vg->serialize();
		h->idx.reset(new Index(h->vg.get()));
		h->idx->serialize(prefix);
		return h;
	} catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}

// ---------------------------------------------------------------- codec / CQF self-checks for tests
int vso_rrr_roundtrip(const uint64_t* words, uint64_t nbits, const char* path) {
	BitVec a; a.nbits = nbits; a.w.assign(words, words + (nbits + 63) / 64);
	if (!codec::write_rrr127(path, a)) return -1;
	BitVec b;
	if (!codec::read_rrr127(path, b)) return -2;
	if (b.nbits != nbits) return -3;
	for (uint64_t i = 0; i < (nbits + 63) / 64; i++) {
		uint64_t m = (i == nbits / 64 && (nbits & 63)) ? ((1ULL << (nbits & 63)) - 1) : ~0ULL;
		if ((a.w[i] & m) != (b.w[i] & m)) return -4;
	}
	return 0;
}

// Encode n vertices described by flat arrays into VariantGraphVertexList bytes (for the
// cross-check against python google.protobuf).  s_info rows: index, sample_id (or -1), phase, gt1, gt2.
char* vso_encode_vertices(uint64_t n, const uint32_t* ids, const uint32_t* offs, const uint32_t* lens,
                          const int64_t* class_ids, const uint32_t* s_begin, const int64_t* s_rows, uint64_t* out_len) {
	std::vector<Vertex> vs(n);
	for (uint64_t i = 0; i < n; i++) {
		vs[i].vertex_id = ids[i]; vs[i].offset = offs[i]; vs[i].length = lens[i];
		if (class_ids[i] >= 0) { vs[i].has_class = true; vs[i].class_id = (uint32_t)class_ids[i]; }
		for (uint32_t k = s_begin[i]; k < s_begin[i + 1]; k++) {
			SampleInfo s; s.index = (uint32_t)s_rows[5 * k];
			if (s_rows[5 * k + 1] >= 0) { s.has_sid = 1; s.sample_id = (uint32_t)s_rows[5 * k + 1]; }
			s.phase = s_rows[5 * k + 2] != 0; s.gt1 = s_rows[5 * k + 3] != 0; s.gt2 = s_rows[5 * k + 4] != 0;
			vs[i].s_info.push_back(s);
		}
	}
	std::string o; codec::encode_vertex_list(vs.data(), n, o);
	*out_len = o.size();
	char* p = (char*)malloc(o.size() + 1); memcpy(p, o.data(), o.size()); return p;
}

// Differential check of the port CQF store against the real gqf (oracle/_ref): replays `nops`
// random add/remove edge operations on two Graphs and compares adjacency + the serialised file.
// Returns 0 equal, 1 files differ byte-wise but enumerate equally, <0 mismatch, -100 no _ref.
int vso_cqf_differential(uint64_t seed, uint64_t nops, uint32_t nverts, int log2_slots, const char* tmp_prefix) {
	try {
		auto ref_probe = make_ref_adjstore(10);
		if (!ref_probe) return -100;
		Graph a(log2_slots, false), b(log2_slots, true);
		Rng rng(seed);
		for (uint64_t i = 0; i < nops; i++) {
			uint32_t s = (uint32_t)rng.below(nverts), d = 1 + (uint32_t)rng.below(nverts);
			if (rng.below(8) == 0) { a.remove_edge(s, d); b.remove_edge(s, d); }
			else { a.add_edge(s, d); b.add_edge(s, d); }
		}
		for (uint32_t v = 0; v < nverts; v++) {
			auto x = a.out_neighbors(v), y = b.out_neighbors(v);
			if (std::vector<uint32_t>(x.begin(), x.end()) != std::vector<uint32_t>(y.begin(), y.end())) return -1;
		}
		if (a.get_num_vertices() != b.get_num_vertices() || a.get_num_edges() != b.get_num_edges()) return -2;
		std::vector<std::array<uint64_t, 3>> ea, eb;
		a.adj->enumerate(ea); b.adj->enumerate(eb);
		if (ea != eb) return -3;
		std::string pa = std::string(tmp_prefix) + "_port", pb = std::string(tmp_prefix) + "_ref";
		mkdir(pa.c_str(), 0755); mkdir(pb.c_str(), 0755);
		a.serialize(pa); b.serialize(pb);
		// each file must be readable by the *other* implementation
		auto la = load_ref_adjstore(pa + "/adj_list.cqf"); auto lb = load_port_adjstore(pb + "/adj_list.cqf");
		if (!la || !lb) return -4;
		std::vector<std::array<uint64_t, 3>> fa, fb;
		la->enumerate(fa); lb->enumerate(fb);
		if (fa != ea || fb != ea) return -5;
		FILE* f1 = fopen((pa + "/adj_list.cqf").c_str(), "rb"); FILE* f2 = fopen((pb + "/adj_list.cqf").c_str(), "rb");
		int same = 1; int c1, c2;
		do { c1 = fgetc(f1); c2 = fgetc(f2); if (c1 != c2) { same = 0; break; } } while (c1 != EOF);
		fclose(f1); fclose(f2);
		return same ? 0 : 1;
	} catch (const std::exception& e) { g_err = e.what(); return -50; }
}

}  // extern "C"
