// libvsgpu host side — on-disk cache of the loaded + flattened index.
//
// vsgpu_open spends its time inflating and parsing the gz-protobuf vertex blocks, decoding the rrr
// vectors and flattening the graph; none of that depends on anything but the files of `ser/`.
// With VSGPU_INDEX_CACHE set, the result is written once as flat arrays and later opens read it
// back instead (a few hundred MB of sequential reads).  The cache is keyed by a fingerprint of the
// ser/ files (names, sizes, modification times) and by the layout version below; anything that
// does not match, is truncated or unreadable is ignored and rebuilt.  Opt-in, because it writes
// beside the user's data:
//   VSGPU_INDEX_CACHE=1       ->  <prefix>/vsgpu_flat.cache
//   VSGPU_INDEX_CACHE=<dir>   ->  <dir>/<fingerprint of the absolute prefix>.vsgpu_cache
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <sys/stat.h>
#include <type_traits>
#include <unistd.h>

#include "host_index.h"

namespace vsgpu {
namespace {

constexpr uint64_t kMagic = 0x3143464750475356ULL;   // "VSGPGFC1"
constexpr uint32_t kLayoutVersion = 2;                // bump whenever SerData / FlatIndex / the walk-entry encoding changes
constexpr uint64_t kEndMark = 0x444E455F43465356ULL;

struct Writer {
	FILE* f; bool ok = true;
	void raw(const void* p, size_t n) { if (ok && n && fwrite(p, 1, n, f) != n) ok = false; }
	template <class T> void pod(T& v) { static_assert(std::is_trivially_copyable<T>::value, "pod"); raw(&v, sizeof v); }
	template <class T> void vec(std::vector<T>& v) { static_assert(std::is_trivially_copyable<T>::value, "pod"); uint64_t n = v.size(); pod(n); raw(v.data(), n * sizeof(T)); }
	void str(std::string& s) { uint64_t n = s.size(); pod(n); raw(s.data(), n); }
	void strs(std::vector<std::string>& v) { uint64_t n = v.size(); pod(n); for (auto& s : v) str(s); }
};

struct Reader {
	FILE* f; bool ok = true; uint64_t left;   // bytes of the file not yet consumed: bounds every length field
	void raw(void* p, size_t n) { if (!ok) return; if (n > left || (n && fread(p, 1, n, f) != n)) { ok = false; return; } left -= n; }
	template <class T> void pod(T& v) { raw(&v, sizeof v); }
	template <class T> void vec(std::vector<T>& v) {
		uint64_t n = 0; pod(n);
		if (!ok || n > left / sizeof(T)) { ok = false; return; }
		v.resize(n); raw(v.data(), n * sizeof(T));
	}
	void str(std::string& s) { uint64_t n = 0; pod(n); if (!ok || n > left) { ok = false; return; } s.resize(n); raw(&s[0], n); }
	void strs(std::vector<std::string>& v) { uint64_t n = 0; pod(n); if (!ok || n > left / 8) { ok = false; return; } v.resize(n); for (auto& s : v) str(s); }
};

// every field of HostIndex that is read after build_host_index, once, for both directions
template <class A>
void archive(A& a, HostIndex& h) {
	SerData& s = h.ser; FlatIndex& f = h.flat;
	a.str(s.chr); a.pod(s.ref_length); a.pod(s.num_samples); a.strs(s.sample_names);
	a.pod(s.index_bits); a.pod(s.num_vertices); a.pod(s.class_mode);
	a.vec(s.v_offset); a.vec(s.v_length); a.vec(s.v_class); a.vec(s.v_sinfo_begin); a.vec(s.s_sample_id); a.vec(s.s_flags);
	a.vec(s.seq); a.pod(s.sample_vector_bits); a.pod(s.cqf_distinct);
	a.pod(f.ref_length); a.pod(f.index_bits); a.pod(f.num_samples); a.pod(f.M); a.pod(f.D); a.pod(f.R); a.pod(f.num_sets); a.pod(f.words_per_set);
	a.pod(f.class_mode); a.pod(f.has_suspect_dups);
	a.vec(f.bb_vertex); a.vec(f.vstart); a.vec(f.vlen); a.vec(f.rec_begin); a.vec(f.cent_begin); a.vec(f.bb_set); a.vec(f.vertex_bb); a.vec(f.bb_nrp); a.vec(f.bb_nref);
	a.vec(f.dstart); a.vec(f.dlev); a.vec(f.dinfo); a.vec(f.t7_lo); a.vec(f.t7_hi);
	a.vec(f.cent); a.vec(f.cent_vertex); a.pod(f.row_words); a.vec(f.marker_bits); a.vec(f.dtin); a.vec(f.cent_anc);
	a.vec(f.rec_k); a.vec(f.rec_vertex); a.vec(f.rec_pos); a.vec(f.rec_refv); a.vec(f.rec_altv); a.vec(f.rec_flags); a.vec(f.rec_hash); a.vec(f.rec_dup_prefix);
	a.vec(f.bitmap); a.vec(f.list_begin); a.vec(f.list_ids);
	a.pod(h.last_end); a.pod(h.t1_fallback_pos);
	a.vec(f.nrp1); a.vec(f.first_reach); a.vec(f.cent_seq); a.pod(f.t2_ok); a.str(f.t2_why);
}

uint64_t mix(uint64_t h, const void* p, size_t n) { return fnv1a(h, p, n); }

// names, sizes and mtimes of everything vsgpu_open reads under the prefix
bool fingerprint(const std::string& prefix, uint64_t& fp) {
	uint64_t h = kFnvInit;
	auto add = [&](const std::string& name, bool required) {
		struct stat st;
		if (stat((prefix + "/" + name).c_str(), &st) != 0) return !required;
		h = mix(h, name.data(), name.size());
		const int64_t v[3] = {(int64_t)st.st_size, (int64_t)st.st_mtim.tv_sec, (int64_t)st.st_mtim.tv_nsec};
		h = mix(h, v, sizeof v);
		return true;
	};
	for (const char* n : {"index.sdsl", "ref_node_id.sdsl", "adj_list.cqf", "aux_vertex_list.sdsl", "aux_vertex_list_lengths.sdsl", "seq_buffer.sdsl", "sample_vector.sdsl", "sampleid_map.lst"})
		if (!add(n, true)) return false;
	for (uint64_t b = 0;; b++) { struct stat st; const std::string n = "vertex_list_" + std::to_string(b) + ".proto"; if (stat((prefix + "/" + n).c_str(), &st) != 0) break; add(n, true); }
	fp = h;
	return true;
}

std::string cache_path(const std::string& prefix) {
	const char* e = getenv("VSGPU_INDEX_CACHE");
	if (!e || !*e || !strcmp(e, "0")) return "";
	if (!strcmp(e, "1")) return prefix + "/vsgpu_flat.cache";
	char real[PATH_MAX];
	const std::string abs = realpath(prefix.c_str(), real) ? std::string(real) : prefix;
	char name[40];
	snprintf(name, sizeof name, "%016llx.vsgpu_cache", (unsigned long long)fnv1a(kFnvInit, abs.data(), abs.size()));
	return std::string(e) + "/" + name;
}

struct Header { uint64_t magic; uint32_t version, sizes; uint64_t fingerprint; };
constexpr uint32_t kSizes = (uint32_t)(sizeof(CEntry) | (sizeof(DLevel) << 8) | (sizeof(size_t) << 16));

}  // namespace

bool load_index_cache(const std::string& prefix, HostIndex& h) {
	const std::string path = cache_path(prefix);
	uint64_t fp = 0;
	if (path.empty() || !fingerprint(prefix, fp)) return false;
	FILE* f = fopen(path.c_str(), "rb");
	if (!f) return false;
	struct stat st;
	bool ok = fstat(fileno(f), &st) == 0;
	Reader r{f, ok, ok ? (uint64_t)st.st_size : 0};
	Header hd{};
	r.pod(hd);
	if (r.ok && hd.magic == kMagic && hd.version == kLayoutVersion && hd.sizes == kSizes && hd.fingerprint == fp) {
		archive(r, h);
		uint64_t end = 0; r.pod(end);
		ok = r.ok && end == kEndMark && r.left == 0;
	} else ok = false;
	fclose(f);
	if (ok) {   // cross-checks the kernels rely on; a cache that fails them is treated as absent
		const FlatIndex& x = h.flat;
		ok = x.M > 0 && x.bb_vertex.size() == x.M && x.vstart.size() == x.M && x.rec_begin.size() == (size_t)x.M + 1 && x.dstart.size() == x.D && x.dlev.size() == (size_t)x.D + 1
		     && x.rec_pos.size() == x.R && x.cent_anc.size() == 2 * x.cent.size() && x.cent_seq.size() == 2 * x.cent.size() && x.nrp1.size() == x.M && x.first_reach.size() == (size_t)x.D + 1 && h.ser.sample_names.size() == h.ser.num_samples && h.ser.v_sinfo_begin.size() == (size_t)h.ser.num_vertices + 1;
	}
	if (ok) {   // slices of the sequence buffer and vertex ids the materialiser / render / copy kernels read unchecked
		const FlatIndex& x = h.flat; const SerData& sd = h.ser;
		ok = sd.v_offset.size() == sd.num_vertices && sd.v_length.size() == sd.num_vertices && x.rec_refv.size() == x.R && x.rec_altv.size() == x.R && x.rec_vertex.size() == x.R;
		if (ok) { try { check_seq_ranges(sd); } catch (const std::exception&) { ok = false; } }
		for (size_t c = 0; ok && c < x.cent.size(); c++) ok = (uint64_t)x.cent_seq[2 * c] + x.cent_seq[2 * c + 1] <= sd.seq.size();
		for (uint32_t r = 0; ok && r < x.R; r++) ok = (x.rec_refv[r] == kNone || x.rec_refv[r] < sd.num_vertices) && (x.rec_altv[r] == kNone || x.rec_altv[r] < sd.num_vertices) && x.rec_vertex[r] < sd.num_vertices;
	}
	if (!ok) { h.ser = SerData(); h.flat = FlatIndex(); h.last_end = 0; h.t1_fallback_pos = 0; }
	return ok;
}

// best effort: a cache that cannot be written is not an error of vsgpu_open
void save_index_cache(const std::string& prefix, HostIndex& h) {
	const std::string path = cache_path(prefix);
	uint64_t fp = 0;
	if (path.empty() || !fingerprint(prefix, fp)) return;
	const std::string tmp = path + ".tmp." + std::to_string((long)getpid());
	FILE* f = fopen(tmp.c_str(), "wb");
	if (!f) return;
	Writer w{f};
	Header hd{kMagic, kLayoutVersion, kSizes, fp};
	w.pod(hd);
	archive(w, h);
	uint64_t end = kEndMark; w.pod(end);
	const bool ok = w.ok && fclose(f) == 0;
	if (!w.ok) fclose(f);
	if (!ok || rename(tmp.c_str(), path.c_str()) != 0) unlink(tmp.c_str());   // rename: readers never see a partial file
}

}  // namespace vsgpu
