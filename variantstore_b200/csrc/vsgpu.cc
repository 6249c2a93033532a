// libvsgpu — C ABI (include/vsgpu.h): index lifecycle, batched query entry points, host-side
// materialisation of result rows.  The query path has no CPU implementation: without a CUDA
// device every query entry point fails with VSGPU_ENODEVICE.
#include "../../include/vsgpu.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "host_index.h"
#include "kernels.cuh"

using namespace vsgpu;

namespace {
thread_local std::string g_err;
int set_err(int code, const std::string& m) { g_err = m; return code; }

// VSGPU_TRACE=1: wall-clock phases of a query call on stderr (diagnostics for the end-to-end numbers).
struct Trace {
	const char* name; bool on; std::vector<std::pair<const char*, std::chrono::steady_clock::time_point>> t;
	explicit Trace(const char* n) : name(n), on(getenv("VSGPU_TRACE") != nullptr) { mark("start"); }
	void mark(const char* label) { if (on) t.emplace_back(label, std::chrono::steady_clock::now()); }
	~Trace() {
		if (!on || t.size() < 2) return;
		fprintf(stderr, "[vsgpu trace] %s:", name);
		for (size_t i = 1; i < t.size(); i++) fprintf(stderr, " %s %.1fus", t[i].first, std::chrono::duration<double, std::micro>(t[i].second - t[i - 1].second).count());
		fprintf(stderr, " | total %.1fus\n", std::chrono::duration<double, std::micro>(t.back().second - t.front().second).count());
	}
};

struct DevBuf {
	void* p = nullptr; size_t cap = 0;
	cudaError_t ensure(size_t bytes) {
		if (bytes <= cap) return cudaSuccess;
		if (p) cudaFree(p);
		p = nullptr; cap = 0;
		size_t want = bytes + bytes / 8 + 256;
		cudaError_t e = cudaMalloc(&p, want);
		if (e == cudaSuccess) cap = want;
		return e;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
	template <class T> T* as() const { return (T*)p; }
};
}  // namespace

// Host side of a t4 answer.  Both arrays live in page-locked memory taken from the owning index's
// pool (device->host copies run at full PCIe rate and nothing is re-allocated per call).
struct vsgpu_result {
	vsgpu_index* owner = nullptr;
	uint64_t n = 0;
	uint64_t* offsets = nullptr; size_t offsets_cap = 0;   // bytes
	uint32_t* hits = nullptr; size_t hits_cap = 0;         // bytes
	uint8_t* status = nullptr; size_t status_cap = 0;      // t5 only
	// Host-buffer t4 calls bring back the per-region row counts (4 bytes a region over PCIe instead of 8); the
	// other of offsets / counts is built on the host the first time it is asked for.
	uint32_t* counts = nullptr; size_t counts_cap = 0;     // bytes
	uint64_t total = 0;
	bool have_offsets = false, have_counts = false;
	std::mutex lazy_mu;
	float kernel_ms = 0;                                   // t5 only
};

// Host side of a rendered t6 answer (page-locked, pooled like vsgpu_result).
struct vsgpu_text {
	vsgpu_index* owner = nullptr;
	uint64_t n = 0, nbytes = 0, nrows = 0;
	char* bytes = nullptr; size_t bytes_cap = 0;
	uint64_t* offsets = nullptr; size_t offsets_cap = 0;
	uint8_t* status = nullptr; size_t status_cap = 0;     // t2 only
	float kernel_ms = 0;
	float stage_ms[3] = {0, 0, 0};                        // t2 only: count (+ scan of CTA sums), plan, copy
};

struct vsgpu_index : vsgpu::HostIndex {
	DevIndex dev;
	int device = 0;
	cudaStream_t stream = nullptr;
	bool own_stream = false;
	std::vector<void*> allocs;
	uint64_t device_bytes = 0;
	bool sparse_walk = false;         // per-sample carried-entry lists instead of the hit map (sparse cohorts)
	double hits_per_base = 0;         // walk-entry carriers per base for an average sample
	double entries_per_base = 0;      // walk entries per covered base
	uint32_t* d_status = nullptr;     // two words: status bits, length of the t6 flagged list
	std::mutex mu;
	DevBuf bx, by, bs, bout, boffsets, bhits, bstate, bhash, brec, bflag, bx32, by32;
	// A large call is cut into chunks of regions so that the input copies (s_in), the kernels (stream)
	// and the result copies (s_out) of different chunks overlap; PCIe is full duplex.
	static constexpr int kMaxChunks = 16;
	// Host-buffer calls are synchronous, so they run entirely on internal streams: s_k for the kernels
	// (the caller's stream set with vsgpu_set_stream is for the device-resident batch API) and up to
	// three input streams (one by default).
	static constexpr int kInStreams = 3;
	cudaStream_t s_in[kInStreams] = {}, s_k = nullptr, s_out = nullptr;
	cudaEvent_t ev_in[kMaxChunks][kInStreams] = {}, ev_k[kMaxChunks] = {}, ev_out[kMaxChunks] = {};
	uint64_t* pin_small = nullptr;    // page-locked: kMaxChunks running totals + the two status words
	// device-side row rendering: tables uploaded on first use
	bool render_ready = false;
	RenderTables render{};
	DevBuf bseg, brow_off, bbyte_off, bscratch, btext;
	bool hit_tables_ready = false;
	HitTables hit_tables{};
	std::vector<uint64_t> set_text_bytes; std::vector<uint32_t> set_pop;     // per carrier set: bytes of its printed carrier list, members (class mode)
	// t2 (query_sample_from_ref): tables uploaded on first use
	bool t2_ready = false;
	T2Tables t2{};
	DevBuf bcnt, bst8, brecs, btile, bkeep, bpos5, bspill;
	const uint32_t* d_gsidx = nullptr;   // sample_info.index per s_info entry, s_info order (t5 rows rendered on the device); uploaded on first use
	bool t3_ready = false;
	T3Tables t3{};
	cudaEvent_t ev_t2[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
	cudaEvent_t ev_render[2] = {nullptr, nullptr};
	// page-locked host buffers: a free list for results + two staging areas for t6
	std::mutex pool_mu;
	std::vector<std::pair<void*, size_t>> pinned_free;
	void* stage[2] = {nullptr, nullptr}; size_t stage_cap[2] = {0, 0};
	void* pinned_acquire(size_t bytes, size_t* cap) {
		std::lock_guard<std::mutex> g(pool_mu);
		size_t best = SIZE_MAX;
		for (size_t i = 0; i < pinned_free.size(); i++) if (pinned_free[i].second >= bytes && (best == SIZE_MAX || pinned_free[i].second < pinned_free[best].second)) best = i;
		if (best != SIZE_MAX) { auto b = pinned_free[best]; pinned_free.erase(pinned_free.begin() + best); *cap = b.second; return b.first; }
		size_t want = 4096; while (want < bytes && want < (64u << 20)) want <<= 1;      // powers of two up to 64 MB, then 64 MB steps
		if (want < bytes) want = (bytes + (64u << 20) - 1) / (64u << 20) * (64u << 20);
		void* p = nullptr;
		if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
		*cap = want; return p;
	}
	void pinned_release(void* p, size_t cap) { if (!p) return; std::lock_guard<std::mutex> g(pool_mu); pinned_free.emplace_back(p, cap); }
	void* staging(int which, size_t bytes) {
		if (stage_cap[which] < bytes) { if (stage[which]) cudaFreeHost(stage[which]); stage[which] = nullptr; stage_cap[which] = 0; size_t want = 4096; while (want < bytes) want <<= 1; if (cudaHostAlloc(&stage[which], want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; } stage_cap[which] = want; }
		return stage[which];
	}
	~vsgpu_index() {
		cudaSetDevice(device);
		for (auto& b : pinned_free) cudaFreeHost(b.first);
		for (void* p : stage) if (p) cudaFreeHost(p);
		for (DevBuf* b : {&bx, &by, &bs, &bout, &boffsets, &bhits, &bstate, &bhash, &brec, &bflag, &bx32, &by32, &bseg, &brow_off, &bbyte_off, &bscratch, &btext, &bcnt, &bst8, &brecs, &btile, &bkeep, &bpos5, &bspill}) b->release();
		for (cudaEvent_t e : ev_render) if (e) cudaEventDestroy(e);
		for (cudaEvent_t e : ev_t2) if (e) cudaEventDestroy(e);
		for (int i = 0; i < kMaxChunks; i++) for (cudaEvent_t e : {ev_in[i][0], ev_in[i][1], ev_in[i][2], ev_k[i], ev_out[i]}) if (e) cudaEventDestroy(e);
		for (cudaStream_t st : {s_in[0], s_in[1], s_in[2], s_k, s_out}) if (st) cudaStreamDestroy(st);
		if (pin_small) cudaFreeHost(pin_small);
		for (void* p : allocs) cudaFree(p);
		if (d_status) cudaFree(d_status);
		if (own_stream && stream) cudaStreamDestroy(stream);
	}
};

struct vsgpu_batch {
	vsgpu_index* idx = nullptr;
	int device = 0;                  // copied from the index: freeing a batch must not touch an index that may be gone
	int type = 0; uint64_t n = 0;
	DevBuf x, y, s, hash, out, offsets, hits, state, rec, flag, spill;
	uint64_t hits_cap = 0;
	bool many_rows = false;          // regions expected to have more rows than k_t4p stages: run its spilling instance
	uint32_t launches = 0;
	uint64_t algo_bytes = 0; bool algo_valid = false;
	std::vector<uint64_t> hx, hy;   // host copies kept for the byte accounting
	cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
	uint32_t* d_status = nullptr;    // per batch: several batches may be in flight on one index
	bool wide_regions = false;
	~vsgpu_batch() { if (idx) cudaSetDevice(device); if (d_status) cudaFree(d_status); for (auto e : ev) if (e) cudaEventDestroy(e); for (DevBuf* b : {&x, &y, &s, &hash, &out, &offsets, &hits, &state, &rec, &flag, &spill}) b->release(); }
};

namespace {

#define CU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) throw std::runtime_error(std::string("CUDA: ") + cudaGetErrorString(e__) + " at " #call); } while (0)

template <class T>
const T* upload(vsgpu_index* ix, const std::vector<T>& v) {
	size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
	void* p = nullptr;
	CU(cudaMalloc(&p, bytes));
	ix->allocs.push_back(p); ix->device_bytes += bytes;
	if (!v.empty()) CU(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
	return (const T*)p;
}

void upload_index(vsgpu_index* ix) {
	FlatIndex& f = ix->flat; DevIndex& d = ix->dev;
	memset(&d, 0, sizeof d);
	d.D = f.D; d.M = f.M; d.R = f.R; d.num_cent = (uint32_t)f.cent.size(); d.words_per_set = f.words_per_set;
	d.num_samples = f.num_samples; d.class_mode = f.class_mode ? 1 : 0; d.index_bits = f.index_bits;
	d.last_end = ix->last_end;
	d.t1_fallback_pos = ix->t1_fallback_pos;
	std::vector<uint32_t> bucket;
	build_buckets(f, bucket, d.bucket_shift);
	d.nbuckets = (uint32_t)bucket.size() - 1;
	d.dstart = upload(ix, f.dstart);
	{ std::vector<uint32_t> d4; build_d4(f, d4); d.d4 = (const uint4*)upload(ix, d4); }
	d.bucket = upload(ix, bucket);
	{ const char* e = getenv("VSGPU_T4_ROW64"); d.walk2 = (!e || atoi(e) != 0) ? 1 : 0; }
	static_assert(sizeof(DLevel) == sizeof(uint4) && sizeof(CEntry) == sizeof(uint4), "AoS rows are 16 bytes");
	d.dlev = (const uint4*)upload(ix, f.dlev);
	d.dinfo = upload(ix, f.dinfo);
	std::vector<uint2> t7(f.D);
	for (uint32_t i = 0; i < f.D; i++) t7[i] = make_uint2(f.t7_lo[i], f.t7_hi[i]);
	d.t7rng = upload(ix, t7);
	d.cent = (const uint4*)upload(ix, f.cent);
	d.bb_set = upload(ix, f.bb_set);
	d.vstart = upload(ix, f.vstart);
	d.bitmap = upload(ix, f.bitmap);
	d.list_begin = upload(ix, f.list_begin);
	d.list_ids = upload(ix, f.list_ids);
	d.rec_pos = upload(ix, f.rec_pos);
	d.rec_hash = upload(ix, f.rec_hash);
	d.rec_flags = upload(ix, f.rec_flags);
	d.rec_dup_prefix = f.has_suspect_dups ? upload(ix, f.rec_dup_prefix) : nullptr;
	d.tail_records = f.rec_begin[f.M] > f.rec_begin[f.M - 1] ? 1 : 0;
	{   // carrier density: sum of carrier-set sizes over the walk entries / (samples x covered bases)
		long double carriers = 0;
		std::vector<uint32_t> set_size(f.num_sets, 0);
		if (f.class_mode) { for (uint32_t c = 0; c < f.num_sets; c++) for (uint32_t w = 0; w < f.words_per_set; w++) set_size[c] += (uint32_t)__builtin_popcountll(f.bitmap[(uint64_t)c * f.words_per_set + w]); }
		else for (uint32_t c = 0; c < f.num_sets; c++) set_size[c] = (uint32_t)(f.list_begin[c + 1] - f.list_begin[c]);
		for (const auto& e : f.cent) if (!(e.tgt & kEntMarker)) carriers += set_size[e.set_id];
		const long double span = std::max<long double>(1, (long double)f.dstart.back() - f.dstart[std::min<size_t>(1, f.dstart.size() - 1)] + 1);
		ix->hits_per_base = (double)(carriers / (std::max<uint32_t>(f.num_samples, 2) - 1) / span);
		ix->entries_per_base = (double)f.cent.size() / (double)span;
	}
	d.marker_bits = upload(ix, f.marker_bits);
	d.cent_begin_k = upload(ix, f.cent_begin);
	d.dtin = upload(ix, f.dtin);
	d.cent_anc = (const uint2*)upload(ix, f.cent_anc);
	d.row_words = f.row_words;
	d.hitmap = nullptr;
	CU(cudaMalloc((void**)&ix->d_status, 8));
	CU(cudaMemset(ix->d_status, 0, 8));
	for (cudaStream_t* st : {&ix->s_in[0], &ix->s_in[1], &ix->s_in[2], &ix->s_k, &ix->s_out}) CU(cudaStreamCreateWithFlags(st, cudaStreamNonBlocking));
	for (int i = 0; i < vsgpu_index::kMaxChunks; i++) for (cudaEvent_t* e : {&ix->ev_in[i][0], &ix->ev_in[i][1], &ix->ev_in[i][2], &ix->ev_k[i], &ix->ev_out[i]}) CU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
	CU(cudaHostAlloc((void**)&ix->pin_small, (vsgpu_index::kMaxChunks + 1) * 8, cudaHostAllocDefault));
	// Sample-major hit map: num_samples rows of row_words words.  Built on the device; skipped (the
	// kernels then test class bitmaps per entry) when it would not fit the budget:
	// VSGPU_HITMAP_MAX_GB (default 64) and at most half of the free device memory.
	if (want_sparse_walk(f)) {
		std::vector<uint64_t> car_begin; std::vector<uint32_t> car, marker_list;
		build_sparse_walk(f, car_begin, car, marker_list);
		d.car_begin = upload(ix, car_begin); d.car = upload(ix, car); d.marker_list = upload(ix, marker_list); d.num_markers = (uint32_t)marker_list.size(); d.marker_span = marker_span(f, marker_list);
		std::vector<uint64_t> can_begin; std::vector<uint32_t> can_entry, can_pmax;
		build_canonical_walks(f, car_begin, car, marker_list, can_begin, can_entry, can_pmax);
		d.can_begin = upload(ix, can_begin); d.can_entry = upload(ix, can_entry); d.can_pmax = upload(ix, can_pmax);
		ix->sparse_walk = true;
		return;                                              // no hit map: the lists answer every membership question of the walks
	}
	const uint64_t hm_bytes = (uint64_t)f.num_samples * f.row_words * 4;
	double max_gb = 64.0;
	if (const char* e = getenv("VSGPU_HITMAP_MAX_GB")) max_gb = atof(e);
	size_t free_b = 0, total_b = 0;
	CU(cudaMemGetInfo(&free_b, &total_b));
	if (!getenv("VSGPU_DISABLE_HITMAP") && hm_bytes <= (uint64_t)(max_gb * (1ull << 30)) && hm_bytes <= free_b / 2) {
		uint32_t* hm = nullptr;
		CU(cudaMalloc((void**)&hm, std::max<uint64_t>(hm_bytes, 16)));
		ix->allocs.push_back(hm); ix->device_bytes += hm_bytes;
		CU(cudaMemsetAsync(hm, 0, hm_bytes, ix->stream));
		CU(launch_build_hitmap(d, hm, ix->stream));
		CU(cudaStreamSynchronize(ix->stream));
		d.hitmap = hm;
	}
}

char* dup_text(const std::string& s) { char* p = (char*)malloc(s.size() + 1); if (!p) return nullptr; memcpy(p, s.data(), s.size()); p[s.size()] = 0; return p; }

int check_device(vsgpu_index* ix) {
	cudaError_t e = cudaSetDevice(ix->device);
	if (e != cudaSuccess) return set_err(VSGPU_ENODEVICE, std::string("CUDA: ") + cudaGetErrorString(e));
	return VSGPU_OK;
}

// Reads (and clears, when set) the two status words on `stream` (default: the index's), synchronising it.
uint32_t read_status(vsgpu_index* ix, uint32_t* d_status = nullptr, uint32_t* nflagged = nullptr, cudaStream_t stream = nullptr) {
	if (!d_status) d_status = ix->d_status;
	if (!stream) stream = ix->stream;
	uint32_t st[2] = {0, 0};
	cudaMemcpyAsync(st, d_status, 8, cudaMemcpyDeviceToHost, stream);
	cudaStreamSynchronize(stream);
	if (st[0] | st[1]) { cudaMemsetAsync(d_status, 0, 8, stream); cudaStreamSynchronize(stream); }
	if (nflagged) *nflagged = st[1];
	return st[0];
}

// input arrays of one chunk go out on this many streams (VSGPU_H2D_STREAMS, 1..3; measured on the
// B200 boxes: concurrent host->device copies are slower than one after another, profiles/README.md)
int in_streams() {
	int k = 1;
	if (const char* e = getenv("VSGPU_H2D_STREAMS")) k = atoi(e);
	return std::max(1, std::min(k, (int)vsgpu_index::kInStreams));
}

// regions per chunk of a large call (VSGPU_CHUNK_REGIONS; 0 = never cut) and the resulting plan
int plan_chunks(uint64_t n, uint64_t* per_chunk) {
	uint64_t cr = 262144;
	if (const char* e = getenv("VSGPU_CHUNK_REGIONS")) cr = strtoull(e, nullptr, 10);
	int c = 1;
	if (cr && n > cr + cr / 2) c = (int)std::min<uint64_t>(vsgpu_index::kMaxChunks, (n + cr - 1) / cr);
	*per_chunk = (((n + c - 1) / c) + 255) / 256 * 256;          // whole tiles per chunk
	return (int)((n + *per_chunk - 1) / *per_chunk);
}

}  // namespace

extern "C" {

const char* vsgpu_last_error(void) { return g_err.c_str(); }
void vsgpu_free(void* p) { free(p); }

int vsgpu_open(const char* ser_prefix, int device, vsgpu_index** out) {
	if (!ser_prefix || !out) return set_err(VSGPU_EINVAL, "vsgpu_open: null argument");
	*out = nullptr;
	std::unique_ptr<vsgpu_index> ix(new vsgpu_index);
	ix->device = device;
	int stage = 0;
	try { build_host_index(ser_prefix, *ix, &stage); } catch (const std::exception& e) { return set_err(stage == 0 ? VSGPU_EIO : VSGPU_ESHAPE, e.what()); }
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device) { cudaGetLastError(); return set_err(VSGPU_ENODEVICE, "vsgpu_open: no usable CUDA device (libvsgpu has no CPU path)"); }
	try {
		CU(cudaSetDevice(device));
		CU(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking)); ix->own_stream = true;
		upload_index(ix.get());
	} catch (const std::exception& e) { return set_err(VSGPU_ENODEVICE, e.what()); }
	*out = ix.release();
	return VSGPU_OK;
}

void vsgpu_close(vsgpu_index* idx) { delete idx; }

int vsgpu_info(const vsgpu_index* ix, vsgpu_info_t* o) {
	if (!ix || !o) return set_err(VSGPU_EINVAL, "vsgpu_info: null argument");
	memset(o, 0, sizeof *o);
	o->ref_length = ix->ser.ref_length; o->seq_length = ix->ser.seq.size(); o->num_vertices_cqf = ix->ser.cqf_distinct;
	o->num_vertices = ix->ser.num_vertices; o->num_samples = ix->ser.num_samples;
	o->num_classes = ix->flat.class_mode ? ix->flat.num_sets - 1 : 0; o->class_mode = ix->flat.class_mode;
	o->backbone_vertices = ix->flat.M; o->distinct_starts = ix->flat.D; o->branch_records = ix->flat.R;
	o->walk_entries = (uint32_t)ix->flat.cent.size(); o->has_suspect_dups = ix->flat.has_suspect_dups; o->device_bytes = ix->device_bytes;
	strncpy(o->chr, ix->ser.chr.c_str(), sizeof o->chr - 1);
	o->from_cache = ix->from_cache ? 1 : 0;
	for (const auto& e : ix->flat.cent) { if (e.tgt & kEntMarker) o->walk_markers++; else if ((e.tgt & kEntAlt) && (e.tgt & kEntTgtCarriers)) o->rejoin_carriers++; }
	return VSGPU_OK;
}

int vsgpu_set_stream(vsgpu_index* ix, void* s) {
	if (!ix) return set_err(VSGPU_EINVAL, "vsgpu_set_stream: null index");
	std::lock_guard<std::mutex> g(ix->mu);
	if (ix->own_stream && ix->stream) { cudaSetDevice(ix->device); cudaStreamSynchronize(ix->stream); cudaStreamDestroy(ix->stream); }
	ix->stream = (cudaStream_t)s; ix->own_stream = false;
	return VSGPU_OK;
}

int vsgpu_sample_id(const vsgpu_index* ix, const char* name, uint32_t* id) {
	if (!ix || !name || !id) return set_err(VSGPU_EINVAL, "vsgpu_sample_id: null argument");
	auto it = ix->name2id.find(name);
	if (it == ix->name2id.end()) return set_err(VSGPU_EINVAL, std::string("Sample not found: ") + name);
	*id = it->second; return VSGPU_OK;
}
const char* vsgpu_sample_name(const vsgpu_index* ix, uint32_t id) { return (ix && id < ix->ser.num_samples) ? ix->ser.sample_names[id].c_str() : nullptr; }

// ------------------------------------------------------------------ t6
namespace {
// counts[i] = rows the reference returns for region i: the slice length, except where the slice
// holds suspect duplicates or the region runs past the contig end (then the literal rule decides).
// The kernel writes the slice lengths and lists the exceptions; the host re-counts only those.
bool t6_special(const vsgpu_index* ix) { return ix->flat.has_suspect_dups || ix->dev.tail_records; }

// Reads the status words on `stream` (synchronising it), raises on a bad region and re-counts the
// regions the kernel flagged.
extern "C++" template <class T>
void settle_t6(vsgpu_index* ix, const T* x, const T* y, uint32_t* counts, const uint32_t* d_flag, uint32_t* d_status, cudaStream_t stream) {
	uint32_t nflag = 0;
	const uint32_t st = read_status(ix, d_status, &nflag, stream);
	if (st & kStatusBadRegion) throw std::invalid_argument("Can't find node corresponding to pos 0");   // index.h:151-154 aborts
	if (!nflag || !counts || !d_flag) return;
	std::vector<uint32_t> flagged(nflag), tmp;
	CU(cudaMemcpyAsync(flagged.data(), d_flag, (size_t)nflag * 4, cudaMemcpyDeviceToHost, stream));
	CU(cudaStreamSynchronize(stream));
	for (uint32_t i : flagged) { t6_literal(ix, x[i], y[i], tmp); counts[i] = (uint32_t)tmp.size(); }
}
}  // namespace

namespace {
// Host-buffer calls take the coordinates as 64-bit (the reference's uint64_t arguments) or 32-bit
// (they are parsed with std::stoi, commands.cc:76-80, so they fit): the 32-bit form halves the bytes
// that cross PCIe, and a small kernel widens them in HBM in front of the query kernel.
extern "C++" template <class T>
int query_t6_impl(vsgpu_index* ix, uint64_t n, const T* x, const T* y, uint32_t* rec_lo, uint32_t* rec_hi, uint32_t* counts) {
	constexpr bool k32 = sizeof(T) == 4;
	if (!ix || (n && (!x || !y))) return set_err(VSGPU_EINVAL, "vsgpu_query_t6: null argument");
	if (n == 0) return VSGPU_OK;
	if (int rc = check_device(ix)) return rc;
	std::lock_guard<std::mutex> g(ix->mu);
	Trace tr("t6");
	try {
		const bool flag = counts && t6_special(ix);
		CU(ix->bx.ensure(n * 8)); CU(ix->by.ensure(n * 8)); CU(ix->bout.ensure(n * 12));
		if (k32) { CU(ix->bx32.ensure(n * 4)); CU(ix->by32.ensure(n * 4)); }
		if (flag) CU(ix->bflag.ensure(n * 4));
		uint64_t* dx = ix->bx.as<uint64_t>(); uint64_t* dy = ix->by.as<uint64_t>();
		T* sx = k32 ? ix->bx32.as<T>() : (T*)dx; T* sy = k32 ? ix->by32.as<T>() : (T*)dy;      // where the host arrays land
		uint32_t* d_lo = ix->bout.as<uint32_t>(); uint32_t* d_hi = d_lo + n; uint32_t* d_cnt = d_hi + n;
		uint64_t per = 0;
		const int chunks = plan_chunks(n, &per);
		// inputs on s_in, kernels on s_k, results on s_out: chunk c's results travel
		// while chunk c+1's inputs arrive
		const int ks = in_streams();
		for (int c = 0; c < chunks; c++) {
			const uint64_t a = c * per, m = std::min<uint64_t>(n, a + per) - a;
			CU(cudaMemcpyAsync(sx + a, x + a, m * sizeof(T), cudaMemcpyHostToDevice, ix->s_in[0]));
			CU(cudaMemcpyAsync(sy + a, y + a, m * sizeof(T), cudaMemcpyHostToDevice, ix->s_in[1 % ks]));
			for (int k = 0; k < std::min(ks, 2); k++) CU(cudaEventRecord(ix->ev_in[c][k], ix->s_in[k]));
		}
		for (int c = 0; c < chunks; c++) {
			const uint64_t a = c * per, m = std::min<uint64_t>(n, a + per) - a;
			for (int k = 0; k < std::min(ks, 2); k++) CU(cudaStreamWaitEvent(ix->s_k, ix->ev_in[c][k], 0));
			if (k32) CU(launch_widen(m, (const uint32_t*)sx + a, (const uint32_t*)sy + a, dx + a, dy + a, ix->s_k));
			CU(launch_t6(ix->dev, m, dx + a, dy + a, d_lo + a, d_hi + a, counts ? d_cnt + a : nullptr, flag ? ix->bflag.as<uint32_t>() : nullptr, (uint32_t)a, ix->d_status, ix->s_k));
			CU(cudaEventRecord(ix->ev_k[c], ix->s_k));
			CU(cudaStreamWaitEvent(ix->s_out, ix->ev_k[c], 0));
			if (rec_lo) CU(cudaMemcpyAsync(rec_lo + a, d_lo + a, m * 4, cudaMemcpyDeviceToHost, ix->s_out));
			if (rec_hi) CU(cudaMemcpyAsync(rec_hi + a, d_hi + a, m * 4, cudaMemcpyDeviceToHost, ix->s_out));
			if (counts) CU(cudaMemcpyAsync(counts + a, d_cnt + a, m * 4, cudaMemcpyDeviceToHost, ix->s_out));
		}
		tr.mark("enqueue");
		settle_t6(ix, x, y, counts, flag ? ix->bflag.as<uint32_t>() : nullptr, ix->d_status, ix->s_out);
		tr.mark("h2d+kernel+d2h");
	} catch (const std::invalid_argument& e) { cudaStreamSynchronize(ix->s_out); return set_err(VSGPU_EINVAL, e.what());
	} catch (const std::exception& e) { cudaDeviceSynchronize(); return set_err(VSGPU_ENODEVICE, e.what()); }
	return VSGPU_OK;
}
}  // namespace
int vsgpu_query_t6(vsgpu_index* ix, uint64_t n, const uint64_t* x, const uint64_t* y, uint32_t* rec_lo, uint32_t* rec_hi, uint32_t* counts) { return query_t6_impl(ix, n, x, y, rec_lo, rec_hi, counts); }
int vsgpu_query_t6_u32(vsgpu_index* ix, uint64_t n, const uint32_t* x, const uint32_t* y, uint32_t* rec_lo, uint32_t* rec_hi, uint32_t* counts) { return query_t6_impl(ix, n, x, y, rec_lo, rec_hi, counts); }

// ------------------------------------------------------------------ t4
namespace {
// One pass of t4 over device-resident inputs.  The hit buffer is sized from a guess (4 codes per
// region) and, if the kernel reports an overflow, re-sized from the total it computed and the pass
// repeated — results are deterministic, so a batch pays that at most once.
void run_t4(vsgpu_index* ix, uint64_t n, const uint64_t* dx, const uint64_t* dy, const uint32_t* ds, DevBuf& offsets, DevBuf& state,
            DevBuf& hits, uint64_t& hits_cap, uint32_t* launches, cudaEvent_t* ev = nullptr, uint32_t* d_status = nullptr, bool wide_regions = false, uint32_t* spill = nullptr) {
	if (!d_status) d_status = ix->d_status;
	CU(offsets.ensure((n + 1) * 8)); CU(state.ensure(t4_state_words(n) * 8));
	if (hits_cap == 0) { hits_cap = std::max<uint64_t>((wide_regions ? 64 : 4) * n, 1024); CU(hits.ensure(hits_cap * 4)); }
	CU(cudaMemsetAsync(state.p, 0, t4_state_words(n) * 8, ix->stream));
	if (ev) CU(cudaEventRecord(ev[0], ix->stream));
	CU(launch_t4(ix->dev, n, dx, dy, ds, offsets.as<uint64_t>(), hits.as<uint32_t>(), hits_cap, state.as<uint64_t>(), d_status, wide_regions, ix->stream, nullptr, spill));
	if (ev) CU(cudaEventRecord(ev[1], ix->stream));
	if (launches) *launches = 1;
}

// Synchronise and resolve an overflow of the hit buffer by re-running with the exact size.  The
// decision is taken from the total the kernel computed, not from the (shared) status word.
// Returns the status bits left after that.
uint32_t finish_t4(vsgpu_index* ix, uint64_t n, const uint64_t* dx, const uint64_t* dy, const uint32_t* ds, DevBuf& offsets, DevBuf& state,
                   DevBuf& hits, uint64_t& hits_cap, uint32_t* d_status, bool wide_regions = false, uint64_t known_total = 0) {
	uint64_t total = known_total;
	if (!known_total) CU(cudaMemcpyAsync(&total, offsets.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, ix->stream));
	uint32_t st = read_status(ix, d_status);
	if (total > hits_cap) {
		hits_cap = total + total / 16 + 1024;
		CU(hits.ensure(hits_cap * 4));
		run_t4(ix, n, dx, dy, ds, offsets, state, hits, hits_cap, nullptr, nullptr, d_status, wide_regions);
		st = (st | read_status(ix, d_status));
	}
	return st & ~kStatusOverflow;
}
}  // namespace

namespace {
// The guessed hit capacity was too small: one exact pass over the resident 64-bit inputs on the index's stream.
void hits_overflow_rerun(vsgpu_index* ix, uint64_t n, const uint64_t* dx, const uint64_t* dy, const uint32_t* ds, uint64_t& cap, bool wide, uint64_t known_total) {
	const uint32_t st = finish_t4(ix, n, dx, dy, ds, ix->boffsets, ix->bstate, ix->bhits, cap, ix->d_status, wide, known_total);
	if (st & kStatusBadRegion) throw std::invalid_argument("region start < 1 or sample id out of range");
}
// Does this batch look like "few, wide regions" (the scan-bound end of the width sweep)?  Decided from
// a sample of the region widths and the index's walk-entry density; such batches get a warp per region.
extern "C++" template <class T>
bool expect_wide_regions(const vsgpu_index* ix, uint64_t n, const T* x, const T* y) {
	if (n == 0 || ix->sparse_walk) return false;        // (the warp-per-region kernel scans hit-map rows)
	const uint64_t step = std::max<uint64_t>(1, n / 1024);
	// the widest sampled region, in walk entries: beyond the kernel's threshold the batch is launched
	// with the warp-cooperative path compiled in
	uint64_t wmax = 0;
	for (uint64_t i = 0; i < n; i += step) if (y[i] > x[i]) wmax = std::max<uint64_t>(wmax, std::min<uint64_t>(y[i] - x[i], ix->flat.ref_length));
	// ... and only when a warp per region still fills the GPU sensibly (a thread per region is the
	// better mapping for large batches, measured in profiles/README.md)
	uint64_t max_regions = 150000;
	if (const char* e = getenv("VSGPU_WIDE_MAX_REGIONS")) max_regions = strtoull(e, nullptr, 10);
	return n <= max_regions && (double)wmax * ix->entries_per_base > (double)t4_wide_entries();
}

// Will the regions of this batch typically have more rows than k_t4p stages per region (kScratchHits)?  From a sample of the
// widths and the index's carried walk entries per base of an average sample.  Such batches run the kernel's spilling
// instance (codes beyond the staging go to a scratch and are copied into place, instead of a second walk).
extern "C++" template <class T>
bool expect_many_rows(const vsgpu_index* ix, uint64_t n, const T* x, const T* y) {
	if (n == 0) return false;
	if (const char* e = getenv("VSGPU_T4_SPILL")) return atoi(e) != 0;
	const uint64_t step = std::max<uint64_t>(1, n / 1024);
	double sum = 0; uint64_t k = 0;
	for (uint64_t i = 0; i < n; i += step, k++) if (y[i] > x[i]) sum += (double)std::min<uint64_t>(y[i] - x[i], ix->flat.ref_length);
	const double rows = k ? sum / (double)k * ix->hits_per_base : 0;          // measured (profiles/r2_spill.txt): wins from ~5 to a few dozen rows per region
	return rows > 0.6 * kScratchHits && rows < 64;
}
// the scratch of the spilling instance of k_t4p / the chunk pool of the warp-per-region kernel, allocated on first use
uint32_t* spill_of(DevBuf& b) { CU(b.ensure(t4x_spill_bytes())); return b.as<uint32_t>(); }
uint32_t* pool_of(DevBuf& b) {
	if (const char* e = getenv("VSGPU_T4W_POOL")) if (atoi(e) == 0) return nullptr;
	CU(b.ensure(t4w_pool_bytes())); return b.as<uint32_t>();
}

// copy a finished t4 answer (device offsets[n+1] + hits) into a pooled page-locked result
vsgpu_result* fetch_t4(vsgpu_index* ix, uint64_t n, const DevBuf& offsets, const DevBuf& hits, bool want_hits, Trace* tr = nullptr) {
	std::unique_ptr<vsgpu_result> r(new vsgpu_result);
	r->owner = ix; r->n = n;
	r->offsets = (uint64_t*)ix->pinned_acquire((n + 1) * 8, &r->offsets_cap);
	if (!r->offsets) throw std::runtime_error("CUDA: cannot allocate page-locked result memory");
	r->offsets[0] = 0; r->offsets[n] = 0;
	if (n) {
		CU(cudaMemcpyAsync(r->offsets, offsets.p, (n + 1) * 8, cudaMemcpyDeviceToHost, ix->stream));
		CU(cudaStreamSynchronize(ix->stream));
		if (tr) tr->mark("d2h offsets");
	}
	const uint64_t total = r->offsets[n];
	r->have_offsets = true; r->total = total;
	if (want_hits && total) {
		r->hits = (uint32_t*)ix->pinned_acquire(total * 4, &r->hits_cap);
		if (!r->hits) { ix->pinned_release(r->offsets, r->offsets_cap); throw std::runtime_error("CUDA: cannot allocate page-locked result memory"); }
		CU(cudaMemcpyAsync(r->hits, hits.p, total * 4, cudaMemcpyDeviceToHost, ix->stream));
		CU(cudaStreamSynchronize(ix->stream));
		if (tr) tr->mark("d2h hits");
	}
	return r.release();
}
}  // namespace

namespace {
// t6 outputs of a fused host-buffer call (vsgpu_query_t6t4*): as vsgpu_query_t6's; rec_hi nullable
struct T6Host { uint32_t* rec_lo; uint32_t* rec_hi; uint32_t* counts; };

// t4 (optionally with t6 over the same regions, from the same kernel) with host buffers.  What crosses PCIe per
// region: x, y (4 or 8 bytes each) and the sample id in; the row count (4 bytes), the hit codes and, when fused,
// the t6 slice start and row count out.
extern "C++" template <class T>
int query_t4_impl(vsgpu_index* ix, uint64_t n, const T* x, const T* y, const uint32_t* sample_ids, vsgpu_result** out, const T6Host* t6 = nullptr) {
	constexpr bool k32 = sizeof(T) == 4;
	if (!ix || !out || (n && (!x || !y || !sample_ids)) || (t6 && n && (!t6->rec_lo || !t6->counts))) return set_err(VSGPU_EINVAL, "vsgpu_query_t4: null argument");
	*out = nullptr;
	if (int rc = check_device(ix)) return rc;
	std::lock_guard<std::mutex> g(ix->mu);
	Trace tr(t6 ? "t6t4" : "t4");
	std::unique_ptr<vsgpu_result, void (*)(vsgpu_result*)> r(nullptr, vsgpu_result_free);
	try {
		if (n == 0) { *out = fetch_t4(ix, 0, ix->boffsets, ix->bhits, true); return VSGPU_OK; }
		const bool wide = expect_wide_regions(ix, n, x, y);
		const bool direct = t4x_supported(wide);               // k_t4p: reads the host's coordinate width, writes counts, fuses t6
		const bool flag6 = t6 && t6_special(ix);
		uint32_t* const spill = wide ? pool_of(ix->bspill) : expect_many_rows(ix, n, x, y) ? spill_of(ix->bspill) : nullptr;
		uint64_t per = 0;
		const int chunks = wide ? (per = n, 1) : plan_chunks(n, &per);
		const uint64_t state_words = t4_state_words(per);
		CU(ix->bs.ensure(n * 4));
		if (k32) { CU(ix->bx32.ensure(n * 4)); CU(ix->by32.ensure(n * 4)); }
		if (!k32 || !direct) { CU(ix->bx.ensure(n * 8)); CU(ix->by.ensure(n * 8)); }
		CU(ix->boffsets.ensure((n + 1) * 8)); CU(ix->bstate.ensure(state_words * chunks * 8));
		if (direct) CU(ix->bcnt.ensure(n * 4));
		if (t6) CU(ix->bout.ensure(n * 12));
		if (flag6) CU(ix->bflag.ensure(n * 4));
		uint64_t cap = ix->bhits.cap / 4;
		if (cap == 0) { cap = std::max<uint64_t>((wide ? 64 : 4) * n, 1024); CU(ix->bhits.ensure(cap * 4)); cap = ix->bhits.cap / 4; }
		uint64_t* dx = ix->bx.as<uint64_t>(); uint64_t* dy = ix->by.as<uint64_t>(); uint32_t* ds = ix->bs.as<uint32_t>();
		T* sx = k32 ? ix->bx32.as<T>() : (T*)dx; T* sy = k32 ? ix->by32.as<T>() : (T*)dy;      // where the host arrays land
		uint64_t* d_off = ix->boffsets.as<uint64_t>();
		uint32_t* d_cnt4 = direct ? ix->bcnt.as<uint32_t>() : nullptr;
		uint32_t* d_lo = ix->bout.as<uint32_t>(); uint32_t* d_hi = d_lo + n; uint32_t* d_cnt6 = d_hi + n;
		r.reset(new vsgpu_result);
		r->owner = ix; r->n = n;
		if (direct) r->counts = (uint32_t*)ix->pinned_acquire(n * 4, &r->counts_cap);
		else r->offsets = (uint64_t*)ix->pinned_acquire((n + 1) * 8, &r->offsets_cap);
		// page-locked room for the hit codes is a guess too (the device buffer may be far larger than this
		// batch needs); a batch that outgrows it gets an exact buffer and one copy at the end
		const uint64_t host_cap = std::min<uint64_t>(cap, std::max<uint64_t>(8 * n, 1u << 18));
		r->hits = (uint32_t*)ix->pinned_acquire(host_cap * 4, &r->hits_cap);
		if ((!r->offsets && !r->counts) || !r->hits) throw std::runtime_error("CUDA: cannot allocate page-locked result memory");
		// inputs on s_in, kernels on s_k (chunk c continues the offsets of chunk c-1), results on s_out as
		// soon as their chunk is done
		const int ks = in_streams();
		for (int c = 0; c < chunks; c++) {
			const uint64_t a = c * per, m = std::min<uint64_t>(n, a + per) - a;
			CU(cudaMemcpyAsync(sx + a, x + a, m * sizeof(T), cudaMemcpyHostToDevice, ix->s_in[0]));
			CU(cudaMemcpyAsync(sy + a, y + a, m * sizeof(T), cudaMemcpyHostToDevice, ix->s_in[1 % ks]));
			CU(cudaMemcpyAsync(ds + a, sample_ids + a, m * 4, cudaMemcpyHostToDevice, ix->s_in[2 % ks]));
			for (int k = 0; k < ks; k++) CU(cudaEventRecord(ix->ev_in[c][k], ix->s_in[k]));
		}
		CU(cudaMemsetAsync(ix->bstate.p, 0, state_words * chunks * 8, ix->s_k));
		for (int c = 0; c < chunks; c++) {
			const uint64_t a = c * per, m = std::min<uint64_t>(n, a + per) - a;
			for (int k = 0; k < ks; k++) CU(cudaStreamWaitEvent(ix->s_k, ix->ev_in[c][k], 0));
			if (direct) {
				const T6Out f6{d_lo + a, t6 && t6->rec_hi ? d_hi + a : nullptr, d_cnt6 + a, flag6 ? ix->bflag.as<uint32_t>() : nullptr, (uint32_t)a};
				const T4Launch L{m, sx + a, sy + a, k32, ds + a, d_off + a, d_cnt4 + a, ix->bhits.as<uint32_t>(), cap, ix->bstate.as<uint64_t>() + c * state_words, ix->d_status,
				                 c ? d_off + a : nullptr, t6 ? &f6 : nullptr, spill};
				CU(launch_t4x(ix->dev, L, ix->s_k));
			} else {
				if (k32) CU(launch_widen(m, (const uint32_t*)sx + a, (const uint32_t*)sy + a, dx + a, dy + a, ix->s_k));
				if (t6) CU(launch_t6(ix->dev, m, dx + a, dy + a, d_lo + a, d_hi + a, d_cnt6 + a, flag6 ? ix->bflag.as<uint32_t>() : nullptr, (uint32_t)a, ix->d_status, ix->s_k));
				CU(launch_t4(ix->dev, m, dx + a, dy + a, ds + a, d_off + a, ix->bhits.as<uint32_t>(), cap, ix->bstate.as<uint64_t>() + c * state_words, ix->d_status, wide,
				             ix->s_k, c ? d_off + a : nullptr, spill));
			}
			CU(cudaEventRecord(ix->ev_k[c], ix->s_k));
		}
		tr.mark("enqueue");
		uint64_t done = 0; bool overflow = false, host_small = false;
		for (int c = 0; c < chunks; c++) {
			const uint64_t a = c * per, m = std::min<uint64_t>(n, a + per) - a;
			CU(cudaStreamWaitEvent(ix->s_out, ix->ev_k[c], 0));
			CU(cudaMemcpyAsync(ix->pin_small + c, d_off + a + m, 8, cudaMemcpyDeviceToHost, ix->s_out));
			CU(cudaEventRecord(ix->ev_out[c], ix->s_out));
			if (direct) CU(cudaMemcpyAsync(r->counts + a, d_cnt4 + a, m * 4, cudaMemcpyDeviceToHost, ix->s_out));
			else CU(cudaMemcpyAsync(r->offsets + a, d_off + a, (m + (c == chunks - 1 ? 1 : 0)) * 8, cudaMemcpyDeviceToHost, ix->s_out));
			if (t6) {
				CU(cudaMemcpyAsync(t6->rec_lo + a, d_lo + a, m * 4, cudaMemcpyDeviceToHost, ix->s_out));
				if (t6->rec_hi) CU(cudaMemcpyAsync(t6->rec_hi + a, d_hi + a, m * 4, cudaMemcpyDeviceToHost, ix->s_out));
				CU(cudaMemcpyAsync(t6->counts + a, d_cnt6 + a, m * 4, cudaMemcpyDeviceToHost, ix->s_out));
			}
			CU(cudaEventSynchronize(ix->ev_out[c]));                   // the running total after chunk c: its hits are final
			const uint64_t total = ix->pin_small[c];
			if (total > cap) { overflow = true; continue; }
			if (total > host_cap) { host_small = true; continue; }
			if (!overflow && !host_small && total > done) CU(cudaMemcpyAsync(r->hits + done, ix->bhits.as<uint32_t>() + done, (total - done) * 4, cudaMemcpyDeviceToHost, ix->s_out));
			done = total;
		}
		// status words, bad regions, and (fused) the t6 counts of the regions the kernel flagged for the literal rule
		uint32_t st = 0;
		if (t6) {
			uint32_t nflag = 0;
			st = read_status(ix, ix->d_status, &nflag, ix->s_out);
			if ((st & kStatusBadRegion) == 0 && nflag) {
				std::vector<uint32_t> flagged(nflag), tmp;
				CU(cudaMemcpyAsync(flagged.data(), ix->bflag.p, (size_t)nflag * 4, cudaMemcpyDeviceToHost, ix->s_out));
				CU(cudaStreamSynchronize(ix->s_out));
				for (uint32_t i : flagged) { t6_literal(ix, x[i], y[i], tmp); t6->counts[i] = (uint32_t)tmp.size(); }
			}
		} else st = read_status(ix, ix->d_status, nullptr, ix->s_out);
		tr.mark("pipeline");
		if (st & kStatusBadRegion) return set_err(VSGPU_EINVAL, "region start < 1 or sample id out of range");
		if (overflow) {
			// the guess for the hit buffer was too small: size it from the total the kernels computed and
			// run t4 again in one piece (inputs are resident; results are deterministic; the t6 answers stand)
			uint64_t c2 = ix->bhits.cap / 4;
			CU(cudaStreamSynchronize(ix->s_k)); CU(cudaStreamSynchronize(ix->s_out));
			if (k32 && direct) {                                 // the exact pass runs on 64-bit coordinates
				CU(ix->bx.ensure(n * 8)); CU(ix->by.ensure(n * 8));
				dx = ix->bx.as<uint64_t>(); dy = ix->by.as<uint64_t>();
				CU(launch_widen(n, (const uint32_t*)sx, (const uint32_t*)sy, dx, dy, ix->stream));
			}
			hits_overflow_rerun(ix, n, dx, dy, ds, c2, wide, ix->pin_small[chunks - 1]);
			*out = fetch_t4(ix, n, ix->boffsets, ix->bhits, true, &tr);
			return VSGPU_OK;
		}
		const uint64_t total = ix->pin_small[chunks - 1];
		if (host_small) {
			ix->pinned_release(r->hits, r->hits_cap);
			r->hits = (uint32_t*)ix->pinned_acquire(total * 4, &r->hits_cap);
			if (!r->hits) throw std::runtime_error("CUDA: cannot allocate page-locked result memory");
			CU(cudaMemcpyAsync(r->hits, ix->bhits.p, total * 4, cudaMemcpyDeviceToHost, ix->s_out));
			CU(cudaStreamSynchronize(ix->s_out));
		}
		r->total = total; r->have_counts = direct; r->have_offsets = !direct;
		*out = r.release();
	} catch (const std::exception& e) { cudaDeviceSynchronize(); return set_err(VSGPU_ENODEVICE, e.what()); }
	return VSGPU_OK;
}
}  // namespace
int vsgpu_query_t4(vsgpu_index* ix, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample_ids, vsgpu_result** out) { return query_t4_impl(ix, n, x, y, sample_ids, out); }
int vsgpu_query_t4_u32(vsgpu_index* ix, uint64_t n, const uint32_t* x, const uint32_t* y, const uint32_t* sample_ids, vsgpu_result** out) { return query_t4_impl(ix, n, x, y, sample_ids, out); }
int vsgpu_query_t6t4(vsgpu_index* ix, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample_ids, uint32_t* rec_lo, uint32_t* rec_hi, uint32_t* counts6, vsgpu_result** out) {
	const T6Host t6{rec_lo, rec_hi, counts6};
	return query_t4_impl(ix, n, x, y, sample_ids, out, &t6);
}
int vsgpu_query_t6t4_u32(vsgpu_index* ix, uint64_t n, const uint32_t* x, const uint32_t* y, const uint32_t* sample_ids, uint32_t* rec_lo, uint32_t* rec_hi, uint32_t* counts6, vsgpu_result** out) {
	const T6Host t6{rec_lo, rec_hi, counts6};
	return query_t4_impl(ix, n, x, y, sample_ids, out, &t6);
}
uint64_t vsgpu_result_num_queries(const vsgpu_result* r) { return r ? r->n : 0; }
const uint64_t* vsgpu_result_offsets(const vsgpu_result* cr) {
	vsgpu_result* r = const_cast<vsgpu_result*>(cr);
	if (!r) return nullptr;
	std::lock_guard<std::mutex> g(r->lazy_mu);
	if (!r->have_offsets && r->have_counts) {              // exclusive prefix sums of the counts
		if (!r->offsets) r->offsets = (uint64_t*)r->owner->pinned_acquire((r->n + 1) * 8, &r->offsets_cap);
		if (!r->offsets) return nullptr;
		uint64_t acc = 0;
		for (uint64_t i = 0; i < r->n; i++) { r->offsets[i] = acc; acc += r->counts[i]; }
		r->offsets[r->n] = acc;
		r->have_offsets = true;
	}
	return r->offsets;
}
const uint32_t* vsgpu_result_counts(const vsgpu_result* cr) {
	vsgpu_result* r = const_cast<vsgpu_result*>(cr);
	if (!r) return nullptr;
	std::lock_guard<std::mutex> g(r->lazy_mu);
	if (!r->have_counts && r->offsets) {
		if (!r->counts) r->counts = (uint32_t*)r->owner->pinned_acquire(std::max<uint64_t>(r->n, 1) * 4, &r->counts_cap);
		if (!r->counts) return nullptr;
		for (uint64_t i = 0; i < r->n; i++) r->counts[i] = (uint32_t)(r->offsets[i + 1] - r->offsets[i]);
		r->have_counts = true;
	}
	return r->counts;
}
uint64_t vsgpu_result_total(const vsgpu_result* r) { return r ? r->total : 0; }
const uint32_t* vsgpu_result_hits(const vsgpu_result* r) { return r ? r->hits : nullptr; }
void vsgpu_result_free(vsgpu_result* r) {
	if (!r) return;
	if (r->owner) { r->owner->pinned_release(r->offsets, r->offsets_cap); r->owner->pinned_release(r->hits, r->hits_cap); r->owner->pinned_release(r->status, r->status_cap); r->owner->pinned_release(r->counts, r->counts_cap); }
	delete r;
}

// ------------------------------------------------------------------ t1
int vsgpu_query_t1(vsgpu_index* ix, uint64_t n, const uint64_t* pos, uint32_t* rec_lo, uint32_t* rec_hi) {
	if (!ix || (n && (!pos || !rec_lo || !rec_hi))) return set_err(VSGPU_EINVAL, "vsgpu_query_t1: null argument");
	if (n == 0) return VSGPU_OK;
	if (int rc = check_device(ix)) return rc;
	std::lock_guard<std::mutex> g(ix->mu);
	try {
		CU(ix->bx.ensure(n * 8)); CU(ix->bout.ensure(n * 8));
		CU(cudaMemcpyAsync(ix->bx.p, pos, n * 8, cudaMemcpyHostToDevice, ix->stream));
		uint32_t* d_lo = ix->bout.as<uint32_t>(); uint32_t* d_hi = d_lo + n;
		CU(launch_t1(ix->dev, n, ix->bx.as<uint64_t>(), d_lo, d_hi, ix->d_status, ix->stream));
		CU(cudaMemcpyAsync(rec_lo, d_lo, n * 4, cudaMemcpyDeviceToHost, ix->stream));
		CU(cudaMemcpyAsync(rec_hi, d_hi, n * 4, cudaMemcpyDeviceToHost, ix->stream));
		uint32_t st = read_status(ix);
		if (st & kStatusBadRegion) return set_err(VSGPU_EINVAL, "Can't find node corresponding to pos 0");
	} catch (const std::exception& e) { return set_err(VSGPU_ENODEVICE, e.what()); }
	return VSGPU_OK;
}

// ------------------------------------------------------------------ t7
int vsgpu_query_t7(vsgpu_index* ix, uint64_t n, const uint64_t* pos, const char* const* refs, const char* const* alts, uint32_t* rec) {
	if (!ix || (n && (!pos || !refs || !alts || !rec))) return set_err(VSGPU_EINVAL, "vsgpu_query_t7: null argument");
	if (int rc = check_device(ix)) return rc;
	std::lock_guard<std::mutex> g(ix->mu);
	try {
		std::vector<uint64_t> qh(n);
		parallel_for(n, [&](uint64_t a, uint64_t b) { for (uint64_t i = a; i < b; i++) qh[i] = hash_query(refs[i], alts[i]); });
		CU(ix->bx.ensure(n * 8)); CU(ix->bhash.ensure(n * 8)); CU(ix->brec.ensure(n * 4));
		CU(cudaMemcpyAsync(ix->bx.p, pos, n * 8, cudaMemcpyHostToDevice, ix->stream));
		CU(cudaMemcpyAsync(ix->bhash.p, qh.data(), n * 8, cudaMemcpyHostToDevice, ix->stream));
		CU(launch_t7(ix->dev, n, ix->bx.as<uint64_t>(), ix->bhash.as<uint64_t>(), ix->brec.as<uint32_t>(), ix->d_status, ix->stream));
		CU(cudaMemcpyAsync(rec, ix->brec.p, n * 4, cudaMemcpyDeviceToHost, ix->stream));
		uint32_t st = read_status(ix);
		if (st & kStatusBadRegion) return set_err(VSGPU_EINVAL, "Can't find node corresponding to pos 0");
		parallel_for(n, [&](uint64_t a, uint64_t b) { for (uint64_t i = a; i < b; i++) rec[i] = t7_confirm(ix, pos[i], refs[i], alts[i], rec[i]); });
	} catch (const std::exception& e) { return set_err(VSGPU_ENODEVICE, e.what()); }
	return VSGPU_OK;
}

// ------------------------------------------------------------------ materialisation
int vsgpu_rows_t6(const vsgpu_index* ix, uint32_t lo, uint32_t hi, int with_samples, char** text, uint64_t* nrows) {
	if (!ix || !text || ((lo != VSGPU_NONE || hi != VSGPU_NONE) && (lo > hi || hi > ix->flat.R))) return set_err(VSGPU_EINVAL, "vsgpu_rows_t6: bad record slice");
	std::string s; uint64_t cnt = 0;
	rows_t6(ix, lo, hi, with_samples != 0, s, cnt);
	if (nrows) *nrows = cnt;
	*text = dup_text(s);
	return *text ? VSGPU_OK : set_err(VSGPU_ENOMEM, "out of memory");
}

int vsgpu_rows_t1(const vsgpu_index* ix, uint32_t lo, uint32_t hi, int with_samples, char** text, uint64_t* nrows) {
	if (!ix || !text || (lo != VSGPU_NONE && (lo > hi || hi > ix->flat.R))) return set_err(VSGPU_EINVAL, "vsgpu_rows_t1: bad record range");
	std::string s; uint64_t cnt = 0;
	rows_t1(ix, lo, hi, with_samples != 0, s, cnt);
	if (nrows) *nrows = cnt;
	*text = dup_text(s);
	return *text ? VSGPU_OK : set_err(VSGPU_ENOMEM, "out of memory");
}

int vsgpu_digest_t1(const vsgpu_index* ix, uint64_t n, const uint32_t* lo, const uint32_t* hi, int with_samples, uint64_t* counts, uint64_t* digests) {
	if (!ix || (n && (!lo || !hi))) return set_err(VSGPU_EINVAL, "vsgpu_digest_t1: null argument");
	digests_t1(ix, n, lo, hi, with_samples != 0, counts, digests);
	return VSGPU_OK;
}

int vsgpu_rows_t4(const vsgpu_index* ix, const uint32_t* hits, uint64_t nhits, int with_samples, char** text) {
	if (!ix || !text || (nhits && !hits)) return set_err(VSGPU_EINVAL, "vsgpu_rows_t4: null argument");
	std::string s;
	for (uint64_t i = 0; i < nhits; i++) {
		if ((hits[i] & VSGPU_HIT_ENTRY_MASK) >= ix->flat.cent.size()) return set_err(VSGPU_EINVAL, "vsgpu_rows_t4: hit code out of range");
		t4_row(ix, hits[i], with_samples != 0, s);
	}
	*text = dup_text(s);
	return *text ? VSGPU_OK : set_err(VSGPU_ENOMEM, "out of memory");
}

int vsgpu_rows_t7(const vsgpu_index* ix, uint32_t rec, char** text, uint64_t* ncarriers) {
	if (!ix || !text) return set_err(VSGPU_EINVAL, "vsgpu_rows_t7: null argument");
	if (rec != VSGPU_NONE && rec >= ix->flat.R) return set_err(VSGPU_EINVAL, "vsgpu_rows_t7: record id out of range");
	std::string s;
	uint64_t cnt = t7_carriers(ix, rec, &s, nullptr);
	if (ncarriers) *ncarriers = cnt;
	*text = dup_text(s);
	return *text ? VSGPU_OK : set_err(VSGPU_ENOMEM, "out of memory");
}

int vsgpu_digest_t6(const vsgpu_index* ix, uint64_t n, const uint32_t* lo, const uint32_t* hi, int with_samples, uint64_t* digests) {
	if (!ix || (n && (!lo || !hi || !digests))) return set_err(VSGPU_EINVAL, "vsgpu_digest_t6: null argument");
	bool bad = false;
	digests_t6(ix, n, lo, hi, with_samples != 0, digests, &bad);
	return bad ? set_err(VSGPU_EINVAL, "vsgpu_digest_t6: bad record slice") : VSGPU_OK;
}

namespace {
// hit codes and offsets handed back by a caller: every code names a walk entry, offsets do not decrease
bool csr_ok(const vsgpu_index* ix, uint64_t n, const uint64_t* offsets, const uint32_t* hits) {
	const uint64_t E = ix->flat.cent.size();
	for (uint64_t i = 0; i < n; i++) if (offsets[i + 1] < offsets[i]) return false;
	if (n && offsets[n] > offsets[0] && !hits) return false;
	std::atomic<bool> ok{true};
	parallel_for(n ? offsets[n] - offsets[0] : 0, [&](uint64_t a, uint64_t b) { for (uint64_t j = offsets[0] + a; j < offsets[0] + b; j++) if ((hits[j] & VSGPU_HIT_ENTRY_MASK) >= E) { ok = false; return; } });
	return ok;
}
}  // namespace
int vsgpu_digest_t4(const vsgpu_index* ix, uint64_t n, const uint64_t* offsets, const uint32_t* hits, int with_samples, uint64_t* digests) {
	if (!ix || (n && (!offsets || !digests))) return set_err(VSGPU_EINVAL, "vsgpu_digest_t4: null argument");
	if (!csr_ok(ix, n, offsets, hits)) return set_err(VSGPU_EINVAL, "vsgpu_digest_t4: offsets decrease or a hit code names no walk entry");
	digests_t4(ix, n, offsets, hits, with_samples != 0, digests);
	return VSGPU_OK;
}

int vsgpu_digest_t7(const vsgpu_index* ix, uint64_t n, const uint32_t* rec, uint64_t* ncarriers, uint64_t* digests) {
	if (!ix || (n && !rec)) return set_err(VSGPU_EINVAL, "vsgpu_digest_t7: null argument");
	digests_t7(ix, n, rec, ncarriers, digests);
	return VSGPU_OK;
}

// ------------------------------------------------------------------ t6 rows rendered on the device
namespace {
// Row lengths and per-record lookups for k_render.  Everything is a function of the index alone.
void ensure_render_tables(vsgpu_index* ix) {
	if (ix->render_ready) return;
	const FlatIndex& f = ix->flat; const SerData& s = ix->ser;
	RenderTables& rt = ix->render;
	std::vector<uint32_t> name_off(s.num_samples + 1, 0);
	std::string chars;
	for (uint32_t i = 0; i < s.num_samples; i++) { chars += s.sample_names[i]; name_off[i + 1] = (uint32_t)chars.size(); }
	// bytes of the carrier list of every class: sum over non-ref members of len(name) + len("(g|g) ")
	std::vector<uint64_t> set_bytes; std::vector<uint32_t> set_pop;
	if (f.class_mode) {
		set_bytes.assign(f.num_sets, 0); set_pop.assign(f.num_sets, 0);
		std::atomic<bool> beyond{false};
		parallel_for(f.num_sets, [&](uint64_t a, uint64_t b) {
			for (uint64_t c = a; c < b; c++) {
				uint64_t bytes = 0; uint32_t pop = 0;
				for (uint32_t w = 0; w < f.words_per_set; w++)
					for (uint64_t m = f.bitmap[c * f.words_per_set + w]; m; m &= m - 1) {
						const uint32_t id = w * 64 + (uint32_t)__builtin_ctzll(m); pop++;
						if (id >= s.num_samples) { beyond = true; continue; }
						if (id) bytes += name_off[id + 1] - name_off[id] + 6;
					}
				set_bytes[c] = bytes; set_pop[c] = pop;
			}
		});
		if (beyond) throw std::runtime_error("a sample class names a sample beyond sampleid_map.lst");
	}
	std::vector<uint4> rec_seq(f.R), rec_car(f.R);
	std::vector<uint64_t> tp0(f.R + 1, 0), tp1(f.R + 1, 0);
	for (uint32_t r = 0; r < f.R; r++) {
		const uint32_t v = f.rec_vertex[r], rv = f.rec_refv[r], av = f.rec_altv[r];
		rec_seq[r] = make_uint4(rv == kNone ? 0 : s.v_offset[rv], rv == kNone ? 0 : s.v_length[rv], av == kNone ? 0 : s.v_offset[av], av == kNone ? 0 : s.v_length[av]);
		const uint64_t sb = s.v_sinfo_begin[v], sc = s.v_sinfo_begin[v + 1] - sb;
		if (sc >= (1u << 28)) throw std::runtime_error("vertex " + std::to_string(v) + " has too many s_info entries to render");
		const uint32_t set = f.class_mode ? s.v_class[v] : 0;
		rec_car[r] = make_uint4(set, (uint32_t)sc | ((uint32_t)f.rec_flags[r] << 28), (uint32_t)sb, (uint32_t)(sb >> 32));
		uint32_t digits = 1; for (uint32_t q = f.rec_pos[r]; q >= 10; q /= 10) digits++;
		const uint64_t hdr = digits + 1 + rec_seq[r].y + 1 + rec_seq[r].w + 1 + 1;
		uint64_t car = 0;
		if (!(f.rec_flags[r] & 4)) {
			if (f.class_mode) {
				if (set_pop[set] != sc) throw std::runtime_error("vertex " + std::to_string(v) + ": s_info entries differ from the members of its sample class");
				car = set_bytes[set];
			} else for (uint64_t i = sb; i < sb + sc; i++) {
				const uint32_t id = s.s_sample_id[i];
				if (id >= s.num_samples) throw std::runtime_error("vertex " + std::to_string(v) + " names a sample beyond sampleid_map.lst");
				if (id) car += name_off[id + 1] - name_off[id] + 6;
			}
		}
		tp0[r + 1] = tp0[r] + hdr; tp1[r + 1] = tp1[r] + hdr + car;
	}
	ix->set_text_bytes = set_bytes; ix->set_pop = set_pop;
	rt.rec_seq = upload(ix, rec_seq); rt.rec_car = upload(ix, rec_car);
	rt.text_prefix[0] = upload(ix, tp0); rt.text_prefix[1] = upload(ix, tp1);
	rt.seq = upload(ix, s.seq); rt.s_flags = upload(ix, s.s_flags); rt.s_sample_id = upload(ix, s.s_sample_id);
	rt.name_off = upload(ix, name_off);
	std::vector<uint4> item16(s.num_samples, make_uint4(0, 0, 0, 0));
	for (uint32_t i = 0; i < s.num_samples; i++) {
		const std::string item = s.sample_names[i] + "(0|0) ";
		if (item.size() > 15) continue;
		uint8_t b[16] = {0};
		memcpy(b, item.data(), item.size()); b[15] = (uint8_t)item.size();
		memcpy(&item16[i], b, 16);
	}
	rt.item16 = upload(ix, item16);
	std::vector<char> cv(chars.begin(), chars.end());
	rt.name_chars = upload(ix, cv);
	for (auto& e : ix->ev_render) CU(cudaEventCreate(&e));
	ix->render_ready = true;
}
}  // namespace

namespace {
// Per walk entry and variant, the row t4_row would print, as numbers k_render_hits can use (kernels.cuh: HitTables).
void ensure_hit_tables(vsgpu_index* ix) {
	ensure_render_tables(ix);
	if (ix->hit_tables_ready) return;
	const FlatIndex& f = ix->flat; const SerData& s = ix->ser;
	const size_t E = f.cent.size();
	std::vector<uint32_t> name_len(s.num_samples);
	for (uint32_t i = 0; i < s.num_samples; i++) name_len[i] = (uint32_t)s.sample_names[i].size();
	for (int v = 0; v < 3; v++) {
		std::vector<uint32_t> pos(E, 0), len0(E, 0), len1(E, 0);
		std::vector<uint4> seq(E, make_uint4(0, 0, 0, 0)), car(E, make_uint4(0, 4u << 28, 0, 0));
		parallel_for(E, [&](uint64_t a, uint64_t b) {
			for (uint64_t c = a; c < b; c++) {
				const CEntry& e = f.cent[c];
				if (e.tgt & kEntMarker) continue;
				if (v == 2 && !((e.tgt & kEntAlt) && (e.tgt & kEntTgtCarriers) && (e.tgt & kEntTgtMask) != kEntTgtNone)) continue;   // only such entries produce REJOIN codes
				const uint32_t code = (uint32_t)c | (v == 1 ? VSGPU_HIT_START : v == 2 ? VSGPU_HIT_REJOIN : 0);
				uint64_t p; uint32_t refv, altv, u; bool ins = false;
				hit_row_parts(ix, code, kNone, p, refv, altv, u, &ins);
				pos[c] = (uint32_t)p;
				seq[c] = make_uint4(refv == kNone ? 0 : s.v_offset[refv], refv == kNone ? 0 : s.v_length[refv], altv == kNone ? 0 : s.v_offset[altv], altv == kNone ? 0 : s.v_length[altv]);
				const uint64_t sb = s.v_sinfo_begin[u], sc = s.v_sinfo_begin[u + 1] - sb;
				const uint32_t set = f.class_mode ? s.v_class[u] : 0;
				uint64_t carb = 0;
				if (f.class_mode) { if (ix->set_pop[set] != sc) throw std::runtime_error("vertex " + std::to_string(u) + ": s_info entries differ from the members of its sample class"); carb = ix->set_text_bytes[set]; }
				else for (uint64_t i = sb; i < sb + sc; i++) { const uint32_t id = s.s_sample_id[i]; if (id >= s.num_samples) throw std::runtime_error("vertex names a sample beyond sampleid_map.lst"); if (id) carb += name_len[id] + 6; }
				if (sc >= (1u << 28)) throw std::runtime_error("vertex " + std::to_string(u) + " has too many s_info entries to render");
				car[c] = make_uint4(set, (uint32_t)sc | (ins ? 1u << 31 : 0), (uint32_t)sb, (uint32_t)(sb >> 32));     // bit 31: an insertion row (t5 prints ref_pos)
				uint32_t digits = 1; for (uint64_t q = p; q >= 10; q /= 10) digits++;
				const uint64_t hdr = digits + 1 + seq[c].y + 1 + seq[c].w + 1 + 1;
				if (hdr + carb > 0xFFFFFFFFull) throw std::runtime_error("a t4 row is longer than 4 GiB");
				len0[c] = (uint32_t)hdr; len1[c] = (uint32_t)(hdr + carb);
			}
		});
		ix->hit_tables.pos[v] = upload(ix, pos); ix->hit_tables.seq[v] = upload(ix, seq); ix->hit_tables.car[v] = upload(ix, car);
		ix->hit_tables.len[0][v] = upload(ix, len0); ix->hit_tables.len[1][v] = upload(ix, len1);
	}
	ix->hit_tables_ready = true;
}
}  // namespace

namespace {
// Rows of a finished t4 / t5 answer (CSR in boffsets / bhits, on the index's stream) as text into t: row lengths -> text offsets of
// every row and region, then the rows in chunks whose copies overlap the rendering of the next.  t5_samples: device sample ids (t5 rows).
void render_hit_rows(vsgpu_index* ix, vsgpu_text* t, uint64_t n, int with_samples, const uint32_t* t5_samples) {
	cudaStream_t st = ix->stream;
	Trace tr("hit rows");
	uint64_t nh = 0;
	CU(cudaMemcpyAsync(&nh, ix->boffsets.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	// row lengths -> text offsets of every row and of every region
	CU(ix->brow_off.ensure((nh + 1) * 8)); CU(ix->bbyte_off.ensure((n + 1) * 8)); CU(ix->bscratch.ensure(((nh + 1023) / 1024 + 2) * 16));
	CU(cudaEventRecord(ix->ev_render[0], st));
	const uint32_t* pos5 = nullptr;
	if (t5_samples) {                                  // t5: the position column of every row, from the row's sample
		CU(ix->bpos5.ensure(std::max<uint64_t>(nh, 1) * 4));
		CU(launch_t5_row_pos(ix->dev, ix->render, ix->hit_tables, ix->d_gsidx, n, ix->boffsets.as<uint64_t>(), t5_samples, ix->bhits.as<uint32_t>(), nh, ix->bpos5.as<uint32_t>(), st));
		pos5 = ix->bpos5.as<uint32_t>();
	}
	CU(launch_hit_offsets(ix->hit_tables, ix->bhits.as<uint32_t>(), nh, with_samples, n, ix->boffsets.as<uint64_t>(), ix->brow_off.as<uint64_t>(), ix->bbyte_off.as<uint64_t>(),
	                      ix->bscratch.as<uint64_t>(), st, pos5));
	CU(cudaMemcpyAsync(t->offsets, ix->bbyte_off.p, (n + 1) * 8, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	const uint64_t total = t->offsets[n];
	tr.mark("row offsets");
	uint64_t max_bytes = 2ull << 30;
	if (const char* e = getenv("VSGPU_RENDER_MAX_BYTES")) max_bytes = strtoull(e, nullptr, 10);
	if (total > max_bytes) throw std::invalid_argument("the rows of this batch take " + std::to_string(total) + " bytes (limit VSGPU_RENDER_MAX_BYTES = " + std::to_string(max_bytes) + "); split the batch");
	t->nrows = nh; t->nbytes = total;
	t->bytes = (char*)ix->pinned_acquire(total + 1, &t->bytes_cap);
	if (!t->bytes) throw std::runtime_error("CUDA: cannot allocate page-locked result memory");
	tr.mark("page-locked text buffer");
	if (total) {
		CU(ix->btext.ensure(total));
		tr.mark("device text buffer");
		// rows in chunks of about 32 MB of text: the copy of one chunk overlaps the rendering of the next
		uint64_t chunk_bytes = 32ull << 20;
		if (const char* e = getenv("VSGPU_RENDER_CHUNK_BYTES")) chunk_bytes = std::max<uint64_t>(1, strtoull(e, nullptr, 10));
		const int chunks = (int)std::min<uint64_t>(vsgpu_index::kMaxChunks, (total + chunk_bytes - 1) / chunk_bytes);
		// region boundaries as chunk boundaries (their row / byte offsets are on the host already)
		std::vector<uint64_t> roff(n + 1);
		CU(cudaMemcpyAsync(roff.data(), ix->boffsets.p, (n + 1) * 8, cudaMemcpyDeviceToHost, st));
		CU(cudaStreamSynchronize(st));
		uint64_t r0 = 0;
		for (int c = 0; c < chunks && r0 < n; c++) {
			uint64_t r1 = n;
			if (c + 1 < chunks) { const uint64_t target = total / chunks * (c + 1); r1 = std::lower_bound(t->offsets + r0 + 1, t->offsets + n, target) - t->offsets; }
			if (r1 <= r0) continue;
			CU(launch_render_hits(ix->dev, ix->render, ix->hit_tables, ix->bhits.as<uint32_t>(), with_samples, ix->brow_off.as<uint64_t>(), roff[r0], roff[r1], ix->btext.as<char>(), st, pos5));
			CU(cudaEventRecord(ix->ev_k[c], st));
			CU(cudaStreamWaitEvent(ix->s_out, ix->ev_k[c], 0));
			if (t->offsets[r1] > t->offsets[r0]) CU(cudaMemcpyAsync(t->bytes + t->offsets[r0], ix->btext.as<char>() + t->offsets[r0], t->offsets[r1] - t->offsets[r0], cudaMemcpyDeviceToHost, ix->s_out));
			r0 = r1;
		}
		CU(cudaEventRecord(ix->ev_render[1], st));
		CU(cudaStreamSynchronize(ix->s_out));
		CU(cudaStreamSynchronize(st));
	} else { CU(cudaEventRecord(ix->ev_render[1], st)); CU(cudaStreamSynchronize(st)); }
	t->bytes[total] = 0;
	tr.mark("render + copy");
	CU(cudaEventElapsedTime(&t->kernel_ms, ix->ev_render[0], ix->ev_render[1]));
}
}  // namespace

// get_sample_var_in_ref(vg, idx, x, y, sample, print = true, outfile) for a batch: t4 on the device, then its rows as text
int vsgpu_render_t4(vsgpu_index* ix, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample_ids, int with_samples, vsgpu_text** out) {
	if (!ix || !out || (n && (!x || !y || !sample_ids))) return set_err(VSGPU_EINVAL, "vsgpu_render_t4: null argument");
	*out = nullptr;
	if (int rc = check_device(ix)) return rc;
	std::lock_guard<std::mutex> g(ix->mu);
	std::unique_ptr<vsgpu_text, void (*)(vsgpu_text*)> t(new vsgpu_text, vsgpu_text_free);
	t->owner = ix; t->n = n;
	try {
		ensure_hit_tables(ix);
		cudaStream_t st = ix->stream;                       // run_t4 / finish_t4 work on the index's stream
		t->offsets = (uint64_t*)ix->pinned_acquire((n + 1) * 8, &t->offsets_cap);
		if (!t->offsets) throw std::runtime_error("CUDA: cannot allocate page-locked result memory");
		t->offsets[0] = 0;
		if (n == 0) { t->bytes = (char*)ix->pinned_acquire(1, &t->bytes_cap); if (!t->bytes) throw std::runtime_error("CUDA: cannot allocate page-locked result memory"); t->bytes[0] = 0; *out = t.release(); return VSGPU_OK; }
		CU(ix->bx.ensure(n * 8)); CU(ix->by.ensure(n * 8)); CU(ix->bs.ensure(n * 4));
		CU(cudaMemcpyAsync(ix->bx.p, x, n * 8, cudaMemcpyHostToDevice, st));
		CU(cudaMemcpyAsync(ix->by.p, y, n * 8, cudaMemcpyHostToDevice, st));
		CU(cudaMemcpyAsync(ix->bs.p, sample_ids, n * 4, cudaMemcpyHostToDevice, st));
		const bool wide = expect_wide_regions(ix, n, x, y);
		uint64_t cap = ix->bhits.cap / 4;
		if (cap == 0) { cap = std::max<uint64_t>((wide ? 64 : 4) * n, 1024); CU(ix->bhits.ensure(cap * 4)); cap = ix->bhits.cap / 4; }
		run_t4(ix, n, ix->bx.as<uint64_t>(), ix->by.as<uint64_t>(), ix->bs.as<uint32_t>(), ix->boffsets, ix->bstate, ix->bhits, cap, nullptr, nullptr, ix->d_status, wide,
		       wide ? pool_of(ix->bspill) : expect_many_rows(ix, n, x, y) ? spill_of(ix->bspill) : nullptr);
		const uint32_t status = finish_t4(ix, n, ix->bx.as<uint64_t>(), ix->by.as<uint64_t>(), ix->bs.as<uint32_t>(), ix->boffsets, ix->bstate, ix->bhits, cap, ix->d_status, wide);
		if (status & kStatusBadRegion) return set_err(VSGPU_EINVAL, "region start < 1 or sample id out of range");
		render_hit_rows(ix, t.get(), n, with_samples, nullptr);
		*out = t.release();
	} catch (const std::invalid_argument& e) { return set_err(VSGPU_ESHAPE, e.what());
	} catch (const std::exception& e) { cudaDeviceSynchronize(); return set_err(VSGPU_ENODEVICE, e.what()); }
	return VSGPU_OK;
}

int vsgpu_render_t6(vsgpu_index* ix, uint64_t n, const uint64_t* x, const uint64_t* y, int with_samples, vsgpu_text** out) {
	if (!ix || !out || (n && (!x || !y))) return set_err(VSGPU_EINVAL, "vsgpu_render_t6: null argument");
	*out = nullptr;
	if (int rc = check_device(ix)) return rc;
	std::lock_guard<std::mutex> g(ix->mu);
	std::unique_ptr<vsgpu_text, void (*)(vsgpu_text*)> t(new vsgpu_text, vsgpu_text_free);
	t->owner = ix; t->n = n;
	try {
		ensure_render_tables(ix);
		cudaStream_t st = ix->s_k;
		t->offsets = (uint64_t*)ix->pinned_acquire((n + 1) * 8, &t->offsets_cap);
		if (!t->offsets) throw std::runtime_error("CUDA: cannot allocate page-locked result memory");
		t->offsets[0] = 0;
		uint64_t nseg = n; const uint32_t* seg_lo = nullptr; const uint32_t* seg_hi = nullptr;
		std::vector<uint64_t> seg_first;                       // only when a region needs several segments
		if (n) {
			CU(ix->bx.ensure(n * 8)); CU(ix->by.ensure(n * 8)); CU(ix->bout.ensure(n * 12));
			CU(cudaMemcpyAsync(ix->bx.p, x, n * 8, cudaMemcpyHostToDevice, st));
			CU(cudaMemcpyAsync(ix->by.p, y, n * 8, cudaMemcpyHostToDevice, st));
			uint32_t* d_lo = ix->bout.as<uint32_t>(); uint32_t* d_hi = d_lo + n;
			CU(launch_t6(ix->dev, n, ix->bx.as<uint64_t>(), ix->by.as<uint64_t>(), d_lo, d_hi, nullptr, nullptr, 0, ix->d_status, st));
			seg_lo = d_lo; seg_hi = d_hi;
			if (t6_special(ix)) {
				// Slices holding a repeated record, or regions past the contig end over tail records, are not
				// printed whole: the literal rule (query.h:397-414, 758-771) decides their rows, so such a
				// region becomes several segments (runs of kept records), built on the host.
				std::vector<uint32_t> lo(n), hi(n), slo, shi, vars;
				CU(cudaMemcpyAsync(lo.data(), d_lo, n * 4, cudaMemcpyDeviceToHost, st));
				CU(cudaMemcpyAsync(hi.data(), d_hi, n * 4, cudaMemcpyDeviceToHost, st));
				CU(cudaStreamSynchronize(st));
				seg_first.assign(n + 1, 0);
				for (uint64_t i = 0; i < n; i++) {
					seg_first[i] = slo.size();
					if (lo[i] == kNone || hi[i] <= lo[i]) continue;
					if (!t6_needs_literal(ix, y[i], lo[i], hi[i])) { slo.push_back(lo[i]); shi.push_back(hi[i]); continue; }
					t6_literal(ix, x[i], y[i], vars);
					for (size_t a = 0; a < vars.size();) { size_t b = a + 1; while (b < vars.size() && vars[b] == vars[b - 1] + 1) b++; slo.push_back(vars[a]); shi.push_back(vars[b - 1] + 1); a = b; }
				}
				seg_first[n] = slo.size();
				nseg = slo.size();
				CU(ix->bseg.ensure(std::max<uint64_t>(nseg, 1) * 8));
				if (nseg) {
					CU(cudaMemcpyAsync(ix->bseg.p, slo.data(), nseg * 4, cudaMemcpyHostToDevice, st));
					CU(cudaMemcpyAsync(ix->bseg.as<uint32_t>() + nseg, shi.data(), nseg * 4, cudaMemcpyHostToDevice, st));
					CU(cudaStreamSynchronize(st));                 // slo / shi go out of scope below
				}
				seg_lo = ix->bseg.as<uint32_t>(); seg_hi = seg_lo + nseg;
			}
		}
		uint64_t totals[2] = {0, 0};
		if (nseg) {
			CU(ix->brow_off.ensure((nseg + 1) * 8)); CU(ix->bbyte_off.ensure((nseg + 1) * 8)); CU(ix->bscratch.ensure(((nseg + 1023) / 1024 + 1) * 16));
			CU(cudaEventRecord(ix->ev_render[0], st));
			CU(launch_render_offsets(ix->dev, ix->render, nseg, seg_lo, seg_hi, with_samples, ix->brow_off.as<uint64_t>(), ix->bbyte_off.as<uint64_t>(), ix->bscratch.as<uint64_t>(), st));
			CU(cudaMemcpyAsync(&ix->pin_small[0], ix->brow_off.as<uint64_t>() + nseg, 8, cudaMemcpyDeviceToHost, st));
			CU(cudaMemcpyAsync(&ix->pin_small[1], ix->bbyte_off.as<uint64_t>() + nseg, 8, cudaMemcpyDeviceToHost, st));
			const uint32_t status = read_status(ix, ix->d_status, nullptr, st);
			if (status & kStatusBadRegion) return set_err(VSGPU_EINVAL, "Can't find node corresponding to pos 0");
			totals[0] = ix->pin_small[0]; totals[1] = ix->pin_small[1];
		}
		uint64_t max_bytes = 2ull << 30;
		if (const char* e = getenv("VSGPU_RENDER_MAX_BYTES")) max_bytes = strtoull(e, nullptr, 10);
		if (totals[1] > max_bytes) return set_err(VSGPU_ESHAPE, "vsgpu_render_t6: the rows of this batch take " + std::to_string(totals[1]) + " bytes (limit VSGPU_RENDER_MAX_BYTES = " + std::to_string(max_bytes) + "); split the batch");
		t->nrows = totals[0]; t->nbytes = totals[1];
		t->bytes = (char*)ix->pinned_acquire(totals[1] + 1, &t->bytes_cap);
		if (!t->bytes) throw std::runtime_error("CUDA: cannot allocate page-locked result memory");
		// segment offsets on the host: the result's region offsets, and where to cut the text into chunks
		// whose device->host copy overlaps the rendering of the next one
		size_t ro_cap = 0, bo_cap = 0;
		uint64_t* ro = nullptr; uint64_t* bo = nullptr;
		struct Temps { vsgpu_index* ix; uint64_t*& a; size_t& ac; uint64_t*& b; size_t& bc; ~Temps() { ix->pinned_release(a, ac); ix->pinned_release(b, bc); } } temps{ix, ro, ro_cap, bo, bo_cap};
		if (nseg) {
			ro = (uint64_t*)ix->pinned_acquire((nseg + 1) * 8, &ro_cap); bo = (uint64_t*)ix->pinned_acquire((nseg + 1) * 8, &bo_cap);
			if (!ro || !bo) throw std::runtime_error("CUDA: cannot allocate page-locked result memory");
			CU(cudaMemcpyAsync(ro, ix->brow_off.p, (nseg + 1) * 8, cudaMemcpyDeviceToHost, st));
			CU(cudaMemcpyAsync(bo, ix->bbyte_off.p, (nseg + 1) * 8, cudaMemcpyDeviceToHost, st));
			CU(cudaStreamSynchronize(st));
		}
		if (totals[1]) {
			CU(ix->btext.ensure(totals[1]));
			uint64_t chunk_bytes = 32ull << 20;
			if (const char* e = getenv("VSGPU_RENDER_CHUNK_BYTES")) chunk_bytes = std::max<uint64_t>(1, strtoull(e, nullptr, 10));
			const int chunks = (int)std::min<uint64_t>(vsgpu_index::kMaxChunks, (totals[1] + chunk_bytes - 1) / chunk_bytes);
			uint64_t s0 = 0;
			for (int c = 0; c < chunks && s0 < nseg; c++) {
				uint64_t s1 = nseg;
				if (c + 1 < chunks) { const uint64_t target = totals[1] / chunks * (c + 1); s1 = std::lower_bound(bo + s0 + 1, bo + nseg, target) - bo; }
				if (s1 <= s0) continue;
				CU(launch_render(ix->dev, ix->render, nseg, seg_lo, with_samples, ix->brow_off.as<uint64_t>(), ix->bbyte_off.as<uint64_t>(), ro[s0], ro[s1], ix->btext.as<char>(), st));
				CU(cudaEventRecord(ix->ev_k[c], st));
				CU(cudaStreamWaitEvent(ix->s_out, ix->ev_k[c], 0));
				if (bo[s1] > bo[s0]) CU(cudaMemcpyAsync(t->bytes + bo[s0], ix->btext.as<char>() + bo[s0], bo[s1] - bo[s0], cudaMemcpyDeviceToHost, ix->s_out));
				s0 = s1;
			}
			CU(cudaEventRecord(ix->ev_render[1], st));
			CU(cudaStreamSynchronize(ix->s_out));
			CU(cudaStreamSynchronize(st));
		} else if (nseg) { CU(cudaEventRecord(ix->ev_render[1], st)); CU(cudaStreamSynchronize(st)); }
		if (!nseg) for (uint64_t i = 0; i <= n; i++) t->offsets[i] = 0;
		else if (seg_first.empty()) memcpy(t->offsets, bo, (n + 1) * 8);
		else for (uint64_t i = 0; i <= n; i++) t->offsets[i] = bo[seg_first[i]];
		t->bytes[totals[1]] = 0;
		if (nseg) CU(cudaEventElapsedTime(&t->kernel_ms, ix->ev_render[0], ix->ev_render[1]));
		*out = t.release();
	} catch (const std::exception& e) { cudaDeviceSynchronize(); return set_err(VSGPU_ENODEVICE, e.what()); }
	return VSGPU_OK;
}
const char* vsgpu_text_bytes(const vsgpu_text* t) { return t ? t->bytes : nullptr; }
const uint64_t* vsgpu_text_offsets(const vsgpu_text* t) { return t ? t->offsets : nullptr; }
uint64_t vsgpu_text_num_rows(const vsgpu_text* t) { return t ? t->nrows : 0; }
float vsgpu_text_kernel_ms(const vsgpu_text* t) { return t ? t->kernel_ms : 0.f; }
const uint8_t* vsgpu_text_status(const vsgpu_text* t) { return t ? t->status : nullptr; }
const float* vsgpu_text_stage_ms(const vsgpu_text* t) { return t ? t->stage_ms : nullptr; }
void vsgpu_text_free(vsgpu_text* t) {
	if (!t) return;
	if (t->owner) { t->owner->pinned_release(t->bytes, t->bytes_cap); t->owner->pinned_release(t->offsets, t->offsets_cap); t->owner->pinned_release(t->status, t->status_cap); }
	delete t;
}

// ------------------------------------------------------------------ t2: query_sample_from_ref
namespace {
void ensure_t2_tables(vsgpu_index* ix) {
	if (ix->t2_ready) return;
	const FlatIndex& f = ix->flat;
	if (!f.t2_ok) throw std::invalid_argument("vsgpu_query_t2: " + f.t2_why);
	T2Tables& t = ix->t2;
	std::vector<uint32_t> bbs(f.vstart.begin(), f.vstart.end());
	bbs.push_back(ix->last_end);
	t.bbs = upload(ix, bbs); t.nrp1 = upload(ix, f.nrp1); t.first_reach = upload(ix, f.first_reach);
	t.cent_seq = (const uint2*)upload(ix, f.cent_seq);
	static const char kBase[8] = {'A', 'C', 'T', 'G', 'N', 5, 5, 5};   // util.cc:32-41 (map_int)
	// 64 bytes of padding on both sides: the copy kernel reads whole aligned 16-byte vectors around a piece
	std::vector<char> ascii(ix->ser.seq.size() + 128, 0);
	parallel_for(ix->ser.seq.size(), [&](uint64_t a, uint64_t b) { for (uint64_t i = a; i < b; i++) ascii[64 + i] = kBase[ix->ser.seq[i] & 7]; });
	t.seq_ascii = upload(ix, ascii) + 64;
	for (auto& e : ix->ev_t2) CU(cudaEventCreate(&e));
	ix->t2_ready = true;
}
}  // namespace

namespace {
// t3 on top of the t2 tables: sample_info.index of every carrier of every walk entry's target vertex.
// vsgpu_open keeps only the first index of a vertex, so the vertex blocks are decoded a second time.
void ensure_t3_tables(vsgpu_index* ix) {
	ensure_t2_tables(ix);
	if (ix->t3_ready) return;
	const FlatIndex& f = ix->flat; const SerData& sd = ix->ser;
	std::vector<uint32_t>& sindex = ix->sindex;            // kept on the host too: t5 rows print the sample's position
	try { load_sample_indexes(ix->prefix, sd.v_sinfo_begin.back(), sindex); }
	catch (const std::exception& e) { throw std::invalid_argument(std::string("vsgpu_query_t3: ") + e.what()); }
	const size_t E = f.cent.size();
	std::vector<uint64_t> begin(E + 1, 0);
	for (size_t c = 0; c < E; c++) { const uint32_t v = f.cent_vertex[c]; begin[c + 1] = begin[c] + (v == kNone ? 0 : sd.v_sinfo_begin[v + 1] - sd.v_sinfo_begin[v]); }
	std::vector<uint32_t> sidx(begin[E]), sid(f.class_mode ? 0 : begin[E]);
	parallel_for(E, [&](uint64_t a, uint64_t b) {
		for (uint64_t c = a; c < b; c++) {
			const uint32_t v = f.cent_vertex[c];
			if (v == kNone) continue;
			const uint64_t s0 = sd.v_sinfo_begin[v], cnt = sd.v_sinfo_begin[v + 1] - s0;
			memcpy(sidx.data() + begin[c], sindex.data() + s0, cnt * 4);
			if (!f.class_mode) memcpy(sid.data() + begin[c], sd.s_sample_id.data() + s0, cnt * 4);
		}
	});
	ix->t3.sidx_begin = upload(ix, begin); ix->t3.sidx = upload(ix, sidx);
	ix->t3.sid = f.class_mode ? nullptr : upload(ix, sid);
	ix->t3.first_index = f.vstart[f.dlev[0].k];
	ix->t3_ready = true;
}

int query_seq_impl(vsgpu_index* ix, bool t3, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample_ids, vsgpu_text** out) {
	if (!ix || !out || (n && (!x || !y || !sample_ids))) return set_err(VSGPU_EINVAL, "vsgpu_query_t2: null argument");
	*out = nullptr;
	if (int rc = check_device(ix)) return rc;
	std::lock_guard<std::mutex> g(ix->mu);
	std::unique_ptr<vsgpu_text, void (*)(vsgpu_text*)> t(new vsgpu_text, vsgpu_text_free);
	t->owner = ix; t->n = n;
	try {
		if (t3) ensure_t3_tables(ix); else ensure_t2_tables(ix);
		const T3Tables* t3p = t3 ? &ix->t3 : nullptr;
		cudaStream_t st = ix->s_k;
		t->offsets = (uint64_t*)ix->pinned_acquire((n + 1) * 8, &t->offsets_cap);
		t->status = (uint8_t*)ix->pinned_acquire(n + 1, &t->status_cap);
		if (!t->offsets || !t->status) throw std::runtime_error("CUDA: cannot allocate page-locked result memory");
		uint64_t totals[2] = {0, 0};
		const uint64_t nctas = t2_ctas(n);
		if (n) {
			CU(ix->bx.ensure(n * 8)); CU(ix->by.ensure(n * 8)); CU(ix->bs.ensure(n * 4));
			CU(ix->bcnt.ensure(n * 8)); CU(ix->bst8.ensure(n)); CU(ix->bkeep.ensure(n * 8 * kT2Keep)); CU(ix->bscratch.ensure((nctas + 1) * 16)); CU(ix->bbyte_off.ensure((n + 1) * 8));
			CU(cudaMemcpyAsync(ix->bx.p, x, n * 8, cudaMemcpyHostToDevice, st));
			CU(cudaMemcpyAsync(ix->by.p, y, n * 8, cudaMemcpyHostToDevice, st));
			CU(cudaMemcpyAsync(ix->bs.p, sample_ids, n * 4, cudaMemcpyHostToDevice, st));
			CU(cudaEventRecord(ix->ev_t2[0], st));
			CU(launch_t2_count(ix->dev, ix->t2, n, ix->bx.as<uint64_t>(), ix->by.as<uint64_t>(), ix->bs.as<uint32_t>(), ix->bcnt.as<uint2>(), ix->bkeep.as<uint2>(), ix->bst8.as<uint8_t>(),
			                   ix->bscratch.as<uint64_t>(), ix->d_status, st, t3p));
			CU(cudaEventRecord(ix->ev_t2[1], st));
			CU(cudaMemcpyAsync(ix->pin_small, ix->bscratch.as<uint64_t>() + 2 * nctas, 16, cudaMemcpyDeviceToHost, st));
			const uint32_t status = read_status(ix, ix->d_status, nullptr, st);
			if (status & kStatusBadRegion) return set_err(VSGPU_EINVAL, "vsgpu_query_t2: sample id out of range");
			totals[0] = ix->pin_small[0]; totals[1] = ix->pin_small[1];
		}
		uint64_t max_bytes = 2ull << 30;
		if (const char* e = getenv("VSGPU_RENDER_MAX_BYTES")) max_bytes = strtoull(e, nullptr, 10);
		if (totals[1] > max_bytes) return set_err(VSGPU_ESHAPE, "vsgpu_query_t2: the sequences of this batch take " + std::to_string(totals[1]) + " bytes (limit VSGPU_RENDER_MAX_BYTES = " + std::to_string(max_bytes) + "); split the batch");
		t->nbytes = totals[1]; t->nrows = n;
		t->bytes = (char*)ix->pinned_acquire(totals[1] + 1, &t->bytes_cap);
		if (!t->bytes) throw std::runtime_error("CUDA: cannot allocate page-locked result memory");
		if (n) {
			CU(ix->brecs.ensure(std::max<uint64_t>(totals[0], 1) * 16)); CU(ix->btext.ensure(totals[1] + kT2Tile));
			CU(ix->btile.ensure((totals[1] / kT2Tile + 2) * 4));
			CU(cudaEventRecord(ix->ev_t2[2], st));
			CU(launch_t2_plan(ix->dev, ix->t2, n, ix->bx.as<uint64_t>(), ix->by.as<uint64_t>(), ix->bs.as<uint32_t>(), ix->bcnt.as<uint2>(), ix->bkeep.as<uint2>(), ix->bscratch.as<uint64_t>(),
			                  ix->bbyte_off.as<uint64_t>(), ix->brecs.as<uint4>(), ix->btile.as<uint32_t>(), st, t3p));
			CU(cudaEventRecord(ix->ev_t2[3], st));
			CU(launch_t2_copy(ix->t2, ix->brecs.as<uint4>(), ix->btile.as<uint32_t>(), ix->bscratch.as<uint64_t>() + 2 * nctas, totals[0], totals[1], ix->btext.as<char>(), st));
			CU(cudaEventRecord(ix->ev_t2[4], st));
			CU(cudaMemcpyAsync(t->offsets, ix->bbyte_off.p, (n + 1) * 8, cudaMemcpyDeviceToHost, st));
			CU(cudaMemcpyAsync(t->status, ix->bst8.p, n, cudaMemcpyDeviceToHost, st));
			if (totals[1]) CU(cudaMemcpyAsync(t->bytes, ix->btext.p, totals[1], cudaMemcpyDeviceToHost, st));
			CU(cudaStreamSynchronize(st));
			CU(cudaEventElapsedTime(&t->stage_ms[0], ix->ev_t2[0], ix->ev_t2[1]));
			CU(cudaEventElapsedTime(&t->stage_ms[1], ix->ev_t2[2], ix->ev_t2[3]));
			CU(cudaEventElapsedTime(&t->stage_ms[2], ix->ev_t2[3], ix->ev_t2[4]));
			t->kernel_ms = t->stage_ms[0] + t->stage_ms[1] + t->stage_ms[2];
		} else t->offsets[0] = 0;
		t->bytes[totals[1]] = 0;
		*out = t.release();
	} catch (const std::invalid_argument& e) { return set_err(VSGPU_ESHAPE, e.what());
	} catch (const std::exception& e) { cudaDeviceSynchronize(); return set_err(VSGPU_ENODEVICE, e.what()); }
	return VSGPU_OK;
}
}  // namespace
// ------------------------------------------------------------------ t5: get_sample_var_in_sample
namespace {
// t5 of a batch on stream st: inputs up, count, scan, write.  Leaves the CSR in boffsets / bhits, the per-region status in
// bst8, the sample ids in bs; returns the number of hit codes; *bad = a sample id was out of range.
uint64_t t5_on_device(vsgpu_index* ix, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample_ids, cudaStream_t st, bool* bad) {
	const uint64_t nctas = (n + 255) / 256;
	*bad = false;
	CU(ix->bx.ensure(n * 8)); CU(ix->by.ensure(n * 8)); CU(ix->bs.ensure(n * 4));
	CU(ix->bcnt.ensure(n * 4)); CU(ix->bst8.ensure(n)); CU(ix->bscratch.ensure((nctas + 1) * 16)); CU(ix->boffsets.ensure((n + 1) * 8)); CU(ix->bkeep.ensure(t5_keep_words(n) * 4));
	CU(cudaMemcpyAsync(ix->bx.p, x, n * 8, cudaMemcpyHostToDevice, st));
	CU(cudaMemcpyAsync(ix->by.p, y, n * 8, cudaMemcpyHostToDevice, st));
	CU(cudaMemcpyAsync(ix->bs.p, sample_ids, n * 4, cudaMemcpyHostToDevice, st));
	CU(cudaEventRecord(ix->ev_t2[0], st));
	CU(launch_t5_count(ix->dev, ix->t2, ix->t3, n, ix->bx.as<uint64_t>(), ix->by.as<uint64_t>(), ix->bs.as<uint32_t>(), ix->bcnt.as<uint32_t>(), ix->bst8.as<uint8_t>(),
	                   ix->bscratch.as<uint64_t>(), ix->d_status, ix->bkeep.as<uint32_t>(), st));
	CU(cudaEventRecord(ix->ev_t2[1], st));
	CU(cudaMemcpyAsync(ix->pin_small, ix->bscratch.as<uint64_t>() + 2 * nctas, 8, cudaMemcpyDeviceToHost, st));
	const uint32_t status = read_status(ix, ix->d_status, nullptr, st);
	if (status & kStatusBadRegion) { *bad = true; return 0; }
	const uint64_t total = ix->pin_small[0];
	CU(ix->bhits.ensure(std::max<uint64_t>(total, 1) * 4));
	CU(cudaEventRecord(ix->ev_t2[2], st));
	CU(launch_t5_write(ix->dev, ix->t2, ix->t3, n, ix->bx.as<uint64_t>(), ix->by.as<uint64_t>(), ix->bs.as<uint32_t>(), ix->bcnt.as<uint32_t>(), ix->bscratch.as<uint64_t>(),
	                   ix->boffsets.as<uint64_t>(), ix->bhits.as<uint32_t>(), ix->bkeep.as<uint32_t>(), st));
	CU(cudaEventRecord(ix->ev_t2[3], st));
	return total;
}
}  // namespace
int vsgpu_query_t5(vsgpu_index* ix, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample_ids, vsgpu_result** out) {
	if (!ix || !out || (n && (!x || !y || !sample_ids))) return set_err(VSGPU_EINVAL, "vsgpu_query_t5: null argument");
	*out = nullptr;
	if (int rc = check_device(ix)) return rc;
	std::lock_guard<std::mutex> g(ix->mu);
	std::unique_ptr<vsgpu_result, void (*)(vsgpu_result*)> r(new vsgpu_result, vsgpu_result_free);
	r->owner = ix; r->n = n;
	try {
		ensure_t3_tables(ix);
		cudaStream_t st = ix->s_k;
		r->offsets = (uint64_t*)ix->pinned_acquire((n + 1) * 8, &r->offsets_cap);
		r->status = (uint8_t*)ix->pinned_acquire(n + 1, &r->status_cap);
		if (!r->offsets || !r->status) throw std::runtime_error("CUDA: cannot allocate page-locked result memory");
		r->offsets[0] = 0;
		if (n) {
			bool bad = false;
			const uint64_t total = t5_on_device(ix, n, x, y, sample_ids, st, &bad);
			if (bad) return set_err(VSGPU_EINVAL, "vsgpu_query_t5: sample id out of range");
			r->hits = (uint32_t*)ix->pinned_acquire(std::max<uint64_t>(total, 1) * 4, &r->hits_cap);
			if (!r->hits) throw std::runtime_error("CUDA: cannot allocate page-locked result memory");
			CU(cudaMemcpyAsync(r->offsets, ix->boffsets.p, (n + 1) * 8, cudaMemcpyDeviceToHost, st));
			CU(cudaMemcpyAsync(r->status, ix->bst8.p, n, cudaMemcpyDeviceToHost, st));
			if (total) CU(cudaMemcpyAsync(r->hits, ix->bhits.p, total * 4, cudaMemcpyDeviceToHost, st));
			CU(cudaStreamSynchronize(st));
			float a = 0, b = 0;
			CU(cudaEventElapsedTime(&a, ix->ev_t2[0], ix->ev_t2[1])); CU(cudaEventElapsedTime(&b, ix->ev_t2[2], ix->ev_t2[3]));
			r->kernel_ms = a + b;
			r->total = r->offsets[n];
		}
		r->have_offsets = true;
		*out = r.release();
	} catch (const std::invalid_argument& e) { return set_err(VSGPU_ESHAPE, e.what());
	} catch (const std::exception& e) { cudaDeviceSynchronize(); return set_err(VSGPU_ENODEVICE, e.what()); }
	return VSGPU_OK;
}
// get_sample_var_in_sample(vg, idx, x, y, sample, print = true, outfile) for a batch: t5 on the device, then its rows as text
// (vsgpu_rows_t5's, written by the kernels of vsgpu_render_t4 with the position column taken from the sample's own coordinate)
int vsgpu_render_t5(vsgpu_index* ix, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample_ids, int with_samples, vsgpu_text** out) {
	if (!ix || !out || (n && (!x || !y || !sample_ids))) return set_err(VSGPU_EINVAL, "vsgpu_render_t5: null argument");
	*out = nullptr;
	if (int rc = check_device(ix)) return rc;
	std::lock_guard<std::mutex> g(ix->mu);
	std::unique_ptr<vsgpu_text, void (*)(vsgpu_text*)> t(new vsgpu_text, vsgpu_text_free);
	t->owner = ix; t->n = n;
	try {
		ensure_t3_tables(ix);
		ensure_hit_tables(ix);
		if (!ix->d_gsidx) ix->d_gsidx = upload(ix, ix->sindex);
		cudaStream_t st = ix->stream;
		t->offsets = (uint64_t*)ix->pinned_acquire((n + 1) * 8, &t->offsets_cap);
		t->status = (uint8_t*)ix->pinned_acquire(n + 1, &t->status_cap);
		if (!t->offsets || !t->status) throw std::runtime_error("CUDA: cannot allocate page-locked result memory");
		t->offsets[0] = 0;
		if (n == 0) { t->bytes = (char*)ix->pinned_acquire(1, &t->bytes_cap); if (!t->bytes) throw std::runtime_error("CUDA: cannot allocate page-locked result memory"); t->bytes[0] = 0; *out = t.release(); return VSGPU_OK; }
		bool bad = false;
		t5_on_device(ix, n, x, y, sample_ids, st, &bad);
		if (bad) return set_err(VSGPU_EINVAL, "vsgpu_render_t5: sample id out of range");
		CU(cudaMemcpyAsync(t->status, ix->bst8.p, n, cudaMemcpyDeviceToHost, st));
		render_hit_rows(ix, t.get(), n, with_samples, ix->bs.as<uint32_t>());
		float a = 0, b = 0;
		CU(cudaEventElapsedTime(&a, ix->ev_t2[0], ix->ev_t2[1])); CU(cudaEventElapsedTime(&b, ix->ev_t2[2], ix->ev_t2[3]));
		t->stage_ms[0] = a; t->stage_ms[1] = b; t->stage_ms[2] = t->kernel_ms;          // count, write, rows
		*out = t.release();
	} catch (const std::invalid_argument& e) { return set_err(VSGPU_ESHAPE, e.what());
	} catch (const std::exception& e) { cudaDeviceSynchronize(); return set_err(VSGPU_ENODEVICE, e.what()); }
	return VSGPU_OK;
}
const uint8_t* vsgpu_result_status(const vsgpu_result* r) { return r ? r->status : nullptr; }
float vsgpu_result_kernel_ms(const vsgpu_result* r) { return r ? r->kernel_ms : 0.f; }
int vsgpu_rows_t5(const vsgpu_index* ix, const uint32_t* hits, uint64_t nhits, uint32_t sample_id, int with_samples, char** text) {
	if (!ix || !text || (nhits && !hits)) return set_err(VSGPU_EINVAL, "vsgpu_rows_t5: null argument");
	if (ix->sindex.empty() && nhits) return set_err(VSGPU_EINVAL, "vsgpu_rows_t5: no t5 query has run on this index");
	if (sample_id == 0 || sample_id >= ix->ser.num_samples) return set_err(VSGPU_EINVAL, "vsgpu_rows_t5: sample id out of range");
	std::string s;
	for (uint64_t i = 0; i < nhits; i++) {
		if ((hits[i] & VSGPU_HIT_ENTRY_MASK) >= ix->flat.cent.size()) return set_err(VSGPU_EINVAL, "vsgpu_rows_t5: hit code out of range");
		t5_row(ix, hits[i], sample_id, with_samples != 0, s);
	}
	*text = dup_text(s);
	return *text ? VSGPU_OK : set_err(VSGPU_ENOMEM, "out of memory");
}
int vsgpu_digest_t5(const vsgpu_index* ix, uint64_t n, const uint64_t* offsets, const uint32_t* hits, const uint32_t* sample_ids, int with_samples, uint64_t* digests) {
	if (!ix || (n && (!offsets || !digests || !sample_ids))) return set_err(VSGPU_EINVAL, "vsgpu_digest_t5: null argument");
	if (ix->sindex.empty() && n && offsets[n]) return set_err(VSGPU_EINVAL, "vsgpu_digest_t5: no t5 query has run on this index");
	if (!csr_ok(ix, n, offsets, hits)) return set_err(VSGPU_EINVAL, "vsgpu_digest_t5: offsets decrease or a hit code names no walk entry");
	for (uint64_t i = 0; i < n; i++) if (sample_ids[i] == 0 || sample_ids[i] >= ix->ser.num_samples) return set_err(VSGPU_EINVAL, "vsgpu_digest_t5: sample id out of range");
	digests_t5(ix, n, offsets, hits, sample_ids, with_samples != 0, digests);
	return VSGPU_OK;
}

int vsgpu_query_t2(vsgpu_index* ix, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample_ids, vsgpu_text** out) { return query_seq_impl(ix, false, n, x, y, sample_ids, out); }
int vsgpu_query_t3(vsgpu_index* ix, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample_ids, vsgpu_text** out) { return query_seq_impl(ix, true, n, x, y, sample_ids, out); }

// ------------------------------------------------------------------ device-resident batches
int vsgpu_batch_create(vsgpu_index* ix, int type, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample_ids,
                       const char* const* refs, const char* const* alts, vsgpu_batch** out) {
	if (!ix || !out || !x || n == 0) return set_err(VSGPU_EINVAL, "vsgpu_batch_create: null argument");
	if (type != 4 && type != 6 && type != 7 && type != 46) return set_err(VSGPU_EINVAL, "vsgpu_batch_create: type must be 4, 6, 7 or 46");
	if ((type != 7 && !y) || ((type == 4 || type == 46) && !sample_ids) || (type == 7 && (!refs || !alts))) return set_err(VSGPU_EINVAL, "vsgpu_batch_create: missing input array");
	if (int rc = check_device(ix)) return rc;
	std::lock_guard<std::mutex> g(ix->mu);
	try {
		std::unique_ptr<vsgpu_batch> b(new vsgpu_batch);
		b->idx = ix; b->device = ix->device; b->type = type; b->n = n;
		CU(cudaMalloc((void**)&b->d_status, 8)); CU(cudaMemset(b->d_status, 0, 8));
		b->hx.assign(x, x + n); if (y) b->hy.assign(y, y + n);
		CU(b->x.ensure(n * 8));
		CU(cudaMemcpyAsync(b->x.p, x, n * 8, cudaMemcpyHostToDevice, ix->stream));
		if (type != 7) { CU(b->y.ensure(n * 8)); CU(cudaMemcpyAsync(b->y.p, y, n * 8, cudaMemcpyHostToDevice, ix->stream)); }
		if (type == 6) { CU(b->out.ensure(n * 12)); if (t6_special(ix)) CU(b->flag.ensure(n * 4)); }
		if (type == 4 || type == 46) { b->wide_regions = expect_wide_regions(ix, n, x, y); b->many_rows = !b->wide_regions && expect_many_rows(ix, n, x, y); CU(b->s.ensure(n * 4)); CU(cudaMemcpyAsync(b->s.p, sample_ids, n * 4, cudaMemcpyHostToDevice, ix->stream)); CU(b->out.ensure(n * 12)); }
		if (type == 46) {
			if (t6_special(ix)) CU(b->flag.ensure(n * 4));           // (a batch of few, wide regions runs k_t6 and the warp-per-region t4 kernel: two launches)
		}
		if (type == 7) {
			std::vector<uint64_t> qh(n);
			for (uint64_t i = 0; i < n; i++) qh[i] = hash_query(refs[i], alts[i]);
			CU(b->hash.ensure(n * 8)); CU(b->rec.ensure(n * 4));
			CU(cudaMemcpyAsync(b->hash.p, qh.data(), n * 8, cudaMemcpyHostToDevice, ix->stream));
			CU(cudaStreamSynchronize(ix->stream));
		}
		CU(cudaStreamSynchronize(ix->stream));
		*out = b.release();
	} catch (const std::exception& e) { return set_err(VSGPU_ENODEVICE, e.what()); }
	return VSGPU_OK;
}

int vsgpu_batch_run(vsgpu_batch* b) {
	if (!b) return set_err(VSGPU_EINVAL, "vsgpu_batch_run: null batch");
	vsgpu_index* ix = b->idx;
	if (int rc = check_device(ix)) return rc;
	try {
		for (auto& e : b->ev) if (!e) CU(cudaEventCreate(&e));
		if (b->type == 6) { CU(cudaMemsetAsync(b->d_status, 0, 8, ix->stream)); CU(cudaEventRecord(b->ev[0], ix->stream)); CU(launch_t6(ix->dev, b->n, b->x.as<uint64_t>(), b->y.as<uint64_t>(), b->out.as<uint32_t>(), b->out.as<uint32_t>() + b->n, b->out.as<uint32_t>() + 2 * b->n, b->flag.as<uint32_t>(), 0, b->d_status, ix->stream)); CU(cudaEventRecord(b->ev[1], ix->stream)); b->launches = 1; }
		else if (b->type == 7) { CU(cudaEventRecord(b->ev[0], ix->stream)); CU(launch_t7(ix->dev, b->n, b->x.as<uint64_t>(), b->hash.as<uint64_t>(), b->rec.as<uint32_t>(), b->d_status, ix->stream)); CU(cudaEventRecord(b->ev[1], ix->stream)); b->launches = 1; }
		else if (b->type == 46 && !t4x_supported(b->wide_regions)) {
			const uint64_t n = b->n;
			uint32_t* o = b->out.as<uint32_t>();
			CU(cudaMemsetAsync(b->d_status, 0, 8, ix->stream));
			CU(cudaEventRecord(b->ev[0], ix->stream));
			CU(launch_t6(ix->dev, n, b->x.as<uint64_t>(), b->y.as<uint64_t>(), o, o + n, o + 2 * n, b->flag.as<uint32_t>(), 0, b->d_status, ix->stream));
			CU(cudaEventRecord(b->ev[1], ix->stream));
			run_t4(ix, n, b->x.as<uint64_t>(), b->y.as<uint64_t>(), b->s.as<uint32_t>(), b->offsets, b->state, b->hits, b->hits_cap, nullptr, nullptr, b->d_status, b->wide_regions,
			       b->wide_regions ? pool_of(b->spill) : nullptr);
			CU(cudaEventRecord(b->ev[2], ix->stream));
			b->launches = 2;
		}
		else if (b->type == 46) {
			// t6 + t4 from one launch of k_t4p<kFuse6>: the t6 slice comes from the two ranks of the t4 setup
			const uint64_t n = b->n;
			CU(b->offsets.ensure((n + 1) * 8)); CU(b->state.ensure(t4_state_words(n) * 8));
			if (b->hits_cap == 0) { b->hits_cap = std::max<uint64_t>(4 * n, 1024); CU(b->hits.ensure(b->hits_cap * 4)); }
			CU(cudaMemsetAsync(b->d_status, 0, 8, ix->stream));
			CU(cudaMemsetAsync(b->state.p, 0, t4_state_words(n) * 8, ix->stream));
			CU(cudaEventRecord(b->ev[0], ix->stream));
			uint32_t* o = b->out.as<uint32_t>();
			const T6Out f6{o, o + n, o + 2 * n, b->flag.as<uint32_t>(), 0};
			const T4Launch L{n, b->x.p, b->y.p, false, b->s.as<uint32_t>(), b->offsets.as<uint64_t>(), nullptr, b->hits.as<uint32_t>(), b->hits_cap, b->state.as<uint64_t>(), b->d_status, nullptr, &f6,
			                 b->many_rows ? spill_of(b->spill) : nullptr};
			CU(launch_t4x(ix->dev, L, ix->stream));
			CU(cudaEventRecord(b->ev[1], ix->stream));
			b->launches = 1;
		}
		else run_t4(ix, b->n, b->x.as<uint64_t>(), b->y.as<uint64_t>(), b->s.as<uint32_t>(), b->offsets, b->state, b->hits, b->hits_cap, &b->launches, b->ev, b->d_status, b->wide_regions,
		            b->wide_regions ? pool_of(b->spill) : b->many_rows ? spill_of(b->spill) : nullptr);
	} catch (const std::exception& e) { return set_err(VSGPU_ENODEVICE, e.what()); }
	return VSGPU_OK;
}

int vsgpu_batch_fetch(vsgpu_batch* b, uint32_t* rec_lo, uint32_t* rec_hi, uint32_t* counts, vsgpu_result** out) {
	if (!b) return set_err(VSGPU_EINVAL, "vsgpu_batch_fetch: null batch");
	vsgpu_index* ix = b->idx;
	if (int rc = check_device(ix)) return rc;
	try {
		const uint64_t n = b->n;
		if (b->type == 6) {
			const uint32_t* o = b->out.as<uint32_t>();
			if (rec_lo) CU(cudaMemcpyAsync(rec_lo, o, n * 4, cudaMemcpyDeviceToHost, ix->stream));
			if (rec_hi) CU(cudaMemcpyAsync(rec_hi, o + n, n * 4, cudaMemcpyDeviceToHost, ix->stream));
			if (counts) CU(cudaMemcpyAsync(counts, o + 2 * n, n * 4, cudaMemcpyDeviceToHost, ix->stream));
			settle_t6(ix, b->hx.data(), b->hy.data(), counts, b->flag.as<uint32_t>(), b->d_status, ix->stream);
		} else if (b->type == 7) {
			if (rec_lo) CU(cudaMemcpyAsync(rec_lo, b->rec.p, n * 4, cudaMemcpyDeviceToHost, ix->stream));
			uint32_t st = read_status(ix, b->d_status);
			if (st & kStatusBadRegion) return set_err(VSGPU_EINVAL, "Can't find node corresponding to pos 0");
		} else if (b->type == 46) {
			// t6 part first (it owns the flagged list in the status words), then the CSR; counts = the t6 row counts
			const uint32_t* o = b->out.as<uint32_t>();
			if (rec_lo) CU(cudaMemcpyAsync(rec_lo, o, n * 4, cudaMemcpyDeviceToHost, ix->stream));
			if (rec_hi) CU(cudaMemcpyAsync(rec_hi, o + n, n * 4, cudaMemcpyDeviceToHost, ix->stream));
			if (counts) CU(cudaMemcpyAsync(counts, o + 2 * n, n * 4, cudaMemcpyDeviceToHost, ix->stream));
			uint64_t total = 0;
			CU(cudaMemcpyAsync(&total, b->offsets.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, ix->stream));
			uint32_t nflag = 0;
			const uint32_t st = read_status(ix, b->d_status, &nflag, ix->stream);
			if (st & kStatusBadRegion) return set_err(VSGPU_EINVAL, "region start < 1 or sample id out of range");
			if (nflag && counts) {
				std::vector<uint32_t> flagged(nflag), tmp;
				CU(cudaMemcpyAsync(flagged.data(), b->flag.p, (size_t)nflag * 4, cudaMemcpyDeviceToHost, ix->stream));
				CU(cudaStreamSynchronize(ix->stream));
				for (uint32_t i : flagged) { t6_literal(ix, b->hx[i], b->hy[i], tmp); counts[i] = (uint32_t)tmp.size(); }
			}
			if (total > b->hits_cap) {                           // the hit buffer was a guess: exact size, t4 alone again
				b->hits_cap = total + total / 16 + 1024;
				CU(b->hits.ensure(b->hits_cap * 4));
				run_t4(ix, n, b->x.as<uint64_t>(), b->y.as<uint64_t>(), b->s.as<uint32_t>(), b->offsets, b->state, b->hits, b->hits_cap, nullptr, nullptr, b->d_status, b->wide_regions);
				read_status(ix, b->d_status);
			}
			vsgpu_result* r = fetch_t4(ix, n, b->offsets, b->hits, out != nullptr);
			if (out) *out = r; else vsgpu_result_free(r);
		} else {
			uint32_t st = finish_t4(ix, n, b->x.as<uint64_t>(), b->y.as<uint64_t>(), b->s.as<uint32_t>(), b->offsets, b->state, b->hits, b->hits_cap, b->d_status, b->wide_regions);
			if (st & kStatusBadRegion) return set_err(VSGPU_EINVAL, "region start < 1 or sample id out of range");
			vsgpu_result* r = fetch_t4(ix, n, b->offsets, b->hits, out != nullptr);
			if (counts) for (uint64_t i = 0; i < n; i++) counts[i] = (uint32_t)(r->offsets[i + 1] - r->offsets[i]);
			if (out) *out = r; else vsgpu_result_free(r);
		}
	} catch (const std::invalid_argument& e) { return set_err(VSGPU_EINVAL, e.what());
	} catch (const std::exception& e) { return set_err(VSGPU_ENODEVICE, e.what()); }
	return VSGPU_OK;
}

// Algorithmic bytes per SURVEY.md §8d:  t6 288 B/region;  t4 292 + 20 v + 4 h (v = branch records
// of the region's slice, h = hits);  t7 148 + 16 r (r = records compared).
int vsgpu_batch_stats(vsgpu_batch* b, uint64_t* algorithmic_bytes, uint32_t* kernel_launches) {
	if (!b) return set_err(VSGPU_EINVAL, "vsgpu_batch_stats: null batch");
	vsgpu_index* ix = b->idx;
	if (kernel_launches) *kernel_launches = b->launches;
	if (!algorithmic_bytes) return VSGPU_OK;
	if (!b->algo_valid) {
		if (int rc = check_device(ix)) return rc;
		try {
			const uint64_t n = b->n; uint64_t bytes = 0;
			if (b->type == 6) bytes = 288 * n;
			else if (b->type == 46) {
				// t6 (288) + t4 (292 + 20 v + 4 h) by the section 8(d) convention; v = the slice lengths the launch itself wrote
				std::vector<uint32_t> o(n); uint64_t total = 0;
				CU(cudaMemcpyAsync(o.data(), b->out.as<uint32_t>() + 2 * n, n * 4, cudaMemcpyDeviceToHost, ix->stream));
				CU(cudaMemcpyAsync(&total, b->offsets.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, ix->stream));
				CU(cudaStreamSynchronize(ix->stream));
				uint64_t v = 0; for (uint64_t i = 0; i < n; i++) v += o[i];
				bytes = (288 + 292) * n + 20 * v + 4 * total;
			}
			else if (b->type == 7) {
				const FlatIndex& f = ix->flat;
				for (uint64_t i = 0; i < n; i++) {
					uint32_t rk = b->hx[i] >= f.index_bits ? f.D : host_rank(f, b->hx[i]); if (rk < 1) rk = 1;
					bytes += 148 + 16ull * (f.t7_hi[rk - 1] - f.t7_lo[rk - 1]);
				}
			} else {
				CU(launch_t6(ix->dev, n, b->x.as<uint64_t>(), b->y.as<uint64_t>(), b->out.as<uint32_t>(), b->out.as<uint32_t>() + n, nullptr, nullptr, 0, b->d_status, ix->stream));
				std::vector<uint32_t> o(2 * n); uint64_t total = 0;
				CU(cudaMemcpyAsync(o.data(), b->out.p, n * 8, cudaMemcpyDeviceToHost, ix->stream));
				CU(cudaMemcpyAsync(&total, b->offsets.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, ix->stream));
				CU(cudaStreamSynchronize(ix->stream));
				uint64_t v = 0; for (uint64_t i = 0; i < n; i++) v += o[n + i] - o[i];
				bytes = 292 * n + 20 * v + 4 * total;
			}
			b->algo_bytes = bytes; b->algo_valid = true;
		} catch (const std::exception& e) { return set_err(VSGPU_ENODEVICE, e.what()); }
	}
	*algorithmic_bytes = b->algo_bytes;
	return VSGPU_OK;
}

// Device time of each kernel of the last vsgpu_batch_run, from CUDA events recorded on the launch
// stream (t6/t7: 1 kernel; t4: walk, scan, gather).  Synchronises on the last event.
int vsgpu_batch_timings(vsgpu_batch* b, float* ms, uint32_t cap, uint32_t* n) {
	if (!b || !ms || !n) return set_err(VSGPU_EINVAL, "vsgpu_batch_timings: null argument");
	if (b->launches == 0 || !b->ev[0]) return set_err(VSGPU_EINVAL, "vsgpu_batch_timings: batch has not run");
	if (int rc = check_device(b->idx)) return rc;
	uint32_t k = b->launches;
	if (cudaEventSynchronize(b->ev[k]) != cudaSuccess) return set_err(VSGPU_ENODEVICE, "CUDA: event synchronize failed");
	*n = 0;
	for (uint32_t i = 0; i < k && i < cap; i++) { if (cudaEventElapsedTime(&ms[i], b->ev[i], b->ev[i + 1]) != cudaSuccess) return set_err(VSGPU_ENODEVICE, "CUDA: event elapsed failed"); (*n)++; }
	return VSGPU_OK;
}

void vsgpu_batch_free(vsgpu_batch* b) { delete b; }

}  // extern "C"
