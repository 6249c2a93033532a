// libvsgpu — several indexes on several GPUs behind one handle (include/vsgpu.h: vsgpu_router_*).
//
// The reference keeps one ser/ directory per contig and queries them as independent processes
// (util.cc:93-96, eval_data_records/evaluation.txt:34); BASELINE.json's north star partitions the work
// "by contig / position shard, with regions routed by the host ... no collective on the hot path".
// This is that host side, in one process: every shard is a vsgpu_index on the GPU the longest-
// processing-time rule gave it (weight = branch records), a call routes every region to the shard that
// owns its (contig, start), one host thread per GPU runs that GPU's shards through the fused host-
// buffer call, and the answers are scattered back into the caller's region order.  Nothing here talks
// to another GPU; the only shared resource is the host's PCIe / memory path.
#include "thread_pool.h"
#include "../../include/vsgpu.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

namespace {
thread_local std::string g_rerr;
int rerr(int code, const std::string& m) { g_rerr = m; return code; }

struct Pinned {          // grow-only page-locked buffer
	void* p = nullptr; size_t cap = 0;
	void* ensure(size_t bytes) {
		if (bytes <= cap) return p;
		if (p) cudaFreeHost(p);
		p = nullptr; cap = 0;
		size_t want = 4096; while (want < bytes) want <<= 1;
		if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); p = nullptr; return nullptr; }
		cap = want; return p;
	}
	~Pinned() { if (p) cudaFreeHost(p); }
};

struct Shard {
	vsgpu_index* ix = nullptr;
	int device = 0;
	uint32_t contig = 0;                 // index into the router's contig table
	uint64_t lo = 0, hi = ~(uint64_t)0;  // owns region starts in [lo, hi)
	uint64_t weight = 0;                 // branch records (the LPT weight)
	std::string prefix;
	// per call
	uint64_t n = 0, first = 0;           // regions routed here, and where they start in the routed order
	Pinned px, py, ps, plo, pc6;
	vsgpu_result* res = nullptr;
	int rc = 0; std::string err;
	double ms = 0;
};
}  // namespace

struct vsgpu_router {
	std::vector<Shard> shards;
	std::vector<std::string> contigs;
	std::map<std::string, uint32_t> contig_id;
	std::vector<std::vector<uint32_t>> by_contig;     // shard ids of a contig, ascending lo
	// the same as flat arrays for the routing loop: contig c owns rt_lo / rt_hi / rt_shard [rt_begin[c], rt_begin[c+1])
	std::vector<uint32_t> rt_begin, rt_shard; std::vector<uint64_t> rt_lo, rt_hi;
	std::vector<int> devices;                         // distinct devices in use
	std::vector<double> device_ms; std::vector<uint64_t> device_regions;
	double route_ms = 0, scatter_ms = 0;
	std::mutex mu;
	// last call: where region i sits (shard_of_last[i], slot[i]); the hit codes stay in the shards' page-locked results and
	// are gathered into one CSR in the caller's order only when vsgpu_router_offsets / _hits ask for it
	std::vector<uint32_t> slot; std::vector<uint32_t> shard_of_last; uint64_t last_n = 0; bool csr_valid = false;
	std::vector<uint32_t> hits; std::vector<uint64_t> offsets;
	// workers of the routing / scatter / gather loops (thread_pool.h): kept for the router's lifetime
	vsgpu::ThreadPool pool{std::min(63u, std::max(1u, std::thread::hardware_concurrency()) - 1)};
	~vsgpu_router() { for (auto& s : shards) { if (s.res) vsgpu_result_free(s.res); if (s.ix) vsgpu_close(s.ix); } }
};

namespace {
double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

extern "C" {

const char* vsgpu_router_last_error(void) { return g_rerr.c_str(); }

int vsgpu_router_open(uint32_t nshards, const char* const* ser_prefixes, const uint64_t* range_lo, const uint64_t* range_hi,
                      const int* devices, int ndevices, vsgpu_router** out) {
	if (!out || !nshards || !ser_prefixes) return rerr(VSGPU_EINVAL, "vsgpu_router_open: null argument");
	*out = nullptr;
	int have = 0;
	if (cudaGetDeviceCount(&have) != cudaSuccess || have <= 0) { cudaGetLastError(); return rerr(VSGPU_ENODEVICE, "vsgpu_router_open: no usable CUDA device (libvsgpu has no CPU path)"); }
	if (ndevices <= 0 || ndevices > have) ndevices = have;
	std::unique_ptr<vsgpu_router> r(new vsgpu_router);
	r->shards.resize(nshards);
	for (uint32_t k = 0; k < nshards; k++) {
		Shard& s = r->shards[k];
		s.prefix = ser_prefixes[k];
		if (range_lo) s.lo = range_lo[k];
		if (range_hi && range_hi[k]) s.hi = range_hi[k];
		if (s.lo >= s.hi) return rerr(VSGPU_EINVAL, "vsgpu_router_open: empty position range for shard " + std::to_string(k));
		s.device = devices ? devices[k] : -1;
		if (devices && (s.device < 0 || s.device >= have)) return rerr(VSGPU_EINVAL, "vsgpu_router_open: device out of range for shard " + std::to_string(k));
	}
	// Weights before placement: the size of the vertex blocks on disk stands in for the record count (the
	// index is not decoded yet); ties by shard order.  Longest-processing-time: heaviest shard first, each
	// to the GPU with the least load so far.
	if (!devices) {
		std::vector<std::pair<uint64_t, uint32_t>> w(nshards);
		for (uint32_t k = 0; k < nshards; k++) {
			uint64_t bytes = 0;
			for (uint64_t b = 0;; b++) { FILE* f = fopen((r->shards[k].prefix + "/vertex_list_" + std::to_string(b) + ".proto").c_str(), "rb"); if (!f) break; fseek(f, 0, SEEK_END); bytes += (uint64_t)ftell(f); fclose(f); }
			w[k] = {bytes, k};
		}
		std::stable_sort(w.begin(), w.end(), [](const auto& a, const auto& b) { return a.first > b.first; });
		std::vector<uint64_t> load(ndevices, 0);
		for (auto& [bytes, k] : w) { int g = (int)(std::min_element(load.begin(), load.end()) - load.begin()); r->shards[k].device = g; load[g] += std::max<uint64_t>(bytes, 1); }
	}
	// open: one host thread per GPU, its shards one after the other (vsgpu_open decodes with all cores itself)
	std::vector<int> devs;
	for (auto& s : r->shards) if (std::find(devs.begin(), devs.end(), s.device) == devs.end()) devs.push_back(s.device);
	std::sort(devs.begin(), devs.end());
	r->devices = devs;
	std::vector<std::thread> th;
	for (int d : devs) th.emplace_back([&, d]() {
		for (auto& s : r->shards) if (s.device == d) { s.rc = vsgpu_open(s.prefix.c_str(), d, &s.ix); if (s.rc) { s.err = vsgpu_last_error(); return; } }
	});
	for (auto& t : th) t.join();
	for (uint32_t k = 0; k < nshards; k++) if (r->shards[k].rc) return rerr(r->shards[k].rc, "shard " + std::to_string(k) + " (" + r->shards[k].prefix + "): " + r->shards[k].err);
	// contig table from the indexes themselves (sampleid_map.lst names the contig, variant_graph.h:541-556)
	for (uint32_t k = 0; k < nshards; k++) {
		Shard& s = r->shards[k];
		vsgpu_info_t inf; vsgpu_info(s.ix, &inf);
		s.weight = inf.branch_records;
		auto it = r->contig_id.find(inf.chr);
		if (it == r->contig_id.end()) { it = r->contig_id.emplace(inf.chr, (uint32_t)r->contigs.size()).first; r->contigs.push_back(inf.chr); r->by_contig.emplace_back(); }
		s.contig = it->second;
		r->by_contig[s.contig].push_back(k);
	}
	for (auto& v : r->by_contig) {
		std::sort(v.begin(), v.end(), [&](uint32_t a, uint32_t b) { return r->shards[a].lo < r->shards[b].lo; });
		for (size_t i = 1; i < v.size(); i++) if (r->shards[v[i]].lo < r->shards[v[i - 1]].hi) return rerr(VSGPU_EINVAL, "vsgpu_router_open: position ranges of contig " + r->contigs[r->shards[v[i]].contig] + " overlap");
	}
	r->rt_begin.assign(1, 0);
	for (auto& v : r->by_contig) {
		for (uint32_t k : v) { r->rt_shard.push_back(k); r->rt_lo.push_back(r->shards[k].lo); r->rt_hi.push_back(r->shards[k].hi); }
		r->rt_begin.push_back((uint32_t)r->rt_shard.size());
	}
	r->device_ms.assign(devs.size(), 0); r->device_regions.assign(devs.size(), 0);
	*out = r.release();
	return VSGPU_OK;
}

void vsgpu_router_close(vsgpu_router* r) { delete r; }
uint32_t vsgpu_router_num_shards(const vsgpu_router* r) { return r ? (uint32_t)r->shards.size() : 0; }
uint32_t vsgpu_router_num_contigs(const vsgpu_router* r) { return r ? (uint32_t)r->contigs.size() : 0; }
const char* vsgpu_router_contig_name(const vsgpu_router* r, uint32_t id) { return (r && id < r->contigs.size()) ? r->contigs[id].c_str() : nullptr; }
int vsgpu_router_contig_id(const vsgpu_router* r, const char* name, uint32_t* id) {
	if (!r || !name || !id) return rerr(VSGPU_EINVAL, "vsgpu_router_contig_id: null argument");
	auto it = r->contig_id.find(name);
	if (it == r->contig_id.end()) return rerr(VSGPU_EINVAL, std::string("no shard holds contig ") + name);
	*id = it->second; return VSGPU_OK;
}
vsgpu_index* vsgpu_router_shard_index(const vsgpu_router* r, uint32_t shard) { return (r && shard < r->shards.size()) ? r->shards[shard].ix : nullptr; }
int vsgpu_router_shard_device(const vsgpu_router* r, uint32_t shard) { return (r && shard < r->shards.size()) ? r->shards[shard].device : -1; }

int vsgpu_router_query_t6t4(vsgpu_router* r, uint64_t n, const uint32_t* contig, const uint32_t* x, const uint32_t* y, const uint32_t* sample_ids,
                            uint32_t* shard_of, uint32_t* rec_lo, uint32_t* counts6, uint32_t* counts4) {
	if (!r || (n && (!contig || !x || !y || !sample_ids || !shard_of || !rec_lo || !counts6 || !counts4))) return rerr(VSGPU_EINVAL, "vsgpu_router_query_t6t4: null argument");
	std::lock_guard<std::mutex> g(r->mu);
	const uint32_t S = (uint32_t)r->shards.size();
	const double t0 = now_ms();
	// ---- route: shard of every region (contig, then the position range holding its start); counting sort keeps the order
	const unsigned NT = 32;
	std::vector<std::vector<uint64_t>> cnt(NT, std::vector<uint64_t>(S, 0));
	std::atomic<int64_t> bad{-1};
	const uint32_t nc = (uint32_t)r->by_contig.size();
	const uint32_t* rtb = r->rt_begin.data(); const uint32_t* rts = r->rt_shard.data(); const uint64_t* rtl = r->rt_lo.data(); const uint64_t* rth = r->rt_hi.data();
	r->pool.par_for(n, NT, [&](unsigned t, uint64_t a, uint64_t b) {
		uint64_t* mycnt = cnt[t].data();
		for (uint64_t i = a; i < b; i++) {
			const uint32_t c = contig[i];
			uint32_t k = VSGPU_NONE;
			if (c < nc) {
				uint32_t lo = rtb[c], hi = rtb[c + 1];
				const uint64_t xi = x[i];
				while (hi - lo > 1) { const uint32_t m = (lo + hi) >> 1; if (rtl[m] <= xi) lo = m; else hi = m; }     // last range starting at or before x
				if (lo < rtb[c + 1] && xi >= rtl[lo] && xi < rth[lo]) k = rts[lo];
			}
			if (k == VSGPU_NONE) { int64_t e = -1; bad.compare_exchange_strong(e, (int64_t)i); k = 0; }
			shard_of[i] = k; mycnt[k]++;
		}
	});
	if (bad.load() >= 0) return rerr(VSGPU_EINVAL, "region " + std::to_string(bad.load()) + ": no shard owns this contig / start position");
	// exclusive offsets per (thread, shard) in routed order
	std::vector<uint64_t> total(S, 0);
	for (uint32_t k = 0; k < S; k++) for (unsigned t = 0; t < NT; t++) { const uint64_t c = cnt[t][k]; cnt[t][k] = total[k]; total[k] += c; }
	for (uint32_t k = 0; k < S; k++) {
		Shard& s = r->shards[k];
		s.n = total[k]; s.rc = 0; s.ms = 0;
		if (s.res) { vsgpu_result_free(s.res); s.res = nullptr; }
		if (!s.n) continue;
		cudaSetDevice(s.device);
		if (!s.px.ensure(s.n * 4) || !s.py.ensure(s.n * 4) || !s.ps.ensure(s.n * 4) || !s.plo.ensure(s.n * 4) || !s.pc6.ensure(s.n * 4)) return rerr(VSGPU_ENOMEM, "cannot allocate page-locked routing buffers");
	}
	if (r->slot.size() < n) r->slot.resize(n);     // position of region i inside its shard's batch
	uint32_t* slot = r->slot.data();
	std::vector<uint32_t*> bx(S), by(S), bs(S);
	for (uint32_t k = 0; k < S; k++) { bx[k] = (uint32_t*)r->shards[k].px.p; by[k] = (uint32_t*)r->shards[k].py.p; bs[k] = (uint32_t*)r->shards[k].ps.p; }
	r->pool.par_for(n, NT, [&](unsigned t, uint64_t a, uint64_t b) {
		uint64_t* mycnt = cnt[t].data();
		for (uint64_t i = a; i < b; i++) {
			const uint32_t k = shard_of[i];
			const uint64_t j = mycnt[k]++;
			slot[i] = (uint32_t)j;
			bx[k][j] = x[i]; by[k][j] = y[i]; bs[k][j] = sample_ids[i];
		}
	});
	const double t1 = now_ms();
	r->route_ms = t1 - t0;
	// ---- host threads per GPU (VSGPU_ROUTER_THREADS, default 4) take that GPU's shards from a queue: every shard is its
	// own index with its own streams, so the copies of one shard's call overlap the kernels of another's
	unsigned per_gpu = 4;
	if (const char* e = getenv("VSGPU_ROUTER_THREADS")) per_gpu = (unsigned)std::max(1, atoi(e));
	std::vector<std::atomic<uint32_t>> next(r->devices.size());
	std::vector<std::atomic<uint64_t>> dregions(r->devices.size());
	std::vector<double> dstart(r->devices.size(), 0), dend(r->devices.size(), 0);
	std::vector<std::vector<uint32_t>> queue(r->devices.size());
	for (size_t di = 0; di < r->devices.size(); di++) {
		next[di] = 0; dregions[di] = 0;
		for (uint32_t k = 0; k < S; k++) if (r->shards[k].device == r->devices[di] && r->shards[k].n) queue[di].push_back(k);
		std::stable_sort(queue[di].begin(), queue[di].end(), [&](uint32_t a, uint32_t b) { return r->shards[a].n > r->shards[b].n; });   // largest batch first
	}
	std::mutex end_mu;
	// (device, worker) pairs as pool tasks: each drains its device's queue; the calls block on CUDA, the pool's threads wait with them
	std::vector<std::pair<size_t, unsigned>> tasks;
	for (size_t di = 0; di < r->devices.size(); di++) {
		dstart[di] = now_ms();
		for (unsigned w = 0; w < std::min<size_t>(per_gpu, std::max<size_t>(queue[di].size(), 1)); w++) tasks.emplace_back(di, w);
	}
	// interleave the devices so that a pool smaller than the task list still starts every GPU at once
	std::stable_sort(tasks.begin(), tasks.end(), [](const std::pair<size_t, unsigned>& a, const std::pair<size_t, unsigned>& b) { return a.second < b.second; });
	r->pool.run((unsigned)tasks.size(), [&](unsigned t) {
		const size_t di = tasks[t].first;
		for (uint32_t qi; (qi = next[di]++) < queue[di].size();) {
			Shard& s = r->shards[queue[di][qi]];
			const double b = now_ms();
			s.rc = vsgpu_query_t6t4_u32(s.ix, s.n, (const uint32_t*)s.px.p, (const uint32_t*)s.py.p, (const uint32_t*)s.ps.p, (uint32_t*)s.plo.p, nullptr, (uint32_t*)s.pc6.p, &s.res);
			if (s.rc) s.err = vsgpu_last_error();
			s.ms = now_ms() - b;
			dregions[di] += s.n;
		}
		const double e = now_ms();
		std::lock_guard<std::mutex> g2(end_mu);
		dend[di] = std::max(dend[di], e);
	});
	for (size_t di = 0; di < r->devices.size(); di++) { r->device_ms[di] = queue[di].empty() ? 0 : dend[di] - dstart[di]; r->device_regions[di] = dregions[di]; }
	for (uint32_t k = 0; k < S; k++) if (r->shards[k].rc) return rerr(r->shards[k].rc, "shard " + std::to_string(k) + ": " + r->shards[k].err);
	const double t2 = now_ms();
	// ---- scatter the per-region words back into the caller's order (the hit codes stay where the GPUs' copies put them)
	std::vector<const uint32_t*> c4(S, nullptr);
	for (uint32_t k = 0; k < S; k++) if (r->shards[k].n) c4[k] = vsgpu_result_counts(r->shards[k].res);
	std::vector<const uint32_t*> plo(S), pc6(S);
	for (uint32_t k = 0; k < S; k++) { plo[k] = (const uint32_t*)r->shards[k].plo.p; pc6[k] = (const uint32_t*)r->shards[k].pc6.p; }
	r->pool.par_for(n, 64, [&](unsigned, uint64_t a, uint64_t b) {
		for (uint64_t i = a; i < b; i++) {
			const uint32_t k = shard_of[i]; const uint32_t j = slot[i];
			rec_lo[i] = plo[k][j]; counts6[i] = pc6[k][j]; counts4[i] = c4[k][j];
		}
	});
	if (r->shard_of_last.size() < n) r->shard_of_last.resize(n);
	r->pool.par_for(n, 64, [&](unsigned, uint64_t a, uint64_t b) { memcpy(r->shard_of_last.data() + a, shard_of + a, (b - a) * 4); });
	r->csr_valid = false; r->last_n = n;
	r->scatter_ms = now_ms() - t2;
	return VSGPU_OK;
}

namespace {
void gather_csr(vsgpu_router* r) {
	if (r->csr_valid) return;
	const uint64_t n = r->last_n; const uint32_t S = (uint32_t)r->shards.size();
	std::vector<const uint32_t*> c4(S, nullptr), hs(S, nullptr); std::vector<const uint64_t*> so(S, nullptr);
	for (uint32_t k = 0; k < S; k++) if (r->shards[k].n && r->shards[k].res) { c4[k] = vsgpu_result_counts(r->shards[k].res); so[k] = vsgpu_result_offsets(r->shards[k].res); hs[k] = vsgpu_result_hits(r->shards[k].res); }
	r->offsets.resize(n + 1);
	uint64_t acc = 0;
	for (uint64_t i = 0; i < n; i++) { r->offsets[i] = acc; acc += c4[r->shard_of_last[i]][r->slot[i]]; }
	r->offsets[n] = acc;
	r->hits.resize(acc);
	r->pool.par_for(n, 64, [&](unsigned, uint64_t a, uint64_t b) {
		for (uint64_t i = a; i < b; i++) {
			const uint32_t k = r->shard_of_last[i]; const uint32_t j = r->slot[i]; const uint32_t c = c4[k][j];
			if (c) memcpy(r->hits.data() + r->offsets[i], hs[k] + so[k][j], (size_t)c * 4);
		}
	});
	r->csr_valid = true;
}
}  // namespace
const uint64_t* vsgpu_router_offsets(const vsgpu_router* cr) { vsgpu_router* r = const_cast<vsgpu_router*>(cr); if (!r) return nullptr; std::lock_guard<std::mutex> g(r->mu); gather_csr(r); return r->offsets.data(); }
const uint32_t* vsgpu_router_hits(const vsgpu_router* cr) { vsgpu_router* r = const_cast<vsgpu_router*>(cr); if (!r) return nullptr; std::lock_guard<std::mutex> g(r->mu); gather_csr(r); return r->hits.data(); }
int vsgpu_router_region_hits(const vsgpu_router* r, uint64_t i, const uint32_t** hits, uint32_t* count) {
	if (!r || !hits || !count || i >= r->last_n) return rerr(VSGPU_EINVAL, "vsgpu_router_region_hits: bad argument");
	const Shard& s = r->shards[r->shard_of_last[i]];
	const uint64_t j = r->slot[i];
	*count = vsgpu_result_counts(s.res)[j];
	*hits = vsgpu_result_hits(s.res) + vsgpu_result_offsets(s.res)[j];
	return VSGPU_OK;
}

int vsgpu_router_stats(const vsgpu_router* r, uint32_t cap, int* devices, double* device_ms, uint64_t* device_regions, uint32_t* ndev, double* route_ms, double* scatter_ms) {
	if (!r) return rerr(VSGPU_EINVAL, "vsgpu_router_stats: null router");
	const uint32_t nd = (uint32_t)r->devices.size();
	if (ndev) *ndev = nd;
	for (uint32_t i = 0; i < nd && i < cap; i++) { if (devices) devices[i] = r->devices[i]; if (device_ms) device_ms[i] = r->device_ms[i]; if (device_regions) device_regions[i] = r->device_regions[i]; }
	if (route_ms) *route_ms = r->route_ms;
	if (scatter_ms) *scatter_ms = r->scatter_ms;
	return VSGPU_OK;
}

}  // extern "C"
