// libvsgpu device side — hand-written sm_100a kernels for the batched region path.
//
//   K1  rank over the distinct backbone starts (replaces Index::find / Index::is_empty,
//       include/index.h:119-166, i.e. rrr_vector rank/select): direct-mapped position buckets +
//       a short binary search inside one or two 128-byte lines of `dstart`.
//   K2  k_t6: slice bounds of get_var_in_ref (include/query.h:736-784).
//   K3  k_t4: get_prev_vertex_with_sample + get_sample_var_in_ref (include/query.h:57-113, :618-729)
//       fused with K5: every thread walks one region (membership through the sample-major hit map,
//       chain rules resolved in order), the CTA scans its hit counts, a decoupled look-back over the
//       tile states turns them into global offsets, and the hits are written straight to their final
//       position — output order equals the reference's push order, no atomically-ordered output.
//   K4  k_t7: samples_has_var lookup (include/query.h:792-823).
//   K6  k_build_hitmap: transposes (walk entry -> carrier set) into the sample-major hit map, once
//       per vsgpu_open.
// Integer-only; nothing here is a contraction, so no tensor-core path exists.
#include <cstdlib>
#include <type_traits>

#include "device_logic.cuh"

namespace vsgpu {
namespace {
using namespace logic;

// CTA barrier that does not need the warp converged (PTX barrier.sync without .aligned; __syncthreads() is the
// .aligned form).  The per-region walks leave their loops from many places; with the aligned barrier, compute-sanitizer
// synccheck reported "divergent thread(s) in warp" at the barrier behind the walk for batches large enough that a CTA
// takes several tiles, and the staged hits of one tile were then overwritten by the next (tools/debug_t4.py).
__device__ __forceinline__ void cta_sync() { asm volatile("barrier.sync 0;" ::: "memory"); }
// Every kernel below that walks regions and then meets at a CTA barrier uses it.


// counts[i] = rows the reference returns when that is the slice length; regions whose slice needs the
// literal dedup rule (a suspect duplicate inside, or a region running past the contig end over tail
// records — DESIGN.md section 9) are appended to `flagged` (count in status[1]) for the host to re-count.
__global__ void __launch_bounds__(256) k_t6(const DevIndex ix, uint64_t n, const uint64_t* __restrict__ xs,
                                            const uint64_t* __restrict__ ys, uint32_t* __restrict__ lo, uint32_t* __restrict__ hi,
                                            uint32_t* __restrict__ counts, uint32_t* __restrict__ flagged, uint32_t flag_base, uint32_t* status) {
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		bool bad = false;
		const uint64_t y = ys[i];
		const uint2 r = t6_bounds(ix, xs[i], y, &bad);
		lo[i] = r.x; hi[i] = r.y;
		const uint32_t c = r.x == kNoneU32 ? 0 : r.y - r.x;
		if (counts) counts[i] = c;
		if (flagged && c) {
			const bool literal = (ix.rec_dup_prefix && __ldg(ix.rec_dup_prefix + r.y) != __ldg(ix.rec_dup_prefix + r.x)) || (ix.tail_records && y > ix.last_end);
			if (literal) flagged[atomicAdd(status + 1, 1u)] = flag_base + (uint32_t)i;
		}
		if (bad) atomicOr(status, kStatusBadRegion);
	}
}

__global__ void __launch_bounds__(256) k_t7(const DevIndex ix, uint64_t n, const uint64_t* __restrict__ pos,
                                            const uint64_t* __restrict__ qhash, uint32_t* __restrict__ rec, uint32_t* status) {
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		bool bad = false;
		rec[i] = t7_lookup(ix, pos[i], qhash[i], &bad);
		if (bad) atomicOr(status, kStatusBadRegion);
	}
}

__global__ void __launch_bounds__(256) k_t1(const DevIndex ix, uint64_t n, const uint64_t* __restrict__ pos,
                                            uint32_t* __restrict__ lo, uint32_t* __restrict__ hi, uint32_t* status) {
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		bool bad = false;
		const uint2 r = t1_lookup(ix, pos[i], &bad);
		lo[i] = r.x; hi[i] = r.y;
		if (bad) atomicOr(status, kStatusBadRegion);
	}
}

// ------------------------------------------------------------------ t4: walk + ordered compaction
#define kFlagAgg (1ull << 62)
#define kFlagIncl (2ull << 62)
#define kValMask ((1ull << 62) - 1)

template <uint32_t kTile, uint32_t kKeep>
struct SmemSink {            // first kKeep codes of this thread, strided so lanes hit distinct banks
	uint32_t* slot; uint32_t n;
	__device__ __forceinline__ void emit(uint32_t code) { if (n < kKeep) slot[n * kTile] = code; n++; }
};

// The same, for batches whose regions have more rows than the staging holds: codes beyond the first kKeep go to this CTA's
// scratch in global memory, in chunks of 7 codes + the offset of the next chunk (32 bytes, one sector), handed out by a
// shared-memory cursor.  After the look-back the thread copies its chain into place instead of walking the region again.
// A thread that finds the scratch full stops spilling and takes the second walk.  (Chunks that double in size, so that the
// copy follows fewer links, were slower: they fill the scratch sooner — profiles/r2_spill.txt.)
constexpr uint32_t kSpillChunk = 8;
template <uint32_t kTile, uint32_t kKeep>
struct SpillSink {
	uint32_t* slot; uint32_t n;
	uint32_t* spill; uint32_t* cursor; uint32_t cap;
	uint32_t first, cur, r; bool ok;
	__device__ __forceinline__ void emit(uint32_t code) {
		if (n < kKeep) slot[n * kTile] = code;
		else if (ok) {
			if (r == kSpillChunk - 1 || n == kKeep) {
				const uint32_t c = atomicAdd(cursor, kSpillChunk);
				if (c + kSpillChunk > cap) ok = false;
				else { if (n == kKeep) first = c; else spill[cur + kSpillChunk - 1] = c; cur = c; r = 0; }
			}
			if (ok) spill[cur + r++] = code;
		}
		n++;
	}
};

// A batch may be launched as several chunks of regions (so that transfers overlap the kernels); the
// offsets of chunk c continue from the total of chunk c-1, which that launch left in *base_ptr.
__device__ __forceinline__ uint64_t chunk_base(const uint64_t* base_ptr) { return base_ptr ? *(const volatile uint64_t*)base_ptr : 0; }

__device__ __forceinline__ uint64_t warp_sum(uint64_t v) {
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
	return v;
}

// ------------------------------------------------------------------ wide regions: one warp per region
// A region whose scan range is long keeps one thread busy for thousands of dependent loads while the
// rest of its warp idles.  Here the whole warp takes such a region: 32 row words per load, the
// carried entries of the window fetched 32 at a time, and the chain rules resolved in parallel —
// every lane decides "hidden?" from an exclusive max-scan of the detour targets before it, checked
// for self-consistency (a hidden entry must not have contributed), falling back to a lane-by-lane
// pass for the batch when that check or a rare entry kind (rejoin carrier, walk past the bound) says so.
// Output order is kept with ballot/popc ranks.  Same answers as fast_forward.
// Codes beyond the `keep` staged in shared memory go to chunks of kPoolChunk codes taken from a pool in global memory (one
// atomicAdd per chunk by lane 0; the warp's chunk ids sit in shared memory), so that a region with more rows than the staging
// holds is copied into place — coalesced, by its warp — instead of walked twice.  reserve() is called by all lanes before
// positions up to `upto` are written; a warp that finds the pool (or its chunk list) exhausted clears *ok and re-walks.
constexpr uint32_t kPoolChunk = 1024, kMaxWarpChunks = 128;
struct CoopSink {            // all lanes call emit with identical arguments; lane 0 writes
	uint32_t* slot; uint32_t stride, keep; uint32_t* direct; uint32_t n; bool writer;
	uint32_t* pool; unsigned long long* cursor; uint32_t pool_chunks; volatile uint32_t* chunk_ids; volatile uint32_t* nchunks; volatile uint32_t* ok;
	__device__ __forceinline__ void reserve(uint32_t upto) {
		if (direct || !pool || upto <= keep) return;
		while (*ok && upto > keep + *nchunks * kPoolChunk) {
			uint32_t c = 0xFFFFFFFFu;
			if (*nchunks < kMaxWarpChunks && (threadIdx.x & 31) == 0) c = (uint32_t)atomicAdd(cursor, 1ull);
			c = __shfl_sync(0xFFFFFFFFu, c, 0);
			__syncwarp();
			if ((threadIdx.x & 31) == 0) { if (c >= pool_chunks) *ok = 0; else { chunk_ids[*nchunks] = c; *nchunks = *nchunks + 1; } }
			__syncwarp();
		}
	}
	__device__ __forceinline__ void put(uint32_t pos, uint32_t code) {
		if (direct) direct[pos] = code;
		else if (pos < keep) slot[pos * stride] = code;
		else if (pool && *ok) { const uint32_t k = pos - keep; pool[(uint64_t)chunk_ids[k / kPoolChunk] * kPoolChunk + (k % kPoolChunk)] = code; }
	}
	__device__ __forceinline__ void emit(uint32_t code) { reserve(n + 1); if (writer) put(n, code); n++; }
};

__device__ __forceinline__ uint32_t warp_excl_max(uint32_t v, uint32_t lane) {   // exclusive prefix max, identity 0
	uint32_t inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= (uint32_t)d) inc = max(inc, t); }
	const uint32_t ex = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
	return lane ? ex : 0;
}

__device__ __noinline__ uint32_t coop_forward(const DevIndex& ix, FwdState st, uint32_t s, CoopSink sink) {
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t lt = (1u << lane) - 1;
	while (st.c < st.limit) {
		const uint32_t limit0 = st.limit;
		const uint32_t w0 = st.c >> 5, w = w0 + lane;
		uint32_t m = ((uint64_t)w << 5) < limit0 ? (__ldg(st.row + w) | __ldg(ix.marker_bits + w)) : 0;
		if (lane == 0) m &= 0xFFFFFFFFu << (st.c & 31);
		const uint32_t cnt = __popc(m);
		uint32_t incl = cnt;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= (uint32_t)d) incl += t; }
		const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31), excl = incl - cnt;
		bool restart = false;
		for (uint32_t base = 0; base < total && !restart; base += 32) {
			// lane j takes hit number base + j of the window: owner word = number of lanes whose inclusive count <= h
			const uint32_t h = base + lane;
			const bool valid = h < total;
			uint32_t o = 0;
#pragma unroll
			for (int step = 16; step > 0; step >>= 1) { const uint32_t v = __shfl_sync(0xFFFFFFFFu, incl, (o + step - 1) & 31); if (v <= h) o += step; }
			o &= 31;
			const uint32_t m_o = __shfl_sync(0xFFFFFFFFu, m, o), excl_o = __shfl_sync(0xFFFFFFFFu, excl, o);
			const uint32_t ci = valid ? ((w0 + o) << 5) + __fns(m_o, 0, (int)(h - excl_o) + 1) : 0;
			const uint4 e = valid ? __ldg(ix.cent + ci) : make_uint4(0, 0, 0, 0);
			const uint32_t nb = min(32u, total - base);
			// ---- per-lane facts
			const bool marker = (e.y & kEntMarker) != 0, isalt = (e.y & kEntAlt) != 0;
			const uint32_t tk = e.y & kEntTgtMask;
			const uint32_t jump = (valid && !marker) ? tk : 0;             // where the walk continues if this entry is taken
			const bool beyond = valid && ci >= st.limit;
			// ---- hidden? fixed point of  hidden_j = src_j < max(cur_k, max over earlier non-hidden jumps)
			uint32_t M = max(st.cur_k, warp_excl_max(jump == kEntTgtMask ? 0 : jump, lane));
			bool hidden = valid && e.x < M;
			bool ok = false;
			for (int it = 0; it < 3 && !ok; it++) {
				const uint32_t M2 = max(st.cur_k, warp_excl_max((hidden || jump == kEntTgtMask) ? 0 : jump, lane));
				const bool hidden2 = valid && e.x < M2;
				ok = !__any_sync(0xFFFFFFFFu, hidden2 != hidden);
				hidden = hidden2; M = M2;
			}
			const bool stop_before = valid && (beyond || (!hidden && ((e.x > M && e.x >= st.k_end) || e.w >= st.y)));
			const bool live = valid && !hidden && !marker && !stop_before;
			const bool stop_after = live && isalt && (tk == kEntTgtMask || tk >= st.k_end);
			const bool rare = live && ((isalt && (e.y & kEntTgtCarriers)) || (!isalt && tk >= st.k_end));
			const uint32_t stop_mask = __ballot_sync(0xFFFFFFFFu, stop_before || stop_after);
			const uint32_t first_stop = stop_mask ? (uint32_t)__ffs((int)stop_mask) - 1 : 32;
			const uint32_t upto = first_stop < 32 ? (2u << first_stop) - 1 : 0xFFFFFFFFu;   // lanes 0..first_stop
			const uint32_t rare_mask = __ballot_sync(0xFFFFFFFFu, rare) & upto;
			if (!ok || rare_mask) {
				// lane-by-lane pass over this batch with the single-thread rules (all lanes in lock step)
				for (uint32_t j = 0; j < nb; j++) {
					const uint32_t cj = __shfl_sync(0xFFFFFFFFu, ci, j);
					const uint4 ej = make_uint4(__shfl_sync(0xFFFFFFFFu, e.x, j), __shfl_sync(0xFFFFFFFFu, e.y, j), __shfl_sync(0xFFFFFFFFu, e.z, j), __shfl_sync(0xFFFFFFFFu, e.w, j));
					if (fwd_step(ix, st, s, cj, ej, sink)) return sink.n;
				}
				if (st.limit != limit0) { st.c = __shfl_sync(0xFFFFFFFFu, ci, nb - 1) + 1; restart = true; }   // the bound moved: reload the window
				continue;
			}
			// ---- emit in order
			const bool emit = live && e.w >= st.x && lane <= first_stop && !(lane == first_stop && stop_before);
			const uint32_t emit_mask = __ballot_sync(0xFFFFFFFFu, emit);
			sink.reserve(sink.n + __popc(emit_mask));
			if (emit) sink.put(sink.n + __popc(emit_mask & lt), ci);
			sink.n += __popc(emit_mask);
			if (first_stop < 32) return sink.n;
			// ---- new walk position: the largest taken jump
			const uint32_t Mend = max(M, (hidden || jump == kEntTgtMask) ? 0 : jump);
			st.cur_k = max(st.cur_k, __shfl_sync(0xFFFFFFFFu, Mend, nb - 1));
		}
		if (!restart) st.c = (w0 + 32) << 5;
	}
	return sink.n;
}

// kTile = regions per CTA, one per thread; kKeep = hits per region staged in shared memory (a
// region with more walks a second time, straight into its final place)
template <uint32_t kTile, uint32_t kMinCtas, uint32_t kKeep>
__global__ void __launch_bounds__(kTile, kMinCtas) k_t4(const DevIndex ix, uint64_t n, const uint64_t* __restrict__ xs,
                                              const uint64_t* __restrict__ ys, const uint32_t* __restrict__ sample,
                                              uint64_t* __restrict__ offsets, uint32_t* __restrict__ hits, uint64_t cap,
                                              uint64_t* tile_state, uint32_t* status, const uint64_t* base_ptr) {
	__shared__ uint32_t s_hits[kTile * kKeep];
	__shared__ uint64_t s_warp[kTile / 32];
	__shared__ uint64_t s_base;
	__shared__ uint32_t s_tile;
	if (threadIdx.x == 0) s_tile = (uint32_t)atomicAdd((unsigned long long*)&tile_state[0], 1ull);   // tiles start in ticket order
	cta_sync();
	const uint32_t tile = s_tile, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	volatile uint64_t* state = tile_state + 1;
	const uint64_t i = (uint64_t)tile * kTile + threadIdx.x;

	// ---- phase 1: walk this thread's region
	SmemSink<kTile, kKeep> sink{s_hits + threadIdx.x, 0};
	uint64_t x = 0, y = 0; uint32_t s = 0;
	if (i < n) {
		x = xs[i]; y = ys[i]; s = sample[i];
		if (x < 1 || s == 0 || s >= ix.num_samples) atomicOr(status, kStatusBadRegion);
		else walk_any(ix, x, y, s, sink);
	}
	const uint32_t cnt = sink.n;

	// ---- phase 2: CTA scan of the counts
	uint64_t incl = cnt;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const uint64_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= (uint32_t)d) incl += t; }
	if (lane == 31) s_warp[warp] = incl;
	cta_sync();
	uint64_t wpre = 0, agg = 0;
#pragma unroll
	for (uint32_t w = 0; w < kTile / 32; w++) { if (w < warp) wpre += s_warp[w]; agg += s_warp[w]; }

	// ---- phase 3: decoupled look-back (warp 0), 32 predecessor tiles per step
	if (warp == 0) {
		uint64_t excl = 0;
		if (tile == 0) { excl = chunk_base(base_ptr); if (lane == 0) state[0] = kFlagIncl | (excl + agg); }
		else {
			if (lane == 0) state[tile] = kFlagAgg | agg;
			for (int64_t idx = (int64_t)tile - 1;; idx -= 32) {
				const int64_t j = idx - lane;
				uint64_t st = j >= 0 ? state[j] : kFlagIncl;
				while (__any_sync(0xFFFFFFFFu, (st >> 62) == 0)) { if ((st >> 62) == 0) st = state[j]; }
				const uint32_t incl_mask = __ballot_sync(0xFFFFFFFFu, (st >> 62) == 2);
				const uint64_t v = st & kValMask;
				if (incl_mask) {
					const uint32_t first = (uint32_t)__ffs((int)incl_mask) - 1;   // nearest predecessor with a full prefix
					excl += warp_sum(lane <= first ? v : 0);
					break;
				}
				excl += warp_sum(v);
			}
			if (lane == 0) state[tile] = kFlagIncl | (excl + agg);
		}
		if (lane == 0) s_base = excl;
	}
	cta_sync();

	// ---- phase 4: ordered write
	if (i < n) {
		const uint64_t off = s_base + wpre + (incl - cnt);
		offsets[i] = off;
		if (i == n - 1) offsets[n] = off + cnt;
		if (off + cnt > cap) atomicOr(status, kStatusOverflow);
		else if (cnt <= kKeep) { for (uint32_t j = 0; j < cnt; j++) hits[off + j] = s_hits[j * kTile + threadIdx.x]; }
		else { DirectSink direct{hits + off, 0}; walk_any(ix, x, y, s, direct); }   // more hits than the staging holds: walk again, straight into place
	}
}

// ------------------------------------------------------------------ t4, persistent + pipelined
// Same work as k_t4, but each CTA stays resident and takes tile after tile: the look-back and the
// ordered write of tile k are done after the walk of tile k+1, by which time the predecessors of
// tile k have published their counts — the wait that cost k_t4 a third of its warp time is hidden
// behind useful work.  Two staging buffers alternate.
// k32: region bounds are read as 32-bit arrays (they are parsed with std::stoi, commands.cc:76-80).
// kFuse6: the two ranks of the t4 setup also give the region's t6 slice (get_var_in_ref, query.h:736-784),
// written in the same pass — one kernel answers both operators for a batch that asks for both.
template <bool k32> __device__ __forceinline__ uint64_t ld_coord(const void* a, uint64_t i) { return k32 ? (uint64_t)((const uint32_t*)a)[i] : ((const uint64_t*)a)[i]; }

// kSpill: codes beyond the staging go to a per-CTA scratch (SpillSink) — for batches of regions with many rows.
template <uint32_t kTile, uint32_t kMinCtas, uint32_t kKeep, bool k32, bool kFuse6, bool kSpill>
__global__ void __launch_bounds__(kTile, kMinCtas) k_t4p(const DevIndex ix, uint64_t n, const void* __restrict__ xs, const void* __restrict__ ys,
                                                          const uint32_t* __restrict__ sample, uint64_t* __restrict__ offsets, uint32_t* __restrict__ counts,
                                                          uint32_t* __restrict__ hits, uint64_t cap, uint64_t* tile_state, uint32_t* status, const uint64_t* base_ptr,
                                                          const T6Out t6, uint32_t* spill, uint32_t spill_words) {
	__shared__ uint32_t s_hits[2][kTile * kKeep];
	__shared__ uint32_t s_cursor[2];
	uint32_t* const my_spill = kSpill ? spill + (uint64_t)blockIdx.x * 2 * spill_words : nullptr;     // two halves of spill_words each
	uint32_t p_first = 0; bool p_ok = false;
	__shared__ uint32_t s_warp[kTile / 32];
	__shared__ uint64_t s_base;
	__shared__ uint32_t s_tile;
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t ntiles = (uint32_t)((n + kTile - 1) / kTile);
	volatile uint64_t* state = tile_state + 1;
	// what this thread still owes for the previous tile
	uint32_t p_tile = 0xFFFFFFFFu, p_cnt = 0, p_pre = 0; uint64_t p_agg = 0;
	uint32_t buf = 0;
	for (;;) {
		if (threadIdx.x == 0) { s_tile = (uint32_t)atomicAdd((unsigned long long*)&tile_state[0], 1ull); if (kSpill) s_cursor[buf] = 0; }   // tiles start in ticket order
		cta_sync();
		const uint32_t tile = s_tile;
		const bool has = tile < ntiles;
		uint32_t cnt = 0, pre = 0; uint64_t agg = 0;
		uint32_t c_first = 0; bool c_ok = false;
		if (has) {
			// ---- walk this thread's region of the new tile
			const uint64_t i = (uint64_t)tile * kTile + threadIdx.x;
			typename std::conditional<kSpill, SpillSink<kTile, kKeep>, SmemSink<kTile, kKeep>>::type sink;
			sink.slot = s_hits[buf] + threadIdx.x; sink.n = 0;
			if constexpr (kSpill) { sink.spill = my_spill + buf * spill_words; sink.cursor = &s_cursor[buf]; sink.cap = spill_words; sink.first = 0; sink.cur = 0; sink.r = 0; sink.ok = true; }
			if (i < n) {
				const uint64_t x = ld_coord<k32>(xs, i), y = ld_coord<k32>(ys, i); const uint32_t s = sample[i];
				uint2 r = make_uint2(kNoneU32, kNoneU32);
				if (x < 1 || s == 0 || s >= ix.num_samples) atomicOr(status, kStatusBadRegion);
				else walk_any(ix, x, y, s, sink, kFuse6 ? &r : nullptr);
				if (kFuse6) {
					t6.lo[i] = r.x; if (t6.hi) t6.hi[i] = r.y;
					const uint32_t c = r.x == kNoneU32 ? 0 : r.y - r.x;
					if (t6.counts) t6.counts[i] = c;
					if (t6.flagged && c) {
						const bool literal = (ix.rec_dup_prefix && __ldg(ix.rec_dup_prefix + r.y) != __ldg(ix.rec_dup_prefix + r.x)) || (ix.tail_records && y > ix.last_end);
						if (literal) t6.flagged[atomicAdd(status + 1, 1u)] = t6.flag_base + (uint32_t)i;
					}
				}
			}
			cnt = sink.n;
			if constexpr (kSpill) { c_first = sink.first; c_ok = sink.ok; }
			uint32_t incl = cnt;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= (uint32_t)d) incl += t; }
			if (lane == 31) s_warp[warp] = incl;
			cta_sync();
			uint32_t wpre = 0;
#pragma unroll
			for (uint32_t w = 0; w < kTile / 32; w++) { if (w < warp) wpre += s_warp[w]; agg += s_warp[w]; }
			pre = wpre + incl - cnt;
			if (threadIdx.x == 0) state[tile] = tile == 0 ? (kFlagIncl | (chunk_base(base_ptr) + agg)) : (kFlagAgg | agg);      // publish the count now; the prefix later
		}
		// ---- finish the previous tile: look-back, then the ordered write
		if (p_tile != 0xFFFFFFFFu) {
			if (warp == 0) {
				uint64_t excl = 0;
				if (p_tile == 0) excl = chunk_base(base_ptr);
				else {
					for (int64_t idx = (int64_t)p_tile - 1;; idx -= 32) {
						const int64_t j = idx - lane;
						uint64_t st = j >= 0 ? state[j] : kFlagIncl;
						while (__any_sync(0xFFFFFFFFu, (st >> 62) == 0)) { if ((st >> 62) == 0) st = state[j]; }
						const uint32_t incl_mask = __ballot_sync(0xFFFFFFFFu, (st >> 62) == 2);
						const uint64_t v = st & kValMask;
						if (incl_mask) { const uint32_t first = (uint32_t)__ffs((int)incl_mask) - 1; excl += warp_sum(lane <= first ? v : 0); break; }
						excl += warp_sum(v);
					}
					if (lane == 0) state[p_tile] = kFlagIncl | (excl + p_agg);
				}
				if (lane == 0) s_base = excl;
			}
			cta_sync();
			const uint64_t i = (uint64_t)p_tile * kTile + threadIdx.x;
			if (i < n) {
				const uint64_t off = s_base + p_pre;
				offsets[i] = off;
				if (i == n - 1) offsets[n] = off + p_cnt;
				if (counts) counts[i] = p_cnt;
				if (off + p_cnt > cap) atomicOr(status, kStatusOverflow);
				else if (p_cnt <= kKeep) { const uint32_t* src = s_hits[buf ^ 1] + threadIdx.x; for (uint32_t j = 0; j < p_cnt; j++) hits[off + j] = src[j * kTile]; }
				else if (kSpill && p_ok) {                         // the staged codes, then the chain of spilled chunks
					const uint32_t* src = s_hits[buf ^ 1] + threadIdx.x;
					for (uint32_t j = 0; j < kKeep; j++) hits[off + j] = src[j * kTile];
					const uint32_t* sp = my_spill + (buf ^ 1) * spill_words;
					uint32_t c = p_first;
					for (uint32_t j = kKeep; j < p_cnt; c = sp[c + kSpillChunk - 1])
						for (uint32_t q = 0; q < kSpillChunk - 1 && j < p_cnt; q++, j++) hits[off + j] = sp[c + q];
				}
				else { DirectSink direct{hits + off, 0}; walk_any(ix, ld_coord<k32>(xs, i), ld_coord<k32>(ys, i), sample[i], direct); }   // more hits than the staging holds: walk again, straight into place
			}
		}
		if (!has) break;
		p_tile = tile; p_cnt = cnt; p_pre = pre; p_agg = agg; p_first = c_first; p_ok = c_ok;
		buf ^= 1;
		cta_sync();        // s_tile / s_warp / s_base are reused by the next round
	}
}

// ------------------------------------------------------------------ t4, one warp per region
// For batches of few, wide regions (the scan-bound end of the width sweep): eight regions per CTA,
// every warp runs the setup in lock step and then the cooperative scan above; up to kKeepW hits per
// region are staged in shared memory and copied out coalesced once the look-back has the offset.
template <uint32_t kKeepW>
__global__ void __launch_bounds__(256, 4) k_t4w(const DevIndex ix, uint64_t n, const uint64_t* __restrict__ xs,
                                                const uint64_t* __restrict__ ys, const uint32_t* __restrict__ sample,
                                                uint64_t* __restrict__ offsets, uint32_t* __restrict__ hits, uint64_t cap,
                                                uint64_t* tile_state, uint32_t* status, const uint64_t* base_ptr, uint32_t* pool, uint32_t pool_chunks) {
	constexpr uint32_t kWarps = 8;
	__shared__ uint32_t s_hits[kWarps * kKeepW];
	__shared__ uint32_t s_cnt[kWarps];
	__shared__ uint32_t s_chunk[kWarps][kMaxWarpChunks];
	__shared__ uint32_t s_nchunk[kWarps], s_pool_ok[kWarps];
	if ((threadIdx.x & 31) == 0) { s_nchunk[threadIdx.x >> 5] = 0; s_pool_ok[threadIdx.x >> 5] = pool ? 1 : 0; }
	__shared__ uint64_t s_base;
	__shared__ uint32_t s_tile;
	if (threadIdx.x == 0) s_tile = (uint32_t)atomicAdd((unsigned long long*)&tile_state[0], 1ull);
	cta_sync();
	const uint32_t tile = s_tile, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	volatile uint64_t* state = tile_state + 1;
	const uint64_t i = (uint64_t)tile * kWarps + warp;
	uint64_t x = 0, y = 0; uint32_t s = 0, cnt = 0;
	bool valid = false;
	if (i < n) {
		x = xs[i]; y = ys[i]; s = sample[i];
		if (x < 1 || s == 0 || s >= ix.num_samples) { if (lane == 0) atomicOr(status, kStatusBadRegion); }
		else {
			valid = true;
			CoopSink sink{s_hits + warp * kKeepW, 1, kKeepW, nullptr, 0, lane == 0, pool ? pool + 2 : nullptr, (unsigned long long*)pool, pool_chunks, s_chunk[warp], &s_nchunk[warp], &s_pool_ok[warp]};   // (the pool's first 8 bytes are its cursor)
			FwdState st;
			if (!ix.hitmap) walk_region(ix, x, y, s, sink);
			else if (fast_setup(ix, x, y, s, sink, st)) sink.n = coop_forward(ix, st, s, sink);
			cnt = sink.n;
		}
	}
	if (lane == 0) s_cnt[warp] = cnt;
	cta_sync();
	uint64_t wpre = 0, agg = 0;
#pragma unroll
	for (uint32_t w = 0; w < kWarps; w++) { if (w < warp) wpre += s_cnt[w]; agg += s_cnt[w]; }
	if (warp == 0) {
		uint64_t excl = 0;
		if (tile == 0) { excl = chunk_base(base_ptr); if (lane == 0) state[0] = kFlagIncl | (excl + agg); }
		else {
			if (lane == 0) state[tile] = kFlagAgg | agg;
			for (int64_t idx = (int64_t)tile - 1;; idx -= 32) {
				const int64_t j = idx - lane;
				uint64_t stt = j >= 0 ? state[j] : kFlagIncl;
				while (__any_sync(0xFFFFFFFFu, (stt >> 62) == 0)) { if ((stt >> 62) == 0) stt = state[j]; }
				const uint32_t incl_mask = __ballot_sync(0xFFFFFFFFu, (stt >> 62) == 2);
				const uint64_t v = stt & kValMask;
				if (incl_mask) { const uint32_t first = (uint32_t)__ffs((int)incl_mask) - 1; excl += warp_sum(lane <= first ? v : 0); break; }
				excl += warp_sum(v);
			}
			if (lane == 0) state[tile] = kFlagIncl | (excl + agg);
		}
		if (lane == 0) s_base = excl;
	}
	cta_sync();
	if (i < n) {
		const uint64_t off = s_base + wpre;
		if (lane == 0) { offsets[i] = off; if (i == n - 1) offsets[n] = off + cnt; }
		if (off + cnt > cap) { if (lane == 0) atomicOr(status, kStatusOverflow); }
		else if (cnt <= kKeepW) { for (uint32_t j = lane; j < cnt; j += 32) hits[off + j] = s_hits[warp * kKeepW + j]; }
		else if (valid && pool && s_pool_ok[warp]) {             // the staged codes, then the warp's pool chunks, 128 bytes per step
			for (uint32_t j = lane; j < kKeepW; j += 32) hits[off + j] = s_hits[warp * kKeepW + j];
			const uint32_t* codes = pool + 2;
			for (uint32_t k = lane; k < cnt - kKeepW; k += 32) hits[off + kKeepW + k] = codes[(uint64_t)s_chunk[warp][k / kPoolChunk] * kPoolChunk + (k % kPoolChunk)];
		}
		else if (valid) {
			CoopSink direct{nullptr, 0, 0, hits + off, 0, lane == 0};
			FwdState st;
			if (!ix.hitmap) walk_region(ix, x, y, s, direct);
			else if (fast_setup(ix, x, y, s, direct, st)) coop_forward(ix, st, s, direct);
		}
	}
}


// ------------------------------------------------------------------ t6 rows as text, on the device
// print_var (query.h:43-50) for the records of get_var_in_ref: the carriers of a row are the set
// bits of its class bitmap (or its id list), each printed as name(gt) with the phasing flags of the
// matching s_info entry (get_samples, query.h:268-285; get_sample_phasing, variant_graph.h:882-900).
// Row lengths are static, so every row knows where its text starts; a warp writes one row.
struct SegCount { uint64_t rows, bytes; };
__device__ __forceinline__ SegCount seg_count(const uint64_t* __restrict__ tp, const uint32_t* lo, const uint32_t* hi, uint64_t i, uint64_t nseg) {
	SegCount c{0, 0};
	if (i < nseg) {
		const uint32_t a = lo[i], b = hi[i];
		if (a != kNoneU32 && b > a) { c.rows = b - a; c.bytes = __ldg(tp + b) - __ldg(tp + a); }
	}
	return c;
}
__device__ __forceinline__ SegCount cta_scan_1024(SegCount mine, SegCount* s_warp, SegCount* total) {   // 256 threads; returns the exclusive prefix of `mine`
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	SegCount inc = mine;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint64_t r = __shfl_up_sync(0xFFFFFFFFu, inc.rows, d), b = __shfl_up_sync(0xFFFFFFFFu, inc.bytes, d);
		if (lane >= (uint32_t)d) { inc.rows += r; inc.bytes += b; }
	}
	if (lane == 31) s_warp[warp] = inc;
	cta_sync();
	SegCount pre{0, 0}, tot{0, 0};
#pragma unroll
	for (uint32_t w = 0; w < 8; w++) { if (w < warp) { pre.rows += s_warp[w].rows; pre.bytes += s_warp[w].bytes; } tot.rows += s_warp[w].rows; tot.bytes += s_warp[w].bytes; }
	cta_sync();
	*total = tot;
	return SegCount{pre.rows + inc.rows - mine.rows, pre.bytes + inc.bytes - mine.bytes};
}
// pass 1: per-CTA totals (4 segments per thread)
__global__ void __launch_bounds__(256) k_seg_sums(const uint64_t* __restrict__ tp, uint64_t nseg, const uint32_t* __restrict__ lo, const uint32_t* __restrict__ hi, uint64_t* __restrict__ cta_sums) {
	__shared__ SegCount s_warp[8];
	SegCount mine{0, 0};
	const uint64_t base = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * 4;
#pragma unroll
	for (int j = 0; j < 4; j++) { const SegCount c = seg_count(tp, lo, hi, base + j, nseg); mine.rows += c.rows; mine.bytes += c.bytes; }
	SegCount tot;
	cta_scan_1024(mine, s_warp, &tot);
	if (threadIdx.x == 0) { cta_sums[2 * (uint64_t)blockIdx.x] = tot.rows; cta_sums[2 * (uint64_t)blockIdx.x + 1] = tot.bytes; }
}
// pass 2: one CTA turns the per-CTA totals into exclusive bases, in place; entry nctas = grand totals
__global__ void __launch_bounds__(256) k_seg_bases(uint64_t nctas, uint64_t* cta_sums) {
	__shared__ SegCount s_warp[8];
	SegCount carry{0, 0};
	for (uint64_t b0 = 0; b0 < nctas; b0 += 256) {
		const uint64_t b = b0 + threadIdx.x;
		SegCount mine{0, 0};
		if (b < nctas) { mine.rows = cta_sums[2 * b]; mine.bytes = cta_sums[2 * b + 1]; }
		SegCount tot;
		const SegCount ex = cta_scan_1024(mine, s_warp, &tot);
		if (b < nctas) { cta_sums[2 * b] = carry.rows + ex.rows; cta_sums[2 * b + 1] = carry.bytes + ex.bytes; }
		carry.rows += tot.rows; carry.bytes += tot.bytes;
	}
	if (threadIdx.x == 0) { cta_sums[2 * nctas] = carry.rows; cta_sums[2 * nctas + 1] = carry.bytes; }
}
// pass 3: exclusive offsets per segment
__global__ void __launch_bounds__(256) k_seg_offsets(const uint64_t* __restrict__ tp, uint64_t nseg, const uint32_t* __restrict__ lo, const uint32_t* __restrict__ hi,
                                                     const uint64_t* __restrict__ cta_sums, uint64_t nctas, uint64_t* __restrict__ row_off, uint64_t* __restrict__ byte_off) {
	__shared__ SegCount s_warp[8];
	SegCount c[4], mine{0, 0};
	const uint64_t base = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * 4;
#pragma unroll
	for (int j = 0; j < 4; j++) { c[j] = seg_count(tp, lo, hi, base + j, nseg); mine.rows += c[j].rows; mine.bytes += c[j].bytes; }
	SegCount tot;
	SegCount ex = cta_scan_1024(mine, s_warp, &tot);
	ex.rows += cta_sums[2 * (uint64_t)blockIdx.x]; ex.bytes += cta_sums[2 * (uint64_t)blockIdx.x + 1];
#pragma unroll
	for (int j = 0; j < 4; j++) { if (base + j < nseg) { row_off[base + j] = ex.rows; byte_off[base + j] = ex.bytes; } ex.rows += c[j].rows; ex.bytes += c[j].bytes; }
	if (blockIdx.x == 0 && threadIdx.x == 0) { row_off[nseg] = cta_sums[2 * nctas]; byte_off[nseg] = cta_sums[2 * nctas + 1]; }
}

__device__ __forceinline__ uint32_t name_len(const RenderTables& rt, uint32_t id) { return __ldg(rt.name_off + id + 1) - __ldg(rt.name_off + id); }
// The carrier lists are staged in shared memory, a window of kRenderWin bytes per warp, and flushed
// with coalesced 32-bit stores: a lane's items land at byte offsets of their own, and writing them
// straight to global memory costs one sector transaction per byte.
constexpr uint32_t kRenderWin = 2048;
// bytes [b0, b1) of the item "name(g|g) " of sample `id`, to stage[0 .. b1 - b0).  Items of up to 15
// bytes come from a 16-byte template per sample ("name(0|0) ", length in byte 15): one load, the
// genotype characters patched in by XOR, bytes peeled off with static shifts; longer names take the
// character-by-character path.
__device__ __forceinline__ void put_carrier(const RenderTables& rt, uint8_t* stage, uint32_t id, uint8_t fl, uint32_t b0, uint32_t b1) {
	const uint4 t = __ldg(rt.item16 + id);
	const uint32_t len = t.w >> 24;
	if (len) {
		uint64_t lo = t.x | ((uint64_t)t.y << 32), hi = t.z | ((uint64_t)(t.w & 0x00FFFFFFu) << 32);
		const uint32_t delta = ((fl & 2) ? 1u : 0u) | ((fl & 1) ? 0u : (uint32_t)('/' ^ '|') << 8) | ((fl & 4) ? 1u << 16 : 0u);
		const uint32_t sh = 8 * (len - 5);                          // the genotype starts right after "name("
		if (sh < 64) { lo ^= (uint64_t)delta << sh; if (sh > 40) hi ^= (uint64_t)delta >> (64 - sh); }
		else hi ^= (uint64_t)delta << (sh - 64);
#pragma unroll
		for (uint32_t j = 0; j < 15; j++)
			if (j >= b0 && j < b1) stage[j - b0] = (uint8_t)(j < 8 ? lo >> (8 * j) : hi >> (8 * (j - 8)));
		return;
	}
	const uint32_t a = __ldg(rt.name_off + id), n = __ldg(rt.name_off + id + 1) - a;
	for (uint32_t b = b0; b < b1; b++) {
		char c;
		if (b < n) c = __ldg(rt.name_chars + a + b);
		else { const uint32_t k = b - n; c = k == 0 ? '(' : k == 1 ? ((fl & 2) ? '1' : '0') : k == 2 ? ((fl & 1) ? '|' : '/') : k == 3 ? ((fl & 4) ? '1' : '0') : k == 4 ? ')' : ' '; }
		stage[b - b0] = (uint8_t)c;
	}
}
// stage[sh .. sh + len) -> dst[0 .. len), where sh = dst & 3 so that 32-bit words line up on both sides
__device__ __forceinline__ void flush_window(const uint8_t* stage, uint32_t sh, uint32_t len, char* dst, uint32_t lane) {
	const uint32_t head = min(len, (4 - sh) & 3);
	if (lane < head) dst[lane] = (char)stage[sh + lane];
	const uint32_t nwords = (len - head) >> 2;
	const uint32_t* sw = (const uint32_t*)(stage + sh + head);
	uint32_t* dw = (uint32_t*)(dst + head);
	for (uint32_t i = lane; i < nwords; i += 32) dw[i] = sw[i];
	const uint32_t done = head + (nwords << 2);
	if (lane < len - done) dst[done + lane] = (char)stage[sh + done + lane];
}
__device__ __forceinline__ void warp_excl2(uint32_t lane, uint32_t a, uint32_t b, uint32_t& ea, uint32_t& eb, uint32_t& ta, uint32_t& tb) {
	uint32_t ia = a, ib = b;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, ia, d), y = __shfl_up_sync(0xFFFFFFFFu, ib, d); if (lane >= (uint32_t)d) { ia += x; ib += y; } }
	ea = ia - a; eb = ib - b;
	ta = __shfl_sync(0xFFFFFFFFu, ia, 31); tb = __shfl_sync(0xFFFFFFFFu, ib, 31);
}

// One row of print_var text (query.h:43-50) written by a warp at its final position: "pos\tref\talt\t", then the carriers of
// the row's vertex expanded from its class bitmap / id list (get_samples, query.h:268-285).  sq = {ref_off, ref_len, alt_off,
// alt_len} into rt.seq, cr = {carrier set id, s_info count | row flags << 28, s_info begin lo, hi}.  Rows of t6 records
// (k_render) and rows of t4 hit codes (k_render_hits) are the same text.
__device__ __forceinline__ void render_row(const DevIndex& ix, const RenderTables& rt, uint32_t pos_value, const uint4 sq, const uint4 cr, char* out, uint32_t row_bytes,
                                           int ws, uint8_t* stage, uint32_t lane) {
	const uint64_t kBase = 0x0505054E47544341ULL;                      // "ACTGN" + 5,5,5 by 3-bit code: map_int, src/util.cc:32-41
	// ---- "pos\t"
	uint32_t at = 0;
	{
		uint32_t p = pos_value, nd = 1;
		for (uint32_t q = p; q >= 10; q /= 10) nd++;
		if (lane == 0) { uint32_t q = p; for (uint32_t i = nd; i-- > 0;) { out[i] = (char)('0' + q % 10); q /= 10; } out[nd] = '\t'; }
		at = nd + 1;
	}
	// ---- ref \t alt \t
	for (uint32_t i = lane; i < sq.y; i += 32) out[at + i] = (char)(kBase >> (8 * (__ldg(rt.seq + sq.x + i) & 7)));
	at += sq.y;
	if (lane == 0) out[at] = '\t';
	at += 1;
	for (uint32_t i = lane; i < sq.w; i += 32) out[at + i] = (char)(kBase >> (8 * (__ldg(rt.seq + sq.z + i) & 7)));
	at += sq.w;
	if (lane == 0) { out[at] = '\t'; out[row_bytes - 1] = '\n'; }
	at += 1;
	if (!ws || ((cr.y >> 28) & 4)) return;                        // no carrier list asked for / the row has none
	// ---- carriers
	const uint64_t sbegin = (uint64_t)cr.z | ((uint64_t)cr.w << 32);
	uint32_t slot0 = 0;                                              // s_info entries consumed so far
	const uint32_t rounds = ix.class_mode ? (ix.words_per_set + 31) / 32 : ((cr.y & 0x0FFFFFFFu) + 31) / 32;
	for (uint32_t rd = 0; rd < rounds; rd++) {
		// this lane's items of the round: class mode = the members of one bitmap word (the ref bit owns an
		// s_info entry but is not printed); explicit-id mode = one s_info entry
		uint64_t all = 0; uint32_t base_id = 0, nslots = 0, bytes = 0;
		if (ix.class_mode) {
			const uint32_t w = rd * 32 + lane;
			all = w < ix.words_per_set ? __ldg(ix.bitmap + (uint64_t)cr.x * ix.words_per_set + w) : 0;
			base_id = w * 64; nslots = (uint32_t)__popcll(all);
			for (uint64_t m = (w == 0 ? all & ~1ULL : all); m; m &= m - 1) bytes += name_len(rt, base_id + (uint32_t)__ffsll((long long)m) - 1) + 6;
		} else {
			const uint32_t j = rd * 32 + lane;
			if (j < (cr.y & 0x0FFFFFFFu)) { base_id = __ldg(rt.s_sample_id + sbegin + j); all = 1; nslots = 1; if (base_id) bytes = name_len(rt, base_id) + 6; }
		}
		uint32_t eslot, ebytes, tslot, tbytes;
		warp_excl2(lane, nslots, bytes, eslot, ebytes, tslot, tbytes);
		for (uint32_t win0 = 0; win0 < tbytes; win0 += kRenderWin) {
			const uint32_t wlen = min(kRenderWin, tbytes - win0);
			char* dst = out + at + win0;
			const uint32_t sh = (uint32_t)((uintptr_t)dst & 3);
			if (bytes && ebytes < win0 + wlen && ebytes + bytes > win0) {
				uint32_t slot = slot0 + eslot, o = ebytes;
				for (uint64_t m = all; m; m &= m - 1) {
					const uint32_t id = ix.class_mode ? base_id + (uint32_t)__ffsll((long long)m) - 1 : base_id;
					if (id != 0) {
						const uint32_t len = name_len(rt, id) + 6;
						if (o + len > win0 && o < win0 + wlen) {
							const uint32_t b0 = o < win0 ? win0 - o : 0, b1 = min(len, win0 + wlen - o);
							put_carrier(rt, stage + sh + (o + b0 - win0), id, __ldg(rt.s_flags + sbegin + slot), b0, b1);
						}
						o += len;
					}
					slot++;
				}
			}
			__syncwarp();
			flush_window(stage, sh, wlen, dst, lane);
			__syncwarp();
		}
		slot0 += tslot; at += tbytes;
	}
}

__global__ void __launch_bounds__(256) k_render(const DevIndex ix, const RenderTables rt, const uint64_t* __restrict__ tp, uint64_t nseg, const uint32_t* __restrict__ seg_lo, int ws,
                                                const uint64_t* __restrict__ row_off, const uint64_t* __restrict__ byte_off, uint64_t row_begin, uint64_t row_end, char* __restrict__ text) {
	__shared__ __align__(16) uint8_t s_stage[8][kRenderWin + 16];
	const uint32_t lane = threadIdx.x & 31;
	uint8_t* stage = s_stage[threadIdx.x >> 5];
	const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	for (uint64_t row = row_begin + warp0; row < row_end; row += nwarps) {
		// segment of this row: last s with row_off[s] <= row (empty segments share an offset with their successor)
		uint64_t a = 0, b = nseg;
		while (b - a > 1) { const uint64_t m = (a + b) >> 1; if (__ldg(row_off + m) <= row) a = m; else b = m; }
		const uint32_t lo = __ldg(seg_lo + a);
		const uint32_t r = lo + (uint32_t)(row - __ldg(row_off + a));
		const uint64_t t0 = __ldg(tp + r);
		char* out = text + __ldg(byte_off + a) + (t0 - __ldg(tp + lo));
		const uint32_t row_bytes = (uint32_t)(__ldg(tp + r + 1) - t0);
		const uint4 sq = __ldg(rt.rec_seq + r), cr = __ldg(rt.rec_car + r);
		render_row(ix, rt, __ldg(ix.rec_pos + r), sq, cr, out, row_bytes, ws, stage, lane);
	}
}

// ------------------------------------------------------------------ t4 rows as text, on the device
// A t4 row depends on the walk-entry code alone (entry, and whether the walk started on it / the row is the vertex the
// alt edge rejoins): HitTables hold, per entry and variant, what t4_row of materialize.cc computes — position, ref / alt
// slices, the vertex whose carriers are printed, and the row's text length.  k_hits_sums / k_hits_offsets scan the row
// lengths of the batch's hit codes, k_render_hits writes the rows (a warp each), k_region_text_offsets reads the text
// offset of every region off its first row.
__device__ __forceinline__ uint32_t hit_variant(uint32_t code) { return (code & kHitRejoin) ? 2u : (code & kHitStart) ? 1u : 0u; }
__device__ __forceinline__ uint32_t ndigits(uint32_t p) { uint32_t nd = 1; for (; p >= 10; p /= 10) nd++; return nd; }
// pos5 (nullable): the rows are t5's — same text with the position column replaced, so the length moves by the digit counts
__device__ __forceinline__ uint32_t hit_len(const HitTables& ht, const uint32_t* __restrict__ hits, uint64_t h, uint64_t nh, int ws, const uint32_t* __restrict__ pos5) {
	if (h >= nh) return 0;
	const uint32_t code = hits[h], v = hit_variant(code), c = code & 0x3FFFFFFFu;
	uint32_t len = __ldg(ht.len[ws][v] + c);
	if (pos5) len = len - ndigits(__ldg(ht.pos[v] + c)) + ndigits(pos5[h]);
	return len;
}
// t5 rows (get_sample_var_in_sample, query.h:553-590): var_pos of every hit code of a t5 answer — ref_pos for an insertion
// (one more than t4 prints), else the sample's own `index` in the vertex whose carriers the row lists (0 when the sample is
// not among them, as the reference's default-constructed sample_info).  One thread per row; its region (and so its sample)
// by a search over the CSR offsets.  gsidx = sample_info.index of every s_info entry, in s_info order.
__global__ void __launch_bounds__(256) k_t5_row_pos(const DevIndex ix, const RenderTables rt, const HitTables ht, const uint32_t* __restrict__ gsidx, uint64_t n,
                                                    const uint64_t* __restrict__ offsets, const uint32_t* __restrict__ samples, const uint32_t* __restrict__ hits, uint64_t nh,
                                                    uint32_t* __restrict__ pos5) {
	for (uint64_t h = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; h < nh; h += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t a = 0, b = n;                                     // last region with offsets[a] <= h (empty regions share an offset with their successor)
		while (b - a > 1) { const uint64_t m = (a + b) >> 1; if (__ldg(offsets + m) <= h) a = m; else b = m; }
		const uint32_t s = __ldg(samples + a);
		const uint32_t code = hits[h], v = hit_variant(code), c = code & 0x3FFFFFFFu;
		const uint4 cr = __ldg(ht.car[v] + c);
		uint32_t p = 0;
		if (cr.y >> 31) p = __ldg(ht.pos[v] + c) + 1;
		else {
			const uint64_t sb = (uint64_t)cr.z | ((uint64_t)cr.w << 32);
			const uint32_t cnt = cr.y & 0x0FFFFFFFu;
			if (ix.class_mode) {
				const uint64_t* row = ix.bitmap + (uint64_t)cr.x * ix.words_per_set;
				uint32_t r = 0;
				for (uint32_t w = 0; w < (s >> 6); w++) r += (uint32_t)__popcll(__ldg(row + w));
				r += (uint32_t)__popcll(__ldg(row + (s >> 6)) & (((uint64_t)1 << (s & 63)) - 1));
				if (r < cnt) p = __ldg(gsidx + sb + r);
			} else {
				for (uint32_t i = 0; i < cnt; i++) if (__ldg(rt.s_sample_id + sb + i) == s) { p = __ldg(gsidx + sb + i); break; }
			}
		}
		pos5[h] = p;
	}
}
__global__ void __launch_bounds__(256) k_hits_sums(const HitTables ht, const uint32_t* __restrict__ hits, uint64_t nh, int ws, const uint32_t* __restrict__ pos5, uint64_t* __restrict__ cta_sums) {
	__shared__ SegCount s_warp[8];
	SegCount mine{0, 0};
	const uint64_t base = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * 4;
#pragma unroll
	for (int j = 0; j < 4; j++) mine.bytes += hit_len(ht, hits, base + j, nh, ws, pos5);
	SegCount tot;
	cta_scan_1024(mine, s_warp, &tot);
	if (threadIdx.x == 0) { cta_sums[2 * (uint64_t)blockIdx.x] = 0; cta_sums[2 * (uint64_t)blockIdx.x + 1] = tot.bytes; }
}
__global__ void __launch_bounds__(256) k_hits_offsets(const HitTables ht, const uint32_t* __restrict__ hits, uint64_t nh, int ws, const uint32_t* __restrict__ pos5, const uint64_t* __restrict__ cta_sums, uint64_t nctas,
                                                      uint64_t* __restrict__ byte_off) {
	__shared__ SegCount s_warp[8];
	uint32_t c[4]; SegCount mine{0, 0};
	const uint64_t base = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * 4;
#pragma unroll
	for (int j = 0; j < 4; j++) { c[j] = hit_len(ht, hits, base + j, nh, ws, pos5); mine.bytes += c[j]; }
	SegCount tot;
	SegCount ex = cta_scan_1024(mine, s_warp, &tot);
	ex.bytes += cta_sums[2 * (uint64_t)blockIdx.x + 1];
#pragma unroll
	for (int j = 0; j < 4; j++) { if (base + j < nh) byte_off[base + j] = ex.bytes; ex.bytes += c[j]; }
	if (blockIdx.x == 0 && threadIdx.x == 0) byte_off[nh] = cta_sums[2 * nctas + 1];
}
__global__ void __launch_bounds__(256) k_region_text_offsets(uint64_t n, const uint64_t* __restrict__ offsets, const uint64_t* __restrict__ byte_off, uint64_t* __restrict__ out) {
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i <= n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = byte_off[offsets[i]];
}
__global__ void __launch_bounds__(256) k_render_hits(const DevIndex ix, const RenderTables rt, const HitTables ht, const uint32_t* __restrict__ hits, int ws, const uint32_t* __restrict__ pos5,
                                                     const uint64_t* __restrict__ byte_off, uint64_t row_begin, uint64_t row_end, char* __restrict__ text) {
	__shared__ __align__(16) uint8_t s_stage[8][kRenderWin + 16];
	const uint32_t lane = threadIdx.x & 31;
	uint8_t* stage = s_stage[threadIdx.x >> 5];
	const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	for (uint64_t row = row_begin + warp0; row < row_end; row += nwarps) {
		const uint32_t code = hits[row], v = hit_variant(code), c = code & 0x3FFFFFFFu;
		const uint64_t b0 = __ldg(byte_off + row);
		render_row(ix, rt, pos5 ? pos5[row] : __ldg(ht.pos[v] + c), __ldg(ht.seq[v] + c), __ldg(ht.car[v] + c), text + b0, (uint32_t)(__ldg(byte_off + row + 1) - b0), ws, stage, lane);
	}
}

// ------------------------------------------------------------------ hit map build (once, at vsgpu_open)
// one warp per walk entry: lanes sweep the words of the entry's carrier set and scatter its bits
// into the sample rows.
__global__ void __launch_bounds__(256) k_build_hitmap(const DevIndex ix, uint32_t* __restrict__ hitmap) {
	const uint32_t lane = threadIdx.x & 31;
	const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	for (uint64_t c = warp; c < ix.num_cent; c += nwarps) {
		const uint4 e = __ldg(ix.cent + c);
		if (e.y & kEntMarker) continue;
		const uint32_t bit = 1u << (c & 31); const uint64_t col = c >> 5;
		if (ix.class_mode) {
			for (uint32_t w = lane; w < ix.words_per_set; w += 32) {
				uint64_t bits = __ldg(ix.bitmap + (uint64_t)e.z * ix.words_per_set + w);
				while (bits) {
					const uint32_t s = w * 64 + (uint32_t)__ffsll((long long)bits) - 1; bits &= bits - 1;
					if (s != 0) atomicOr(hitmap + (uint64_t)s * ix.row_words + col, bit);
				}
			}
		} else {
			for (uint64_t i = __ldg(ix.list_begin + e.z) + lane, end = __ldg(ix.list_begin + e.z + 1); i < end; i += 32)
				atomicOr(hitmap + (uint64_t)__ldg(ix.list_ids + i) * ix.row_words + col, bit);
		}
	}
}

__global__ void __launch_bounds__(256) k_widen(uint64_t n, const uint32_t* __restrict__ x32, const uint32_t* __restrict__ y32, uint64_t* __restrict__ x, uint64_t* __restrict__ y) {
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) { x[i] = x32[i]; y[i] = y32[i]; }
}

// ------------------------------------------------------------------ t2: query_sample_from_ref (query.h:120-189)
// count: one thread per region walks the sample's path (logic::t2_walk) and counts copy records and
// bytes; the CTA reduces them into cta_sums (k_seg_bases then scans those in place).
template <bool kT3>
__global__ void __launch_bounds__(256) k_t2_count(const DevIndex ix, const T2Tables t2, const T3Tables t3, uint64_t n, const uint64_t* __restrict__ xs, const uint64_t* __restrict__ ys,
                                                  const uint32_t* __restrict__ sample, uint2* __restrict__ cnt, uint2* __restrict__ keep, uint8_t* __restrict__ status,
                                                  uint64_t* __restrict__ cta_sums, uint32_t* gstatus) {
	__shared__ SegCount s_warp[8];
	const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
	SegCount mine{0, 0};
	if (i < n) {
		const uint32_t s = sample[i];
		uint32_t st = 0;
		T2CountSink sink{0, 0, 0, 0, keep + i, n};
		if (s == 0 || s >= ix.num_samples) atomicOr(gstatus, kStatusBadRegion);
		else { st = kT3 ? t3_walk(ix, t2, t3, xs[i], ys[i], s, sink) : t2_walk(ix, t2, xs[i], ys[i], s, sink); sink.flush(); }
		if (st) { sink.nrec = 0; sink.bytes = 0; }
		cnt[i] = make_uint2(sink.nrec, (uint32_t)sink.bytes);
		status[i] = (uint8_t)st;
		mine.rows = sink.nrec; mine.bytes = sink.bytes;
	}
	SegCount tot;
	cta_scan_1024(mine, s_warp, &tot);
	if (threadIdx.x == 0) { cta_sums[2 * (uint64_t)blockIdx.x] = tot.rows; cta_sums[2 * (uint64_t)blockIdx.x + 1] = tot.bytes; }
}
// plan: byte offset of every region (exclusive scan of the counts) and its copy records, written at
// their final index so the records of the batch are in region order.
template <bool kT3>
__global__ void __launch_bounds__(256) k_t2_plan(const DevIndex ix, const T2Tables t2, const T3Tables t3, uint64_t n, const uint64_t* __restrict__ xs, const uint64_t* __restrict__ ys,
                                                 const uint32_t* __restrict__ sample, const uint2* __restrict__ cnt, const uint2* __restrict__ keep, const uint64_t* __restrict__ cta_sums,
                                                 uint64_t nctas, uint64_t* __restrict__ offsets, uint4* __restrict__ recs, uint32_t* __restrict__ tile_first) {
	__shared__ SegCount s_warp[8];
	const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
	SegCount mine{0, 0};
	if (i < n) { const uint2 c = cnt[i]; mine.rows = c.x; mine.bytes = c.y; }
	SegCount tot;
	SegCount ex = cta_scan_1024(mine, s_warp, &tot);
	ex.rows += cta_sums[2 * (uint64_t)blockIdx.x]; ex.bytes += cta_sums[2 * (uint64_t)blockIdx.x + 1];
	if (i < n) {
		offsets[i] = ex.bytes;
		if (mine.rows) {
			T2WriteSink sink{0, 0, recs + ex.rows, ex.bytes, recs, tile_first};
			if (mine.rows <= kT2Keep) {                               // the pieces the count pass kept
				for (uint32_t k = 0; k < (uint32_t)mine.rows; k++) { const uint2 p = __ldg(keep + k * n + i); sink.off = p.x; sink.len = p.y; sink.flush(); }
			} else {
				if (kT3) t3_walk(ix, t2, t3, xs[i], ys[i], sample[i], sink); else t2_walk(ix, t2, xs[i], ys[i], sample[i], sink);
				sink.flush();
			}
		}
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) offsets[n] = cta_sums[2 * nctas + 1];
}
// ------------------------------------------------------------------ t5: get_sample_var_in_sample (query.h:490-612)
// count: one thread per region walks (logic::t5_walk), counts its rows and keeps the first kT5Keep hit codes; write: offsets
// from the scan of the counts, then the kept codes are copied to their final place (region order = the reference's push
// order) — only a region with more rows than were kept walks a second time.
constexpr uint32_t kT5Keep = 8;
struct T5CountSink { uint32_t n; uint32_t* keep; __device__ __forceinline__ void emit(uint32_t code) { if (n < kT5Keep) keep[n] = code; n++; } };
__global__ void __launch_bounds__(256) k_t5_count(const DevIndex ix, const T2Tables t2, const T3Tables t3, uint64_t n, const uint64_t* __restrict__ xs, const uint64_t* __restrict__ ys,
                                                  const uint32_t* __restrict__ sample, uint32_t* __restrict__ cnt, uint8_t* __restrict__ status,
                                                  uint64_t* __restrict__ cta_sums, uint32_t* gstatus, uint32_t* __restrict__ keep) {
	__shared__ SegCount s_warp[8];
	const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
	SegCount mine{0, 0};
	if (i < n) {
		const uint32_t s = sample[i];
		uint32_t st = 0;
		T5CountSink sink{0, keep + i * kT5Keep};
		if (s == 0 || s >= ix.num_samples) atomicOr(gstatus, kStatusBadRegion);
		else st = t5_walk(ix, t2, t3, xs[i], ys[i], s, sink);
		if (st) sink.n = 0;
		cnt[i] = sink.n; status[i] = (uint8_t)st;
		mine.rows = sink.n;
	}
	SegCount tot;
	cta_scan_1024(mine, s_warp, &tot);
	if (threadIdx.x == 0) { cta_sums[2 * (uint64_t)blockIdx.x] = tot.rows; cta_sums[2 * (uint64_t)blockIdx.x + 1] = 0; }
}
__global__ void __launch_bounds__(256) k_t5_write(const DevIndex ix, const T2Tables t2, const T3Tables t3, uint64_t n, const uint64_t* __restrict__ xs, const uint64_t* __restrict__ ys,
                                                  const uint32_t* __restrict__ sample, const uint32_t* __restrict__ cnt, const uint64_t* __restrict__ cta_sums,
                                                  uint64_t nctas, uint64_t* __restrict__ offsets, uint32_t* __restrict__ hits, const uint32_t* __restrict__ keep) {
	__shared__ SegCount s_warp[8];
	const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
	SegCount mine{0, 0};
	if (i < n) mine.rows = cnt[i];
	SegCount tot;
	SegCount ex = cta_scan_1024(mine, s_warp, &tot);
	ex.rows += cta_sums[2 * (uint64_t)blockIdx.x];
	if (i < n) {
		offsets[i] = ex.rows;
		if (mine.rows > kT5Keep) { DirectSink sink{hits + ex.rows, 0}; t5_walk(ix, t2, t3, xs[i], ys[i], sample[i], sink); }
		else for (uint32_t j = 0; j < (uint32_t)mine.rows; j++) hits[ex.rows + j] = keep[i * kT5Keep + j];
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) offsets[n] = cta_sums[2 * nctas];
}

// copy: the output is cut into tiles of kT2Tile bytes, one per warp and step; lane l owns bytes
// [16 l, 16 l + 16) of the tile.  tile_first names the record covering the tile's first byte; the warp
// loads the 32 records from there with one coalesced load, turns them into tile-relative 32-bit
// (start, end, source base) and every lane finds the record of its chunk from a bit mask of the
// chunks in which a record starts (one warp OR-reduction; no dependent loads, no 64-bit compares).
// A lane reads the aligned 16 bytes under its source position; the following 16 come from the next
// lane's load when that lane continues the same record (a shuffle), else from a second load.  The
// chunk is cut out of that 32-byte window and stored with one 128-bit store — a warp stores 512
// contiguous bytes.  Chunks in which a record starts are left out of that pass (a few lanes per tile
// would serialise the whole warp); a second loop, one thread per record, builds exactly those chunks
// from the two records that meet there (byte by byte when more than two do).
__device__ __forceinline__ uint64_t rec_dst(const uint4& r) { return r.z | ((uint64_t)r.w << 32); }
// bytes [off, off + 16) of the 32-byte window a | b
__device__ __forceinline__ uint4 cut16(const uint4& a, const uint4& b, uint32_t off) {
	const bool s2 = off & 8, s1 = off & 4;
	const uint32_t sh = (off & 3) * 8;
	const uint32_t x0 = s2 ? a.z : a.x, x1 = s2 ? a.w : a.y, x2 = s2 ? b.x : a.z, x3 = s2 ? b.y : a.w, x4 = s2 ? b.z : b.x, x5 = s2 ? b.w : b.y;
	const uint32_t y0 = s1 ? x1 : x0, y1 = s1 ? x2 : x1, y2 = s1 ? x3 : x2, y3 = s1 ? x4 : x3, y4 = s1 ? x5 : x4;
	return make_uint4(__funnelshift_r(y0, y1, sh), __funnelshift_r(y1, y2, sh), __funnelshift_r(y2, y3, sh), __funnelshift_r(y3, y4, sh));
}
__device__ __forceinline__ uint4 load16u(const char* p) {          // 16 bytes at any alignment (reads the aligned 32 around them)
	const uint32_t off = (uint32_t)((uintptr_t)p & 15);
	const uint4* av = (const uint4*)(p - off);
	const uint4 a = __ldg(av);
	return cut16(a, off ? __ldg(av + 1) : a, off);
}
__device__ __forceinline__ uint32_t low_bytes_mask(uint32_t m, uint32_t lo) {   // bytes of word [lo, lo + 4) that lie below byte m
	return m >= lo + 4 ? 0xFFFFFFFFu : (m <= lo ? 0u : (1u << (8 * (m - lo))) - 1);
}
template <bool kAhead, uint32_t kCtas>
__global__ void __launch_bounds__(256, kCtas) k_t2_copy(const T2Tables t2, const uint4* __restrict__ recs, const uint32_t* __restrict__ tile_first,
                                                        const uint64_t* __restrict__ totals, char* __restrict__ text) {
	const uint32_t nrecs = (uint32_t)totals[0];
	const uint64_t nbytes = totals[1];
	const uint32_t ntiles = (uint32_t)((nbytes + kT2Tile - 1) / kT2Tile);
	const uint32_t last_len = (uint32_t)(nbytes - (uint64_t)(ntiles ? ntiles - 1 : 0) * kT2Tile);
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t c_rel = lane * 16;
	const char* __restrict__ seq = t2.seq_ascii;
	const uint32_t wstride = (gridDim.x * 256u) >> 5;
	uint32_t t = (blockIdx.x * 256u + threadIdx.x) >> 5;
	// software pipeline: tile_first is read two steps ahead and (kAhead) the records one step ahead,
	// so that a step's own dependent chain is just source load -> store
	uint32_t j0_next = t < ntiles ? __ldg(tile_first + t) : 0;
	uint32_t j0_next2 = t + wstride < ntiles ? __ldg(tile_first + t + wstride) : 0;
	uint4 R_next = make_uint4(0, 0, 0, 0);
	if (kAhead && t < ntiles) R_next = __ldg(recs + (j0_next + lane < nrecs ? j0_next + lane : nrecs - 1));
	for (; t < ntiles; t += wstride) {
		const uint32_t j0 = j0_next;
		j0_next = j0_next2;
		if (t + 2 * wstride < ntiles) j0_next2 = __ldg(tile_first + t + 2 * wstride);
		const uint64_t base = (uint64_t)t * kT2Tile;
		const uint32_t tile_len = t + 1 == ntiles ? last_len : kT2Tile;
		const bool have = j0 + lane < nrecs;
		uint4 R;
		if (kAhead) { R = R_next; if (t + wstride < ntiles) R_next = __ldg(recs + (j0_next + lane < nrecs ? j0_next + lane : nrecs - 1)); }
		else R = __ldg(recs + (have ? j0 + lane : nrecs - 1));
		// tile-relative start / end of the loaded record (clamped to the tile; 513 = runs past it) and
		// where byte 0 of the tile would lie in seq_ascii if this record covered it (mod 2^32)
		const uint64_t ds = rec_dst(R) - base, de = ds + R.y;        // lane 0: ds may be "negative", de is not
		const uint32_t rel_s = (lane == 0 || !have) ? (lane ? kT2Tile : 0u) : (ds >= kT2Tile ? kT2Tile : (uint32_t)ds);
		const uint32_t rel_e = (int64_t)de > (int64_t)kT2Tile ? kT2Tile + 1 : (uint32_t)de;
		const uint32_t sb = R.x - (uint32_t)ds;
		const uint32_t in_tile = __ballot_sync(0xFFFFFFFFu, lane > 0 && rel_s < tile_len);
		if (in_tile >> 31) {
			// 32 or more records in one tile (records of a few bytes): every lane searches on its own
			const uint64_t c0 = base + c_rel, end = c0 + 16 < nbytes ? c0 + 16 : nbytes;
			const bool live = c0 < nbytes;
			uint32_t j = j0; uint4 rec = __ldg(recs + j);
			while (live && j + 1 < nrecs) { const uint4 nx = __ldg(recs + j + 1); if (rec_dst(nx) > c0) break; rec = nx; j++; }
			const uint64_t d = rec_dst(rec);
			if (live && d + rec.y >= end) *(uint4*)(text + c0) = load16u(seq + rec.x + (c0 - d));
			continue;
		}
		// idx = how many of the records 1..cnt start at or before my chunk (a SNP allele is a record of
		// one byte, so two records starting in one chunk is the normal case)
		uint32_t idx = 0;
		for (uint32_t k = 1, cnt = __popc(in_tile); k <= cnt; k++) idx += __shfl_sync(0xFFFFFFFFu, rel_s, k) <= c_rel ? 1u : 0u;
		const uint32_t e_i = __shfl_sync(0xFFFFFFFFu, rel_e, idx), sb_i = __shfl_sync(0xFFFFFFFFu, sb, idx);
		const uint32_t end_rel = c_rel + 16 < tile_len ? c_rel + 16 : tile_len;
		if (c_rel < tile_len && e_i >= end_rel) {                    // else a record starts inside: second loop
			const uint32_t so = sb_i + c_rel, off = so & 15;
			const uint4* av = (const uint4*)(seq + (so - off));
			const uint4 a = __ldg(av);
			uint4 b2 = a;
			if (off) b2 = __ldg(av + 1);
			*(uint4*)(text + base + c_rel) = cut16(a, b2, off);
		}
	}
}
// seams: thread r takes the chunk record r starts in, if it starts off a 16-byte boundary and is the
// first record to do so in that chunk (record r - 1 then covers the chunk's first byte): bytes below
// the start of r come from r - 1; then record after record lays its bytes over the rest — each as the
// 16 bytes around a virtual source pointer, masked to the bytes the record owns.
__global__ void __launch_bounds__(256, 8) k_t2_seams(const T2Tables t2, const uint4* __restrict__ recs, const uint64_t* __restrict__ totals, char* __restrict__ text) {
	const uint32_t nrecs = (uint32_t)totals[0];
	const uint64_t nbytes = totals[1];
	const char* __restrict__ seq = t2.seq_ascii;
	for (uint32_t r = blockIdx.x * 256u + threadIdx.x + 1; r < nrecs; r += gridDim.x * 256u) {
		uint4 rec = __ldg(recs + r);
		uint64_t dq = rec_dst(rec);
		const uint32_t m = (uint32_t)dq & 15;
		if (m == 0) continue;
		const uint64_t c0 = dq - m;
		const uint4 prev = __ldg(recs + r - 1);
		const uint64_t dp = rec_dst(prev);
		if (dp > c0) continue;
		const uint32_t need = c0 + 16 < nbytes ? 16u : (uint32_t)(nbytes - c0);
		uint4 o = load16u(seq + prev.x + (uint32_t)(c0 - dp));
		uint32_t lo = m;                                           // record q owns bytes [lo, hi) of the chunk
		for (uint32_t q = r;;) {
			const uint64_t room = (uint64_t)lo + rec.y;
			const uint32_t hi = room < need ? (uint32_t)room : need;
			const uint4 B = load16u(seq + (int64_t)rec.x - (int64_t)lo);     // byte lo of B = first byte of record q
			const uint32_t k0 = low_bytes_mask(lo, 0) | ~low_bytes_mask(hi, 0), k1 = low_bytes_mask(lo, 4) | ~low_bytes_mask(hi, 4);
			const uint32_t k2 = low_bytes_mask(lo, 8) | ~low_bytes_mask(hi, 8), k3 = low_bytes_mask(lo, 12) | ~low_bytes_mask(hi, 12);
			o = make_uint4((o.x & k0) | (B.x & ~k0), (o.y & k1) | (B.y & ~k1), (o.z & k2) | (B.z & ~k2), (o.w & k3) | (B.w & ~k3));
			if (hi >= need) break;
			lo = hi;
			rec = __ldg(recs + ++q);
		}
		*(uint4*)(text + c0) = o;
	}
}

inline uint32_t grid_for(uint64_t n, uint32_t block, int ctas_per_sm) {
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	uint64_t want = (n + block - 1) / block;
	uint64_t cap = (uint64_t)sms * ctas_per_sm;
	return (uint32_t)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace

cudaError_t launch_build_hitmap(const DevIndex& ix, uint32_t* hitmap, cudaStream_t stream) {
	if (ix.num_cent == 0) return cudaSuccess;
	k_build_hitmap<<<grid_for((uint64_t)ix.num_cent * 32, 256, 8), 256, 0, stream>>>(ix, hitmap);
	return cudaGetLastError();
}
cudaError_t launch_render_offsets(const DevIndex& ix, const RenderTables& rt, uint64_t nseg, const uint32_t* seg_lo, const uint32_t* seg_hi, int with_samples,
                                  uint64_t* row_off, uint64_t* byte_off, uint64_t* scratch, cudaStream_t stream) {
	(void)ix;
	const int ws = with_samples ? 1 : 0;
	const uint64_t nctas = (nseg + 1023) / 1024;
	if (nseg == 0) return cudaSuccess;
	k_seg_sums<<<(uint32_t)nctas, 256, 0, stream>>>(rt.text_prefix[ws], nseg, seg_lo, seg_hi, scratch);
	k_seg_bases<<<1, 256, 0, stream>>>(nctas, scratch);
	k_seg_offsets<<<(uint32_t)nctas, 256, 0, stream>>>(rt.text_prefix[ws], nseg, seg_lo, seg_hi, scratch, nctas, row_off, byte_off);
	return cudaGetLastError();
}
cudaError_t launch_render(const DevIndex& ix, const RenderTables& rt, uint64_t nseg, const uint32_t* seg_lo, int with_samples,
                          const uint64_t* row_off, const uint64_t* byte_off, uint64_t row_begin, uint64_t row_end, char* text, cudaStream_t stream) {
	if (row_end <= row_begin) return cudaSuccess;
	k_render<<<grid_for((row_end - row_begin) * 32, 256, 8), 256, 0, stream>>>(ix, rt, rt.text_prefix[with_samples ? 1 : 0], nseg, seg_lo, with_samples ? 1 : 0, row_off, byte_off, row_begin, row_end, text);
	return cudaGetLastError();
}
cudaError_t launch_t5_row_pos(const DevIndex& ix, const RenderTables& rt, const HitTables& ht, const uint32_t* gsidx, uint64_t n, const uint64_t* offsets, const uint32_t* samples,
                              const uint32_t* hits, uint64_t nh, uint32_t* pos5, cudaStream_t stream) {
	if (nh == 0 || n == 0) return cudaSuccess;
	k_t5_row_pos<<<grid_for(nh, 256, 8), 256, 0, stream>>>(ix, rt, ht, gsidx, n, offsets, samples, hits, nh, pos5);
	return cudaGetLastError();
}
cudaError_t launch_hit_offsets(const HitTables& ht, const uint32_t* hits, uint64_t nh, int with_samples, uint64_t n, const uint64_t* offsets, uint64_t* byte_off, uint64_t* region_off,
                               uint64_t* scratch, cudaStream_t stream, const uint32_t* pos5) {
	const int ws = with_samples ? 1 : 0;
	const uint64_t nctas = (nh + 1023) / 1024;
	if (nh) {
		k_hits_sums<<<(uint32_t)nctas, 256, 0, stream>>>(ht, hits, nh, ws, pos5, scratch);
		k_seg_bases<<<1, 256, 0, stream>>>(nctas, scratch);
		k_hits_offsets<<<(uint32_t)nctas, 256, 0, stream>>>(ht, hits, nh, ws, pos5, scratch, nctas, byte_off);
	} else cudaMemsetAsync(byte_off, 0, 8, stream);
	k_region_text_offsets<<<grid_for(n + 1, 256, 8), 256, 0, stream>>>(n, offsets, byte_off, region_off);
	return cudaGetLastError();
}
cudaError_t launch_render_hits(const DevIndex& ix, const RenderTables& rt, const HitTables& ht, const uint32_t* hits, int with_samples, const uint64_t* byte_off,
                               uint64_t row_begin, uint64_t row_end, char* text, cudaStream_t stream, const uint32_t* pos5) {
	if (row_end <= row_begin) return cudaSuccess;
	k_render_hits<<<grid_for((row_end - row_begin) * 32, 256, 8), 256, 0, stream>>>(ix, rt, ht, hits, with_samples ? 1 : 0, pos5, byte_off, row_begin, row_end, text);
	return cudaGetLastError();
}
cudaError_t launch_widen(uint64_t n, const uint32_t* x32, const uint32_t* y32, uint64_t* x, uint64_t* y, cudaStream_t stream) {
	if (n == 0) return cudaSuccess;
	k_widen<<<grid_for(n, 256, 8), 256, 0, stream>>>(n, x32, y32, x, y);
	return cudaGetLastError();
}
cudaError_t launch_t6(const DevIndex& ix, uint64_t n, const uint64_t* x, const uint64_t* y, uint32_t* lo, uint32_t* hi, uint32_t* counts,
                      uint32_t* flagged, uint32_t flag_base, uint32_t* status, cudaStream_t stream) {
	if (n == 0) return cudaSuccess;
	k_t6<<<grid_for(n, 256, 8), 256, 0, stream>>>(ix, n, x, y, lo, hi, counts, flagged, flag_base, status);
	return cudaGetLastError();
}
cudaError_t launch_t1(const DevIndex& ix, uint64_t n, const uint64_t* pos, uint32_t* lo, uint32_t* hi, uint32_t* status, cudaStream_t stream) {
	if (n == 0) return cudaSuccess;
	k_t1<<<grid_for(n, 256, 8), 256, 0, stream>>>(ix, n, pos, lo, hi, status);
	return cudaGetLastError();
}
cudaError_t launch_t7(const DevIndex& ix, uint64_t n, const uint64_t* pos, const uint64_t* qhash, uint32_t* rec, uint32_t* status, cudaStream_t stream) {
	if (n == 0) return cudaSuccess;
	k_t7<<<grid_for(n, 256, 8), 256, 0, stream>>>(ix, n, pos, qhash, rec, status);
	return cudaGetLastError();
}
uint64_t t5_keep_words(uint64_t n) { return n * kT5Keep; }
cudaError_t launch_t5_count(const DevIndex& ix, const T2Tables& t2, const T3Tables& t3, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample,
                            uint32_t* cnt, uint8_t* status, uint64_t* cta_sums, uint32_t* gstatus, uint32_t* keep, cudaStream_t stream) {
	if (n == 0) return cudaSuccess;
	const uint64_t nctas = (n + 255) / 256;
	k_t5_count<<<(uint32_t)nctas, 256, 0, stream>>>(ix, t2, t3, n, x, y, sample, cnt, status, cta_sums, gstatus, keep);
	k_seg_bases<<<1, 256, 0, stream>>>(nctas, cta_sums);
	return cudaGetLastError();
}
cudaError_t launch_t5_write(const DevIndex& ix, const T2Tables& t2, const T3Tables& t3, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample,
                            const uint32_t* cnt, const uint64_t* cta_sums, uint64_t* offsets, uint32_t* hits, const uint32_t* keep, cudaStream_t stream) {
	if (n == 0) return cudaSuccess;
	const uint64_t nctas = (n + 255) / 256;
	k_t5_write<<<(uint32_t)nctas, 256, 0, stream>>>(ix, t2, t3, n, x, y, sample, cnt, cta_sums, nctas, offsets, hits, keep);
	return cudaGetLastError();
}
uint64_t t2_ctas(uint64_t n) { return (n + 255) / 256; }
cudaError_t launch_t2_count(const DevIndex& ix, const T2Tables& t2, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample,
                            uint2* cnt, uint2* keep, uint8_t* status, uint64_t* cta_sums, uint32_t* gstatus, cudaStream_t stream, const T3Tables* t3) {
	if (n == 0) return cudaSuccess;
	const uint64_t nctas = t2_ctas(n);
	if (t3) k_t2_count<true><<<(uint32_t)nctas, 256, 0, stream>>>(ix, t2, *t3, n, x, y, sample, cnt, keep, status, cta_sums, gstatus);
	else k_t2_count<false><<<(uint32_t)nctas, 256, 0, stream>>>(ix, t2, T3Tables{}, n, x, y, sample, cnt, keep, status, cta_sums, gstatus);
	k_seg_bases<<<1, 256, 0, stream>>>(nctas, cta_sums);
	return cudaGetLastError();
}
cudaError_t launch_t2_plan(const DevIndex& ix, const T2Tables& t2, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample,
                           const uint2* cnt, const uint2* keep, const uint64_t* cta_sums, uint64_t* offsets, uint4* recs, uint32_t* tile_first, cudaStream_t stream,
                           const T3Tables* t3) {
	if (n == 0) return cudaSuccess;
	const uint64_t nctas = t2_ctas(n);
	if (t3) k_t2_plan<true><<<(uint32_t)nctas, 256, 0, stream>>>(ix, t2, *t3, n, x, y, sample, cnt, keep, cta_sums, nctas, offsets, recs, tile_first);
	else k_t2_plan<false><<<(uint32_t)nctas, 256, 0, stream>>>(ix, t2, T3Tables{}, n, x, y, sample, cnt, keep, cta_sums, nctas, offsets, recs, tile_first);
	return cudaGetLastError();
}
cudaError_t launch_t2_copy(const T2Tables& t2, const uint4* recs, const uint32_t* tile_first, const uint64_t* totals, uint64_t recs_hint, uint64_t bytes_hint, char* text, cudaStream_t stream) {
	if (bytes_hint == 0) return cudaSuccess;
	static int ahead = -1;                          // tuning knob: records loaded one step ahead, or not
	if (ahead < 0) { const char* e = getenv("VSGPU_T2_AHEAD"); ahead = e ? atoi(e) : 1; }   // measured on 1 M x 1 kb regions: 0.484 ms with, 0.517 ms without
	if (ahead) k_t2_copy<true, 8><<<grid_for((bytes_hint + 15) / 16, 256, 8), 256, 0, stream>>>(t2, recs, tile_first, totals, text);
	else k_t2_copy<false, 8><<<grid_for((bytes_hint + 15) / 16, 256, 8), 256, 0, stream>>>(t2, recs, tile_first, totals, text);
	k_t2_seams<<<grid_for(recs_hint, 256, 8), 256, 0, stream>>>(t2, recs, totals, text);
	return cudaGetLastError();
}
static uint32_t t4_tile() {                // regions per tile (tuning knob, read per launch)
	const char* e = getenv("VSGPU_T4_TILE");
	const uint32_t tile = e ? (uint32_t)atoi(e) : 64;          // measured: 64 -> 101.4 us, 128 -> 102.7, 256 -> 104.5 (1 M regions, k_t4p)
	return (tile == 256 || tile == 128) ? tile : 64;
}
uint32_t t4_wide_entries() {              // scan ranges longer than this many walk entries are taken by a whole warp
	const char* e = getenv("VSGPU_WIDE_ENTRIES");          // test / tuning knob, read per launch
	const uint32_t v = e ? (uint32_t)atoi(e) : 2048;
	return v ? v : 1;
}
uint64_t t4_state_words(uint64_t n) { return 2 + (n + 7) / 8; }

// scratch of the spilling instance: two halves of kSpillWords 32-bit words per CTA of the largest grid launch_t4x uses
constexpr uint32_t kSpillWords = 16384;
static uint32_t spill_words() {               // VSGPU_T4_SPILL_WORDS: a small value makes threads run out of scratch (tests of the second-walk fallback)
	const char* e = getenv("VSGPU_T4_SPILL_WORDS");
	return e ? (uint32_t)std::min<uint64_t>(kSpillWords, std::max<uint64_t>(kSpillChunk, strtoull(e, nullptr, 10) / kSpillChunk * kSpillChunk)) : kSpillWords;
}
// the chunk pool of the warp-per-region kernel: an 8-byte cursor, then kPoolChunks chunks of kPoolChunk codes
constexpr uint32_t kPoolChunks = 262144;
uint64_t t4w_pool_bytes() { return 8 + (uint64_t)kPoolChunks * kPoolChunk * 4; }
uint64_t t4x_spill_bytes() { return (uint64_t)grid_for(~0ull >> 8, 64, 24) * 2 * kSpillWords * 4; }
bool t4x_supported(bool wide_regions) {
	if (wide_regions) return false;
	const char* pe = getenv("VSGPU_T4_PIPE");
	return !pe || atoi(pe) != 0;
}
namespace {
template <uint32_t kTile, uint32_t kMinCtas, bool kAllowSpill = false, uint32_t kKeep = kScratchHits>
cudaError_t launch_t4p_cfg(const DevIndex& ix, const T4Launch& a, cudaStream_t stream) {
	const uint32_t tiles = (uint32_t)((a.n + kTile - 1) / kTile);
	const uint32_t grid = min(tiles, grid_for((uint64_t)tiles * kTile, kTile, kMinCtas));
	const T6Out t6 = a.fuse6 ? *a.fuse6 : T6Out{nullptr, nullptr, nullptr, nullptr, 0};
#define VSGPU_T4P(K32, F6, SP) k_t4p<kTile, kMinCtas, kKeep, K32, F6, SP><<<grid, kTile, 0, stream>>>(ix, a.n, a.x, a.y, a.sample, a.offsets, a.counts, a.hits, a.cap, a.tile_state, a.status, a.base_ptr, t6, a.spill, spill_words())
	if (kAllowSpill && a.spill) {
		if (a.coords32) { if (a.fuse6) VSGPU_T4P(true, true, kAllowSpill); else VSGPU_T4P(true, false, kAllowSpill); }
		else { if (a.fuse6) VSGPU_T4P(false, true, kAllowSpill); else VSGPU_T4P(false, false, kAllowSpill); }
	}
	else if (a.coords32) { if (a.fuse6) VSGPU_T4P(true, true, false); else VSGPU_T4P(true, false, false); }
	else { if (a.fuse6) VSGPU_T4P(false, true, false); else VSGPU_T4P(false, false, false); }
#undef VSGPU_T4P
	return cudaGetLastError();
}
}  // namespace
cudaError_t launch_t4x(const DevIndex& ix, const T4Launch& a, cudaStream_t stream) {
	if (a.n == 0) return cudaSuccess;
	const uint32_t tile = t4_tile();
	if (tile == 128) return launch_t4p_cfg<128, 12>(ix, a, stream);
	if (tile == 256) return launch_t4p_cfg<256, 6>(ix, a, stream);
	return launch_t4p_cfg<64, 24, true>(ix, a, stream);
}
cudaError_t launch_t4(const DevIndex& ix, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample,
                      uint64_t* offsets, uint32_t* hits, uint64_t cap, uint64_t* tile_state, uint32_t* status, bool wide_regions,
                      cudaStream_t stream, const uint64_t* base_ptr, uint32_t* spill) {
	if (n == 0) return cudaSuccess;
	const uint32_t tile = t4_tile();
	const uint32_t grid = (uint32_t)((n + tile - 1) / tile);
	static int min_ctas = 0;                       // tuning knob: registers per thread follow from it
	if (!min_ctas) { const char* e = getenv("VSGPU_T4_MINCTAS"); min_ctas = e ? atoi(e) : 6; }
#define VSGPU_T4_ARGS ix, n, x, y, sample, offsets, hits, cap, tile_state, status, base_ptr
	const char* pe = getenv("VSGPU_T4_PIPE");      // 1: persistent pipelined kernel (default), 0: one CTA per tile
	const int pipe = pe ? atoi(pe) : 1;
	if (wide_regions) {                                                                           // few, wide regions: a warp each
		if (spill) cudaMemsetAsync(spill, 0, 8, stream);                                           // here `spill` is the chunk pool (t4w_pool_bytes()): its cursor
		uint32_t chunks = kPoolChunks;                                                              // VSGPU_T4W_POOL_CHUNKS: a small pool (tests of the second-walk fallback)
		if (const char* e = getenv("VSGPU_T4W_POOL_CHUNKS")) chunks = (uint32_t)std::min<uint64_t>(kPoolChunks, strtoull(e, nullptr, 10));
		k_t4w<1024><<<(uint32_t)((n + 7) / 8), 256, 0, stream>>>(VSGPU_T4_ARGS, spill, spill ? chunks : 0);
	}
	else if (pipe) {
		const T4Launch a{n, x, y, false, sample, offsets, nullptr, hits, cap, tile_state, status, base_ptr, nullptr, spill};
		return launch_t4x(ix, a, stream);
	}
	else if (tile == 64) k_t4<64, 16, kScratchHits><<<grid, 64, 0, stream>>>(VSGPU_T4_ARGS);
	else if (tile == 128) k_t4<128, 10, kScratchHits><<<grid, 128, 0, stream>>>(VSGPU_T4_ARGS);
	else if (min_ctas == 8) k_t4<256, 8, kScratchHits><<<grid, 256, 0, stream>>>(VSGPU_T4_ARGS);
	else if (min_ctas == 5) k_t4<256, 5, kScratchHits><<<grid, 256, 0, stream>>>(VSGPU_T4_ARGS);
	else k_t4<256, 6, kScratchHits><<<grid, 256, 0, stream>>>(VSGPU_T4_ARGS);
#undef VSGPU_T4_ARGS
	return cudaGetLastError();
}

}  // namespace vsgpu
