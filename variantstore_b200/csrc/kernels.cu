// libvsgpu device side — hand-written sm_100a kernels for the batched region path.
//
//   K1  rank search over the distinct backbone starts (replaces Index::find / Index::is_empty,
//       include/index.h:119-166, i.e. rrr_vector rank/select) — static 32-ary sampled hierarchy,
//       top level staged in shared memory, one 128-byte node per level below it.
//   K2  t6 slice bounds (get_var_in_ref, include/query.h:736-784).
//   K3  t4 sample walk (get_prev_vertex_with_sample + get_sample_var_in_ref, include/query.h:57-113,
//       :618-729): back-walk over the position index, then a forward scan of the compact walk
//       entries with a carrier-set membership test per entry and the chain rules (first carrier
//       wins, detours hide what they skip) resolved in order.
//   K4  t7 lookup (samples_has_var, include/query.h:792-823).
//   K5  single-pass decoupled look-back exclusive scan + ordered compaction of the t4 hits, so the
//       output order equals the reference's push order.
// Integer-only; nothing here is a contraction, so no tensor-core path exists.
#include "device_logic.cuh"

namespace vsgpu {
namespace {
using namespace logic;

__device__ __forceinline__ void stage_top(const DevIndex& ix, uint32_t* s_top) {
	const uint32_t n = ix.lvl_n[ix.nlvl - 1];
	const uint32_t* top = ix.lvl[ix.nlvl - 1];
	for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s_top[i] = __ldg(top + i);
	__syncthreads();
}


__global__ void __launch_bounds__(256) k_t6(const DevIndex ix, uint64_t n, const uint64_t* __restrict__ xs,
                                            const uint64_t* __restrict__ ys, uint2* __restrict__ out, uint32_t* status) {
	__shared__ uint32_t s_top[kTopMax];
	stage_top(ix, s_top);
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		bool bad = false;
		out[i] = t6_bounds(ix, s_top, xs[i], ys[i], &bad);
		if (bad) atomicOr(status, kStatusBadRegion);
	}
}

__global__ void __launch_bounds__(256) k_t7(const DevIndex ix, uint64_t n, const uint64_t* __restrict__ pos,
                                            const uint64_t* __restrict__ qhash, uint32_t* __restrict__ rec, uint32_t* status) {
	__shared__ uint32_t s_top[kTopMax];
	stage_top(ix, s_top);
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		bool bad = false;
		rec[i] = t7_lookup(ix, s_top, pos[i], qhash[i], &bad);
		if (bad) atomicOr(status, kStatusBadRegion);
	}
}

__global__ void __launch_bounds__(256) k_t4_walk(const DevIndex ix, uint64_t n, const uint64_t* __restrict__ xs,
                                                 const uint64_t* __restrict__ ys, const uint32_t* __restrict__ sample,
                                                 uint32_t* __restrict__ counts, uint32_t* __restrict__ scratch, uint32_t* status) {
	__shared__ uint32_t s_top[kTopMax];
	stage_top(ix, s_top);
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		CountSink sink{scratch + i * kScratchHits, 0};
		const uint64_t x = xs[i];
		const uint32_t s = sample[i];
		if (x < 1 || s == 0 || s >= ix.num_samples) atomicOr(status, kStatusBadRegion);
		else walk_any(ix, s_top, x, ys[i], s, sink);
		counts[i] = sink.n;
	}
}

__global__ void __launch_bounds__(256) k_t4_gather(const DevIndex ix, uint64_t n, const uint64_t* __restrict__ xs,
                                                   const uint64_t* __restrict__ ys, const uint32_t* __restrict__ sample,
                                                   const uint32_t* __restrict__ counts, const uint32_t* __restrict__ scratch,
                                                   const uint64_t* __restrict__ offsets, uint32_t* __restrict__ hits, uint64_t cap,
                                                   uint32_t* status) {
	__shared__ uint32_t s_top[kTopMax];
	stage_top(ix, s_top);
	if (offsets[n] > cap) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(status, kStatusOverflow); return; }
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint32_t cnt = counts[i];
		const uint64_t off = offsets[i];
		if (cnt <= kScratchHits) { for (uint32_t j = 0; j < cnt; j++) hits[off + j] = scratch[i * kScratchHits + j]; }
		else { DirectSink sink{hits + off, 0}; walk_any(ix, s_top, xs[i], ys[i], sample[i], sink); }
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

// ------------------------------------------------------------------ decoupled look-back scan
constexpr uint32_t kScanThreads = 256, kScanItems = 8, kScanTile = kScanThreads * kScanItems;
constexpr uint64_t kFlagAgg = 1ull << 62, kFlagIncl = 2ull << 62, kValMask = (1ull << 62) - 1;

__global__ void __launch_bounds__(kScanThreads) k_scan(uint64_t n, const uint32_t* __restrict__ counts, uint64_t* __restrict__ offsets,
                                                      uint64_t* tile_state) {
	__shared__ uint32_t s_tile;
	__shared__ uint64_t s_warp[kScanThreads / 32];
	__shared__ uint64_t s_excl;
	if (threadIdx.x == 0) s_tile = (uint32_t)atomicAdd((unsigned long long*)&tile_state[0], 1ull);   // ticket: tiles start in order
	__syncthreads();
	const uint32_t tile = s_tile;
	volatile uint64_t* state = tile_state + 1;
	const uint64_t base = (uint64_t)tile * kScanTile + (uint64_t)threadIdx.x * kScanItems;
	uint32_t v[kScanItems]; uint64_t tsum = 0;
#pragma unroll
	for (uint32_t j = 0; j < kScanItems; j++) { v[j] = base + j < n ? counts[base + j] : 0; tsum += v[j]; }
	uint64_t incl = tsum;
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { uint64_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
	if (lane == 31) s_warp[warp] = incl;
	__syncthreads();
	uint64_t wpre = 0, agg = 0;
#pragma unroll
	for (uint32_t w = 0; w < kScanThreads / 32; w++) { if (w < warp) wpre += s_warp[w]; agg += s_warp[w]; }
	if (threadIdx.x == 0) {
		uint64_t excl = 0;
		if (tile == 0) { __threadfence(); state[0] = kFlagIncl | agg; }
		else {
			state[tile] = kFlagAgg | agg; __threadfence();
			for (int64_t p = (int64_t)tile - 1; p >= 0; p--) {
				uint64_t st;
				do { st = state[p]; } while ((st >> 62) == 0);
				excl += st & kValMask;
				if ((st >> 62) == 2) break;
			}
			state[tile] = kFlagIncl | (excl + agg); __threadfence();
		}
		s_excl = excl;
	}
	__syncthreads();
	uint64_t run = s_excl + wpre + (incl - tsum);
#pragma unroll
	for (uint32_t j = 0; j < kScanItems; j++) { if (base + j < n) offsets[base + j] = run; run += v[j]; }
	if (base <= n - 1 && n - 1 < base + kScanItems) offsets[n] = run;   // thread owning the last item publishes the total
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
cudaError_t launch_t6(const DevIndex& ix, uint64_t n, const uint64_t* x, const uint64_t* y, uint2* out, uint32_t* status, cudaStream_t stream) {
	if (n == 0) return cudaSuccess;
	k_t6<<<grid_for(n, 256, 8), 256, 0, stream>>>(ix, n, x, y, out, status);
	return cudaGetLastError();
}
cudaError_t launch_t7(const DevIndex& ix, uint64_t n, const uint64_t* pos, const uint64_t* qhash, uint32_t* rec, uint32_t* status, cudaStream_t stream) {
	if (n == 0) return cudaSuccess;
	k_t7<<<grid_for(n, 256, 8), 256, 0, stream>>>(ix, n, pos, qhash, rec, status);
	return cudaGetLastError();
}
cudaError_t launch_t4_walk(const DevIndex& ix, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample,
                           uint32_t* counts, uint32_t* scratch, uint32_t* status, cudaStream_t stream) {
	if (n == 0) return cudaSuccess;
	k_t4_walk<<<grid_for(n, 256, 8), 256, 0, stream>>>(ix, n, x, y, sample, counts, scratch, status);
	return cudaGetLastError();
}
uint64_t scan_state_words(uint64_t n) { return 2 + (n + kScanTile - 1) / kScanTile; }
cudaError_t launch_scan(uint64_t n, const uint32_t* counts, uint64_t* offsets, uint64_t* tile_state, cudaStream_t stream) {
	if (n == 0) return cudaSuccess;
	k_scan<<<(uint32_t)((n + kScanTile - 1) / kScanTile), kScanThreads, 0, stream>>>(n, counts, offsets, tile_state);
	return cudaGetLastError();
}
cudaError_t launch_t4_gather(const DevIndex& ix, uint64_t n, const uint64_t* x, const uint64_t* y, const uint32_t* sample,
                             const uint32_t* counts, const uint32_t* scratch, const uint64_t* offsets, uint32_t* hits,
                             uint64_t cap, uint32_t* status, cudaStream_t stream) {
	if (n == 0) return cudaSuccess;
	k_t4_gather<<<grid_for(n, 256, 8), 256, 0, stream>>>(ix, n, x, y, sample, counts, scratch, offsets, hits, cap, status);
	return cudaGetLastError();
}

}  // namespace vsgpu
