// Exercises csrc/thread_pool.h on the CPU: many jobs back to back (every item visited exactly once, any piece count), then
// the cost of four loops as the router runs them, on the pool and with threads created per loop.  Built by tests/test_thread_pool.py
// (also under -fsanitize=thread).
#include <cstdlib>
#include "thread_pool.h"
#include <chrono>
#include <cstdio>
#include <numeric>
using namespace vsgpu;
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
template <class F> void par_for_threads(uint64_t n, unsigned nt, F&& fn) {
	std::vector<std::thread> th; const uint64_t chunk = (n + nt - 1) / nt;
	for (unsigned t = 0; t < nt; t++) { const uint64_t a = t * chunk, b = std::min<uint64_t>(n, a + chunk); if (a < b) th.emplace_back([=, &fn]() { fn(t, a, b); }); }
	for (auto& t : th) t.join();
}
int main(int argc, char** argv) {
	const unsigned workers = argc > 1 ? atoi(argv[1]) : 31;
	ThreadPool pool(workers);
	const uint64_t n = argc > 2 ? strtoull(argv[2], nullptr, 10) : 2'000'000;
	std::vector<uint32_t> v(n); std::iota(v.begin(), v.end(), 0u);
	// correctness: many jobs back to back, every item visited exactly once
	std::vector<std::atomic<uint64_t>> sums(64);
	for (int rep = 0; rep < (argc > 3 ? atoi(argv[3]) : 2000); rep++) {
		for (auto& s : sums) s = 0;
		const uint64_t m = 1000 + rep * 37;
		pool.par_for(m, 1 + rep % 64, [&](unsigned t, uint64_t a, uint64_t b) { uint64_t acc = 0; for (uint64_t i = a; i < b; i++) acc += v[i]; sums[t] += acc; }, 16);
		uint64_t tot = 0; for (auto& s : sums) tot += s;
		if (tot != m * (m - 1) / 2) { printf("FAIL rep %d\n", rep); return 1; }
	}
	// overhead: 4 loops per "call" as the router does, 32 / 32 / 64 / 64 pieces
	std::vector<uint32_t> out(n);
	auto body = [&](unsigned, uint64_t a, uint64_t b) { for (uint64_t i = a; i < b; i++) out[i] = v[i] * 3u; };
	for (int which = 0; which < 2; which++) {
		double best = 1e9;
		for (int rep = 0; rep < 10; rep++) {
			const double t0 = now();
			for (unsigned nt : {32u, 32u, 64u, 64u}) { if (which) pool.par_for(n, nt, body); else par_for_threads(n, std::min(nt, workers + 1), body); }
			best = std::min(best, now() - t0);
		}
		printf("%s: %.3f ms for 4 loops over %lu items\n", which ? "pool" : "threads created per loop", best, (unsigned long)n);
	}
	double best = 1e9;
	for (int rep = 0; rep < 200; rep++) { const double t0 = now(); pool.run(32, [](unsigned) {}); best = std::min(best, now() - t0); }
	printf("empty job on the pool: %.1f us\n", best * 1e3);
	best = 1e9;
	for (int rep = 0; rep < 50; rep++) { const double t0 = now(); par_for_threads(32, 32, [](unsigned, uint64_t, uint64_t) {}); best = std::min(best, now() - t0); }
	printf("32 threads created + joined: %.1f us\n", best * 1e3);
	printf("ok\n");
	return 0;
}
