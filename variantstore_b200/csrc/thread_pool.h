// A small persistent worker pool for the router's host loops (routing, scatter, gather): the loops run a few milliseconds, so
// creating and joining 32-64 threads per loop cost as much as the loops themselves.  run(nt, fn) calls fn(t) for t in [0, nt)
// on the pool's threads (and the caller's) and returns when all are done; one job at a time (the router holds its lock).
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace vsgpu {

class ThreadPool {
public:
	explicit ThreadPool(unsigned workers) {
		for (unsigned i = 0; i < workers; i++) th_.emplace_back([this]() { loop(); });
	}
	~ThreadPool() {
		{ std::lock_guard<std::mutex> g(mu_); stop_ = true; gen_++; }
		cv_.notify_all();
		for (auto& t : th_) t.join();
	}
	ThreadPool(const ThreadPool&) = delete;
	ThreadPool& operator=(const ThreadPool&) = delete;
	unsigned size() const { return (unsigned)th_.size() + 1; }       // workers + the calling thread

	void run(unsigned nt, const std::function<void(unsigned)>& fn) {
		if (nt == 0) return;
		if (nt == 1 || th_.empty()) { for (unsigned t = 0; t < nt; t++) fn(t); return; }
		// every job has its own state: a worker that wakes late holds the job it woke for, finds it exhausted and goes back to sleep
		auto job = std::make_shared<Job>();
		job->fn = &fn; job->ntasks = nt; job->left.store(nt, std::memory_order_relaxed);
		{ std::lock_guard<std::mutex> g(mu_); job_ = job; gen_++; }
		cv_.notify_all();
		work(*job);                                                    // the caller takes tasks too
		std::unique_lock<std::mutex> g(mu_);
		done_cv_.wait(g, [&]() { return job->left.load(std::memory_order_acquire) == 0; });
		job_.reset();
	}

	// fn(t, a, b) over [0, n) cut into at most max_threads contiguous pieces of at least min_piece items
	template <class F>
	void par_for(uint64_t n, unsigned max_threads, F&& fn, uint64_t min_piece = 65536) {
		const unsigned nt = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(std::min(max_threads, size()), (n + min_piece - 1) / min_piece));
		if (nt <= 1) { fn(0u, (uint64_t)0, n); return; }
		const uint64_t chunk = (n + nt - 1) / nt;
		run(nt, [&](unsigned t) { const uint64_t a = t * chunk, b = std::min<uint64_t>(n, a + chunk); if (a < b) fn(t, a, b); });
	}

private:
	struct Job { const std::function<void(unsigned)>* fn = nullptr; unsigned ntasks = 0; std::atomic<unsigned> next{0}, left{0}; };
	void work(Job& j) {
		for (;;) {
			const unsigned t = j.next.fetch_add(1, std::memory_order_relaxed);
			if (t >= j.ntasks) return;
			(*j.fn)(t);                                                  // (fn outlives the job's last task: run() returns only after left == 0)
			if (j.left.fetch_sub(1, std::memory_order_acq_rel) == 1) { std::lock_guard<std::mutex> g(mu_); done_cv_.notify_all(); }
		}
	}
	void loop() {
		uint64_t seen = 0;
		for (;;) {
			std::shared_ptr<Job> job;
			{
				std::unique_lock<std::mutex> g(mu_);
				cv_.wait(g, [&]() { return gen_ != seen; });
				seen = gen_;
				if (stop_) return;
				job = job_;
			}
			if (job) work(*job);
		}
	}
	std::vector<std::thread> th_;
	std::mutex mu_;
	std::condition_variable cv_, done_cv_;
	std::shared_ptr<Job> job_;
	uint64_t gen_ = 0;
	bool stop_ = false;
};

}  // namespace vsgpu
