// g2s_pool.h - a small persistent host thread pool (host-only code, part of the product).  The host stages that
// follow a device pass - subgraph analysis, traceback, statistics lines, splicing - are independent per gap or per
// record once the libc rand() stream has been laid out (g2s_finish.h), so they run as parallel loops on the threads
// the reference would have spent on -nb-cores Dispatcher workers (src/Gap2Seq.cpp:305-434).  Spawning std::threads
// per call costs more than a whole cfg2 device pass, hence the pool.
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace g2s {

class ThreadPool {
 public:
  explicit ThreadPool(int threads) {
    int n = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (n < 1) n = 1;
    if (n > 64) n = 64;
    for (int t = 1; t < n; t++) workers_.emplace_back([this] { loop(); });
  }
  ~ThreadPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto &t : workers_) t.join();
  }
  ThreadPool(const ThreadPool &) = delete;
  ThreadPool &operator=(const ThreadPool &) = delete;
  int size() const { return (int)workers_.size() + 1; }

  // fn(i) for i in [0, n), dynamically scheduled in chunks of `grain`; the caller takes part and returns when all
  // iterations are done.  Not re-entrant (one loop at a time), which is all the engine needs.
  void parallel_for(size_t n, size_t grain, const std::function<void(size_t)> &fn) {
    if (n == 0) return;
    if (grain < 1) grain = 1;
    if (workers_.empty() || n <= grain) { for (size_t i = 0; i < n; i++) fn(i); return; }
    {
      std::lock_guard<std::mutex> lk(mu_);
      job_ = &fn; n_ = n; grain_ = grain; next_.store(0); active_ = 0; gen_++;
    }
    cv_.notify_all();
    work(fn, n, grain);
    std::unique_lock<std::mutex> lk(mu_);
    job_ = nullptr;                       // late wakers of this generation find nothing to do
    done_.wait(lk, [this] { return active_ == 0; });
  }

  // one detached task on a worker (runs inline when the pool has no workers); wait_async() joins it
  void run_async(std::function<void()> fn) {
    if (workers_.empty()) { fn(); return; }
    {
      std::lock_guard<std::mutex> lk(mu_);
      async_ = std::move(fn); async_pending_ = true; async_running_ = true;
    }
    cv_.notify_one();
  }
  void wait_async() {
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return !async_running_; });
  }

 private:
  void work(const std::function<void(size_t)> &fn, size_t n, size_t grain) {
    for (;;) {
      const size_t b = next_.fetch_add(grain);
      if (b >= n) break;
      const size_t e = b + grain < n ? b + grain : n;
      for (size_t i = b; i < e; i++) fn(i);
    }
  }
  void loop() {
    unsigned long long seen = 0;
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
      cv_.wait(lk, [&] { return stop_ || async_pending_ || (gen_ != seen && job_); });
      if (stop_) return;
      if (async_pending_) {
        std::function<void()> f = std::move(async_);
        async_pending_ = false;
        lk.unlock();
        f();
        lk.lock();
        async_running_ = false;
        done_.notify_all();
        continue;
      }
      seen = gen_;
      const std::function<void(size_t)> *fn = job_;
      const size_t n = n_, grain = grain_;
      active_++;
      lk.unlock();
      work(*fn, n, grain);
      lk.lock();
      active_--;
      if (active_ == 0) done_.notify_all();
    }
  }

  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const std::function<void(size_t)> *job_ = nullptr;
  size_t n_ = 0, grain_ = 1;
  std::atomic<size_t> next_{0};
  int active_ = 0;
  unsigned long long gen_ = 0;
  bool stop_ = false;
  std::function<void()> async_;
  bool async_pending_ = false, async_running_ = false;
};

}  // namespace g2s
