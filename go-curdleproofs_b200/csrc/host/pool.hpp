// Minimal persistent thread pool: parallel_for over the independent proofs of a
// batch (host-side Fiat-Shamir and Fr work between GPU stages).
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace cdlh {

class ThreadPool {
 public:
  explicit ThreadPool(unsigned n) {
    if (n < 1) n = 1;
    for (unsigned i = 0; i + 1 < n; i++) workers_.emplace_back([this] { loop(); });
  }
  ~ThreadPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      gen_++;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  unsigned size() const { return (unsigned)workers_.size() + 1; }

  // fn(i) for i in [0, n); the calling thread participates
  void parallel_for(size_t n, const std::function<void(size_t)>& fn) {
    if (n == 0) return;
    if (n == 1 || workers_.empty()) {
      for (size_t i = 0; i < n; i++) fn(i);
      return;
    }
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn;
      n_ = n;
      next_.store(0);
      pending_ = workers_.size();
      gen_++;
    }
    cv_.notify_all();
    run();
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void run() {
    for (;;) {
      size_t i = next_.fetch_add(1);
      if (i >= n_) break;
      (*fn_)(i);
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
      }
      run();
      {
        std::lock_guard<std::mutex> lk(mu_);
        pending_--;
      }
      done_cv_.notify_one();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(size_t)>* fn_ = nullptr;
  size_t n_ = 0;
  std::atomic<size_t> next_{0};
  size_t pending_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

}  // namespace cdlh
