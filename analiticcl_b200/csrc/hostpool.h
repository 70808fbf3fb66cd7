// hostpool.h -- host-side helpers shared by the batch engine and the find_all_matches producer:
// profiling switches, a persistent worker pool, parallel_ranges.
#pragma once
#include <stddef.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <functional>
#include <mutex>
#include <memory>
#include <new>
#include <thread>
#include <utility>
#include <vector>

namespace anl {

// Process-wide recycler of large host blocks (result arrays are hundreds of MB per million queries;
// a fresh mmap per call costs tens of ms of page faults, a recycled block is already mapped).
void* big_block_take(size_t min_bytes, size_t* got_bytes);  // nullptr if nothing suitable is parked
void big_block_give(void* p, size_t bytes);                 // parks or frees

// std::vector whose resize() default-initialises (leaves PODs untouched) instead of zero-filling: the large index
// arrays are sized first and then first-touched / zeroed on all cores (a plain vector would memset them on one
// thread, and writing through data() beyond size() of a reserved vector is undefined behaviour).
template <class T>
struct DefaultInitAlloc : std::allocator<T> {
  template <class U>
  struct rebind {
    using other = DefaultInitAlloc<U>;
  };
  DefaultInitAlloc() = default;
  template <class U>
  DefaultInitAlloc(const DefaultInitAlloc<U>&) noexcept {}
  template <class U>
  void construct(U* p) noexcept(noexcept(::new ((void*)p) U)) {
    ::new ((void*)p) U;
  }
  template <class U, class... A>
  void construct(U* p, A&&... a) {
    ::new ((void*)p) U(std::forward<A>(a)...);
  }
};
template <class T>
using RawVec = std::vector<T, DefaultInitAlloc<T>>;

// Growable array of PODs that does not value-initialise on resize (a std::vector would memset
// hundreds of MB that are overwritten immediately, on one thread).
template <class T>
class PodBuffer {
 public:
  PodBuffer() = default;
  PodBuffer(const PodBuffer&) = delete;
  PodBuffer& operator=(const PodBuffer&) = delete;
  ~PodBuffer() {
    if (p_) big_block_give(p_, cap_ * sizeof(T));
  }
  T* data() { return p_; }
  const T* data() const { return p_; }
  size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }
  void clear() { n_ = 0; }
  void reserve(size_t c) {
    if (c <= cap_) return;
    if (!p_) {
      size_t got = 0;
      if (void* r = big_block_take(c * sizeof(T), &got)) {
        p_ = static_cast<T*>(r);
        cap_ = got / sizeof(T);
        return;
      }
    }
    T* q = static_cast<T*>(realloc(p_, c * sizeof(T)));
    if (!q) throw std::bad_alloc();
    p_ = q;
    cap_ = c;
  }
  void resize(size_t n) {
    if (n > cap_) reserve(std::max(n, cap_ + cap_ / 2));
    n_ = n;
  }
  T& operator[](size_t i) { return p_[i]; }
  const T& operator[](size_t i) const { return p_[i]; }
  const T* begin() const { return p_; }
  const T* end() const { return p_ + n_; }

 private:
  T* p_ = nullptr;
  size_t n_ = 0, cap_ = 0;
};


// Result arrays the GPU writes by DMA (anl_result_set): page-locked blocks from a process-wide recycler (pinning
// hundreds of MB costs tens of ms; a recycled block is pinned already).  Falls back to plain memory when the
// CUDA runtime cannot pin (then the copies are staged by the driver: slower, still correct).  engine.cu.
void* dma_block_take(size_t min_bytes, size_t* got_bytes, bool* pinned);
void dma_block_give(void* p, size_t bytes, bool pinned);

// Growable POD array in DMA-able host memory.  resize() keeps the contents and never initialises new elements.
template <class T>
class DmaBuffer {
 public:
  DmaBuffer() = default;
  DmaBuffer(const DmaBuffer&) = delete;
  DmaBuffer& operator=(const DmaBuffer&) = delete;
  ~DmaBuffer() { release(); }
  T* data() { return p_; }
  const T* data() const { return p_; }
  size_t size() const { return n_; }
  size_t capacity() const { return cap_; }
  bool empty() const { return n_ == 0; }
  void clear() { n_ = 0; }
  void release() {
    if (p_) dma_block_give(p_, cap_ * sizeof(T), pinned_);
    p_ = nullptr;
    n_ = cap_ = 0;
  }
  // NOTE: moves the block when it grows -- the caller makes sure no DMA into the old block is in flight
  void reserve(size_t c) {
    if (c <= cap_) return;
    size_t got = 0;
    bool pinned = false;
    T* q = static_cast<T*>(dma_block_take(c * sizeof(T), &got, &pinned));
    if (!q) throw std::bad_alloc();
    if (p_) {
      if (n_) memcpy(q, p_, n_ * sizeof(T));
      dma_block_give(p_, cap_ * sizeof(T), pinned_);
    }
    p_ = q;
    cap_ = got / sizeof(T);
    pinned_ = pinned;
  }
  void resize(size_t n) {
    if (n > cap_) reserve(std::max(n, cap_ + cap_ / 2));
    n_ = n;
  }
  void assign(size_t n, const T& v) {
    resize(n);
    for (size_t i = 0; i < n; ++i) p_[i] = v;
  }
  T& operator[](size_t i) { return p_[i]; }
  const T& operator[](size_t i) const { return p_[i]; }

 private:
  T* p_ = nullptr;
  size_t n_ = 0, cap_ = 0;
  bool pinned_ = false;
};

inline bool profile_enabled() {
  static int v = -1;
  if (v < 0) v = getenv("ANL_PROFILE") ? 1 : 0;
  return v == 1;
}
struct PhaseTimer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void lap(const char* what) {
    if (!profile_enabled()) return;
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[anl profile] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

inline unsigned host_threads() {
  static unsigned n = 0;
  if (n == 0) {
    n = std::max(1u, std::thread::hardware_concurrency());
    if (const char* e = getenv("ANL_HOST_THREADS")) n = (unsigned)std::max(1, atoi(e));
    n = std::min(n, 64u);
  }
  return n;
}

// Persistent worker pool for the host phases of a batch (encode, post-pass, assembly).  Three
// parallel phases per chunk of the batch call used to mean ~48 thread creations per chunk; the pool's
// workers sleep on a condition variable between phases instead.  One job at a time: a second caller
// (another model / another host thread) that finds the pool busy runs its ranges on fresh threads.
class HostPool {
 public:
  static HostPool& get() {
    static HostPool* p = new HostPool();  // leaked on purpose: workers outlive static destruction
    return *p;
  }
  // runs fn(part) for part in [0, parts); the caller executes part 0.  An exception thrown by any part (std::bad_alloc
  // from a per-thread vector is the realistic one) is caught where it is thrown, the remaining parts still run to
  // completion, and the first exception is rethrown on the calling thread after the join -- so it reaches the
  // function-try-blocks of capi.cpp instead of std::terminate on a worker.
  template <class F>
  void run(unsigned parts, F& fn) {
    if (parts <= 1) {
      fn(0u);
      return;
    }
    std::exception_ptr first_error;
    std::mutex error_m;
    auto guarded = [&fn, &first_error, &error_m](unsigned t) noexcept {
      try {
        fn(t);
      } catch (...) {
        std::lock_guard<std::mutex> lk(error_m);
        if (!first_error) first_error = std::current_exception();
      }
    };
    std::unique_lock<std::mutex> busy(run_m_, std::try_to_lock);
    if (!busy.owns_lock() || parts - 1 > workers_.size()) {
      std::vector<std::thread> th;
      try {
        for (unsigned t = 1; t < parts; ++t) th.emplace_back([&guarded, t]() { guarded(t); });
      } catch (...) {  // thread creation failed: run the parts that got no thread here
        for (unsigned t = 1 + (unsigned)th.size(); t < parts; ++t) guarded(t);
      }
      guarded(0u);
      for (auto& t : th) t.join();
      if (first_error) std::rethrow_exception(first_error);
      return;
    }
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = [&guarded](unsigned t) { guarded(t); };
      parts_ = parts;
      pending_ = parts - 1;
      ++gen_;
    }
    cv_.notify_all();
    guarded(0u);
    {
      std::unique_lock<std::mutex> lk(m_);
      done_.wait(lk, [this]() { return pending_ == 0; });
      job_ = nullptr;
    }
    if (first_error) std::rethrow_exception(first_error);
  }

 private:
  HostPool() {
    const unsigned n = host_threads();
    for (unsigned t = 1; t < n; ++t) workers_.emplace_back([this, t]() { loop(t); });
    for (auto& w : workers_) w.detach();
  }
  void loop(unsigned id) {
    uint64_t seen = 0;
    for (;;) {
      std::function<void(unsigned)> job;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&]() { return gen_ != seen; });
        seen = gen_;
        if (id >= parts_) continue;  // not needed for this job
        job = job_;
      }
      job(id);
      {
        std::lock_guard<std::mutex> lk(m_);
        --pending_;
      }
      done_.notify_one();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_, run_m_;
  std::condition_variable cv_, done_;
  std::function<void(unsigned)> job_;
  unsigned parts_ = 0, pending_ = 0;
  uint64_t gen_ = 0;
};

// A dispatcher thread of a multi-device call sets this: its host phases run inline (every device has its own
// thread already; the shared pool serves one job at a time).
inline bool& serial_ranges_flag() {
  static thread_local bool f = false;
  return f;
}

// fn(thread index, lo, hi) over [0, n) split into contiguous ranges, one per thread
template <class F>
inline unsigned parallel_ranges(uint64_t n, uint64_t min_per_thread, F fn) {
  unsigned nt = serial_ranges_flag() ? 1u : host_threads();
  const uint64_t mp = std::max<uint64_t>(1, min_per_thread);
  if (n / mp < nt) nt = (unsigned)std::max<uint64_t>(1, n / mp);
  if (nt <= 1) {
    fn(0u, (uint64_t)0, n);
    return 1;
  }
  const uint64_t per = (n + nt - 1) / nt;
  std::vector<double> took(profile_enabled() ? nt : 0, 0.0);
  auto part = [&](unsigned t) {
    const uint64_t lo = std::min(n, (uint64_t)t * per), hi = std::min(n, lo + per);
    if (took.empty()) {
      fn(t, lo, hi);
    } else {
      const auto t0 = std::chrono::steady_clock::now();
      fn(t, lo, hi);
      took[t] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
  };
  const auto t0 = std::chrono::steady_clock::now();
  HostPool::get().run(nt, part);
  if (!took.empty()) {
    const double all = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (all > 15.0) {
      double mx = 0;
      for (double v : took) mx = std::max(mx, v);
      fprintf(stderr, "[anl profile] slow parallel phase: %.2f ms wall, slowest part %.2f ms, own part %.2f ms, %u parts\n", all, mx,
              took[0], nt);
    }
  }
  return nt;
}


}  // namespace anl
