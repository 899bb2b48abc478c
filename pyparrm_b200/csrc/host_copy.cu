// Host-to-host copy on a small persistent thread pool, for the staged leg of the host pipeline
// (pageable NumPy array -> pinned staging buffer, pinned staging buffer -> pageable result).
// The reference accepts any ndarray (parrm.py:877-886); only page-locked memory can be the
// source or target of an asynchronous copy, so a pageable recording is moved through a pinned
// ring, and that move -- not PCIe -- bounds the pageable path.  Measured on the GPU box's host
// (scripts/micro/host_copy.c): one thread 9 GB/s, eight 43-50 GB/s; the previous Python
// thread pool used four threads per 29 MB chunk (26 GB/s under the concurrent DMA).
// Streaming (non-temporal) stores for big slices: the destination is read next by the copy
// engine, not by a core, and write-allocate would read every destination line first.
// Host-only C++; no CUDA calls.
#include <emmintrin.h>
#include <pthread.h>
#include <string.h>

#include <new>

#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace parrm {
namespace {

void copy_slice(char* dst, const char* src, size_t n) {
  constexpr size_t kStreamFrom = size_t(256) << 10;
  if (n < kStreamFrom) {
    memcpy(dst, src, n);
    return;
  }
  // head up to a 16-byte boundary of the destination, streamed body, tail
  const size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
  memcpy(dst, src, head);
  dst += head;
  src += head;
  n -= head;
  const size_t blocks = n / 64;
  const __m128i* s = reinterpret_cast<const __m128i*>(src);
  __m128i* d = reinterpret_cast<__m128i*>(dst);
  for (size_t i = 0; i < blocks; ++i) {
    const __m128i a = _mm_loadu_si128(s + 4 * i), b = _mm_loadu_si128(s + 4 * i + 1);
    const __m128i c = _mm_loadu_si128(s + 4 * i + 2), e = _mm_loadu_si128(s + 4 * i + 3);
    _mm_stream_si128(d + 4 * i, a);
    _mm_stream_si128(d + 4 * i + 1, b);
    _mm_stream_si128(d + 4 * i + 2, c);
    _mm_stream_si128(d + 4 * i + 3, e);
  }
  _mm_sfence();
  memcpy(dst + blocks * 64, src + blocks * 64, n - blocks * 64);
}

class CopyPool {
 public:
  // one job at a time (callers are serialised by `call_`); slices are handed out by index
  void run(char* dst, const char* src, size_t n, int n_threads) {
    std::lock_guard<std::mutex> call(call_);
    grow(n_threads - 1);
    const int parts = n_threads;
    const size_t step = (((n + parts - 1) / parts) + 63) & ~size_t(63);  // parts * step >= n
    {
      std::lock_guard<std::mutex> g(m_);
      dst_ = dst;
      src_ = src;
      n_ = n;
      step_ = step;
      next_ = 1;  // slice 0 is the caller's
      parts_ = parts;
      pending_ = parts - 1;
    }
    wake_.notify_all();
    slice(0);
    std::unique_lock<std::mutex> g(m_);
    done_.wait(g, [&] { return pending_ == 0; });
  }

 private:
  void slice(int i) {
    const size_t off = size_t(i) * step_;
    if (off < n_) copy_slice(dst_ + off, src_ + off, n_ - off < step_ ? n_ - off : step_);
  }
  void grow(int workers) {
    while (int(threads_.size()) < workers) {
      threads_.emplace_back([this] { work(); });
      threads_.back().detach();  // parked on the condition variable for the life of the process
    }
  }
  void work() {
    for (;;) {
      int mine = -1;
      {
        // a finished job leaves next_ == parts_; run() resets next_ only after every slice of
        // the previous job has been reported done
        std::unique_lock<std::mutex> g(m_);
        wake_.wait(g, [&] { return next_ < parts_; });
        mine = next_++;
      }
      slice(mine);
      {
        std::lock_guard<std::mutex> g(m_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  std::mutex call_, m_;
  std::condition_variable wake_, done_;
  std::vector<std::thread> threads_;
  char* dst_ = nullptr;
  const char* src_ = nullptr;
  size_t n_ = 0, step_ = 0;
  int next_ = 0, parts_ = 0, pending_ = 0;
};

// Never destroyed: its threads are parked for the life of the process.  A forked child has
// none of the parent's threads, so it starts over with a pool of its own.
CopyPool* g_pool = nullptr;
std::mutex g_pool_mutex;
CopyPool* pool() {
  std::lock_guard<std::mutex> g(g_pool_mutex);
  if (g_pool == nullptr) {
    static bool hooked = false;
    if (!hooked) {
      pthread_atfork(nullptr, nullptr, [] {
        new (&g_pool_mutex) std::mutex();  // may have been held by a thread that is not here
        g_pool = nullptr;
      });
      hooked = true;
    }
    g_pool = new CopyPool();
  }
  return g_pool;
}

}  // namespace
}  // namespace parrm

extern "C" int parrm_host_copy(void* h_dst, const void* h_src, size_t bytes, int n_threads) {
  using namespace parrm;
  PARRM_REQUIRE(bytes == 0 || (h_dst != nullptr && h_src != nullptr), "parrm_host_copy: null pointer");
  PARRM_REQUIRE(n_threads >= 1 && n_threads <= 64, "parrm_host_copy: 1..64 threads, got %d", n_threads);
  if (bytes == 0) return PARRM_OK;
  const size_t per_thread = size_t(1) << 20;  // below 1 MB per thread a pool is not worth waking
  int threads = int(bytes / per_thread);
  threads = threads < 1 ? 1 : (threads > n_threads ? n_threads : threads);
  if (threads == 1) {
    copy_slice(static_cast<char*>(h_dst), static_cast<const char*>(h_src), bytes);
    return PARRM_OK;
  }
  pool()->run(static_cast<char*>(h_dst), static_cast<const char*>(h_src), bytes, threads);
  return PARRM_OK;
}
