// Host side of the packed-tangent hand-off (host code only, no CUDA).
//
// The small-strain tangent is symmetric; the device keeps its 21 unique entries (include/dxm.h).  The reference
// boundary wants the full row-major (n, 36) array on the HOST (quadrature_map.py:334, utils.py:136-143), and that
// copy is PCIe-bound: 36 doubles per point dominate the 49 the call returns.  With the host mirror the device sends
// the 21 packed entries per point (AoS) into a page-locked ring and a few host threads write the 36-entry rows into
// the caller's array while the next chunk is in flight: 120 B per point less on the link, pure data movement on the
// host (no arithmetic: every value is copied bit for bit).
#pragma once
#include <emmintrin.h>

#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace dxm_host {

// entry c = j*6+i of the row-major symmetric 6x6  ->  packed row (same map as dxm::sym6_packed)
inline const int* sym6_map() {
  static const int* map = [] {
    static int m[36];
    for (int c = 0; c < 36; ++c) {
      int j = c / 6, i = c % 6;
      if (j > i) std::swap(i, j);
      m[c] = j * 6 - (j * (j - 1)) / 2 + (i - j);
    }
    return m;
  }();
  return map;
}

// rows [r0, r1): packed (n, 21) -> full (n, 36).  Streaming (non-temporal) stores when the destination is 16-byte
// aligned: the output is written once and not read back by this library, so it should not displace the packed
// chunk (or the caller's data) from the cache, and no read-for-ownership traffic is spent on it.
inline void mirror_rows(const double* __restrict__ packed, double* __restrict__ full, int64_t r0, int64_t r1) {
  if ((reinterpret_cast<uintptr_t>(full) & 15u) == 0) {
#define DXM_PAIR(c, a, b) _mm_stream_pd(d + (c), _mm_set_pd(p[b], p[a]))
    for (int64_t r = r0; r < r1; ++r) {
      const double* p = packed + r * 21;
      double* d = full + r * 36;
      // rows of the symmetric 6x6, two entries per store; packed rows: 0-5 | 6-10 | 11-14 | 15-17 | 18-19 | 20
      DXM_PAIR(0, 0, 1);   DXM_PAIR(2, 2, 3);    DXM_PAIR(4, 4, 5);
      DXM_PAIR(6, 1, 6);   DXM_PAIR(8, 7, 8);    DXM_PAIR(10, 9, 10);
      DXM_PAIR(12, 2, 7);  DXM_PAIR(14, 11, 12); DXM_PAIR(16, 13, 14);
      DXM_PAIR(18, 3, 8);  DXM_PAIR(20, 12, 15); DXM_PAIR(22, 16, 17);
      DXM_PAIR(24, 4, 9);  DXM_PAIR(26, 13, 16); DXM_PAIR(28, 18, 19);
      DXM_PAIR(30, 5, 10); DXM_PAIR(32, 14, 17); DXM_PAIR(34, 19, 20);
    }
#undef DXM_PAIR
    _mm_sfence();
  } else {
    const int* map = sym6_map();
    for (int64_t r = r0; r < r1; ++r) {
      const double* p = packed + r * 21;
      double* d = full + r * 36;
      for (int c = 0; c < 36; ++c) d[c] = p[map[c]];
    }
  }
}

// A small persistent worker pool (created on first use, never destroyed: its threads sleep on a condition variable
// and die with the process; a static destructor joining them at exit would race with the CUDA runtime's teardown).
class Pool {
 public:
  explicit Pool(int nworkers) : nw_(std::max(0, nworkers)) {
    for (int t = 0; t < nw_; ++t) th_.emplace_back([this, t] { loop(t + 1); });
    for (auto& t : th_) t.detach();
  }
  int parties() const { return nw_ + 1; }
  // f(part, nparts) runs on every worker and on the caller (part 0); returns when all parts are done
  void run(const std::function<void(int, int)>& f) {
    if (nw_ == 0) {
      f(0, 1);
      return;
    }
    std::lock_guard<std::mutex> one_caller(run_mu_);  // handles driven from different threads take turns
    {
      std::lock_guard<std::mutex> lk(mu_);
      job_ = &f;
      pending_ = nw_;
      ++gen_;
    }
    cv_.notify_all();
    f(0, nw_ + 1);
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
    job_ = nullptr;
  }

 private:
  void loop(int part) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(int, int)>* f;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        f = job_;
      }
      (*f)(part, nw_ + 1);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  int nw_;
  std::vector<std::thread> th_;
  std::mutex mu_, run_mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int, int)>* job_ = nullptr;
  uint64_t gen_ = 0;
  int pending_ = 0;
};

// DXM_HOST_THREADS=<t> threads take part in the mirror (including the caller).  Default: the cores this rank can
// claim (hardware threads / LOCAL_WORLD_SIZE, minus one for the Python side), at most 8 -- enough to outrun a
// PCIe 5 x16 link (measured, profiles/), few enough not to fight the other ranks of the box.
inline int default_threads() {
  if (const char* e = std::getenv("DXM_HOST_THREADS")) {
    const int v = std::atoi(e);
    if (v > 0) return std::min(v, 64);
  }
  int hw = (int)std::thread::hardware_concurrency();
  if (hw < 1) hw = 1;
  int lw = 1;
  if (const char* e = std::getenv("LOCAL_WORLD_SIZE")) lw = std::max(1, std::atoi(e));
  return std::max(1, std::min(8, hw / lw - 1));
}

inline Pool& pool() {
  static Pool* p = new Pool(default_threads() - 1);
  return *p;
}

// packed (n, 21) -> full (n, 36) on `threads` threads (<= 0: the pool's default)
inline void mirror_sym6(const double* packed, double* full, int64_t n, int threads) {
  if (n <= 0) return;
  if (threads == 1 || n < 4096) {
    mirror_rows(packed, full, 0, n);
    return;
  }
  Pool& p = pool();
  p.run([&](int part, int nparts) {
    // 8-row granularity keeps every part's first output row 64-byte aligned relative to the base
    const int64_t per = ((n + nparts - 1) / nparts + 7) & ~int64_t(7);
    const int64_t r0 = std::min<int64_t>(n, per * part), r1 = std::min<int64_t>(n, r0 + per);
    if (r1 > r0) mirror_rows(packed, full, r0, r1);
  });
}

// Row gather / scatter by index on the pool: dst[r, :] = src[rows[r], :] and dst[rows[r], :] = src[r, :] with rows of
// `row_len` doubles.  These are the two host passes a QuadratureMap over a cell SUBSET cannot avoid (the Function
// arrays span the whole mesh, quadrature_map.py:251-260, utils.py:136-143); numpy's fancy indexing runs them at
// 1-2 GB/s on one core, which is 10x slower than the PCIe copy they sit next to.
inline void gather_rows(const double* src, const int64_t* rows, int64_t n, int64_t row_len, double* dst, int threads) {
  if (n <= 0 || row_len <= 0) return;
  auto body = [&](int64_t r0, int64_t r1) {
    for (int64_t r = r0; r < r1; ++r) std::memcpy(dst + r * row_len, src + rows[r] * row_len, sizeof(double) * row_len);
  };
  if (threads == 1 || n * row_len < 65536) return body(0, n);
  pool().run([&](int part, int nparts) {
    const int64_t per = (n + nparts - 1) / nparts;
    const int64_t r0 = std::min<int64_t>(n, per * part), r1 = std::min<int64_t>(n, r0 + per);
    if (r1 > r0) body(r0, r1);
  });
}

// one row with streaming stores (dst 16-byte aligned, len even): the destination is written once and not read back here,
// so it should neither pull its old contents into the cache (read-for-ownership) nor evict the source
inline void copy_row_stream(double* __restrict__ d, const double* __restrict__ s, int64_t len) {
  for (int64_t c = 0; c < len; c += 2) _mm_stream_pd(d + c, _mm_loadu_pd(s + c));
}

inline void scatter_rows(double* dst, const int64_t* rows, int64_t n, int64_t row_len, const double* src, int threads) {
  if (n <= 0 || row_len <= 0) return;
  // rows of >= 256 B whose starts are all 16-byte aligned take the streaming path (1.6x over memcpy on the build
  // host: 806 MB of tangent rows in 22.5 instead of 36-38 ms)
  const bool stream = (row_len % 2 == 0) && row_len >= 32 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0 &&
                      !(std::getenv("DXM_HOST_NT") && std::atoi(std::getenv("DXM_HOST_NT")) == 0);  // DXM_HOST_NT=0: memcpy
  auto body = [&](int64_t r0, int64_t r1) {
    if (stream) {
      for (int64_t r = r0; r < r1; ++r) copy_row_stream(dst + rows[r] * row_len, src + r * row_len, row_len);
      _mm_sfence();
    } else {
      for (int64_t r = r0; r < r1; ++r) std::memcpy(dst + rows[r] * row_len, src + r * row_len, sizeof(double) * row_len);
    }
  };
  if (threads == 1 || n * row_len < 65536) return body(0, n);
  pool().run([&](int part, int nparts) {  // distinct rows -> disjoint destinations, no synchronisation needed
    const int64_t per = (n + nparts - 1) / nparts;
    const int64_t r0 = std::min<int64_t>(n, per * part), r1 = std::min<int64_t>(n, r0 + per);
    if (r1 > r0) body(r0, r1);
  });
}

}  // namespace dxm_host
