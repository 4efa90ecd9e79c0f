// Stress of the library's host thread pool (csrc/dxm_host_mirror.hpp: packed-tangent mirror, row gather / scatter) driven from
// three caller threads at once; built with -fsanitize=thread and -fsanitize=address,undefined by tests/test_host_pool_sanitizers.py.
// Test scaffolding only.
#include <cstdio>
#include <thread>
#include <vector>
#include <cmath>
#include "../dolfinx_materials_b200/csrc/dxm_host_mirror.hpp"
int main() {
  const int64_t n = 50000;
  int bad = 0;
  auto work = [&](int seed) {
    std::vector<double> packed(n * 21), full(n * 36 + 2), src(n * 40), dst(n * 40), back(n * 40);
    std::vector<int64_t> rows(n);
    for (int64_t i = 0; i < n * 21; ++i) packed[i] = seed + i * 0.5;
    for (int64_t i = 0; i < n; ++i) rows[i] = (i * 7919 + seed) % n;  // 7919 prime, n not a multiple: a permutation
    for (int64_t i = 0; i < n * 40; ++i) src[i] = seed * 3.0 + i;
    for (int rep = 0; rep < 6; ++rep) {
      double* f = full.data() + (rep & 1);  // aligned and unaligned destinations
      dxm_host::mirror_sym6(packed.data(), f, n, 0);
      const int* map = dxm_host::sym6_map();
      for (int64_t r = 0; r < n; r += 997)
        for (int c = 0; c < 36; ++c)
          if (f[r * 36 + c] != packed[r * 21 + map[c]]) ++bad;
      dxm_host::gather_rows(src.data(), rows.data(), n, 40, dst.data(), 0);
      dxm_host::scatter_rows(back.data(), rows.data(), n, 40, dst.data(), 0);
      for (int64_t i = 0; i < n * 40; i += 1013)
        if (back[i] != src[i]) ++bad;
    }
  };
  std::thread a(work, 1), b(work, 2);
  work(3);
  a.join();
  b.join();
  std::printf("bad=%d\n", bad);
  return bad != 0;
}
