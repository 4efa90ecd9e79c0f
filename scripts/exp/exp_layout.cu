// Experiment: HBM ceiling of the 25-read / 49-write per-point traffic pattern under different device
// layouts and access widths.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_layout exp_layout.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// layout: T == 0 -> SoA (c*ld + i); T > 0 -> AoSoA tiles of T points: (i/T)*(NC*T) + c*T + i%T
template <int NC, int T>
__device__ __forceinline__ int64_t addr(int64_t i, int c, int64_t ld) {
  if (T == 0) return (int64_t)c * ld + i;
  return (i / T) * (int64_t)(NC * T) + (int64_t)c * T + (i % T);
}

template <int NR, int NW, int T, int HINT, int MINB>
__global__ void __launch_bounds__(256, MINB) mix1(const double* __restrict__ src, double* __restrict__ dst, int64_t ld, int64_t n) {
  const int64_t ntile = (n + blockDim.x - 1) / blockDim.x;
  for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int64_t i = tile * blockDim.x + threadIdx.x;
    if (i >= n) continue;
    double v[NR];
#pragma unroll
    for (int c = 0; c < NR; ++c) v[c] = HINT ? __ldcs(src + addr<NR, T>(i, c, ld)) : src[addr<NR, T>(i, c, ld)];
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < NR; ++c) s += v[c];
#pragma unroll
    for (int c = 0; c < NW; ++c) {
      if (HINT) __stcs(dst + addr<NW, T>(i, c, ld), s + (double)c); else dst[addr<NW, T>(i, c, ld)] = s + (double)c;
    }
  }
}

// two consecutive points per thread, 16-byte accesses
template <int NR, int NW, int T, int MINB>
__global__ void __launch_bounds__(256, MINB) mix2(const double* __restrict__ src, double* __restrict__ dst, int64_t ld, int64_t n) {
  const int64_t ntile = (n + 2 * blockDim.x - 1) / (2 * blockDim.x);
  for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int64_t i = (tile * blockDim.x + threadIdx.x) * 2;
    if (i >= n) continue;
    double2 v[NR];
#pragma unroll
    for (int c = 0; c < NR; ++c) v[c] = __ldcs(reinterpret_cast<const double2*>(src + addr<NR, T>(i, c, ld)));
    double2 s = make_double2(0.0, 0.0);
#pragma unroll
    for (int c = 0; c < NR; ++c) { s.x += v[c].x; s.y += v[c].y; }
#pragma unroll
    for (int c = 0; c < NW; ++c) __stcs(reinterpret_cast<double2*>(dst + addr<NW, T>(i, c, ld)), make_double2(s.x + c, s.y + c));
  }
}

template <typename K>
int run(const char* name, K kern, int grid, const double* s, double* d, int64_t ld, int64_t n, int nr, int nw) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 7; ++rep) {
    CK(cudaEventRecord(a));
    kern<<<grid, 256>>>(s, d, ld, n);
    CK(cudaGetLastError());
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (rep >= 2 && ms < best) best = ms;
  }
  printf("%-44s grid %5d  %8.3f ms  %8.1f GB/s\n", name, grid, best, 8.0 * (nr + nw) * n / (best * 1e-3) / 1e9);
  return 0;
}

int main() {
  const int64_t n = 40000000 / 256 * 256, ld = n;
  double *s, *d;
  CK(cudaMalloc(&s, sizeof(double) * ld * 25)); CK(cudaMalloc(&d, sizeof(double) * ld * 49));
  CK(cudaMemset(s, 0, sizeof(double) * ld * 25));
  const int S = 148;
#define RUN1(T, H, M, G) run("mix1 T=" #T " hint=" #H " minb=" #M, mix1<25, 49, T, H, M>, S * G, s, d, ld, n, 25, 49)
#define RUN2(T, M, G) run("mix2(16B) T=" #T " minb=" #M, mix2<25, 49, T, M>, S * G, s, d, ld, n, 25, 49)
  RUN1(0, 1, 2, 2); RUN1(0, 1, 2, 3); RUN1(0, 1, 2, 4); RUN1(0, 1, 2, 6); RUN1(0, 1, 2, 8); RUN1(0, 1, 2, 16); RUN1(0, 1, 2, 32); RUN1(0, 1, 2, 64); RUN1(0, 1, 2, 256); RUN1(0, 1, 2, 1024);
  RUN1(64, 1, 2, 2); RUN1(64, 1, 2, 4); RUN1(64, 1, 2, 8); RUN1(64, 1, 2, 16); RUN1(64, 1, 2, 32); RUN1(64, 1, 2, 64); RUN1(64, 1, 2, 256); RUN1(64, 1, 2, 1024);
  RUN1(32, 1, 2, 16); RUN1(32, 1, 2, 64); RUN1(256, 1, 2, 16); RUN1(256, 1, 2, 64);
  RUN2(0, 2, 16); RUN2(0, 2, 64); RUN2(64, 2, 16); RUN2(64, 2, 64);
  return 0;
}
