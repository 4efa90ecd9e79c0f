// Probe: scattered fp64 accumulation into a large array (the CSR assembly's access pattern) with
//   A. scalar red.global.add.f64 (what fe_forms_kernel does: 3 consecutive doubles per (row, column-node)), vs
//   B. cp.reduce.async.bulk.global.shared::cta.add.f64 of one 32-byte span per (row, column-node) (TMA reduction;
//      bulk operations need 16-byte alignment and sizes, so a 24-byte block is padded with a zero to 32 bytes).
// Spans imitate assembly: "cell" c touches 300 spans inside a window of the value array that slides with c.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_bulk_red scripts/probe_bulk_red.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

// span s of cell c: start (in doubles, even => 16-byte aligned) inside [c * stride, c * stride + window)
__device__ __forceinline__ int64_t span_pos(int64_t c, int s, int64_t stride, int64_t window, int64_t n) {
  const int64_t p = c * stride + (int64_t)(mix((uint64_t)c * 1315423911ull + s) % (uint64_t)window);
  return (p % (n - 4)) & ~(int64_t)1;
}

constexpr int kWarps = 4, kSpans = 300;

__global__ void __launch_bounds__(32 * kWarps) scalar_red(double* vals, int64_t n, int64_t cells, int64_t stride, int64_t window) {
  const int64_t c = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= cells) return;
  for (int s = lane; s < kSpans; s += 32) {
    const int64_t p = span_pos(c, s, stride, window, n);
    atomicAdd(vals + p, 1.0);
    atomicAdd(vals + p + 1, 1.0);
    atomicAdd(vals + p + 2, 1.0);
  }
}

__global__ void __launch_bounds__(32 * kWarps) bulk_red(double* vals, int64_t n, int64_t cells, int64_t stride, int64_t window) {
  __shared__ __align__(16) double stage[kWarps][kSpans][4];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t c = (int64_t)blockIdx.x * kWarps + w;
  if (c >= cells) return;
  for (int s = lane; s < kSpans; s += 32) {
    stage[w][s][0] = 1.0; stage[w][s][1] = 1.0; stage[w][s][2] = 1.0; stage[w][s][3] = 0.0;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  for (int s = lane; s < kSpans; s += 32) {
    const int64_t p = span_pos(c, s, stride, window, n);
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(&stage[w][s][0]);
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 32;" ::"l"(vals + p), "r"(src) : "memory");
  }
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

int main(int argc, char** argv) {
  const int64_t cells = argc > 1 ? atoll(argv[1]) : 663552;
  const int64_t n = 232000000, stride = 350, window = argc > 2 ? atoll(argv[2]) : 2000000;
  double* vals;
  CK(cudaMalloc(&vals, n * sizeof(double)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const unsigned grid = (unsigned)((cells + kWarps - 1) / kWarps);
  for (int variant = 0; variant < 2; ++variant) {
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      CK(cudaMemset(vals, 0, n * sizeof(double)));
      CK(cudaEventRecord(e0));
      if (variant == 0) scalar_red<<<grid, 32 * kWarps>>>(vals, n, cells, stride, window);
      else bulk_red<<<grid, 32 * kWarps>>>(vals, n, cells, stride, window);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) best = ms;
    }
    // checksum: every span adds 3.0 in total
    double* h = (double*)malloc(1 << 20);
    printf("{\"variant\": \"%s\", \"cells\": %lld, \"spans\": %lld, \"ms\": %.4f, \"G_doubles_per_s\": %.2f, \"G_spans_per_s\": %.2f}\n",
           variant == 0 ? "scalar red.f64 x3" : "cp.reduce.async.bulk 32 B", (long long)cells, (long long)cells * kSpans, best,
           cells * kSpans * 3.0 / best / 1e6, cells * (double)kSpans / best / 1e6);
    free(h);
  }
  // correctness of the bulk variant: total sum == 3 * spans
  CK(cudaMemset(vals, 0, n * sizeof(double)));
  bulk_red<<<grid, 32 * kWarps>>>(vals, n, cells, stride, window);
  CK(cudaDeviceSynchronize());
  double* hv = (double*)malloc(n * sizeof(double));
  CK(cudaMemcpy(hv, vals, n * sizeof(double), cudaMemcpyDeviceToHost));
  double sum = 0; for (int64_t i = 0; i < n; ++i) sum += hv[i];
  printf("{\"bulk_sum\": %.1f, \"expected\": %.1f}\n", sum, 3.0 * cells * kSpans);
  return 0;
}
