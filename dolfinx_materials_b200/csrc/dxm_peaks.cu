// libdxm_cuda.so -- measurement support: register-resident FP64 FMA peak, device copy bandwidth and the pure-traffic
// twins of the update kernels (the practical HBM ceiling for a given read/write stream mix).
#include "dxm_internal.cuh"

namespace dxm {

// ---- measurement kernels ---------------------------------------------------------------------
__global__ void fp64_fma_kernel(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = __fma_rn(a0, m, c);
    a1 = __fma_rn(a1, m, c);
    a2 = __fma_rn(a2, m, c);
    a3 = __fma_rn(a3, m, c);
    a4 = __fma_rn(a4, m, c);
    a5 = __fma_rn(a5, m, c);
    a6 = __fma_rn(a6, m, c);
    a7 = __fma_rn(a7, m, c);
  }
  const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 123.456) out[0] = s;
}

__global__ void copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst, int64_t n2) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2;
       i += (int64_t)gridDim.x * blockDim.x)
    __stcs(dst + i, __ldcs(src + i));
}

// pure-traffic twin of the constitutive kernels: NR coalesced read streams, NW coalesced write
// streams, one point per thread, no arithmetic to speak of -- the practical HBM ceiling for that mix
template <int NR, int NW>
__global__ void __launch_bounds__(256, 2)
    stream_mix_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t ld, int64_t n) {
  const int64_t ntile = (n + blockDim.x - 1) / blockDim.x;
  for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int64_t i = tile * blockDim.x + threadIdx.x;
    if (i >= n) continue;
    double v[NR];
#pragma unroll
    for (int c = 0; c < NR; ++c) v[c] = __ldcs(src + (int64_t)c * ld + i);
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < NR; ++c) s += v[c];
#pragma unroll
    for (int c = 0; c < NW; ++c) __stcs(dst + (int64_t)c * ld + i, s + (double)c);
  }
}

}  // namespace dxm

using namespace dxm;

extern "C" {

int dxm_fp64_peak(int device, double* tflops) {
  if (!tflops) return fail("dxm_fp64_peak: NULL argument");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop{};
  CK(cudaGetDeviceProperties(&prop, device));
  double* d = nullptr;
  CK(cudaMalloc(&d, 8));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  const int iters = 1 << 15, block = 256, grid = prop.multiProcessorCount * 8;
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(a));
    fp64_fma_kernel<<<grid, block>>>(d, iters, 1.0);
    LAUNCH_CHECK();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    const double fl = 2.0 * 8.0 * iters * (double)block * grid;
    best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d);
  *tflops = best;
  return 0;
}

int dxm_copy_peak(int device, int64_t bytes, double* gbs) {
  if (!gbs || bytes < 16) return fail("dxm_copy_peak: bad argument");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop{};
  CK(cudaGetDeviceProperties(&prop, device));
  double2 *s = nullptr, *d = nullptr;
  CK(cudaMalloc(&s, bytes));
  CK(cudaMalloc(&d, bytes));
  CK(cudaMemset(s, 0, bytes));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  double best = 0;
  for (int rep = 0; rep < 6; ++rep) {
    CK(cudaEventRecord(a));
    copy_kernel<<<prop.multiProcessorCount * 16, 256>>>(s, d, bytes / 16);
    LAUNCH_CHECK();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    best = std::max(best, 2.0 * bytes / (ms * 1e-3) / 1e9);
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(s);
  cudaFree(d);
  *gbs = best;
  return 0;
}


int dxm_stream_peak(int device, int64_t n, int nread, int nwrite, double* gbs) {
  if (!gbs || n < 1) return fail("dxm_stream_peak: bad argument");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop{};
  CK(cudaGetDeviceProperties(&prop, device));
  const int64_t ld = (n + 63) & ~int64_t(63);
  double *s = nullptr, *d = nullptr;
  CK(cudaMalloc(&s, sizeof(double) * ld * nread));
  CK(cudaMalloc(&d, sizeof(double) * ld * nwrite));
  CK(cudaMemset(s, 0, sizeof(double) * ld * nread));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  const int grid = prop.multiProcessorCount * 2;
  double best = 0;
  for (int rep = 0; rep < 6; ++rep) {
    CK(cudaEventRecord(a));
    if (nread == 25 && nwrite == 49)
      stream_mix_kernel<25, 49><<<grid, 256>>>(s, d, ld, n);
    else if (nread == 25 && nwrite == 34)
      stream_mix_kernel<25, 34><<<grid, 256>>>(s, d, ld, n);
    else if (nread == 25 && nwrite == 97)
      stream_mix_kernel<25, 97><<<grid, 256>>>(s, d, ld, n);
    else if (nread == 37 && nwrite == 37)
      stream_mix_kernel<37, 37><<<grid, 256>>>(s, d, ld, n);
    else if (nread == 1 && nwrite == 1)
      stream_mix_kernel<1, 1><<<grid, 256>>>(s, d, ld, n);
    else
      return fail("dxm_stream_peak: supported mixes are 25/49, 25/34, 25/97, 37/37, 1/1");
    LAUNCH_CHECK();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    best = std::max(best, 8.0 * (nread + nwrite) * n / (ms * 1e-3) / 1e9);
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(s);
  cudaFree(d);
  *gbs = best;
  return 0;
}

}  // extern "C"
