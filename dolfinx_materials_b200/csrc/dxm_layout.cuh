// Layout helpers at the boundary: the reference hands over / consumes C-contiguous (n, dim) AoS
// arrays (quadrature_map.py:313, :331-334), the kernels work on SoA.  Both directions are
// smem-tiled transposes so that global accesses on either side are contiguous; plus the
// counter-based synthetic gradient generator (twin of oracle/synth.py) and small fills.
#pragma once
#include "dxm_canon.cuh"

namespace dxm {

constexpr int kTile = 128;

// AoS rows [0,count) with row stride `rs` and column offset `c0`  ->  SoA dst[c*ld + d0 + i], c<D
__global__ void __launch_bounds__(256)
    aos_to_soa_kernel(const double* __restrict__ src, int64_t rs, int c0, double* __restrict__ dst,
                      int64_t ld, int64_t d0, int64_t count, int D) {
  extern __shared__ double sm[];
  const int stride = D | 1;
  const int64_t ntile = (count + kTile - 1) / kTile;
  for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int64_t base = tile * kTile;
    const int m = (int)((count - base) < kTile ? (count - base) : kTile);
    if (rs == D) {  // fully contiguous tile
      const double* s = src + base * rs;
      for (int idx = threadIdx.x; idx < m * D; idx += blockDim.x)
        sm[(idx / D) * stride + (idx % D)] = __ldcs(s + idx);
    } else {
      for (int idx = threadIdx.x; idx < m * D; idx += blockDim.x) {
        const int p = idx / D, c = idx % D;
        sm[p * stride + c] = __ldcs(src + (base + p) * rs + c0 + c);
      }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < kTile * D; idx += blockDim.x) {
      const int c = idx / kTile, p = idx % kTile;
      if (p < m) dst[(int64_t)c * ld + d0 + base + p] = sm[p * stride + c];
    }
    __syncthreads();
  }
}

// SoA src[c*ld + s0 + i], c<D  ->  AoS rows [0,count) with row stride `rs`, column offset `c0`.
// sym6 != 0: the source holds the 21 packed rows of a symmetric 6x6 tangent, expanded to D = 36 columns on the way out
__global__ void __launch_bounds__(256)
    soa_to_aos_kernel(const double* __restrict__ src, int64_t ld, int64_t s0,
                      double* __restrict__ dst, int64_t rs, int c0, int64_t count, int D, int sym6) {
  extern __shared__ double sm[];
  const int stride = D | 1;
  const int Ds = sym6 ? kSym6Rows : D;
  const int64_t ntile = (count + kTile - 1) / kTile;
  for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int64_t base = tile * kTile;
    const int m = (int)((count - base) < kTile ? (count - base) : kTile);
    for (int idx = threadIdx.x; idx < kTile * Ds; idx += blockDim.x) {
      const int c = idx / kTile, p = idx % kTile;
      if (p < m) sm[p * stride + c] = __ldcs(src + (int64_t)c * ld + s0 + base + p);
    }
    __syncthreads();
    if (rs == D) {
      double* d = dst + base * rs;
      for (int idx = threadIdx.x; idx < m * D; idx += blockDim.x) {
        const int c = idx % D;
        __stcs(d + idx, sm[(idx / D) * stride + (sym6 ? sym6_packed(c) : c)]);
      }
    } else {
      for (int idx = threadIdx.x; idx < m * D; idx += blockDim.x) {
        const int p = idx / D, c = idx % D;
        __stcs(dst + (base + p) * rs + c0 + c, sm[p * stride + (sym6 ? sym6_packed(c) : c)]);
      }
    }
    __syncthreads();
  }
}

__global__ void fill_kernel(double* __restrict__ dst, int64_t count, double v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = v;
}

// splitmix64 of seed + (16*idx + comp + 1) * GOLD  ->  u in [0,1)
__device__ __forceinline__ double synth_uniform(uint64_t seed, uint64_t idx, int comp) {
  uint64_t z = seed + (idx * 16ull + (uint64_t)(comp + 1)) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * 1.1102230246251565e-16;  // 2^-53, exact
}

// recipe 0: eps = ((k/K)*(amp*u_D)) * (2u_c - 1);  recipe 1: F = I + same, D = 9
__global__ void synth_kernel(double* __restrict__ dst, int64_t ld, int64_t count, int D, int recipe,
                             uint64_t seed, double amp, double kfrac, int64_t start) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t g = (uint64_t)(start + i);
    const double a = amp * synth_uniform(seed, g, D);
    const double scale = kfrac * a;
    for (int c = 0; c < D; ++c) {
      const double d = 2.0 * synth_uniform(seed, g, c) - 1.0;
      const double ident = (recipe == 1 && c < 3) ? 1.0 : 0.0;
      dst[(int64_t)c * ld + i] = ident + scale * d;
    }
  }
}

}  // namespace dxm
