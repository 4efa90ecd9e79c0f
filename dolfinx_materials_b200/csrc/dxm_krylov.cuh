// Device-resident Krylov solve of the assembled tangent system: BiCGStab with a (bs x bs) block-Jacobi
// preconditioner, CSR SpMV with a sub-warp per row.  This is NOT part of the drop-in path (the reference hands
// the assembled system to PETSc: solvers.py:182-196, KSP options e.g. finite_strain_elastoplasticity.py:192-200);
// it exists so that config 5's Newton loop can run without DOLFINx / PETSc in this image (scripts/newton_bar.py),
// with u -> gradients -> update -> assembly -> solve all resident on one B200.
//
// All scalars of the recurrence live in device memory: every thread derives alpha / beta / omega from the dot
// product slots, so one iteration is six kernels and no host round trip; the host polls |r| every few iterations.
#pragma once
#include "dxm_fe_forms.cuh"

namespace dxm {

// dot-product slots of one BiCGStab iteration (two banks, ping-pong by iteration parity)
enum { KS_RHO = 0, KS_RV = 1, KS_TS = 2, KS_TT = 3, KS_RR = 4, KS_N = 8 };

struct KrylovVecs {
  double *x, *r, *rhat, *p, *v, *s, *t, *y, *z;
  const double* minv;  // [nblocks][bs*bs] inverse diagonal blocks
  double* slots;       // [2][KS_N]
  double* scal;        // [4]: rho_prev, alpha_prev, omega_prev, unused
  int64_t n;
  int bs;
};

template <int G>
__global__ void __launch_bounds__(256) spmv_dot_kernel(const int64_t* __restrict__ rowptr,
                                                       const int32_t* __restrict__ colidx,
                                                       const double* __restrict__ vals, const double* __restrict__ x,
                                                       double* __restrict__ y, int64_t nrows,
                                                       const double* __restrict__ d1, double* slot1,
                                                       const double* __restrict__ d2, double* slot2) {
  // y = A x; optionally slot1 += (d1 . y), slot2 += (d2 . y)   (d2 == nullptr -> (y . y))
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / G;
  const int lane = threadIdx.x % G;
  double acc = 0.0;
  if (row < nrows) {
    const int64_t lo = rowptr[row], hi = rowptr[row + 1];
    for (int64_t k = lo + lane; k < hi; k += G) acc += vals[k] * x[colidx[k]];
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, G);
  double a1 = 0.0, a2 = 0.0;
  if (row < nrows && lane == 0) {
    y[row] = acc;
    if (slot1) a1 = d1[row] * acc;
    if (slot2) a2 = (d2 ? d2[row] : acc) * acc;
  }
  if (slot1 || slot2) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    __shared__ double sh[2][8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
      sh[0][w] = a1;
      sh[1][w] = a2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double b1 = 0.0, b2 = 0.0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
        b1 += sh[0][i];
        b2 += sh[1][i];
      }
      if (slot1) atomicAdd(slot1, b1);
      if (slot2) atomicAdd(slot2, b2);
    }
  }
}

__device__ __forceinline__ void block_add(double a, double b, double* sa, double* sb) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  __shared__ double sh[2][8];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    sh[0][w] = a;
    sh[1][w] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double b1 = 0.0, b2 = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
      b1 += sh[0][i];
      b2 += sh[1][i];
    }
    if (sa) atomicAdd(sa, b1);
    if (sb) atomicAdd(sb, b2);
  }
}

// inverse of the (bs x bs) diagonal blocks (bs = 1, 2, 3); singular blocks fall back to the identity
__global__ void block_jacobi_kernel(const int64_t* rowptr, const int32_t* colidx, const double* vals, int64_t nblocks,
                                    int bs, double* minv) {
  const int64_t nb = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (nb >= nblocks) return;
  double D[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int i = 0; i < bs; ++i) {
    const int64_t row = nb * bs + i;
    for (int j = 0; j < bs; ++j) {
      const int64_t pos = csr_find(colidx, rowptr[row], rowptr[row + 1], (int32_t)(nb * bs + j));
      D[i][j] = pos >= 0 ? vals[pos] : 0.0;
    }
  }
  double I[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  bool ok = true;
  if (bs == 1) {
    ok = D[0][0] != 0.0;
    if (ok) I[0][0] = 1.0 / D[0][0];
  } else if (bs == 2) {
    const double det = D[0][0] * D[1][1] - D[0][1] * D[1][0];
    ok = det != 0.0 && isfinite(det);
    if (ok) {
      I[0][0] = D[1][1] / det;
      I[0][1] = -D[0][1] / det;
      I[1][0] = -D[1][0] / det;
      I[1][1] = D[0][0] / det;
    }
  } else {
    double c[3][3];
    c[0][0] = D[1][1] * D[2][2] - D[1][2] * D[2][1];
    c[0][1] = D[0][2] * D[2][1] - D[0][1] * D[2][2];
    c[0][2] = D[0][1] * D[1][2] - D[0][2] * D[1][1];
    c[1][0] = D[1][2] * D[2][0] - D[1][0] * D[2][2];
    c[1][1] = D[0][0] * D[2][2] - D[0][2] * D[2][0];
    c[1][2] = D[0][2] * D[1][0] - D[0][0] * D[1][2];
    c[2][0] = D[1][0] * D[2][1] - D[1][1] * D[2][0];
    c[2][1] = D[0][1] * D[2][0] - D[0][0] * D[2][1];
    c[2][2] = D[0][0] * D[1][1] - D[0][1] * D[1][0];
    const double det = (D[0][0] * c[0][0] + D[0][1] * c[1][0]) + D[0][2] * c[2][0];
    ok = det != 0.0 && isfinite(det);
    if (ok)
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) I[i][j] = c[i][j] / det;
  }
  for (int i = 0; i < bs; ++i)
    for (int j = 0; j < bs; ++j) minv[(nb * bs + i) * bs + j] = ok ? I[i][j] : (i == j ? 1.0 : 0.0);
}

__device__ __forceinline__ double apply_minv(const KrylovVecs& k, const double* vec, int64_t i) {
  const int64_t nb = i / k.bs;
  const int r = (int)(i - nb * k.bs);
  double acc = 0.0;
  for (int j = 0; j < k.bs; ++j) acc += k.minv[(nb * k.bs + r) * k.bs + j] * vec[nb * k.bs + j];
  return acc;
}

// r = rhat = b (x0 = 0), p = v = 0, slot RHO (bank 0) = (b . b), scal = {1, 1, 1}
__global__ void __launch_bounds__(256) bicg_init_kernel(KrylovVecs k, const double* __restrict__ b) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  double d = 0.0;
  if (i < k.n) {
    const double bi = b[i];
    k.x[i] = 0.0;
    k.r[i] = bi;
    k.rhat[i] = bi;
    k.p[i] = 0.0;
    k.v[i] = 0.0;
    d = bi * bi;
  }
  block_add(d, d, k.slots + KS_RHO, k.slots + KS_RR);
}

// p = r + beta (p - omega v), beta = (rho/rho_prev)(alpha_prev/omega_prev); clears the other slot bank
__global__ void __launch_bounds__(256) bicg_p_kernel(KrylovVecs k, int bank) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const double rho = k.slots[bank * KS_N + KS_RHO];
  const double beta = (rho / k.scal[0]) * (k.scal[1] / k.scal[2]);
  if (i < k.n) k.p[i] = k.r[i] + beta * (k.p[i] - k.scal[2] * k.v[i]);
  if (i < KS_N) k.slots[(1 - bank) * KS_N + i] = 0.0;
}

// out = M^-1 in
__global__ void __launch_bounds__(256) bicg_prec_kernel(KrylovVecs k, const double* __restrict__ in,
                                                        double* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < k.n) out[i] = apply_minv(k, in, i);
}

// s = r - alpha v, alpha = rho / (rhat . v)
__global__ void __launch_bounds__(256) bicg_s_kernel(KrylovVecs k, int bank) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const double alpha = k.slots[bank * KS_N + KS_RHO] / k.slots[bank * KS_N + KS_RV];
  if (i < k.n) k.s[i] = k.r[i] - alpha * k.v[i];
}

// x += alpha y + omega z ; r = s - omega t ; next bank: RHO = (rhat . r), RR = (r . r); scal <- rho, alpha, omega
__global__ void __launch_bounds__(256) bicg_x_kernel(KrylovVecs k, int bank) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const double* S = k.slots + bank * KS_N;
  const double rho = S[KS_RHO];
  const double alpha = rho / S[KS_RV];
  const double tt = S[KS_TT];
  const double omega = tt != 0.0 ? S[KS_TS] / tt : 0.0;
  double d1 = 0.0, d2 = 0.0;
  if (i < k.n) {
    k.x[i] = k.x[i] + (alpha * k.y[i] + omega * k.z[i]);
    const double ri = k.s[i] - omega * k.t[i];
    k.r[i] = ri;
    d1 = k.rhat[i] * ri;
    d2 = ri * ri;
  }
  double* N = k.slots + (1 - bank) * KS_N;
  block_add(d1, d2, N + KS_RHO, N + KS_RR);
}

__global__ void bicg_scal_kernel(KrylovVecs k, int bank) {
  const double* S = k.slots + bank * KS_N;
  const double rho = S[KS_RHO];
  const double tt = S[KS_TT];
  k.scal[0] = rho;
  k.scal[1] = rho / S[KS_RV];
  k.scal[2] = tt != 0.0 ? S[KS_TS] / tt : 0.0;
}

}  // namespace dxm
