/*
 * dxm_fma_vec.c -- element-wise correctly rounded fused multiply-add for the numpy oracle (oracle/canon.py: fma / fms /
 * fnma).  TEST INFRASTRUCTURE ONLY.  numpy has no fma ufunc and Python 3.12 no math.fma; this is the C99 fma() of
 * <math.h> over arrays.  Strides are in elements and may be 0 (a broadcast scalar operand).
 */
#include <math.h>
#include <stdint.h>

void dxo_fma_vec(int64_t n, const double* a, int64_t sa, const double* b, int64_t sb, const double* c, int64_t sc,
                 double* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = fma(a[i * sa], b[i * sb], c[i * sc]);
}
