/*
 * dxm_canon.h -- canonical fp64 primitives of the plain-C oracle.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * The canonical arithmetic is a fixed sequence of correctly rounded IEEE-754 operations: + - * / sqrt rint and the
 * EXPLICIT fused multiply-add FMA / FMS / FNMA below, written by hand at the same positions as in the CUDA kernels
 * (dolfinx_materials_b200/csrc/dxm_canon.cuh: fma_c / fms_c / fnma_c) and in the numpy oracle (oracle/canon.py: fma /
 * fms / fnma).  The file is built with -ffp-contract=off (oracle/Makefile), so the compiler never fuses on its own;
 * -DDXO_UNFUSED turns the three helpers back into two roundings, i.e. the round-1 arithmetic (liboracle_unfused.so:
 * tests assert that both agree to the north star's rtol 1e-10 with identical active sets).
 */
#ifndef DXM_CANON_H
#define DXM_CANON_H
#include <math.h>
#include <stdint.h>

#ifdef DXO_UNFUSED
#define FMA(a, b, c) ((a) * (b) + (c))  /* a*b + c */
#define FMS(a, b, c) ((a) * (b) - (c))  /* a*b - c */
#define FNMA(a, b, c) ((c) - (a) * (b)) /* c - a*b */
#else
#define FMA(a, b, c) fma((a), (b), (c))
#define FMS(a, b, c) fma((a), (b), -(c))
#define FNMA(a, b, c) fma(-(a), (b), (c))
#endif

/* exp(x): Cody-Waite reduction, degree-13 Horner in fused steps, exact 2^k scaling (twin of dxm::exp_c / exp_hd) */
static inline double exp_c(double x) {
  const double LOG2E = 1.4426950408889634, LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
  if (x != x) return x;
  if (x < -700.0) return 0.0;
  if (x > 700.0) return INFINITY;
  const double k = rint(x * LOG2E);
  const double r = FNMA(k, LN2_LO, FNMA(k, LN2_HI, x));
  double y = 1.0 / 6227020800.0;
  y = FMA(y, r, 1.0 / 479001600.0);
  y = FMA(y, r, 1.0 / 39916800.0);
  y = FMA(y, r, 1.0 / 3628800.0);
  y = FMA(y, r, 1.0 / 362880.0);
  y = FMA(y, r, 1.0 / 40320.0);
  y = FMA(y, r, 1.0 / 5040.0);
  y = FMA(y, r, 1.0 / 720.0);
  y = FMA(y, r, 1.0 / 120.0);
  y = FMA(y, r, 1.0 / 24.0);
  y = FMA(y, r, 1.0 / 6.0);
  y = FMA(y, r, 0.5);
  y = FMA(y, r, 1.0);
  y = FMA(y, r, 1.0);
  return ldexp(y, (int)k);
}

/* (a0 b0 + a1 b1) + a2 b2 as one product and two fused steps */
static inline double dot3(double a0, double b0, double a1, double b1, double a2, double b2) {
  return FMA(a2, b2, FMA(a1, b1, a0 * b0));
}
#endif
