/* C (not C++) consumer of include/dxm.h: proves the header is plain C99 and the library is usable without Python.
 *   gcc -std=c99 -Wall -Wextra -Werror -I include tests/abi_c/abi_smoke.c -L dolfinx_materials_b200/lib -ldxm_cuda -lm
 * Without arguments it only checks what needs no GPU (symbols resolve, error paths); with "gpu" it runs a small
 * J2 + linear hardening history through dxm_integrate and checks it against the closed form of
 * tests/mfront/IsotropicLinearHardeningPlasticity.mfront:49-77 evaluated here in C. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dxm.h"

#define CHECK(cond, msg)                                          \
  do {                                                            \
    if (!(cond)) {                                                \
      fprintf(stderr, "FAIL %s (%s)\n", msg, dxm_last_error());  \
      return 1;                                                   \
    }                                                             \
  } while (0)

int main(int argc, char** argv) {
  dxm_handle* h = NULL;
  CHECK(strstr(dxm_version(), "sm_100a") != NULL, "version string");
  CHECK(dxm_create(99, 0, 8, &h) < 0 && h == NULL, "unknown behaviour is rejected");
  CHECK(dxm_create(DXM_J2_LINEAR, 0, 0, &h) < 0, "n = 0 is rejected");
  CHECK(dxm_integrate(NULL, NULL, DXM_MEM_HOST, 0.0, NULL, NULL, NULL, DXM_MEM_HOST, NULL) < 0, "NULL handle");
  CHECK(strlen(dxm_last_error()) > 0, "error text");
  {
    double packed[21], full[36];
    int i;
    for (i = 0; i < 21; ++i) packed[i] = (double)i;
    CHECK(dxm_host_mirror_sym6(packed, full, 1, 1) == 0, "host mirror");
    CHECK(full[0] == 0.0 && full[1] == 1.0 && full[6] == 1.0 && full[7] == 6.0 && full[35] == 20.0, "mirror layout");
  }
  if (argc < 2 || strcmp(argv[1], "gpu") != 0) {
    printf("abi_smoke: C header + host-side entry points ok (no GPU part requested)\n");
    return 0;
  }

  {
    enum { N = 1000 };
    const double E = 70e3, nu = 0.3, sig0 = 250.0, H = 5e3;
    const double mu = E / 2 / (1 + nu), lam = E * nu / (1 + nu) / (1 - 2 * nu);
    double *eps = malloc(sizeof(double) * N * 6), *sig = malloc(sizeof(double) * N * 6);
    double *isv = malloc(sizeof(double) * N * 7), *ct = malloc(sizeof(double) * N * 36);
    dxm_stats st;
    int i, c, rc, n_plastic = 0;
    CHECK(dxm_create(DXM_J2_LINEAR, 0, N, &h) == 0, "dxm_create");
    CHECK(dxm_set_property(h, "E", &E, 1, DXM_MEM_HOST) == 0 && dxm_set_property(h, "nu", &nu, 1, DXM_MEM_HOST) == 0 &&
              dxm_set_property(h, "sig0", &sig0, 1, DXM_MEM_HOST) == 0 && dxm_set_property(h, "H", &H, 1, DXM_MEM_HOST) == 0,
          "dxm_set_property");
    CHECK(dxm_set_property(h, "nope", &E, 1, DXM_MEM_HOST) < 0, "unknown property is rejected");
    for (i = 0; i < N; ++i) { /* isochoric extension of growing amplitude + a little shear */
      const double a = 1e-2 * (double)i / N;
      eps[i * 6 + 0] = a; eps[i * 6 + 1] = -0.5 * a; eps[i * 6 + 2] = -0.5 * a;
      eps[i * 6 + 3] = 0.1 * a; eps[i * 6 + 4] = 0.0; eps[i * 6 + 5] = 0.0;
    }
    rc = dxm_integrate(h, eps, DXM_MEM_HOST, 0.0, sig, isv, ct, DXM_MEM_HOST, &st);
    CHECK(rc == 0, "dxm_integrate");
    for (i = 0; i < N; ++i) { /* closed form from the virgin state */
      double s[6], tr = eps[i * 6] + eps[i * 6 + 1] + eps[i * 6 + 2], ss = 0, seq, dp = 0, ref[6];
      for (c = 0; c < 6; ++c) s[c] = 2 * mu * (eps[i * 6 + c] - (c < 3 ? tr / 3 : 0));
      for (c = 0; c < 6; ++c) ss += s[c] * s[c];
      seq = sqrt(1.5 * ss);
      if (seq - sig0 > 0) { dp = (seq - sig0) / (3 * mu + H); ++n_plastic; }
      for (c = 0; c < 6; ++c) {
        const double n = seq > 0 ? 1.5 * s[c] / seq : 0.0;
        ref[c] = (c < 3 ? (lam + 2 * mu / 3) * tr : 0) + s[c] - 2 * mu * dp * n;
        CHECK(fabs(sig[i * 6 + c] - ref[c]) <= 1e-10 * (fabs(ref[c]) + sig0), "stress vs closed form");
      }
      CHECK(fabs(isv[i * 7] - dp) <= 1e-12 * (dp + 1e-6), "p vs closed form");
      CHECK(ct[i * 36 + 1] == ct[i * 36 + 6], "tangent symmetry");
    }
    CHECK(st.n_points == N && st.n_plastic == n_plastic && st.n_fail == 0 && n_plastic > 100 && n_plastic < N, "statistics");
    CHECK(dxm_update(h) == 0 && dxm_revert(h) == 0 && dxm_destroy(h) == 0, "update / revert / destroy");
    free(eps); free(sig); free(isv); free(ct);
    printf("abi_smoke: %d points through the C ABI from C, %d plastic, closed form matched\n", N, n_plastic);
  }
  return 0;
}
