/* C translation unit that includes include/udales_gpu.h (compiled with gcc as C11, -Wall -Werror -pedantic): proves the
 * header is plain C, prints the layout of udgpu_cfg for comparison with the ctypes / Fortran bind(C) mirrors, and — with a
 * GPU — drives the library the way the Fortran shim does: init, push, 3 substeps, divergence, pull, finalize.
 *   usage: harness layout | harness run                                                                          */
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "udales_gpu.h"

#define OFF(f) printf("  \"%s\": %zu,\n", #f, offsetof(udgpu_cfg, f))

static int layout(void) {
  printf("{\n  \"sizeof\": %zu,\n", sizeof(udgpu_cfg));
  OFF(abi_version); OFF(itot); OFF(imax); OFF(ih); OFF(ihc); OFF(nsv); OFF(zstart); OFF(nprocx); OFF(myidx); OFF(BCxm); OFF(BCtopm);
  OFF(BCzp); OFF(ipoiss); OFF(iadv_mom); OFF(iadv_sv); OFF(lles); OFF(loneeqn); OFF(ltempeq); OFF(lmoist); OFF(dx); OFF(dy); OFF(dzf);
  OFF(dzh); OFF(delta); OFF(numol); OFF(prandtlmoli); OFF(prandtli); OFF(c_vreman); OFF(cs); OFF(Uinf); OFF(Vinf); OFF(e12min);
  OFF(device); OFF(flags); OFF(iadv_thl);
  printf("  \"abi\": %d,\n  \"nfields\": %d\n}\n", UDGPU_ABI_VERSION, (int)UDGPU_NFIELDS);
  return 0;
}

static int run(void) {
  enum { N = 16 };
  static double dzf[N + 2], dzh[N + 1];
  udgpu_cfg c;
  udgpu_t *h = NULL;
  memset(&c, 0, sizeof c);
  c.abi_version = UDGPU_ABI_VERSION;
  c.itot = c.jtot = c.ktot = c.imax = c.jmax = c.kmax = N;
  c.ih = c.jh = c.kh = c.ihc = c.jhc = c.khc = 1;
  c.zstart[0] = c.zstart[1] = c.zstart[2] = 1;
  c.nprocx = c.nprocy = 1;
  c.BCxm = c.BCym = c.BCtopm = c.BCzp = 1;
  c.iadv_mom = 2; c.iadv_sv = 7; c.lles = c.lvreman = 1;
  c.dx = c.dy = 0.5;
  for (int k = 0; k < N + 2; k++) dzf[k] = 0.5;
  for (int k = 0; k < N + 1; k++) dzh[k] = 0.5;
  c.dzf = dzf; c.dzh = dzh;
  c.numol = 1.5e-5; c.prandtlmoli = 1. / 0.71; c.prandtli = 1. / 0.333; c.c_vreman = 0.07; c.cs = -1.; c.e12min = 5e-5;
  c.device = -1;
  int rc = udgpu_init(&c, NULL, &h);
  if (rc == UDGPU_ENODEV) { printf("ENODEV: %s\n", udgpu_last_error()); return 3; }
  if (rc != UDGPU_OK) { printf("init failed %d: %s\n", rc, udgpu_last_error()); return 1; }
  size_t n = 0; int dims[3];
  if (udgpu_field_count(h, UDGPU_U0, &n, dims) != UDGPU_OK || n != (size_t)(N + 2) * (N + 2) * (N + 2)) return 1;
  double *u = (double *)calloc(n, sizeof(double));
  unsigned s = 12345u;
  for (int k = 1; k <= N; k++) for (int j = 1; j <= N; j++) for (int i = 1; i <= N; i++) {
    s = s * 1664525u + 1013904223u;
    u[i + (N + 2) * (j + (size_t)(N + 2) * k)] = 1.0 + 0.05 * ((double)(s >> 8) / 16777216.0 - 0.5);
  }
  if (udgpu_push(h, UDGPU_U0, 0, u) || udgpu_halos(h) || udgpu_boundary(h) || udgpu_pull(h, UDGPU_U0, 0, u) || udgpu_push(h, UDGPU_UM, 0, u)) {
    printf("setup failed: %s\n", udgpu_last_error()); return 1;
  }
  double dt = 0.05; int rk3 = 0;
  for (int q = 0; q < 3; q++)
    if (udgpu_substep(h, &dt, &rk3, 0.05, 0, 1.0, 0.25)) { printf("substep failed: %s\n", udgpu_last_error()); return 1; }
  double dmax, dtot, drms;
  if (udgpu_divergence(h, &dmax, &dtot, &drms)) return 1;
  if (udgpu_pull(h, UDGPU_U0, 0, u)) return 1;
  printf("ok rk3step %d launches %ld divrms %.3e u(8,8,8) %.6f\n", rk3, udgpu_launch_count(h), drms, u[8 + (N + 2) * (8 + (size_t)(N + 2) * 8)]);
  free(u);
  if (udgpu_finalize(h)) return 1;
  return (rk3 == 3 && drms < 1e-12 && isfinite(dmax)) ? 0 : 2;
}

int main(int argc, char **argv) {
  if (argc > 1 && !strcmp(argv[1], "layout")) return layout();
  if (argc > 1 && !strcmp(argv[1], "run")) return run();
  fprintf(stderr, "usage: harness layout|run\n");
  return 64;
}
