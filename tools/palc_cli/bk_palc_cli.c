/* bk_palc_cli.c -- a plain C99 caller of the C ABI (include/bk200.h): the localized-front branch of examples/SH2d-fronts.jl through
 * bk_palc_run, one process per GPU (the native driver SURVEY.md 8(b) names beside the Julia adapter and the Python harness).
 *
 *   gcc -std=c99 -O2 -I include tools/palc_cli/bk_palc_cli.c -o bk_palc_cli -L bifurcationkit.jl_b200 -lbk200 -lm \
 *       -Wl,-rpath,'$ORIGIN/../../bifurcationkit.jl_b200'
 *   ./bk_palc_cli [--grid 1024] [--steps 200] [--device $LOCAL_RANK]        (rows as JSON lines on stdout)
 *
 * Start-up as the example does it (:44-80): Newton from sol0 to the hexagons, front guess 0.4 u_hexa exp(-(x + lx)^2 / 25), Newton to
 * the front -- both through bk_palc_run stopped by its callback at step 0 (its start-up Newton solve is src/Newton.jl:66-114). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bk200.h"

static int32_t stop_at_step0(void* user, int32_t step, const double* row, const double* z_u, double z_p) {
  (void)user; (void)row; (void)z_u; (void)z_p;
  return step < 0;  /* 0 = stop: only the start-up Newton solves run */
}

static int fail(bk_ctx* c, const char* what, int st) {
  fprintf(stderr, "bk_palc_cli: %s failed (%d): %s\n", what, st, c ? bk_last_error(c) : "no context");
  return 2;
}

int main(int argc, char** argv) {
  int n = 1024, steps = 200, device = 0;
  for (int i = 1; i + 1 < argc; i += 2) {
    if (!strcmp(argv[i], "--grid")) n = atoi(argv[i + 1]);
    else if (!strcmp(argv[i], "--steps")) steps = atoi(argv[i + 1]);
    else if (!strcmp(argv[i], "--device")) device = atoi(argv[i + 1]);
  }
  /* the domain grows with the grid (mesh width of examples/SH2d-fronts-cuda.jl:66-69), as in bench.py */
  const double pi = 3.14159265358979323846, s = n > 256 ? n / 256.0 : 1.0;
  const double lx = 8 * pi * s, ly = 4 * pi / sqrt(3.0) * s;
  const int64_t dims[3] = {n, n, 1};
  const double lengths[3] = {lx, ly, 1.0};
  const double par[2] = {-0.1, 1.3};
  const int64_t N = (int64_t)n * n;
  bk_ctx* c = NULL;
  int st = bk_ctx_create(device, BK_SH2D, dims, lengths, 100, &c);
  if (st < 0) return fail(c, "bk_ctx_create", st);
  if ((st = bk_set_params(c, par, 2)) < 0) return fail(c, "bk_set_params", st);
  if ((st = bk_precond_setup(c, BK_PC_SH_DCT, 1.0, 1.0)) < 0) return fail(c, "bk_precond_setup", st);   /* (L1 + I)^-1, :121 */
  bk_gmres_opts g = {1e-5, 0.0, 100, 100, BK_SIDE_RIGHT, BK_ORTH_CGS, 1, 0};                          /* :122 */
  bk_palc_opts o;
  memset(&o, 0, sizeof o);
  o.ds = -1e-3, o.dsmin = 1e-4, o.dsmax = 5e-3, o.a = 0.5, o.p_min = -1.0, o.p_max = 0.0, o.theta = 0.5, o.eta = 150.0;  /* :86 */
  o.newton_tol = 1e-9, o.newton_maxit = 30, o.max_steps = 1, o.lens = 0, o.normc = 1;
  double* u = (double*)malloc(sizeof(double) * (size_t)N);
  double* rows = (double*)malloc(sizeof(double) * BK_PALC_ROW * (size_t)(steps + 8));
  if (!u || !rows) return 2;
  /* sol0 (:44-51) */
  double mn = 1e300, mx = -1e300;
  for (int64_t j = 0; j < n; ++j)
    for (int64_t i = 0; i < n; ++i) {
      const double x = -lx + 2 * lx / n * i, y = -ly + 2 * ly / n * j;
      const double v = cos(x) + cos(x / 2) * cos(sqrt(3.0) * y / 2);
      u[i + j * n] = v;
      if (v < mn) mn = v;
      if (v > mx) mx = v;
    }
  for (int64_t k = 0; k < N; ++k) u[k] = ((u[k] - mn) / (mx - mn) - 0.25) * 1.7;
  bk_palc_result r;
  if ((st = bk_palc_run(c, &o, &g, u, par[0], NULL, 0.0, rows, steps + 8, stop_at_step0, NULL, u, &r)) < 0) return fail(c, "hexagons", st);
  for (int64_t j = 0; j < n; ++j)
    for (int64_t i = 0; i < n; ++i) {
      const double x = -lx + 2 * lx / n * i;
      u[i + j * n] *= 0.4 * exp(-(x + lx) * (x + lx) / 25.0);                                          /* :75 */
    }
  if ((st = bk_palc_run(c, &o, &g, u, par[0], NULL, 0.0, rows, steps + 8, stop_at_step0, NULL, u, &r)) < 0) return fail(c, "front", st);
  /* the branch */
  o.newton_maxit = 15, o.max_steps = steps;
  if ((st = bk_palc_run(c, &o, &g, u, par[0], NULL, 0.0, rows, steps + 8, NULL, NULL, NULL, &r)) < 0) return fail(c, "bk_palc_run", st);
  for (int k = 0; k < r.nrows; ++k) {
    const double* q = rows + (size_t)k * BK_PALC_ROW;
    printf("{\"step\": %d, \"param\": %.17g, \"x\": %.17g, \"itnewton\": %d, \"itlinear\": %d, \"ds\": %.6g}\n", (int)q[5], q[0], q[1], (int)q[2],
           (int)q[3], q[4]);
  }
  fprintf(stderr, "bk_palc_cli: %d steps, %d rejected, %lld Newton / %lld Krylov iterations, lambda = %.10f\n", r.steps, r.nfail,
          (long long)r.work_newton, (long long)r.work_linear, r.p_final);
  free(u);
  free(rows);
  bk_ctx_destroy(c);
  return 0;
}
