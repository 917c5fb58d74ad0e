// bk_bls.cu -- S3/S4: bordered linear solvers of src/LinearBorderSolver.jl on device vectors.
//   BorderingBLS  (:88-123) = BEC (:125-144) [+ residualBEC (:146-166) refinement rounds]
//   MatrixFreeBLS (:404-437) = one GMRES on the (N+1)-system through MatrixFreeBLSmap (:299-335),
//                              rhs = vcat(R, n) (use_bordered_array = false, :414,:434)
// dotp(x, y) = dotscale * <x, y>  (PALC passes 1/N: src/continuation/Palc.jl:4, LinearBorderSolver.jl:22).
#include <cmath>
#include "bk_common.cuh"

// BEC: x1, dx = (shift I + J)^-1 R, (shift I + J)^-1 dR ; dl = (n - xiu dotp(dzu,x1)) / (xip dzp - xiu dotp(dzu,dx));
// dX = x1 - dl dx.   All pointers device.  dX receives the result; tmp_dx is scratch.
static int bec(bk_ctx* c, const double* dR, const double* dzu, double dzp, const double* R, double n, double xiu, double xip,
               double a0, double dotscale, const bk_gmres_opts* o, double* dX, double* tmp_dx, double* dl, int* cv,
               int iters[2]) {
  const long long N = c->N;
  OpDesc op = bk_make_op(c, a0, 1.0);
  int c1 = 0, c2 = 0, i1 = 0, i2 = 0;
  int st = bk_gmres_dev(c, op, R, dX, o, &c1, &i1, nullptr);
  if (st < 0) return st;
  st = bk_gmres_dev(c, op, dR, tmp_dx, o, &c2, &i2, nullptr);
  if (st < 0) return st;
  double d1 = 0, d2 = 0;
  BK_TRY(bk_dev_dot(c, dzu, dX, N, &d1));
  BK_TRY(bk_dev_dot(c, dzu, tmp_dx, N, &d2));
  double l = (n - dotscale * d1 * xiu) / (dzp * xip - dotscale * d2 * xiu);
  BK_TRY(bk_dev_axpby(c, dX, -l, tmp_dx, 1.0, N));  // dX = x1 - dl dx
  *dl = l;
  *cv = c1 & c2;
  iters[0] = i1;
  iters[1] = i2;
  return BK_OK;
}

extern "C" int32_t bk_bls_bordering(bk_ctx* c, const double* dR, const double* dzu, double dzp, const double* R, double n,
                                    double xiu, double xip, int32_t has_shift, double shift, double dotscale,
                                    const bk_gmres_opts* opts, int32_t check_precision, int32_t kmax, double tol, double* dX,
                                    double* dl, int32_t* converged, int32_t iters[2]) {
  BK_ENTER(c);
  BkRange nvtx_range("bk_bls_bordering");
  BK_CHECK(c, c->have_state, "bk_jac_set_state must be called first");
  BK_CHECK(c, opts != nullptr, "opts required");
  const long long N = c->N;
  double *d_dR, *d_dzu, *d_R, *d_dX;
  BK_TRY(bk_stage_in(c, dR, N, 6, true, &d_dR));
  BK_TRY(bk_stage_in(c, dzu, N, 7, true, &d_dzu));
  BK_TRY(bk_stage_in(c, R, N, 8, true, &d_R));
  BK_TRY(bk_stage_in(c, dX, N, 9, false, &d_dX));
  double *t_dx, *t_res, *t_dX1;
  BK_TRY(bk_tmp(c, 0, &t_dx));
  const double a0 = has_shift ? shift : 0.0;
  double l = 0;
  int cv = 0, it[2] = {0, 0};
  BK_TRY(bec(c, d_dR, d_dzu, dzp, d_R, n, xiu, xip, a0, dotscale, opts, d_dX, t_dx, &l, &cv, it));
  int k = 0;
  bool fail = true;
  while (check_precision && k < kmax && fail) {
    // residualBEC: dXr = R - (shift I + J) dX - dl dR ; dlr = n - xip dzp dl - xiu dotp(dzu, dX)
    BK_TRY(bk_tmp(c, 1, &t_res));
    BK_TRY(bk_tmp(c, 2, &t_dX1));
    OpDesc op = bk_make_op(c, a0, 1.0);
    BK_TRY(bk_launch_apply(c, op, d_dX, nullptr, t_res));
    BK_TRY(bk_dev_axpby(c, t_res, l, d_dR, 1.0, N));
    BK_TRY(bk_dev_axpby(c, t_res, 1.0, d_R, -1.0, N));
    double dd = 0, nr = 0;
    BK_TRY(bk_dev_dot(c, d_dzu, d_dX, N, &dd));
    double rl = n - xip * dzp * l - xiu * dotscale * dd;
    BK_TRY(bk_dev_dot(c, t_res, t_res, N, &nr));
    fail = sqrt(nr) > tol || fabs(rl) > tol;
    if (fail) {
      double l1 = 0;
      BK_TRY(bec(c, d_dR, d_dzu, dzp, t_res, rl, xiu, xip, a0, dotscale, opts, t_dX1, t_dx, &l1, &cv, it));
      BK_TRY(bk_dev_axpby(c, d_dX, 1.0, t_dX1, 1.0, N));
      l += l1;
      ++k;
    }
  }
  if (dl) *dl = l;
  if (converged) *converged = cv;
  if (iters) {
    iters[0] = it[0];
    iters[1] = it[1];
  }
  BK_TRY(bk_stage_out(c, dX, N, d_dX));
  BK_CUDA(c, cudaStreamSynchronize(c->stream));
  return cv ? BK_OK : BK_NOT_CONVERGED;
}

static __global__ void k_set_tail(double* v, long long idx, double val) { v[idx] = val; }

extern "C" int32_t bk_bls_matrixfree(bk_ctx* c, const double* dR, const double* dzu, double dzp, const double* R, double n,
                                     double xiu, double xip, int32_t has_shift, double shift, double dotscale,
                                     const bk_gmres_opts* opts, double* dX, double* dl, int32_t* converged, int32_t* iters) {
  BK_ENTER(c);
  BkRange nvtx_range("bk_bls_matrixfree");
  BK_CHECK(c, c->have_state, "bk_jac_set_state must be called first");
  BK_CHECK(c, opts != nullptr, "opts required");
  const long long N = c->N;
  double *d_dR, *d_dzu, *d_R;
  BK_TRY(bk_stage_in(c, dR, N, 6, true, &d_dR));
  BK_TRY(bk_stage_in(c, dzu, N, 7, true, &d_dzu));
  BK_TRY(bk_stage_in(c, R, N, 8, true, &d_R));
  double *rhs, *sol;
  BK_TRY(bk_tmp(c, 0, &rhs));
  BK_TRY(bk_tmp(c, 1, &sol));
  BK_TRY(bk_dev_copy(c, rhs, d_R, N));
  k_set_tail<<<1, 1, 0, c->stream>>>(rhs, N, n);
  c->stats.kernel_launches++;
  // linearmap = MatrixFreeBLSmap(J, dR, xiu*dzu, dzp*xip, shift, dotp)  (:433)
  OpDesc op = bk_make_op(c, 0.0, 1.0);
  op.bordered = 1;
  op.ba = d_dR;
  op.bb = d_dzu;
  op.bscale = dotscale * xiu;
  op.bc = dzp * xip;
  op.bshift = has_shift ? shift : 0.0;
  int cv = 0, it = 0;
  int st = bk_gmres_dev(c, op, rhs, sol, opts, &cv, &it, nullptr);
  if (st < 0) return st;
  double tail = 0;
  BK_CUDA(c, cudaMemcpyAsync(c->red_pinned + 1, sol + N, 8, cudaMemcpyDeviceToHost, c->stream));
  BK_CUDA(c, cudaStreamSynchronize(c->stream));
  tail = c->red_pinned[1];
  if (dl) *dl = tail;
  if (converged) *converged = cv;
  if (iters) *iters = it;
  if (bk_is_device_ptr(dX)) {
    BK_TRY(bk_dev_copy(c, dX, sol, N));
    BK_CUDA(c, cudaStreamSynchronize(c->stream));
  } else {
    BK_TRY(bk_stage_out(c, dX, N, sol));
  }
  return cv ? BK_OK : BK_NOT_CONVERGED;
}
