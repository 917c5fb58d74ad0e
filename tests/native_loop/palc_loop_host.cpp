// palc_loop_host.cpp -- TEST HARNESS (not product code): instantiates the library's PALC loop template
// (bifurcationkit.jl_b200/csrc/bk_palc_loop.hpp, the body of bk_palc_run) with a host backend whose problem and solvers are
// C callbacks, so that the CPU suite can compare the native loop row by row with the Python host loop (palc.py) and with the
// reference's known answers (test/continuation/test-cont-non-vector.jl:22-45, simple_continuation.jl) without a GPU.
// Built by tests/test_native_loop_cpu.py with g++; the product library never links this file.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include "../../bifurcationkit.jl_b200/csrc/bk_palc_loop.hpp"

extern "C" {
struct host_opts {
  double ds, dsmin, dsmax, a, p_min, p_max, theta, eta, newton_tol, fd_eps;
  int32_t max_steps, newton_maxit, tangent, normc;
};
struct host_callbacks {
  void (*residual)(const double* x, double p, double* out);
  void (*jacobian)(const double* x, double p);
  int32_t (*linsolve)(const double* rhs, double* out, int32_t* iters);
  int32_t (*bls)(const double* dR, const double* dzu, double dzp, const double* R, double n, double xiu, double xip, double dotscale,
                 double* dX, double* dl, int32_t* iters);
  int32_t (*on_step)(int32_t step, const double* row, const double* z_u, double z_p);  // may be NULL
};
struct host_result {
  int32_t nrows, steps, nfail, stopped;
  int64_t work_newton, work_linear;
  double p_final, ds_final;
};
}

namespace {
struct HostBackend {
  using Vec = double*;
  long long n;
  int normc;
  const host_callbacks* cb;
  long long size() const { return n; }
  Vec alloc() { return (double*)calloc((size_t)n, sizeof(double)); }
  void release(Vec v) { free(v); }
  void copy(Vec d, Vec s) { if (d != s) memcpy(d, s, sizeof(double) * (size_t)n); }
  void zero(Vec x) { memset(x, 0, sizeof(double) * (size_t)n); }
  void axpby(Vec y, double a, Vec x, double b) { for (long long i = 0; i < n; ++i) y[i] = a * x[i] + b * y[i]; }
  void scale(Vec x, double a) { for (long long i = 0; i < n; ++i) x[i] *= a; }
  double dot(Vec x, Vec y) { double s = 0; for (long long i = 0; i < n; ++i) s += x[i] * y[i]; return s; }
  double diffdot(Vec x, Vec x0, Vec t) { double s = 0; for (long long i = 0; i < n; ++i) s += (x[i] - x0[i]) * t[i]; return s; }
  double norm2(Vec x) { return std::sqrt(dot(x, x)); }
  double normC(Vec x) {
    if (normc == 0) return norm2(x);
    double m = 0;
    for (long long i = 0; i < n; ++i) m = bkpalc::nanmax2(m, std::fabs(x[i]));
    return m;
  }
  void residual(Vec x, double p, Vec out) { cb->residual(x, p, out); }
  void jacobian(Vec x, double p) { cb->jacobian(x, p); }
  bool linsolve(Vec rhs, Vec out, int& it) { int32_t k = 0; int32_t ok = cb->linsolve(rhs, out, &k); it = k; return ok != 0; }
  bool bls(Vec dR, Vec dzu, double dzp, Vec R, double nn, double xiu, double xip, double dotscale, Vec dX, double& dl, int& it) {
    int32_t k = 0;
    int32_t ok = cb->bls(dR, dzu, dzp, R, nn, xiu, xip, dotscale, dX, &dl, &k);
    it = k;
    return ok != 0;
  }
};
bool step_thunk(void* u, int step, const double* row, double* z_u, double z_p) {
  return static_cast<const host_callbacks*>(u)->on_step(step, row, z_u, z_p) != 0;
}
}  // namespace

// returns 0 ok, -3 start-up Newton failure (the reference throws there, src/Continuation.jl:375-393)
extern "C" int32_t palc_loop_host_run(const host_opts* ho, int64_t n, const host_callbacks* cb, const double* u0, double p0,
                                      const double* u1, double p1, double* rows, int32_t max_rows, double* u_final, host_result* res) {
  bkpalc::Opts o;
  o.ds = ho->ds, o.dsmin = ho->dsmin, o.dsmax = ho->dsmax, o.a = ho->a, o.p_min = ho->p_min, o.p_max = ho->p_max;
  o.theta = ho->theta, o.eta = ho->eta, o.newton_tol = ho->newton_tol, o.fd_eps = ho->fd_eps;
  o.max_steps = ho->max_steps, o.newton_maxit = ho->newton_maxit, o.tangent = ho->tangent;
  HostBackend be{n, ho->normc, cb};
  bkpalc::Result R;
  int32_t status = 0;
  try {
    bkpalc::Loop<HostBackend> loop(be, o);
    R = loop.run(const_cast<double*>(u0), p0, const_cast<double*>(u1), u1 != nullptr, p1, rows, max_rows,
                 cb->on_step ? &step_thunk : nullptr, const_cast<host_callbacks*>(cb), u_final);
  } catch (const bkpalc::StartupFailure&) {
    status = -3;
  }
  if (res) {
    res->nrows = R.nrows, res->steps = R.steps, res->nfail = R.nfail, res->stopped = R.stopped;
    res->work_newton = R.work_newton, res->work_linear = R.work_linear, res->p_final = R.z_p, res->ds_final = R.ds;
  }
  return status;
}

// step_size_control alone (src/continuation/Contbase.jl:77-102)
extern "C" double palc_loop_host_step_size(const host_opts* ho, double ds, int32_t converged, int32_t itnewton, int32_t* stop) {
  bkpalc::Opts o;
  o.dsmin = ho->dsmin, o.dsmax = ho->dsmax, o.a = ho->a, o.newton_maxit = ho->newton_maxit;
  HostBackend be{1, 0, nullptr};
  bkpalc::Loop<HostBackend> loop(be, o);
  bool s = false;
  double r = loop.step_size_control(ds, converged != 0, itnewton, s);
  *stop = s ? 1 : 0;
  return r;
}
