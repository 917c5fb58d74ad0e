// bk_palc_loop.hpp -- the PALC continuation loop as host C++ over an abstract vector / solver backend.
//
// What the reference runs in Julia around the device hot path:
//   continuation / iterate / continuation!     src/Continuation.jl:349-504, 506-601   (two start-up Newton solves, first
//                                              tangent from two points, the step loop, `done` :254-257, save :259-272)
//   newton_palc                                src/continuation/Palc.jl:187-305       (linesearch = false)
//   arc_length_eq / DotTheta                   src/continuation/Palc.jl:1-56
//   secant / Bordered tangents, addtangent!    src/continuation/Tangents.jl:8-42, 71-104
//   step_size_control!                         src/continuation/Contbase.jl:77-102
//   _newton                                    src/Newton.jl:66-114
//   corrector at the parameter bounds          src/continuation/Palc.jl:157-160 (Natural corrector)
//
// Pure C++17, no CUDA: libbk200.so instantiates it with the device backend (bk_palc.cu -> bk_palc_run); the CPU tests
// instantiate it with a callback backend (tests/native_loop/) and compare it row by row with the Python host loop
// (bifurcationkit.jl_b200/palc.py), whose floating-point expressions it repeats operation by operation -- on the device
// both loops therefore produce bit-identical branches (all reductions are deterministic).
//
// Backend concept B:
//   using Vec = <handle>;              Vec alloc(); void release(Vec);   long long size();
//   void copy(Vec dst, Vec src); void zero(Vec); void axpby(Vec y, double a, Vec x, double b);   // y = a x + b y
//   void scale(Vec x, double a); double dot(Vec, Vec); double diffdot(Vec x, Vec x0, Vec tau);   // <x - x0, tau>
//   double norm2(Vec); double normC(Vec);                                                        // record / Newton norm
//   void residual(Vec x, double p, Vec out);  void jacobian(Vec x, double p);                    // J = jacobian(prob, x, p)
//   bool linsolve(Vec rhs, Vec out, int& iters);                                                 // J out = rhs
//   bool bls(Vec dR, Vec dzu, double dzp, Vec R, double n, double xiu, double xip, double dotscale, Vec dX, double& dl, int& iters);
// Backend failures are reported by throwing (the ABI entry point catches; nothing crosses the boundary).
#pragma once
#include <cmath>
#include <limits>
#include <stdexcept>
#include <utility>
#include <vector>

namespace bkpalc {

struct Opts {
  double ds = 1e-2, dsmin = 1e-4, dsmax = 1e-1, a = 0.5, p_min = -1.0, p_max = 1.0, theta = 0.5, eta = 150.0;
  double newton_tol = 1e-10;
  double fd_eps = 0.0;   // finite-difference step of dF/dp (Palc.jl:239-240); 0: sqrt(eps) (src/Problems.jl:69)
  int max_steps = 400, newton_maxit = 25;
  int tangent = 0;       // 0 secant, 1 bordered
};

enum { ROW_PARAM = 0, ROW_X = 1, ROW_ITNEWTON = 2, ROW_ITLINEAR = 3, ROW_DS = 4, ROW_STEP = 5, ROW_LEN = 6 };

struct Result {
  int nrows = 0, steps = 0, nfail = 0, stopped = 0;  // stopped: 1 ds fell to dsmin, 2 callback, 3 row buffer full
  long long work_newton = 0, work_linear = 0;        // all corrector work, rejected attempts included
  double z_p = 0, ds = 0;
};

struct StartupFailure : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// max that propagates NaN like Julia's max / norm(x, Inf): a NaN residual must not pass `res < tol`
static inline double nanmax2(double x, double y) {
  return (x == x && y == y) ? (x > y ? x : y) : std::numeric_limits<double>::quiet_NaN();
}

// min(max(v, lo), hi) with the comparison order of the host loops (a NaN v stays NaN and the step is rejected)
static inline double clamp(double v, double lo, double hi) {
  double t = (lo > v) ? lo : v;
  return (hi < t) ? hi : t;
}

template <class B>
class Loop {
 public:
  using Vec = typename B::Vec;
  // callback(step, row, z_u, z_p) -> false stops the run (called at step 0 and after every accepted step)
  using Callback = bool (*)(void* user, int step, const double* row, Vec z_u, double z_p);

  Loop(B& be, const Opts& o) : be_(be), o_(o), N_((double)be.size()) {
    eps_ = o.fd_eps > 0 ? o.fd_eps : std::sqrt(std::numeric_limits<double>::epsilon());
  }
  ~Loop() {
    for (Vec v : owned_) be_.release(v);
  }

  struct Newton {
    bool converged = false;
    int it = 0, itlin = 0;
    double res = 0;
  };

  // src/Newton.jl:66-114: x is updated in place; fx, du are work vectors
  Newton newton(Vec x, double p, Vec fx, Vec du) {
    Newton r;
    be_.residual(x, p, fx);
    r.res = be_.normC(fx);
    while (r.it < o_.newton_maxit && r.res > o_.newton_tol) {
      be_.jacobian(x, p);
      int it = 0;
      be_.linsolve(fx, du, it);
      r.itlin += it;
      be_.axpby(x, -1.0, du, 1.0);  // minus!!(x, u)
      be_.residual(x, p, fx);
      r.res = be_.normC(fx);
      ++r.it;
    }
    r.converged = r.res < o_.newton_tol;
    return r;
  }

  // run from u0 (Newton-corrected at p0, second point at p0 + ds / eta) or from the two points (u0, p0), (u1, p1)
  // (iterate_from_two_points, src/Continuation.jl:408-456).  rows: max_rows x ROW_LEN doubles.
  Result run(Vec u0_in, double p0, Vec u1_in, bool two_points, double p1, double* rows, int max_rows, Callback cb, void* user,
             Vec u_final) {
    Result R;
    Vec u0 = make(), u1 = make(), fx = make(), du = make();
    be_.copy(u0, u0_in);
    if (!two_points) {
      if (!(o_.p_min <= p0 && p0 <= o_.p_max)) throw StartupFailure("p0 outside [p_min, p_max]");
      Newton s0 = newton(u0, p0, fx, du);
      if (!s0.converged) throw StartupFailure("Newton failed to converge for the initial guess");
      p1 = p0 + o_.ds / o_.eta;
      be_.copy(u1, u0);
      Newton s1 = newton(u1, p1, fx, du);
      if (!s1.converged) throw StartupFailure("Newton failed to converge for the initial tangent");
    } else {
      be_.copy(u1, u1_in);
    }
    // state.z = z1, z_old = z0 -> secant tangent; then z <- z0 (initialize!, Palc.jl:112-123)
    z_ = u1;
    zold_ = u0;
    z_p_ = p1;
    zold_p_ = p0;
    tau_ = make();
    zpred_ = make();
    x_ = make();
    dFdp_ = make();
    be_.zero(tau_);
    be_.zero(zpred_);
    tau_p_ = 0.0;
    ds_ = o_.ds;
    secant();
    be_.copy(z_, u0);  // z_u is its own buffer: z_old keeps u0, z is overwritten with it
    z_p_ = p0;
    predict();

    int step = 0, itnewton = 0, itlinear = 0;
    bool converged = true, stop = false;
    auto save = [&]() {
      if (R.nrows >= max_rows) {
        stop = true;
        R.stopped = 3;
        return;
      }
      double* r = rows + (size_t)R.nrows * ROW_LEN;
      r[ROW_PARAM] = z_p_;
      r[ROW_X] = be_.norm2(z_);  // record_from_solution default = norm(x) (src/Problems.jl:286)
      r[ROW_ITNEWTON] = itnewton;
      r[ROW_ITLINEAR] = itlinear;
      r[ROW_DS] = ds_;
      r[ROW_STEP] = step;
      ++R.nrows;
      if (cb && !cb(user, step, r, z_, z_p_)) {
        stop = true;
        R.stopped = 2;
      }
    };
    save();  // step 0
    bool first = true;
    for (;;) {
      if (!first && converged && step <= o_.max_steps && step > 0) save();
      first = false;
      // done, src/Continuation.jl:254-257
      if (!((step <= o_.max_steps) && ((o_.p_min < z_p_ && z_p_ < o_.p_max) || step == 0) && !stop)) break;
      Newton sol;
      double sol_p;
      if (zpred_p_ <= o_.p_min || zpred_p_ >= o_.p_max) {  // Palc.jl:157-160: Natural corrector at the bound
        zpred_p_ = clamp(zpred_p_, o_.p_min, o_.p_max);
        be_.copy(x_, zpred_);
        sol = newton(x_, zpred_p_, fx, du);
        sol_p = zpred_p_;
      } else {
        sol = newton_palc(fx, du, sol_p);
      }
      converged = sol.converged;
      itnewton = sol.it;
      itlinear = sol.itlin;
      R.work_newton += sol.it;
      R.work_linear += sol.itlin;
      if (!sol.converged) ++R.nfail;
      if (sol.converged) {
        std::swap(zold_, z_);  // z_old <- z (buffers swapped), z <- corrected point
        zold_p_ = z_p_;
        be_.copy(z_, x_);
        z_p_ = sol_p;
        ++step;
      }
      if (!stop) {
        bool s = false;
        ds_ = step_size_control(ds_, converged, itnewton, s);
        if (s) {
          stop = true;
          R.stopped = 1;
        }
      }
      if (converged) {
        if (o_.tangent == 0)
          secant();
        else
          bordered_tangent(fx, du);
      }
      predict();
    }
    R.steps = step;
    R.z_p = z_p_;
    R.ds = ds_;
    if (u_final_set(u_final)) be_.copy(u_final, z_);
    return R;
  }

  // src/continuation/Contbase.jl:77-102
  double step_size_control(double ds, bool converged, int itnewton, bool& stop) const {
    stop = false;
    double dsnew;
    if (!converged) {
      if (std::fabs(ds) <= o_.dsmin) {
        stop = true;
        return ds;
      }
      dsnew = std::copysign(std::fmax(std::fabs(ds) / 2, o_.dsmin), ds);
    } else {
      const double Nmax = o_.newton_maxit;
      const double factor = (Nmax - itnewton) / Nmax;
      dsnew = ds * (1 + o_.a * (factor * factor));
    }
    return std::copysign(std::fmin(std::fmax(std::fabs(dsnew), o_.dsmin), o_.dsmax), dsnew);
  }

 private:
  static bool u_final_set(Vec v) { return !(v == Vec()); }
  Vec make() {
    Vec v = be_.alloc();
    owned_.push_back(v);
    return v;
  }
  double dot_theta(Vec u1, Vec u2, double p1, double p2) {  // DotTheta, Palc.jl:1-34
    return be_.dot(u1, u2) / N_ * o_.theta + p1 * p2 * (1.0 - o_.theta);
  }
  // src/continuation/Tangents.jl:28-42: tau = (z - z_old) sign(ds) / ||.||_theta
  void secant() {
    be_.copy(tau_, z_);
    be_.axpby(tau_, -1.0, zold_, 1.0);
    tau_p_ = z_p_ - zold_p_;
    const double alpha = std::copysign(1.0, ds_) / std::sqrt(dot_theta(tau_, tau_, tau_p_, tau_p_));
    be_.scale(tau_, alpha);
    tau_p_ *= alpha;
  }
  // src/continuation/Tangents.jl:71-104
  void bordered_tangent(Vec fx, Vec du) {
    be_.residual(z_, z_p_ + eps_, dFdp_);
    be_.residual(z_, z_p_, fx);
    be_.axpby(dFdp_, -1.0 / eps_, fx, 1.0 / eps_);
    be_.jacobian(z_, z_p_);
    be_.zero(fx);
    double tp = 0;
    int it = 0;
    be_.bls(dFdp_, tau_, tau_p_, fx, 1.0, o_.theta, 1.0 - o_.theta, 1.0 / N_, du, tp, it);
    double alpha = 1.0 / std::sqrt(dot_theta(du, du, tp, tp));
    alpha *= std::copysign(1.0, dot_theta(tau_, du, tau_p_, tp));
    be_.copy(tau_, du);
    be_.scale(tau_, alpha);
    tau_p_ = tp * alpha;
  }
  void predict() {  // addtangent!, Tangents.jl:8-15
    be_.copy(zpred_, z_);
    be_.axpby(zpred_, ds_, tau_, 1.0);
    zpred_p_ = z_p_ + ds_ * tau_p_;
  }
  double arc_length_eq(Vec u, double p) {  // Palc.jl:44-56
    return o_.theta * be_.diffdot(u, z_, tau_) / N_ + (1.0 - o_.theta) * (p - z_p_) * tau_p_ - ds_;
  }
  // src/continuation/Palc.jl:187-305 (linesearch = false); the corrected point is left in x_
  Newton newton_palc(Vec res_f, Vec du, double& p_out) {
    Newton r;
    be_.copy(x_, zpred_);
    double p = zpred_p_;
    be_.residual(x_, p, res_f);
    double res_n = arc_length_eq(x_, p);
    r.res = nanmax2(be_.normC(res_f), std::fabs(res_n));
    while (r.it < o_.newton_maxit && r.res > o_.newton_tol) {
      be_.residual(x_, p + eps_, dFdp_);
      be_.axpby(dFdp_, -1.0 / eps_, res_f, 1.0 / eps_);  // (F(x, p + eps) - F(x, p)) / eps
      be_.jacobian(x_, p);
      double up = 0;
      int it = 0;
      be_.bls(dFdp_, tau_, tau_p_, res_f, res_n, o_.theta, 1.0 - o_.theta, 1.0 / N_, du, up, it);
      r.itlin += it;
      be_.axpby(x_, -1.0, du, 1.0);
      p = clamp(p - up, o_.p_min, o_.p_max);
      be_.residual(x_, p, res_f);
      res_n = arc_length_eq(x_, p);
      r.res = nanmax2(be_.normC(res_f), std::fabs(res_n));
      ++r.it;
    }
    r.converged = r.res < o_.newton_tol;
    p_out = p;
    return r;
  }

  B& be_;
  const Opts& o_;
  double N_, eps_ = 0;
  std::vector<Vec> owned_;
  Vec z_{}, zold_{}, tau_{}, zpred_{}, x_{}, dFdp_{};
  double z_p_ = 0, zold_p_ = 0, tau_p_ = 0, zpred_p_ = 0, ds_ = 0;
};

}  // namespace bkpalc
