// bk_cpu_baseline.cpp -- CPU restatement of the reference's Newton-Krylov PALC path for 2-D Swift-Hohenberg in C++17 / OpenMP.
//
// TEST / MEASUREMENT INFRASTRUCTURE ONLY (oracle/): this is the CPU baseline SURVEY.md section 8(d) and BASELINE.md section 3
// prescribe ("CSR SpMV with the kron-assembled L1, MGS GMRES, identical tolerances / preconditioner, all host cores"); the
// product (bifurcationkit.jl_b200/) never links or calls it.  The Julia reference itself cannot run in this image
// (no julia binary; IterativeSolvers.jl / KrylovKit.jl are not vendored), so this file restates, line by line, what the
// reference's CPU path computes for examples/SH2d-fronts.jl:
//   L1 = (I + Lap)^2, Lap assembled with the Neumann closure of examples/SH2d-fronts.jl:13-29  (CSR, sparse product)
//   F(u, l) = -L1 u + l u + nu u^2 - u^3,  dF(u) v = -L1 v + (l + 2 nu u - 3 u^2) v            (:31-34, 124-127)
//   GMRES: IterativeSolvers.gmres semantics as called at src/LinearSolver.jl:186-206 (modified Gram-Schmidt, right
//          preconditioner, tolerance reltol * ||r0|| on the Givens residual, restart, maxiter)   (oracle/krylov.py)
//   MatrixFreeBLS: one GMRES solve on the (N+1) bordered map, src/LinearBorderSolver.jl:299-335,404-437 (oracle/bls.py)
//   newton_palc + secant predictor + step-size control: src/continuation/Palc.jl:187-305, Tangents.jl:28-42,
//          Contbase.jl:77-102, Continuation.jl:349-504 (oracle/palc.py, which is pinned against the reference's tests)
//   preconditioner: (L1 + shift I)^-1 through the DCT-II diagonalisation (examples/SH2d-fronts.jl:121 uses lu(L1 + I);
//          same operator, see tests/test_oracle_palc.py::test_dct_symbol_diagonalises_L1)
// It is checked row by row against the NumPy oracle in tests/test_cpu_baseline.py.
//
// Build: make -C oracle/c   ->  oracle/c/libbkcpu.so  (g++ -O3 -march=x86-64-v3 -fopenmp; portable between this container and the GPU box)
#include <omp.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {
typedef std::complex<double> cpx;

// Large arrays: 64-byte aligned, zeroed by the threads that will work on them (first-touch page placement: on a two-socket
// host a std::vector zeroed by one thread would put every page on one NUMA node and halve the memory bandwidth).
template <class T>
struct PVec {
  T* p = nullptr;
  size_t n = 0;
  PVec() {}
  explicit PVec(size_t n_) { resize(n_); }
  PVec(const T* a, const T* b) { assign(a, b); }
  PVec(const PVec& o) { assign(o.p, o.p + o.n); }
  PVec& operator=(const PVec& o) {
    if (this != &o) assign(o.p, o.p + o.n);
    return *this;
  }
  ~PVec() { free(p); }
  void resize(size_t m) {
    if (m == n) return;
    T* q = m ? (T*)aligned_alloc(64, ((m * sizeof(T) + 63) / 64) * 64) : nullptr;
    const size_t keep = m < n ? m : n;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)m; ++i) q[i] = (size_t)i < keep ? p[i] : T();
    free(p);
    p = q;
    n = m;
  }
  // rows x len array whose row slices are first-touched by the thread that will work on them (the Krylov basis: every BLAS-1
  // loop runs over ONE row with a static schedule, so slice t of every row must live on thread t's NUMA node; a flat parallel
  // first touch would put whole rows on single nodes)
  void resize_rows(size_t rows, size_t len) {
    free(p);
    n = rows * len;
    p = n ? (T*)aligned_alloc(64, ((n * sizeof(T) + 63) / 64) * 64) : nullptr;
    for (size_t r = 0; r < rows; ++r) {
      T* q = p + r * len;
#pragma omp parallel for schedule(static)
      for (long i = 0; i < (long)len; ++i) q[i] = T();
    }
  }
  void assign(size_t m, T v) {
    resize(m);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)m; ++i) p[i] = v;
  }
  void assign(const T* a, const T* b) {
    resize((size_t)(b - a));
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; ++i) p[i] = a[i];
  }
  void swap(PVec& o) {
    std::swap(p, o.p);
    std::swap(n, o.n);
  }
  T* data() { return p; }
  const T* data() const { return p; }
  size_t size() const { return n; }
  T& operator[](size_t i) { return p[i]; }
  const T& operator[](size_t i) const { return p[i]; }
};
typedef PVec<double> vec;

// ------------------------------------------------------------------------------------------------ BLAS-1 (OpenMP)
// algorithmic bytes moved by the BLAS-1 sweeps and the SpMVs since the last bkcpu_reset_counters (control thread only)
static double g_blas1_bytes = 0, g_spmv_bytes = 0;
inline double dot(const double* a, const double* b, long n) {
  g_blas1_bytes += 16.0 * (double)n;
  double s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (long i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}
inline void axpy(double* y, double a, const double* x, long n) {  // y += a x
  g_blas1_bytes += 24.0 * (double)n;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) y[i] += a * x[i];
}
inline void scal_copy(double* y, double a, const double* x, long n) {  // y = a x
  g_blas1_bytes += 16.0 * (double)n;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) y[i] = a * x[i];
}
inline double norminf(const double* a, long n) {
  double m = 0;
#pragma omp parallel for reduction(max : m) schedule(static)
  for (long i = 0; i < n; ++i) m = std::max(m, std::fabs(a[i]));
  return m;
}

// ------------------------------------------------------------------------------------------------ CSR
struct Csr {
  long n = 0;
  PVec<long> ptr;
  PVec<int> col;
  vec val;
  void spmv(const double* x, double* y) const {
    g_spmv_bytes += 12.0 * (double)ptr[n] + 8.0 * (double)(n + 1) + 16.0 * (double)n;  // val + col, row pointers, x (once) and y
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
      double s = 0;
      for (long k = ptr[i]; k < ptr[i + 1]; ++k) s += val[k] * x[col[k]];
      y[i] = s;
    }
  }
};

// I + Lap with the Neumann closure (corner diagonal -1/h^2): examples/SH2d-fronts.jl:13-29, x fastest
Csr assemble_I_plus_lap(int nx, int ny, double lx, double ly) {
  const double hx = 2 * lx / nx, hy = 2 * ly / ny, cx = 1 / (hx * hx), cy = 1 / (hy * hy);
  Csr A;
  A.n = (long)nx * ny;
  A.ptr.assign(A.n + 1, 0L);
  auto nnz_row = [&](int i, int j) { return 1 + (j > 0) + (i > 0) + (i < nx - 1) + (j < ny - 1); };
  for (int j = 0; j < ny; ++j)   // prefix sum (serial: 8 bytes per row)
    for (int i = 0; i < nx; ++i) {
      const long r = i + (long)j * nx;
      A.ptr[r + 1] = A.ptr[r] + nnz_row(i, j);
    }
  A.col.resize(A.ptr[A.n]);
  A.val.resize(A.ptr[A.n]);
#pragma omp parallel for schedule(static)
  for (long r = 0; r < A.n; ++r) {
    const int i = (int)(r % nx), j = (int)(r / nx);
    long o = A.ptr[r];
    // column-sorted entries: (i, j-1), (i-1, j), (i, j), (i+1, j), (i, j+1)
    const double dx = (i == 0 || i == nx - 1) ? -1.0 : -2.0, dy = (j == 0 || j == ny - 1) ? -1.0 : -2.0;
    if (j > 0) { A.col[o] = (int)(r - nx); A.val[o++] = cy; }
    if (i > 0) { A.col[o] = (int)(r - 1); A.val[o++] = cx; }
    A.col[o] = (int)r;
    A.val[o++] = 1.0 + dx * cx + dy * cy;
    if (i < nx - 1) { A.col[o] = (int)(r + 1); A.val[o++] = cx; }
    if (j < ny - 1) { A.col[o] = (int)(r + nx); A.val[o++] = cy; }
  }
  return A;
}
// C = A * A (row-wise sparse accumulator), columns sorted
Csr square(const Csr& A) {
  Csr C;
  C.n = A.n;
  C.ptr.assign(A.n + 1, 0L);
  std::vector<std::vector<std::pair<int, double>>> rows(A.n);
#pragma omp parallel
  {
    std::vector<std::pair<int, double>> acc;
#pragma omp for schedule(static)
    for (long i = 0; i < A.n; ++i) {
      acc.clear();
      for (long k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
        const int m = A.col[k];
        const double a = A.val[k];
        for (long q = A.ptr[m]; q < A.ptr[m + 1]; ++q) {
          const int c = A.col[q];
          bool found = false;
          for (auto& e : acc)
            if (e.first == c) { e.second += a * A.val[q]; found = true; break; }
          if (!found) acc.push_back({c, a * A.val[q]});
        }
      }
      std::sort(acc.begin(), acc.end());
      rows[i] = acc;
    }
  }
  for (long i = 0; i < A.n; ++i) C.ptr[i + 1] = C.ptr[i] + (long)rows[i].size();
  C.col.resize(C.ptr[A.n]);
  C.val.resize(C.ptr[A.n]);
#pragma omp parallel for schedule(static)
  for (long i = 0; i < A.n; ++i) {
    long o = C.ptr[i];
    for (auto& e : rows[i]) { C.col[o] = e.first; C.val[o] = e.second; ++o; }
  }
  return C;
}

// ------------------------------------------------------------------------------------------------ DCT preconditioner
// radix-2 complex FFT (in place, precomputed twiddles, bit reversal); two real lines per transform (z = v1 + i v2)
struct Fft {
  int n = 0, logn = 0;
  std::vector<cpx> tw;     // exp(-2 pi i k / n), k < n/2
  std::vector<int> rev;
  std::vector<cpx> om;     // exp(-i pi k / 2n)
  void init(int n_) {
    n = n_;
    logn = 0;
    while ((1 << logn) < n) ++logn;
    tw.resize(n / 2);
    rev.resize(n);
    om.resize(n);
    const long double PI = 3.14159265358979323846264338327950288L;
    for (int k = 0; k < n / 2; ++k) tw[k] = cpx((double)cosl(-2 * PI * k / n), (double)sinl(-2 * PI * k / n));
    for (int k = 0; k < n; ++k) om[k] = cpx((double)cosl(-PI * k / (2.0L * n)), (double)sinl(-PI * k / (2.0L * n)));
    for (int i = 0; i < n; ++i) {
      int r = 0;
      for (int b = 0; b < logn; ++b) r |= ((i >> b) & 1) << (logn - 1 - b);
      rev[i] = r;
    }
  }
  void run(cpx* a, bool inverse) const {
    for (int i = 0; i < n; ++i)
      if (i < rev[i]) std::swap(a[i], a[rev[i]]);
    for (int len = 2; len <= n; len <<= 1) {
      const int half = len >> 1, step = n / len;
      for (int s = 0; s < n; s += len)
        for (int k = 0; k < half; ++k) {
          cpx w = tw[k * step];
          if (inverse) w = std::conj(w);
          const cpx t = a[s + k + half] * w;
          a[s + k + half] = a[s + k] - t;
          a[s + k] += t;
        }
    }
  }
};
bool is_pow2(int n) { return n >= 2 && (n & (n - 1)) == 0; }

struct DctPrecond {
  int nx = 0, ny = 0;
  Fft fx, fy;
  std::vector<double> lamx, lamy;
  double shift = 1.0;
  vec work, workT;
  void init(int nx_, int ny_, double lx, double ly, double shift_) {
    nx = nx_;
    ny = ny_;
    shift = shift_;
    fx.init(nx);
    fy.init(ny);
    lamx.resize(nx);
    lamy.resize(ny);
    const double hx = 2 * lx / nx, hy = 2 * ly / ny;
    for (int k = 0; k < nx; ++k) lamx[k] = (2 * std::cos(M_PI * k / nx) - 2) / (hx * hx);
    for (int k = 0; k < ny; ++k) lamy[k] = (2 * std::cos(M_PI * k / ny) - 2) / (hy * hy);
    work.assign((size_t)nx * ny, 0.0);
    workT.assign((size_t)nx * ny, 0.0);
  }
  // rows of length n (contiguous), nrows rows: forward (2 C) or inverse (n x) DCT-II of every row, two rows per FFT
  static void rows_dct(const Fft& f, double* a, int n, int nrows, bool inverse) {
#pragma omp parallel
    {
      std::vector<cpx> z(n), zz(n);
#pragma omp for schedule(static)
      for (int r = 0; r < nrows; r += 2) {
        double* x1 = a + (size_t)r * n;
        double* x2 = (r + 1 < nrows) ? x1 + n : nullptr;
        if (!inverse) {
          for (int m = 0; m < n / 2; ++m) {
            z[m] = cpx(x1[2 * m], x2 ? x2[2 * m] : 0.0);
            z[n - 1 - m] = cpx(x1[2 * m + 1], x2 ? x2[2 * m + 1] : 0.0);
          }
          f.run(z.data(), false);
          for (int k = 0; k < n; ++k) {
            const cpx zk = z[k], zc = std::conj(z[(n - k) & (n - 1)]);
            const cpx v1 = zk + zc, v2 = cpx(0, -1) * (zk - zc);
            x1[k] = (f.om[k] * v1).real();
            if (x2) x2[k] = (f.om[k] * v2).real();
          }
        } else {
          for (int k = 0; k < n; ++k) {
            const double c1n = k ? x1[n - k] : 0.0, c2 = x2 ? x2[k] : 0.0, c2n = (k && x2) ? x2[n - k] : 0.0;
            zz[k] = std::conj(f.om[k]) * cpx(x1[k] + c2n, c2 - c1n);
          }
          f.run(zz.data(), true);
          for (int m = 0; m < n / 2; ++m) {
            x1[2 * m] = zz[m].real();
            x1[2 * m + 1] = zz[n - 1 - m].real();
            if (x2) {
              x2[2 * m] = zz[m].imag();
              x2[2 * m + 1] = zz[n - 1 - m].imag();
            }
          }
        }
      }
    }
  }
  static void transpose(const double* a, double* b, int rows, int cols) {  // b[c][r] = a[r][c]
    const int B = 32;
#pragma omp parallel for collapse(2) schedule(static)
    for (int r0 = 0; r0 < rows; r0 += B)
      for (int c0 = 0; c0 < cols; c0 += B)
        for (int r = r0; r < std::min(rows, r0 + B); ++r)
          for (int c = c0; c < std::min(cols, c0 + B); ++c) b[(size_t)c * rows + r] = a[(size_t)r * cols + c];
  }
  void apply(const double* in, double* out) {
    const size_t N = (size_t)nx * ny;
    std::memcpy(work.data(), in, 8 * N);
    rows_dct(fx, work.data(), nx, ny, false);            // x
    transpose(work.data(), workT.data(), ny, nx);         // -> [x][y]
    rows_dct(fy, workT.data(), ny, nx, false);            // y
    const double scale = 1.0 / (4.0 * nx * ny);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nx; ++i)
      for (int j = 0; j < ny; ++j) {
        const double t = 1.0 + lamx[i] + lamy[j];
        workT[(size_t)i * ny + j] *= scale / (t * t + shift);
      }
    rows_dct(fy, workT.data(), ny, nx, true);
    transpose(workT.data(), work.data(), nx, ny);
    rows_dct(fx, work.data(), nx, ny, true);
    std::memcpy(out, work.data(), 8 * N);
  }
};

// ------------------------------------------------------------------------------------------------ problem
struct SH2d {
  int nx, ny;
  long N;
  double nu;
  Csr L1;
  DctPrecond pc;
  vec tmp;
  void F(const double* u, double l, double* out) {
    L1.spmv(u, out);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < N; ++i) out[i] = -out[i] + (l * u[i] + nu * u[i] * u[i] - u[i] * u[i] * u[i]);
  }
  void dF(const double* u, double l, const double* v, double* out) {
    L1.spmv(v, out);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < N; ++i) out[i] = -out[i] + (l + 2.0 * nu * u[i] - 3.0 * u[i] * u[i]) * v[i];
  }
};

// operator of a linear solve: plain Jacobian (n = N) or the bordered map (n = N + 1)
struct LinOp {
  SH2d* p;
  const double* u;
  double l;
  bool bordered = false;
  const double* a = nullptr;   // dR
  const double* b = nullptr;   // xiu * dzu
  double c = 0, dotscale = 1;  // xip * dzp, 1/N
  long n() const { return p->N + (bordered ? 1 : 0); }
  void apply(const double* x, double* out) const {
    p->dF(u, l, x, out);
    if (bordered) {
      const long N = p->N;
      const double xp = x[N];
      axpy(out, xp, a, N);
      out[N] = dot(b, x, N) * dotscale + c * xp;
    }
  }
  void precond(const double* x, double* out) const {  // Pr \ x : DCT inverse on the first N entries, identity on the border
    p->pc.apply(x, out);
    if (bordered) out[p->N] = x[p->N];
  }
};

struct GmresOpts {
  double reltol = 1e-5;
  int restart = 100, maxiter = 100;
};

// right-preconditioned GMRES, modified Gram-Schmidt, initially_zero (oracle/krylov.py::gmres)
int gmres(const LinOp& A, const double* b, double* x, const GmresOpts& o, bool* converged, vec& V, vec& w, vec& z) {
  const long n = A.n();
  const int restart = (int)std::min<long>(o.restart, n);
  if ((long)V.size() < (long)(restart + 1) * n) V.resize_rows((size_t)(restart + 1), (size_t)n);
  if ((long)w.size() < n) w.resize(n);
  if ((long)z.size() < n) z.resize(n);
  std::vector<double> H((size_t)(restart + 1) * restart, 0.0), g(restart + 1), cs(restart), sn(restart), y(restart);
  std::memset(x, 0, 8 * n);
  auto Hm = [&](int i, int k) -> double& { return H[(size_t)k * (restart + 1) + i]; };
  auto init_residual = [&](bool first) {
    if (first) scal_copy(w.data(), 1.0, b, n);
    else {
      A.apply(x, w.data());
#pragma omp parallel for schedule(static)
      for (long i = 0; i < n; ++i) w[i] = b[i] - w[i];
    }
    const double beta = std::sqrt(dot(w.data(), w.data(), n));
    scal_copy(V.data(), beta > 0 ? 1.0 / beta : 1.0, w.data(), n);
    return beta;
  };
  double beta = init_residual(true);
  const double tol = o.reltol * beta;
  double res = beta;
  int total = 0;
  while (total < o.maxiter && res > tol) {
    std::fill(g.begin(), g.end(), 0.0);
    g[0] = beta;
    int k = 0;
    while (k < restart && total < o.maxiter && res > tol) {
      A.precond(V.data() + (size_t)k * n, z.data());
      A.apply(z.data(), w.data());
      for (int i = 0; i <= k; ++i) {
        const double h = dot(V.data() + (size_t)i * n, w.data(), n);
        Hm(i, k) = h;
        axpy(w.data(), -h, V.data() + (size_t)i * n, n);
      }
      const double hk1 = std::sqrt(dot(w.data(), w.data(), n));
      Hm(k + 1, k) = hk1;
      if (hk1 != 0) scal_copy(V.data() + (size_t)(k + 1) * n, 1.0 / hk1, w.data(), n);
      for (int i = 0; i < k; ++i) {
        const double t = cs[i] * Hm(i, k) + sn[i] * Hm(i + 1, k);
        Hm(i + 1, k) = -sn[i] * Hm(i, k) + cs[i] * Hm(i + 1, k);
        Hm(i, k) = t;
      }
      const double d = std::hypot(Hm(k, k), Hm(k + 1, k));
      cs[k] = Hm(k, k) / d;
      sn[k] = Hm(k + 1, k) / d;
      Hm(k, k) = d;
      Hm(k + 1, k) = 0;
      g[k + 1] = -sn[k] * g[k];
      g[k] = cs[k] * g[k];
      res = std::fabs(g[k + 1]);
      ++k;
      ++total;
    }
    if (k > 0) {
      for (int i = k - 1; i >= 0; --i) {
        double t = g[i];
        for (int q = i + 1; q < k; ++q) t -= Hm(i, q) * y[q];
        y[i] = t / Hm(i, i);
      }
      std::memset(w.data(), 0, 8 * n);
      for (int i = 0; i < k; ++i) axpy(w.data(), y[i], V.data() + (size_t)i * n, n);
      A.precond(w.data(), z.data());
      axpy(x, 1.0, z.data(), n);
    }
    if (total < o.maxiter && res > tol) {
      beta = init_residual(false);
      res = beta;
    }
  }
  *converged = res <= tol;
  return total;
}

struct Work {
  vec V, w, z, rhs, sol, fx, fx2, dFdp;
};

// src/Newton.jl:66-114
bool newton(SH2d& P, double* x, double l, double tol, int maxit, const GmresOpts& go, Work& W, int* itn, int* itl) {
  const long N = P.N;
  if ((long)W.fx.size() < N) W.fx.resize(N);
  if ((long)W.sol.size() < N + 1) W.sol.resize(N + 1);
  P.F(x, l, W.fx.data());
  double res = norminf(W.fx.data(), N);
  int step = 0, itlin = 0;
  while (step < maxit && res > tol) {
    LinOp A{&P, x, l};
    bool cv;
    itlin += gmres(A, W.fx.data(), W.sol.data(), go, &cv, W.V, W.w, W.z);
    axpy(x, -1.0, W.sol.data(), N);
    P.F(x, l, W.fx.data());
    res = norminf(W.fx.data(), N);
    ++step;
  }
  *itn = step;
  *itl = itlin;
  return res < tol;
}

double dot_theta(const double* u1, const double* u2, double p1, double p2, double theta, long N) {
  return dot(u1, u2, N) / (double)N * theta + p1 * p2 * (1.0 - theta);
}
}  // namespace

extern "C" {

struct bkcpu_opts {
  double ds, dsmin, dsmax, p_min, p_max, a, theta, eta;
  int32_t max_steps;
  double newton_tol;
  int32_t newton_maxit;
  double gmres_reltol;
  int32_t gmres_restart, gmres_maxiter;
  double pc_shift;
  int32_t nthreads;   // 0: OpenMP default (all cores)
};

// Newton solve F(u, l) = 0 from the guess in u (in place).  Returns 1 when converged.
int32_t bkcpu_sh2d_newton(int32_t nx, int32_t ny, double lx, double ly, double l, double nu, double* u, double tol, int32_t maxit,
                          const bkcpu_opts* o, int32_t* itn, int32_t* itl) {
  if (!is_pow2(nx) || !is_pow2(ny)) return -1;
  if (o->nthreads > 0) omp_set_num_threads(o->nthreads);
  SH2d P;
  P.nx = nx;
  P.ny = ny;
  P.N = (long)nx * ny;
  P.nu = nu;
  P.L1 = square(assemble_I_plus_lap(nx, ny, lx, ly));
  P.pc.init(nx, ny, lx, ly, o->pc_shift);
  GmresOpts go{o->gmres_reltol, o->gmres_restart, o->gmres_maxiter};
  Work W;
  int a = 0, b = 0;
  const bool ok = newton(P, u, l, tol, maxit, go, W, &a, &b);
  if (itn) *itn = a;
  if (itl) *itl = b;
  return ok ? 1 : 0;
}

// PALC continuation (secant predictor, MatrixFreeBLS corrector) from the converged state u_start at l = p_start.
// rows: (max_steps + 1) x 5 doubles (param, ||u||_2, itnewton, itlinear, ds); loop_seconds: the continuation! loop only
// (the two start-up Newton solves are excluded, src/Continuation.jl:370-393); step_seconds (may be NULL): time stamp of every
// accepted step relative to the start of the loop.
int32_t bkcpu_sh2d_palc(int32_t nx, int32_t ny, double lx, double ly, double nu, const double* u_start, double p_start,
                        const bkcpu_opts* o, double* rows, int32_t* nrows, double* loop_seconds, double* step_seconds,
                        double* u_final, int32_t* work_newton, int32_t* work_linear) {
  if (!is_pow2(nx) || !is_pow2(ny)) return -1;
  if (o->nthreads > 0) omp_set_num_threads(o->nthreads);
  SH2d P;
  P.nx = nx;
  P.ny = ny;
  P.N = (long)nx * ny;
  P.nu = nu;
  P.L1 = square(assemble_I_plus_lap(nx, ny, lx, ly));
  P.pc.init(nx, ny, lx, ly, o->pc_shift);
  const long N = P.N;
  GmresOpts go{o->gmres_reltol, o->gmres_restart, o->gmres_maxiter};
  Work W;
  const double theta = o->theta, eps = std::sqrt(2.220446049250313e-16);
  vec z_u(u_start, u_start + N), zold_u(N), tau_u(N), zpred_u(N), x(N), u1(N), res_f(N), dFdp(N), rhs(N + 1), sol(N + 1), bvec(N);
  double z_p = p_start, zold_p = p_start, tau_p = 0, zpred_p = 0, ds = o->ds;
  int itn = 0, itl = 0, wn = 0, wl = 0;
  // start-up: two Newton solves (src/Continuation.jl:370-393)
  if (!newton(P, z_u.data(), z_p, o->newton_tol, o->newton_maxit, go, W, &itn, &itl)) return -2;
  u1 = z_u;
  const double p1 = z_p + o->ds / o->eta;
  if (!newton(P, u1.data(), p1, o->newton_tol, o->newton_maxit, go, W, &itn, &itl)) return -3;
  auto secant = [&](const double* a_u, double a_p, const double* b_u, double b_p) {  // tangent from b to a
#pragma omp parallel for schedule(static)
    for (long i = 0; i < N; ++i) tau_u[i] = a_u[i] - b_u[i];
    tau_p = a_p - b_p;
    const double al = (ds > 0 ? 1.0 : -1.0) / std::sqrt(dot_theta(tau_u.data(), tau_u.data(), tau_p, tau_p, theta, N));
    scal_copy(tau_u.data(), al, tau_u.data(), N);
    tau_p *= al;
  };
  secant(u1.data(), p1, z_u.data(), z_p);
  zold_u = z_u;
  auto predict = [&]() {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < N; ++i) zpred_u[i] = z_u[i] + ds * tau_u[i];
    zpred_p = z_p + ds * tau_p;
  };
  predict();
  int nr = 0, step = 0;
  auto save = [&](int a, int b) {
    double* r = rows + 5 * (size_t)nr;
    r[0] = z_p;
    r[1] = std::sqrt(dot(z_u.data(), z_u.data(), N));
    r[2] = a;
    r[3] = b;
    r[4] = ds;
    ++nr;
  };
  save(0, 0);
  const auto t0 = std::chrono::steady_clock::now();
  auto now = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
  bool stop = false, converged = true, first = true;
  while (true) {
    if (!first && converged && step <= o->max_steps && step > 0) {
      save(itn, itl);
      if (step_seconds) step_seconds[nr - 1] = now();
    }
    first = false;
    if (!(step <= o->max_steps && ((o->p_min < z_p && z_p < o->p_max) || step == 0) && !stop)) break;
    if (step == o->max_steps) break;  // the rows asked for are complete
    // ---- corrector: newton_palc (src/continuation/Palc.jl:187-305)
    bool natural = (zpred_p <= o->p_min || zpred_p >= o->p_max);
    double p = zpred_p;
    x = zpred_u;
    int nstep = 0, itlin = 0;
    if (natural) {
      p = std::min(std::max(zpred_p, o->p_min), o->p_max);
      converged = newton(P, x.data(), p, o->newton_tol, o->newton_maxit, go, W, &nstep, &itlin);
    } else {
      auto Nfun = [&](const double* u, double pp) {
        // arc_length_eq: dot_theta(u - z0, tau) - ds with the parameter part (p - z0.p) tau_p
        double s = 0;
        const double* z0 = z_u.data();
        const double* tu = tau_u.data();
#pragma omp parallel for reduction(+ : s) schedule(static)
        for (long i = 0; i < N; ++i) s += (u[i] - z0[i]) * tu[i];
        return theta * s / (double)N + (1.0 - theta) * (pp - z_p) * tau_p - ds;
      };
      P.F(x.data(), p, res_f.data());
      double res_n = Nfun(x.data(), p);
      double res = std::max(norminf(res_f.data(), N), std::fabs(res_n));
      while (nstep < o->newton_maxit && res > o->newton_tol) {
        P.F(x.data(), p + eps, dFdp.data());
#pragma omp parallel for schedule(static)
        for (long i = 0; i < N; ++i) dFdp[i] = (dFdp[i] - res_f[i]) / eps;
        // MatrixFreeBLS: (N+1) bordered map with a = dFdp, b = theta * tau_u (dot / N), c = (1 - theta) tau_p
        scal_copy(bvec.data(), theta, tau_u.data(), N);
        LinOp A{&P, x.data(), p, true, dFdp.data(), bvec.data(), (1.0 - theta) * tau_p, 1.0 / (double)N};
        std::memcpy(rhs.data(), res_f.data(), 8 * N);
        rhs[N] = res_n;
        bool cv;
        itlin += gmres(A, rhs.data(), sol.data(), go, &cv, W.V, W.w, W.z);
        axpy(x.data(), -1.0, sol.data(), N);
        p = std::min(std::max(p - sol[N], o->p_min), o->p_max);
        P.F(x.data(), p, res_f.data());
        res_n = Nfun(x.data(), p);
        res = std::max(norminf(res_f.data(), N), std::fabs(res_n));
        ++nstep;
      }
      converged = res < o->newton_tol;
    }
    itn = nstep;
    itl = itlin;
    wn += nstep;
    wl += itlin;
    if (converged) {
      zold_u.swap(z_u);
      zold_p = z_p;
      z_u = x;
      z_p = p;
      ++step;
    }
    // ---- step size control (src/continuation/Contbase.jl:77-102)
    if (!stop) {
      double dsnew;
      if (!converged) {
        if (std::fabs(ds) <= o->dsmin) { stop = true; dsnew = ds; }
        else dsnew = std::copysign(std::max(std::fabs(ds) / 2, o->dsmin), ds);
      } else {
        const double f = (double)(o->newton_maxit - itn) / o->newton_maxit;
        dsnew = ds * (1 + o->a * f * f);
      }
      if (!stop) ds = std::copysign(std::min(std::max(std::fabs(dsnew), o->dsmin), o->dsmax), dsnew);
    }
    // ---- predictor (secant, src/continuation/Tangents.jl:28-42)
    if (converged) secant(z_u.data(), z_p, zold_u.data(), zold_p);
    predict();
  }
  *loop_seconds = now();
  *nrows = nr;
  if (u_final) std::memcpy(u_final, z_u.data(), 8 * N);
  if (work_newton) *work_newton = wn;
  if (work_linear) *work_linear = wl;
  return 0;
}

// measurement aids for bench.py's CPU arm: algorithmic bytes of the BLAS-1 sweeps (dot 16 N, axpy 24 N, scaled copy 16 N) and of
// the CSR SpMVs (12 B per stored entry + row pointers + both vectors) since the last reset, and the host's own streaming
// bandwidth (STREAM triad a = b + s c over three 256 MB arrays first-touched by the same team) to set them against
void bkcpu_reset_counters() { g_blas1_bytes = g_spmv_bytes = 0; }
void bkcpu_counters(double out[2]) {
  out[0] = g_blas1_bytes;
  out[1] = g_spmv_bytes;
}
double bkcpu_triad_gbs(int32_t nthreads) {
  if (nthreads > 0) omp_set_num_threads(nthreads);
  const long n = 32L * 1024 * 1024;
  vec a((size_t)n), b((size_t)n), c((size_t)n);
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    a[i] = 0.0;
    b[i] = 1.0;
    c[i] = 2.0;
  }
  double best = 0;
  for (int rep = 0; rep < 4; ++rep) {
    const auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) a[i] = b[i] + 3.0 * c[i];
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rep > 0) best = std::max(best, 24.0 * (double)n / dt * 1e-9);
  }
  return best + 0.0 * a[n / 2];
}

// processors available to this process (affinity-aware); NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1
int32_t bkcpu_max_threads() { return omp_get_num_procs(); }

// The MGS sweep of GMRES does not scale to every hardware thread of a big host (barrier cost, SMT siblings sharing a core's
// load ports, NUMA): time one sweep over a 48-vector basis (larger than the caches, like the real solves) for a few thread
// counts and return the fastest -- "all the host threads it can use" in the sense of the fastest configuration the host offers.
int32_t bkcpu_calibrate_threads(int64_t n) {
  const int maxt = omp_get_num_procs();
  const int nvec = 48;
  int best = maxt;
  double best_t = 1e300;
  for (int t : {8, 16, 32, 48, 64, 96, 128, 192, 256}) {
    if (t > maxt) t = maxt;
    omp_set_num_threads(t);
    vec Vb, w((size_t)n);   // first-touched by THIS team, row by row like the solver's basis
    Vb.resize_rows(nvec, (size_t)n);
    for (int i = 0; i < nvec; ++i) scal_copy(Vb.data() + (size_t)i * n, 0.0, Vb.data() + (size_t)i * n, n);
    w.assign((size_t)n, 1.0);
    for (int rep = 0; rep < 2; ++rep) {
      const auto t0 = std::chrono::steady_clock::now();
      for (int i = 0; i < nvec; ++i) {
        const double h = dot(Vb.data() + (size_t)i * n, w.data(), n);
        axpy(w.data(), -1e-9 * h, Vb.data() + (size_t)i * n, n);
      }
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (rep == 1 && dt < best_t) {
        best_t = dt;
        best = t;
      }
    }
    if (t == maxt) break;
  }
  omp_set_num_threads(maxt);
  return best;
}
}
