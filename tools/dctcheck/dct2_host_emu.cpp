// Host emulation of bk_dct2.cuh (k_dct2v2): one emulated thread per CTA executes every phase to completion, which is exact
// because the items of a phase are independent and phases are separated by __syncthreads().  Checks forward / inverse / fused
// kernels against the naive O(n^2) DCT-II in both directions, with ragged line counts.
//   g++ -O2 -std=c++17 -DBK_DCT_HOST_EMU -I../../bifurcationkit.jl_b200/csrc dct2_host_emu.cpp -o dct2_host_emu && ./dct2_host_emu
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct double2 { double x, y; };
static inline double2 make_double2(double a, double b) { return double2{a, b}; }
struct Idx3 { int x = 0, y = 0, z = 0; };
static Idx3 threadIdx, blockIdx, blockDim, gridDim;
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __syncthreads() ((void)0)
#define __ldg(p) (*(p))
using std::min;
static inline unsigned __brev(unsigned v) {
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
  return r;
}
static double2 sdct2[1 << 18];
// interface structs of bk_dct.cuh
struct LineGeom { int n; long long es; int nx; long long os; int nouter; };
struct DctTables { const double2* tw; const double2* wn; const double2* dtw; };
struct SymbolArgs { const double* lam_e; const double* lam_x; const double* lam_o; double shift; const double* tail_src; double* tail_dst; int tail_n; };
#include "bk_dct2.cuh"

static const long double PI = 3.14159265358979323846264338327950288L;
struct Tables {
  std::vector<double2> tw, wn, dtw; std::vector<double> lam;
  explicit Tables(int n) {
    const int M = n / 2;
    tw.resize(M / 2 > 0 ? M / 2 : 1); wn.resize(M + 1); dtw.resize(M + 1); lam.resize(n);
    for (int k = 0; k < M / 2; ++k) tw[k] = make_double2((double)cosl(-2.0L * PI * k / M), (double)sinl(-2.0L * PI * k / M));
    for (int k = 0; k <= M; ++k) {
      wn[k] = make_double2((double)cosl(-2.0L * PI * k / n), (double)sinl(-2.0L * PI * k / n));
      dtw[k] = make_double2((double)cosl(-PI * k / (2.0L * n)), (double)sinl(-PI * k / (2.0L * n)));
    }
    for (int k = 0; k < n; ++k) lam[k] = (double)(2.0L * cosl(PI * k / n) - 2.0L) * 0.37;
  }
};
static void naive_fwd(const double* x, long long es, int n, double* C) {
  for (int k = 0; k < n; ++k) { long double a = 0; for (int e = 0; e < n; ++e) a += (long double)x[e * es] * cosl(PI * (2 * e + 1) * k / (2.0L * n)); C[k] = (double)a; }
}
static void naive_inv(const double* C, int n, double* x, long long es) {
  for (int e = 0; e < n; ++e) { long double a = C[0]; for (int k = 1; k < n; ++k) a += 2.0L * C[k] * cosl(PI * (2 * e + 1) * k / (2.0L * n)); x[e * es] = (double)(a / n); }
}
static double rel(const std::vector<double>& a, const std::vector<double>& b) {
  double m = 0, s = 0;
  for (size_t i = 0; i < a.size(); ++i) { m = std::max(m, std::fabs(a[i] - b[i])); s = std::max(s, std::fabs(b[i])); }
  return m / (s > 0 ? s : 1);
}

template <int LOGM, int LOGW, bool STRIDED, int MODE>
static void launch(const double* in, double* out, LineGeom g, DctTables tb, SymbolArgs sy) {
  const int W = 1 << LOGW;
  blockDim.x = 1; threadIdx.x = 0;
  const int gx = STRIDED ? (g.nx + W - 1) / W : (g.nouter + W - 1) / W, gy = STRIDED ? g.nouter : 1;
  for (int by = 0; by < gy; ++by)
    for (int bx = 0; bx < gx; ++bx) { blockIdx.x = bx; blockIdx.y = by; k_dct2v2<LOGM, LOGW, STRIDED, MODE>(in, out, g, tb, sy); }
}

template <int LOGM, int LOGW>
static int check(int lines) {   // `lines` = extent of the other dimension (ragged w.r.t. W on purpose)
  const int n = 2 << LOGM;
  int fails = 0;
  Tables T(n);
  DctTables tb{T.tw.data(), T.wn.data(), T.dtw.data()};
  SymbolArgs none{nullptr, nullptr, nullptr, 0.0, nullptr, nullptr, 0};
  std::vector<double> lam_other(lines);
  for (int i = 0; i < lines; ++i) lam_other[i] = -0.01 * i;
  for (int strided = 0; strided < 2; ++strided) {
    // contiguous: `lines` lines of length n, x fastest (nx = n);  strided: lines along y, `lines` columns (nx = lines)
    const int nx = strided ? lines : n, ny = strided ? n : lines;
    const long long N = (long long)nx * ny;
    LineGeom g = strided ? LineGeom{n, nx, nx, N, 1} : LineGeom{n, 1, 1, nx, ny};
    std::vector<double> x(N), a(N, -7.0), b(N, -7.0), ref(N), C(n), tmp(n);
    srand(1 + LOGM * 7 + LOGW + strided);
    for (auto& v : x) v = rand() / (double)RAND_MAX - 0.5;
    const long long es = strided ? nx : 1;
    for (int l = 0; l < lines; ++l) { const long long off = strided ? l : (long long)l * nx; naive_fwd(x.data() + off, es, n, C.data()); for (int k = 0; k < n; ++k) ref[off + k * es] = C[k]; }
    if (strided) launch<LOGM, LOGW, true, 0>(x.data(), a.data(), g, tb, none); else launch<LOGM, LOGW, false, 0>(x.data(), a.data(), g, tb, none);
    double e1 = rel(a, ref);
    if (strided) launch<LOGM, LOGW, true, 1>(a.data(), b.data(), g, tb, none); else launch<LOGM, LOGW, false, 1>(a.data(), b.data(), g, tb, none);
    double e2 = rel(b, x);
    // fused: forward, divide by (1 + lam_e[k] + lam_x[line])^2 + shift, inverse
    SymbolArgs sy{T.lam.data(), strided ? lam_other.data() : nullptr, nullptr, 0.8, nullptr, nullptr, 0};
    std::vector<double> f(N, -7.0), fr(N);
    double e3 = 0;
    if (strided) {
      launch<LOGM, LOGW, true, 2>(x.data(), f.data(), g, tb, sy);
      for (int l = 0; l < lines; ++l) {
        for (int k = 0; k < n; ++k) { double t = 1.0 + T.lam[k] + lam_other[l]; C[k] = ref[l + (long long)k * es] / (t * t + 0.8); }
        naive_inv(C.data(), n, fr.data() + l, es);
      }
      e3 = rel(f, fr);
    }
    printf("n=%5d W=%2d lines=%3d %s  fwd %.1e  roundtrip %.1e  fused %.1e\n", n, 1 << LOGW, lines, strided ? "strided   " : "contiguous", e1, e2, e3);
    if (!(e1 < 1e-12 && e2 < 1e-12 && e3 < 1e-12)) ++fails;
  }
  return fails;
}

int main() {
  int f = 0;
  f += check<2, 1>(5);    // n = 8
  f += check<3, 2>(7);    // n = 16, odd log
  f += check<4, 0>(3);    // n = 32, W = 1
  f += check<5, 3>(13);   // n = 64
  f += check<6, 2>(9);    // n = 128
  f += check<7, 4>(19);   // n = 256, W = 16
  f += check<8, 3>(11);   // n = 512, W = 8
  f += check<9, 2>(6);    // n = 1024, W = 4
  f += check<10, 1>(3);   // n = 2048, W = 2
  printf(f ? "FAILED (%d)\n" : "ALL OK\n", f);
  return f ? 1 : 0;
}
