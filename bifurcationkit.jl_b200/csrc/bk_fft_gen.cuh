// bk_fft_gen.cuh -- K6, general line lengths: DCT-II (Neumann) and DST-I (Dirichlet) of ANY length n through a
// mixed-radix Stockham FFT in shared memory with a run-time radix list (prime factors of the FFT length; a prime factor p
// costs p complex multiply-adds per point, so a prime n degenerates gracefully into the dense transform).
//
// Replaces the O(n^2)-per-line dense kernel and the cuBLAS DGEMMs of round 1 for every size the register-resident
// power-of-two kernels (bk_fft_fast.cuh) do not cover: the example grid 151 x 100 (examples/SH2d-fronts.jl:8-9), 96 x 64,
// 48 x 32, and the Dirichlet grids of examples/cGL2d.jl (DST-I of length n <-> odd extension of length 2 (n + 1):
// n = 512 gives 1026 = 2 * 3^3 * 19).
//
// As in the fast path two real lines travel as one complex line z = x1 + i x2; every transform used here is REAL-linear,
// so T(z) = T(x1) + i T(x2) and no pair splitting is needed when the line is extended to a full symmetric sequence:
//   DCT-II : y[j] = z[j], y[2n-1-j] = z[j]            (L = 2n)      Y = FFT_L(y),  2 C[k] = exp(-i pi k / 2n) Y[k]
//   inverse: Y[k] = exp(+i pi k / 2n) C[k] (k < n), Y[n] = 0, Y[2n-k] = exp(-i pi k / 2n) C[k]     n x[j] = IFFT_L(Y)[j]
//   DST-I  : y[0] = 0, y[j+1] = z[j], y[n+1] = 0, y[L-1-j] = -z[j]   (L = 2n + 2)   S[k] = sqrt(2/(n+1)) (i/2) Y[k+1]
// (conventions: forward returns 2 C with C[k] = sum_j x[j] cos(pi (2j+1) k / 2n); the inverse returns n x with
//  x = (C0 + 2 sum_k C[k] cos) / n; the DST-I is orthonormal and its own inverse -- the same conventions as bk_fft_fast.cuh.)
// Stockham pass of radix R after Ns = product of the earlier radices: output (j, q), j < L/R, k = j mod Ns:
//   out[(j - k) R + k + q Ns] = sum_r in[j + r L/R] W_{Ns R}^{r (k + q Ns)}
// One thread computes one output ("output task"): R multiply-adds with twiddles from a table of W_L^t.
#pragma once
#include "bk_common.cuh"
#include "bk_fft_fast.cuh"  // Geom, complex helpers

namespace bkg {

#ifdef __CUDACC__
// STRIDED: lines along a strided dimension, pairs = neighbouring columns; else contiguous lines, pairs = neighbouring lines.
// MODE 0: DCT-II forward (2 C), 1: DCT-II inverse (n x), 2: DST-I (orthonormal).  PPG pairs per CTA.
template <bool STRIDED, int MODE>
static __global__ void __launch_bounds__(512) k_gen(const double* __restrict__ in, double* __restrict__ out, bkf::Geom g, Plan pl,
                                                    int PPG) {
  bk_pdl_sync();
  extern __shared__ __align__(16) double2 sm_gen[];
  const int L = pl.L, n = pl.n;
  double2* A = sm_gen;
  double2* B = sm_gen + (size_t)L * PPG;
  const int tid = threadIdx.x, nth = blockDim.x;
  // element (pair pr, index j) lives at buf[j * PPG + pr]: pairs fastest, so neighbouring threads of the strided kernels touch
  // neighbouring columns and the shared-memory accesses of a pass are contiguous
  const long long b0 = (long long)blockIdx.x * PPG;  // first pair of this CTA
  const long long obase = STRIDED ? (long long)blockIdx.y * g.os : 0;
  auto gaddr = [&](int pr, int which, int e) -> long long {
    // which = 0/1: first / second line of the pair; returns -1 when the line does not exist
    const long long idx = 2 * (b0 + pr) + which;
    if (idx >= g.nb) return -1;
    return STRIDED ? obase + idx + (long long)e * g.es : idx * g.os + e;
  };
  // ---- load + extend
  for (int t = tid; t < L * PPG; t += nth) {
    const int pr = STRIDED ? (t % PPG) : (t / L), j = STRIDED ? (t / PPG) : (t % L);
    double2 v = make_double2(0.0, 0.0);
    if (MODE == 0) {
      const int e = j < n ? j : 2 * n - 1 - j;
      const long long a0 = gaddr(pr, 0, e), a1 = gaddr(pr, 1, e);
      v = make_double2(a0 >= 0 ? __ldg(in + a0) : 0.0, a1 >= 0 ? __ldg(in + a1) : 0.0);
    } else if (MODE == 1) {
      // Y[k] = ph[k]^* C[k] (k < n), 0 (k = n), ph[2n-k] C[2n-k] (k > n);   ph[k] = exp(-i pi k / 2n)
      if (j != n) {
        const int k = j < n ? j : 2 * n - j;
        const long long a0 = gaddr(pr, 0, k), a1 = gaddr(pr, 1, k);
        const double2 c = make_double2(a0 >= 0 ? __ldg(in + a0) : 0.0, a1 >= 0 ? __ldg(in + a1) : 0.0);
        const double2 p = __ldg(pl.ph + k);
        v = j < n ? bkf::cmulc(c, p) : bkf::cmul(c, p);
      }
    } else {
      if (j != 0 && j != n + 1) {
        const int e = j <= n ? j - 1 : L - 1 - j;
        const long long a0 = gaddr(pr, 0, e), a1 = gaddr(pr, 1, e);
        v = make_double2(a0 >= 0 ? __ldg(in + a0) : 0.0, a1 >= 0 ? __ldg(in + a1) : 0.0);
        if (j > n) v = make_double2(-v.x, -v.y);
      }
    }
    A[(size_t)j * PPG + pr] = v;
  }
  __syncthreads();
  // ---- Stockham passes (forward sign for MODE 0 / 2, inverse sign for MODE 1)
  int Ns = 1;
  for (int p = 0; p < pl.npass; ++p) {
    const int R = pl.radix[p], LR = L / R, NR = Ns * R, tstride = L / NR;
    for (int t = tid; t < L * PPG; t += nth) {
      const int pr = t % PPG, o = t / PPG;       // output task o = j * R + q in (j, q) order is not needed: enumerate outputs
      const int q = o / LR, j = o - q * LR;      // q slowest: neighbouring threads share q and walk j
      const int k = j % Ns;
      const int c = k + q * Ns;                  // twiddle exponent step, < NR
      double2 acc = A[(size_t)j * PPG + pr];
      int idx = 0;
      for (int r = 1; r < R; ++r) {
        idx += c;
        if (idx >= NR) idx -= NR;
        double2 w = __ldg(pl.wl + (size_t)idx * tstride);
        if (MODE == 1) w.y = -w.y;
        const double2 x = A[(size_t)(j + r * LR) * PPG + pr];
        acc.x = fma(x.x, w.x, fma(-x.y, w.y, acc.x));
        acc.y = fma(x.x, w.y, fma(x.y, w.x, acc.y));
      }
      B[(size_t)((j - k) * R + k + q * Ns) * PPG + pr] = acc;
    }
    __syncthreads();
    double2* tmp = A;
    A = B;
    B = tmp;
    Ns = NR;
  }
  // ---- post-process + store (A holds the spectrum / the signal in natural order)
  for (int t = tid; t < n * PPG; t += nth) {
    const int pr = STRIDED ? (t % PPG) : (t / n), k = STRIDED ? (t / PPG) : (t % n);
    double2 r;
    if (MODE == 0) {
      r = bkf::cmul(A[(size_t)k * PPG + pr], __ldg(pl.ph + k));
    } else if (MODE == 1) {
      r = A[(size_t)k * PPG + pr];
    } else {
      const double2 y = A[(size_t)(k + 1) * PPG + pr];  // (i/2) Y = (-Y.y, Y.x) / 2
      r = make_double2(-y.y * pl.dst_scale, y.x * pl.dst_scale);
    }
    const long long a0 = gaddr(pr, 0, k), a1 = gaddr(pr, 1, k);
    if (a0 >= 0) out[a0] = r.x;
    if (a1 >= 0) out[a1] = r.y;
  }
}
#endif
}  // namespace bkg
