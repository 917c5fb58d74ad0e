// bk_precond.cu -- K6: preconditioners honouring the reference's Pl/Pr contract
// (ldiv!(y, P, x), src/Preconditioner.jl:11-37).
//
// BK_PC_SH_DCT: the examples precondition Swift-Hohenberg with a sparse factorisation of L1 + I
//   (examples/SH2d-fronts.jl:120-122  Pl = lu(par.L1 + I); examples/SH3d.jl:88 cholesky(L1)).  The
//   Neumann-closure Laplacian (SH2d-fronts.jl:13-29) is diagonalised by the DCT-II, so
//   (L1 + shift I)^-1 r = IDCT( DCT(r) / ((1 + lx_i + ly_j [+ lz_k])^2 + shift) )  exactly
//   (identity pinned in tests/test_oracle_palc.py::test_dct_symbol_diagonalises_L1).
//   Each 1-D DCT of a power-of-two length is a shared-memory radix-2 complex FFT (Makhoul's
//   reordering: v[m] = x[2m], v[n-1-m] = x[2m+1]; C[k] = Re(e^{-i pi k/2n} FFT(v)[k])); other lengths
//   use a dense n x n transform.  Lines along x are contiguous; lines along y/z are processed in
//   batches of W consecutive x so every global access stays coalesced.
// BK_PC_CHAN_TRIDIAG: lu(P) of examples/chan.jl:108-111 (Thomas algorithm, one thread: n = 1e3 plumbing).
// BK_PC_CGL_DST: per-component (a0 I + a1 Lap_dirichlet)^-1 by dense DST-I (stand-in for the ILU of
//   examples/cGL2d.jl:209-213); for potrap contexts it is applied slice by slice (block Jacobi, cf.
//   jacobian_block_diag, src/periodicorbit/PeriodicOrbitTrapeze.jl:619-643).
#include <cmath>
#include <cstdlib>
#include <utility>
#include <vector>
#include <cublas_v2.h>
#include "bk_common.cuh"

#include "bk_dct.cuh"
#include "bk_dct2.cuh"

// Opt-in second version of the DCT kernels (bk_dct2.cuh; BK_DCT_V2=1).  Returns false when (n, W) has no instantiation.
template <int LOGM, int LOGW>
static bool launch_dct_v2(bk_ctx* c, int d, int mode, const double* in, double* out, const LineGeom& g, int nthr, const DctTables& tb,
                          const SymbolArgs& sy) {
  constexpr int W = 1 << LOGW;
  const size_t sm = 16 * ((size_t)Dct2Cfg<LOGM>::MP * W + Dct2Cfg<LOGM>::TWN) + (mode == 2 ? 8 * (size_t)Dct2Cfg<LOGM>::N * W : 0);
  static bool attr = false;
  if (!attr) {
    const int mx = 160 * 1024;
    cudaFuncSetAttribute(k_dct2v2<LOGM, LOGW, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    cudaFuncSetAttribute(k_dct2v2<LOGM, LOGW, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    cudaFuncSetAttribute(k_dct2v2<LOGM, LOGW, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    cudaFuncSetAttribute(k_dct2v2<LOGM, LOGW, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    cudaFuncSetAttribute(k_dct2v2<LOGM, LOGW, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    attr = true;
  }
  if (d == 0) {
    dim3 grid((g.nouter + W - 1) / W);
    if (mode == 0)
      bk_launch_pdl(k_dct2v2<LOGM, LOGW, false, 0>, grid, dim3(nthr), sm, c->stream, in, out, g, tb, sy);
    else if (mode == 1)
      bk_launch_pdl(k_dct2v2<LOGM, LOGW, false, 1>, grid, dim3(nthr), sm, c->stream, in, out, g, tb, sy);
    else
      return false;
  } else {
    dim3 grid((g.nx + W - 1) / W, g.nouter);
    if (mode == 0)
      bk_launch_pdl(k_dct2v2<LOGM, LOGW, true, 0>, grid, dim3(nthr), sm, c->stream, in, out, g, tb, sy);
    else if (mode == 1)
      bk_launch_pdl(k_dct2v2<LOGM, LOGW, true, 1>, grid, dim3(nthr), sm, c->stream, in, out, g, tb, sy);
    else
      bk_launch_pdl(k_dct2v2<LOGM, LOGW, true, 2>, grid, dim3(nthr), sm, c->stream, in, out, g, tb, sy);
  }
  return true;
}
static bool try_dct_v2(bk_ctx* c, int d, int mode, const double* in, double* out, const LineGeom& g, int W, int nthr, const DctTables& tb,
                       const SymbolArgs& sy) {
  static int on = -1;
  if (on < 0) on = getenv("BK_DCT_V2") ? 1 : 0;
  if (!on || nthr > 512) return false;
  if (g.n == 2048 && W == 2) return launch_dct_v2<10, 1>(c, d, mode, in, out, g, nthr, tb, sy);
  if (g.n == 1024 && W == 4) return launch_dct_v2<9, 2>(c, d, mode, in, out, g, nthr, tb, sy);
  if (g.n == 1024 && W == 2) return launch_dct_v2<9, 1>(c, d, mode, in, out, g, nthr, tb, sy);
  if (g.n == 512 && W == 8) return launch_dct_v2<8, 3>(c, d, mode, in, out, g, nthr, tb, sy);
  if (g.n == 512 && W == 4) return launch_dct_v2<8, 2>(c, d, mode, in, out, g, nthr, tb, sy);
  if (g.n == 256 && W == 16) return launch_dct_v2<7, 4>(c, d, mode, in, out, g, nthr, tb, sy);
  if (g.n == 256 && W == 8) return launch_dct_v2<7, 3>(c, d, mode, in, out, g, nthr, tb, sy);
  return false;
}

// dense transform of every line: out[line, k] = sum_e M[k*n + e] in[line, e]
static __global__ void __launch_bounds__(256) k_dense_lines(const double* __restrict__ in, double* __restrict__ out, LineGeom g,
                                                            const double* __restrict__ M) {
  const long long nlines = (long long)g.nx * g.nouter;
  const long long total = nlines * g.n;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    long long line = q % nlines;
    int k = (int)(q / nlines);
    long long x = line % g.nx, o = line / g.nx;
    long long base = x + o * g.os;
    const double* Mk = M + (long long)k * g.n;
    double acc = 0.0;
    for (int e = 0; e < g.n; ++e) acc = fma(__ldg(Mk + e), in[base + (long long)e * g.es], acc);
    out[base + (long long)k * g.es] = acc;
  }
}

// divide by the symbol
static __global__ void __launch_bounds__(256) k_sh_symbol_div(double* __restrict__ a, int nx, int ny, int nz,
                                                              const double* __restrict__ lx, const double* __restrict__ ly,
                                                              const double* __restrict__ lz, double shift) {
  const long long total = (long long)nx * ny * nz;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    int i = (int)(q % nx), j = (int)((q / nx) % ny), k = (int)(q / ((long long)nx * ny));
    double t = 1.0 + lx[i] + ly[j] + (lz ? lz[k] : 0.0);
    a[q] = a[q] / (t * t + shift);
  }
}
static __global__ void __launch_bounds__(256) k_helmholtz_symbol_div(double* __restrict__ a, int nx, int ny, long long nblocks,
                                                                     const double* __restrict__ lx,
                                                                     const double* __restrict__ ly, double a0, double a1) {
  const long long n = (long long)nx * ny, total = n * nblocks;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    long long g = q % n;
    int i = (int)(g % nx), j = (int)(g / nx);
    a[q] = a[q] / (a0 + a1 * (lx[i] + ly[j]));
  }
}

// ---- potrap circulant preconditioner: time direction -------------------------------------------------------------------
// B holds the DST-I coefficients of all 2M slice components (field f = 2*slice + comp, n values each).  One thread per
// spatial mode: w+- = u1 +- i u2 over the K = M-1 cyclic slices, DFT in time, divide by
//   s+-_k = (1 - g_k) - h/2 (1 + g_k)(lambda + r +- i nu),   g_k = exp(-2 pi i k/K),
// inverse DFT, back to (u1, u2).  In place.
#define BK_PO_KMAX 64
static __global__ void __launch_bounds__(128) k_potrap_time(double* __restrict__ B, long long n, int nx, int K,
                                                            const double* __restrict__ lamx, const double* __restrict__ lamy,
                                                            double h, double r, double nu, const double2* __restrict__ tw) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const double lam = lamx[g % nx] + lamy[g / nx];
  double2 wp[BK_PO_KMAX], wm[BK_PO_KMAX];
  for (int i = 0; i < K; ++i) {
    const double a = B[(long long)(2 * i) * n + g], b = B[(long long)(2 * i + 1) * n + g];
    wp[i] = make_double2(a, b);
    wm[i] = make_double2(a, -b);
  }
  double2 yp[BK_PO_KMAX], ym[BK_PO_KMAX];
  const double invK = 1.0 / K;
  for (int k = 0; k < K; ++k) {
    double2 ap = make_double2(0, 0), am = make_double2(0, 0);
    int idx = 0;
    for (int i = 0; i < K; ++i) {
      const double2 t = tw[idx];  // exp(-2 pi i k i / K)
      ap.x += wp[i].x * t.x - wp[i].y * t.y;
      ap.y += wp[i].x * t.y + wp[i].y * t.x;
      am.x += wm[i].x * t.x - wm[i].y * t.y;
      am.y += wm[i].x * t.y + wm[i].y * t.x;
      idx += k;
      if (idx >= K) idx -= K;
    }
    const double2 gk = tw[k];
    // s = (1 - g) - h/2 (1 + g) (lam + r +- i nu)
    const double2 omg = make_double2(1.0 - gk.x, -gk.y), opg = make_double2(1.0 + gk.x, gk.y);
    const double cr = lam + r;
    double2 sp = make_double2(omg.x - 0.5 * h * (opg.x * cr - opg.y * nu), omg.y - 0.5 * h * (opg.x * nu + opg.y * cr));
    double2 sm = make_double2(omg.x - 0.5 * h * (opg.x * cr + opg.y * nu), omg.y - 0.5 * h * (-opg.x * nu + opg.y * cr));
    const double dp = 1.0 / (sp.x * sp.x + sp.y * sp.y), dm = 1.0 / (sm.x * sm.x + sm.y * sm.y);
    yp[k] = make_double2((ap.x * sp.x + ap.y * sp.y) * dp * invK, (ap.y * sp.x - ap.x * sp.y) * dp * invK);
    ym[k] = make_double2((am.x * sm.x + am.y * sm.y) * dm * invK, (am.y * sm.x - am.x * sm.y) * dm * invK);
  }
  for (int i = 0; i < K; ++i) {
    double2 ap = make_double2(0, 0), am = make_double2(0, 0);
    int idx = 0;
    for (int k = 0; k < K; ++k) {
      const double2 t = tw[idx];  // conj -> exp(+2 pi i k i / K)
      ap.x += yp[k].x * t.x + yp[k].y * t.y;
      ap.y += yp[k].y * t.x - yp[k].x * t.y;
      am.x += ym[k].x * t.x + ym[k].y * t.y;
      am.y += ym[k].y * t.x - ym[k].x * t.y;
      idx += i;
      if (idx >= K) idx -= K;
    }
    B[(long long)(2 * i) * n + g] = 0.5 * (ap.x + am.x);      // Re((yp + ym)/2)
    B[(long long)(2 * i + 1) * n + g] = 0.5 * (ap.y - am.y);  // Re((yp - ym)/(2i)) = Im(yp - ym)/2
  }
}
// closure row of the preconditioner: x_M = r_M + x_1; the period entry passes through
static __global__ void __launch_bounds__(256) k_potrap_close(const double* __restrict__ in, double* __restrict__ out,
                                                             long long Ns, int M) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Ns) out[(long long)(M - 1) * Ns + i] = in[(long long)(M - 1) * Ns + i] + out[i];
  if (i == 0) out[(long long)M * Ns] = in[(long long)M * Ns];
}

// Thomas solve with precomputed factors: tri = [cprime (n) | denom_inv (n) | lower (n)]
static __global__ void k_thomas(const double* __restrict__ tri, const double* __restrict__ in, double* __restrict__ out, int n) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const double* cp = tri;
  const double* di = tri + n;
  const double* lo = tri + 2 * n;
  double prev = in[0] * di[0];
  out[0] = prev;
  for (int i = 1; i < n; ++i) {
    prev = (in[i] - lo[i] * prev) * di[i];
    out[i] = prev;
  }
  for (int i = n - 2; i >= 0; --i) out[i] -= cp[i] * out[i + 1];
}

// ------------------------------------------------------------------------------------------------ host
static bool is_pow2(long long n) { return n >= 16 && n <= 4096 && (n & (n - 1)) == 0; }
static int ilog2(long long n) {
  int l = 0;
  while ((1LL << l) < n) ++l;
  return l;
}
static inline int lin_grid(bk_ctx* c, long long n) {
  long long g = (n + 255) / 256, cap = (long long)c->nsm * 8;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

static int upload(bk_ctx* c, void** dst, const void* src, size_t bytes) {
  if (*dst) cudaFree(*dst);
  BK_CUDA(c, cudaMalloc(dst, bytes));
  BK_CUDA(c, cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
  return BK_OK;
}

// transform tables for dimension d of length n. type 0: DCT-II (Neumann), 1: DST-I (Dirichlet; dense only)
static int setup_dim(bk_ctx* c, int d, long long n, double inv_h2, int type) {
  Precond& pc = c->pc;
  std::vector<double> lam(n);
  const long double PI = 3.14159265358979323846264338327950288L;
  for (long long k = 0; k < n; ++k)
    lam[k] = (type == 0) ? (double)((2.0L * cosl(PI * k / n) - 2.0L)) * inv_h2
                         : (double)(-(2.0L - 2.0L * cosl(PI * (k + 1) / (n + 1)))) * inv_h2;
  BK_TRY(upload(c, (void**)&pc.lam[d], lam.data(), 8 * n));
  pc.pow2[d] = (type == 0 && is_pow2(n)) ? 1 : 0;
  if (pc.pow2[d]) {
    const long long M = n / 2;
    std::vector<double2> tw(M / 2), wn(M + 1), dtw(M + 1);
    for (long long k = 0; k < M / 2; ++k) tw[k] = make_double2((double)cosl(-2.0L * PI * k / M), (double)sinl(-2.0L * PI * k / M));
    for (long long k = 0; k <= M; ++k) {
      wn[k] = make_double2((double)cosl(-2.0L * PI * k / n), (double)sinl(-2.0L * PI * k / n));
      dtw[k] = make_double2((double)cosl(-PI * k / (2.0L * n)), (double)sinl(-PI * k / (2.0L * n)));
    }
    BK_TRY(upload(c, (void**)&pc.tw[d], tw.data(), 16 * (M / 2)));
    BK_TRY(upload(c, (void**)&pc.wn[d], wn.data(), 16 * (M + 1)));
    BK_TRY(upload(c, (void**)&pc.dtw[d], dtw.data(), 16 * (M + 1)));
  } else {
    // dense forward F (n x n) followed by dense inverse Finv (n x n)
    std::vector<double> M(2 * n * n);
    for (long long k = 0; k < n; ++k)
      for (long long e = 0; e < n; ++e) {
        if (type == 0) {
          long double cv = cosl(PI * (2 * e + 1) * k / (2.0L * n));
          M[k * n + e] = (double)cv;                                        // C[k] = sum_e x[e] cos(pi (2e+1) k / 2n)
          M[n * n + e * n + k] = (double)((k == 0 ? 1.0L : 2.0L) * cv / n);  // x[e] = (C0 + 2 sum_k>0 C[k] cos)/n
        } else {
          long double sv = sinl(PI * (e + 1) * (k + 1) / (n + 1.0L)) * sqrtl(2.0L / (n + 1.0L));
          M[k * n + e] = (double)sv;  // orthonormal DST-I is its own inverse
          M[n * n + e * n + k] = (double)sv;
        }
      }
    BK_TRY(upload(c, (void**)&pc.dense[d], M.data(), 8 * 2 * n * n));
  }
  return BK_OK;
}

extern "C" int32_t bk_precond_setup(bk_ctx* c, int32_t kind, double a0, double a1) {
  BK_ENTER(c);
  BK_CUDA(c, cudaStreamSynchronize(c->stream));
  Precond& pc = c->pc;
  if (kind == BK_PC_NONE) {
    pc.kind = BK_PC_NONE;
    return BK_OK;
  }
  if (!pc.work) BK_CUDA(c, cudaMalloc(&pc.work, 8 * (size_t)c->ld));
  if (!pc.work2) BK_CUDA(c, cudaMalloc(&pc.work2, 8 * (size_t)c->ld));
  if (kind == BK_PC_SH_DCT) {
    BK_CHECK(c, c->kind == BK_SH2D || c->kind == BK_SH3D, "BK_PC_SH_DCT needs a Swift-Hohenberg context");
    int nd = c->kind == BK_SH3D ? 3 : 2;
    for (int d = 0; d < nd; ++d) {
      double h = 2 * c->lengths[d] / c->dims[d];
      BK_TRY(setup_dim(c, d, c->dims[d], 1.0 / (h * h), 0));
    }
  } else if (kind == BK_PC_CGL_DST) {
    BK_CHECK(c, c->kind == BK_CGL2D || c->kind == BK_POTRAP_CGL2D, "BK_PC_CGL_DST needs a cGL context");
    for (int d = 0; d < 2; ++d) {
      double h = 2 * c->lengths[d] / c->dims[d];
      BK_TRY(setup_dim(c, d, c->dims[d], 1.0 / (h * h), 1));
    }
    if (getenv("BK_CGL_DST_GEMM") && !pc.blas) {  // opt-in: dense DST-I through cuBLAS instead of k_dense_lines (see apply)
      cublasHandle_t hnd;
      BK_CHECK(c, cublasCreate(&hnd) == CUBLAS_STATUS_SUCCESS, "cublasCreate failed");
      cublasSetStream(hnd, c->stream);
      pc.blas = (void*)hnd;
    }
  } else if (kind == BK_PC_POTRAP_CIRC) {
    BK_CHECK(c, c->kind == BK_POTRAP_CGL2D, "BK_PC_POTRAP_CIRC needs a Trapeze (potrap) context");
    const int K = (int)c->dims[2] - 1;
    BK_CHECK(c, K >= 1 && K <= BK_PO_KMAX, "BK_PC_POTRAP_CIRC supports 2 <= M <= 65 time slices");
    BK_CHECK(c, a0 > 0, "BK_PC_POTRAP_CIRC: a0 must be the period T > 0");
    for (int d = 0; d < 2; ++d) {
      double h = 2 * c->lengths[d] / c->dims[d];
      BK_TRY(setup_dim(c, d, c->dims[d], 1.0 / (h * h), 1));
    }
    std::vector<double2> tw(K);
    const long double PI = 3.14159265358979323846264338327950288L;
    for (int j = 0; j < K; ++j) tw[j] = make_double2((double)cosl(-2.0L * PI * j / K), (double)sinl(-2.0L * PI * j / K));
    BK_TRY(upload(c, (void**)&pc.tdft, tw.data(), 16 * (size_t)K));
    if (!pc.blas) {
      cublasHandle_t hnd;
      BK_CHECK(c, cublasCreate(&hnd) == CUBLAS_STATUS_SUCCESS, "cublasCreate failed");
      cublasSetStream(hnd, c->stream);
      pc.blas = (void*)hnd;
    }
    pc.po_T = a0;
    pc.po_r = c->par[0];   // (r, mu, nu, c3, c5)
    pc.po_nu = c->par[2];
  } else if (kind == BK_PC_CHAN_TRIDIAG) {
    BK_CHECK(c, c->kind == BK_CHAN, "BK_PC_CHAN_TRIDIAG needs a chan context");
    long long n = c->N;
    double s = (double)(n - 1) * (double)(n - 1);
    std::vector<double> lo(n, s), di(n, -2 * s), up(n, s), tri(3 * n);
    di[0] = 1;
    up[0] = 0;
    lo[n - 1] = 0;
    di[n - 1] = 1;  // P[1,1:2] = [1,0]; P[end,end-1:end] = [0,1]  (chan.jl:109)
    lo[0] = 0;
    up[n - 1] = 0;
    // forward elimination factors
    double denom = di[0];
    tri[n + 0] = 1.0 / denom;
    tri[0] = up[0] / denom;
    for (long long i = 1; i < n; ++i) {
      denom = di[i] - lo[i] * tri[i - 1];
      tri[n + i] = 1.0 / denom;
      tri[i] = up[i] / denom;
      tri[2 * n + i] = lo[i];
    }
    BK_TRY(upload(c, (void**)&pc.tri, tri.data(), 8 * 3 * n));
  } else {
    return bk_fail(c, BK_ERR_ARG, "unknown preconditioner kind", __FILE__, __LINE__);
  }
  pc.kind = kind;
  pc.a0 = a0;
  pc.a1 = a1;
  return BK_OK;
}

// one 1-D transform pass along dimension d over `nblocks` consecutive blocks of nx*ny(*nz) values
static int transform_pass(bk_ctx* c, int d, int dir, const double* in, double* out, int nx, int ny, int nz,
                          const SymbolArgs* fused_sym = nullptr) {
  Precond& pc = c->pc;
  LineGeom g;
  const int dims[3] = {nx, ny, nz};
  g.n = dims[d];
  if (d == 0) {
    g.es = 1;
    g.nx = 1;
    g.os = nx;
    g.nouter = ny * nz;
  } else if (d == 1) {
    g.es = nx;
    g.nx = nx;
    g.os = (long long)nx * ny;
    g.nouter = nz;
  } else {
    g.es = (long long)nx * ny;
    g.nx = nx;
    g.os = nx;
    g.nouter = ny;
  }
  if (pc.pow2[d]) {
    static int env_w = -1, env_t = -1;
    if (env_w < 0) {
      const char* a = getenv("BK_DCT_W");
      const char* b = getenv("BK_DCT_THREADS");
      env_w = a ? atoi(a) : 0;
      env_t = b ? atoi(b) : 0;
    }
    int W = 4096 / g.n;
    if (W < 1) W = 1;
    if (W > 16) W = 16;
    if (env_w > 0) W = env_w;
    while (W & (W - 1)) W &= W - 1;  // power of two (shift/mask indexing in the kernels)
    const int nthr = env_t > 0 ? env_t : 512;
    const int mode = fused_sym ? 2 : (dir > 0 ? 0 : 1);
    size_t sm = sizeof(double2) * (size_t)DCT_PADDED(g.n / 2) * W + (mode == 2 ? sizeof(double) * (size_t)g.n * W : 0);
    int logM = ilog2(g.n / 2);
    DctTables tb{pc.tw[d], pc.wn[d], pc.dtw[d]};
    SymbolArgs sy{nullptr, nullptr, nullptr, 0.0, nullptr, nullptr, 0};
    if (fused_sym) sy = *fused_sym;
    if (try_dct_v2(c, d, mode, in, out, g, W, nthr, tb, sy)) {
      c->stats.kernel_launches++;
      BK_CUDA(c, cudaGetLastError());
      return BK_OK;
    }
    static bool attr = false;
    if (!attr) {
      const int mx = 160 * 1024;
      cudaFuncSetAttribute(k_dct2<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
      cudaFuncSetAttribute(k_dct2<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
      cudaFuncSetAttribute(k_dct2<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
      cudaFuncSetAttribute(k_dct2<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
      cudaFuncSetAttribute(k_dct2<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
      attr = true;
    }
    if (d == 0) {
      int grid = (g.nouter + W - 1) / W;
      if (mode == 0)
        bk_launch_pdl(k_dct2<false, 0>, dim3(grid), dim3(nthr), sm, c->stream, in, out, g, logM, W, ilog2(W), tb, sy);
      else
        bk_launch_pdl(k_dct2<false, 1>, dim3(grid), dim3(nthr), sm, c->stream, in, out, g, logM, W, ilog2(W), tb, sy);
    } else {
      dim3 grid((g.nx + W - 1) / W, g.nouter);
      if (mode == 0)
        bk_launch_pdl(k_dct2<true, 0>, dim3(grid), dim3(nthr), sm, c->stream, in, out, g, logM, W, ilog2(W), tb, sy);
      else if (mode == 1)
        bk_launch_pdl(k_dct2<true, 1>, dim3(grid), dim3(nthr), sm, c->stream, in, out, g, logM, W, ilog2(W), tb, sy);
      else
        bk_launch_pdl(k_dct2<true, 2>, dim3(grid), dim3(nthr), sm, c->stream, in, out, g, logM, W, ilog2(W), tb, sy);
    }
  } else {
    const double* M = pc.dense[d] + (dir > 0 ? 0 : (size_t)g.n * g.n);
    LineGeom gg = g;
    if (d == 0) {  // express x-lines in the (x, outer) form used by the dense kernel: x index is the line element
      gg.nx = 1;
      gg.os = nx;
      gg.nouter = ny * nz;
    }
    long long total = (long long)nx * ny * nz;
    k_dense_lines<<<lin_grid(c, total), 256, 0, c->stream>>>(in, out, gg, M);
  }
  c->stats.kernel_launches++;
  BK_CUDA(c, cudaGetLastError());
  return BK_OK;
}

int bk_precond_apply_dev(bk_ctx* c, const double* in, double* out, long long n) {
  Precond& pc = c->pc;
  BK_CHECK(c, pc.kind != BK_PC_NONE, "no preconditioner set up (bk_precond_setup)");
  BK_CHECK(c, in != out, "preconditioner: in-place application is not supported");
  const long long N = c->N;
  bool tail_done = false;
  if (pc.kind == BK_PC_SH_DCT) {
    const int nx = (int)c->dims[0], ny = (int)c->dims[1], nz = c->kind == BK_SH3D ? (int)c->dims[2] : 1;
    const int nd = c->kind == BK_SH3D ? 3 : 2;
    double* A = pc.work;
    double* B = pc.work2;
    const int last = nd - 1;
    if (pc.pow2[last]) {
      // x fwd, [y fwd,] (last dim: fwd + symbol + inverse in one kernel), [y inv,] x inv
      SymbolArgs sy{pc.lam[last], pc.lam[0], nd == 3 ? pc.lam[1] : nullptr, pc.a0, nullptr, nullptr, 0};
      if (n > N && n - N <= 32) {  // border entries ride along with the fused kernel
        sy.tail_src = in + N;
        sy.tail_dst = out + N;
        sy.tail_n = (int)(n - N);
        tail_done = true;
      }
      BK_TRY(transform_pass(c, 0, +1, in, A, nx, ny, nz));
      if (nd == 3) {
        BK_TRY(transform_pass(c, 1, +1, A, B, nx, ny, nz));
        BK_TRY(transform_pass(c, 2, +1, B, A, nx, ny, nz, &sy));
        BK_TRY(transform_pass(c, 1, -1, A, B, nx, ny, nz));
        BK_TRY(transform_pass(c, 0, -1, B, out, nx, ny, nz));
      } else {
        BK_TRY(transform_pass(c, 1, +1, A, B, nx, ny, nz, &sy));
        BK_TRY(transform_pass(c, 0, -1, B, out, nx, ny, nz));
      }
    } else {
      BK_TRY(transform_pass(c, 0, +1, in, A, nx, ny, nz));
      BK_TRY(transform_pass(c, 1, +1, A, B, nx, ny, nz));
      double* cur = B;
      double* oth = A;
      if (nd == 3) {
        BK_TRY(transform_pass(c, 2, +1, B, A, nx, ny, nz));
        cur = A;
        oth = B;
      }
      k_sh_symbol_div<<<lin_grid(c, N), 256, 0, c->stream>>>(cur, nx, ny, nz, pc.lam[0], pc.lam[1],
                                                            nd == 3 ? pc.lam[2] : nullptr, pc.a0);
      c->stats.kernel_launches++;
      BK_CUDA(c, cudaGetLastError());
      if (nd == 3) {
        BK_TRY(transform_pass(c, 2, -1, cur, oth, nx, ny, nz));
        std::swap(cur, oth);
      }
      BK_TRY(transform_pass(c, 1, -1, cur, oth, nx, ny, nz));
      BK_TRY(transform_pass(c, 0, -1, oth, out, nx, ny, nz));
    }
  } else if (pc.kind == BK_PC_CGL_DST) {
    const int nx = (int)c->dims[0], ny = (int)c->dims[1];
    const long long nblk = (c->kind == BK_POTRAP_CGL2D) ? 2 * c->dims[2] : 2;  // components x slices
    double* A = pc.work;
    double* B = pc.work2;
    if (pc.blas && getenv("BK_CGL_DST_GEMM")) {
      // Opt-in (not yet run on a GPU): the four dense DST-I passes as GEMMs, as BK_PC_POTRAP_CIRC does.  k_dense_lines costs
      // 2.4 ms per apply at 512^2 x 2 fields (Floquet: 21 s for 48 monodromy applications); the GEMMs should be ~0.15 ms.
      cublasHandle_t hnd = (cublasHandle_t)pc.blas;
      const double one = 1.0, zero = 0.0;
      const long long nn = (long long)nx * ny;
      const int nf = (int)nblk;
      const double* Sx = pc.dense[0];
      const double* Sy = pc.dense[1];
      BK_CHECK(c, cublasDgemm(hnd, CUBLAS_OP_N, CUBLAS_OP_N, nx, ny * nf, nx, &one, Sx, nx, in, nx, &zero, A, nx) ==
                      CUBLAS_STATUS_SUCCESS, "cublasDgemm failed");
      BK_CHECK(c, cublasDgemmStridedBatched(hnd, CUBLAS_OP_N, CUBLAS_OP_N, nx, ny, ny, &one, A, nx, nn, Sy, ny, 0, &zero, B, nx, nn,
                                            nf) == CUBLAS_STATUS_SUCCESS, "cublasDgemmStridedBatched failed");
      k_helmholtz_symbol_div<<<lin_grid(c, nn * nblk), 256, 0, c->stream>>>(B, nx, ny, nblk, pc.lam[0], pc.lam[1], pc.a0, pc.a1);
      BK_CUDA(c, cudaGetLastError());
      BK_CHECK(c, cublasDgemmStridedBatched(hnd, CUBLAS_OP_N, CUBLAS_OP_N, nx, ny, ny, &one, B, nx, nn, Sy, ny, 0, &zero, A, nx, nn,
                                            nf) == CUBLAS_STATUS_SUCCESS, "cublasDgemmStridedBatched failed");
      BK_CHECK(c, cublasDgemm(hnd, CUBLAS_OP_N, CUBLAS_OP_N, nx, ny * nf, nx, &one, Sx, nx, A, nx, &zero, out, nx) ==
                      CUBLAS_STATUS_SUCCESS, "cublasDgemm failed");
      c->stats.kernel_launches += 5;
      if (c->kind == BK_POTRAP_CGL2D) BK_CUDA(c, cudaMemcpyAsync(out + N - 1, in + N - 1, 8, cudaMemcpyDeviceToDevice, c->stream));
    } else {
    BK_TRY(transform_pass(c, 0, +1, in, A, nx, ny, (int)nblk));
    BK_TRY(transform_pass(c, 1, +1, A, B, nx, ny, (int)nblk));
    k_helmholtz_symbol_div<<<lin_grid(c, (long long)nx * ny * nblk), 256, 0, c->stream>>>(B, nx, ny, nblk, pc.lam[0], pc.lam[1],
                                                                                         pc.a0, pc.a1);
    c->stats.kernel_launches++;
    BK_CUDA(c, cudaGetLastError());
    BK_TRY(transform_pass(c, 1, -1, B, A, nx, ny, (int)nblk));
    BK_TRY(transform_pass(c, 0, -1, A, out, nx, ny, (int)nblk));
    if (c->kind == BK_POTRAP_CGL2D) BK_CUDA(c, cudaMemcpyAsync(out + N - 1, in + N - 1, 8, cudaMemcpyDeviceToDevice, c->stream));
    }
  } else if (pc.kind == BK_PC_POTRAP_CIRC) {
    const int nx = (int)c->dims[0], ny = (int)c->dims[1], M = (int)c->dims[2];
    const long long nn = (long long)nx * ny, Ns = 2 * nn;
    const int nf = 2 * M;
    cublasHandle_t hnd = (cublasHandle_t)pc.blas;
    const double one = 1.0, zero = 0.0;
    double* A = pc.work;
    double* Bf = pc.work2;
    const double* Sx = pc.dense[0];
    const double* Sy = pc.dense[1];
    // DST-I in space as dense GEMMs (the DST-I of 512 points needs a 1026-point FFT; fp64 GEMM is ~2 ms at 512^2 x 60):
    // column-major view of a field = nx x ny;  A = Sx * in (all fields at once),  B_f = A_f * Sy (batched)
    BK_CHECK(c, cublasDgemm(hnd, CUBLAS_OP_N, CUBLAS_OP_N, nx, ny * nf, nx, &one, Sx, nx, in, nx, &zero, A, nx) ==
                    CUBLAS_STATUS_SUCCESS, "cublasDgemm failed");
    BK_CHECK(c, cublasDgemmStridedBatched(hnd, CUBLAS_OP_N, CUBLAS_OP_N, nx, ny, ny, &one, A, nx, nn, Sy, ny, 0, &zero, Bf, nx, nn,
                                          nf) == CUBLAS_STATUS_SUCCESS, "cublasDgemmStridedBatched failed");
    k_potrap_time<<<(unsigned)((nn + 127) / 128), 128, 0, c->stream>>>(Bf, nn, nx, M - 1, pc.lam[0], pc.lam[1], pc.po_T / M,
                                                                    pc.po_r, pc.po_nu, pc.tdft);
    BK_CUDA(c, cudaGetLastError());
    BK_CHECK(c, cublasDgemmStridedBatched(hnd, CUBLAS_OP_N, CUBLAS_OP_N, nx, ny, ny, &one, Bf, nx, nn, Sy, ny, 0, &zero, A, nx, nn,
                                          nf) == CUBLAS_STATUS_SUCCESS, "cublasDgemmStridedBatched failed");
    BK_CHECK(c, cublasDgemm(hnd, CUBLAS_OP_N, CUBLAS_OP_N, nx, ny * nf, nx, &one, Sx, nx, A, nx, &zero, out, nx) ==
                    CUBLAS_STATUS_SUCCESS, "cublasDgemm failed");
    k_potrap_close<<<(unsigned)((Ns + 255) / 256), 256, 0, c->stream>>>(in, out, Ns, M);
    c->stats.kernel_launches += 6;
    BK_CUDA(c, cudaGetLastError());
  } else if (pc.kind == BK_PC_CHAN_TRIDIAG) {
    k_thomas<<<1, 32, 0, c->stream>>>(pc.tri, in, out, (int)N);
    c->stats.kernel_launches++;
    BK_CUDA(c, cudaGetLastError());
  }
  if (n > N && !tail_done)
    BK_CUDA(c, cudaMemcpyAsync(out + N, in + N, 8 * (size_t)(n - N), cudaMemcpyDeviceToDevice, c->stream));
  return BK_OK;
}

extern "C" int32_t bk_precond_apply(bk_ctx* c, const double* in, double* out) {
  BK_ENTER(c);
  double *din, *dout;
  BK_TRY(bk_stage_in(c, in, c->N, 10, true, &din));
  BK_TRY(bk_stage_in(c, out, c->N, 11, false, &dout));
  BK_TRY(bk_precond_apply_dev(c, din, dout, c->N));
  return bk_stage_out(c, out, c->N, dout);
}
