// Stand-alone check + timing of the shared-memory DCT kernels (bk_dct.cuh): the preconditioner's three launches are 40 % of a
// PALC step, so this is the loop to iterate on:  edit bk_dct.cuh, rebuild this file, run (< 5 s).
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -I../../bifurcationkit.jl_b200/csrc dct_check.cu -o dct_check
//   ./dct_check [n=1024] [W=4096/n] [threads=512] [reps=50]
// Conventions (same as the dense fallback in bk_precond.cu::setup_dim): forward C[k] = sum_e x[e] cos(pi (2e+1) k / 2n),
// inverse x[e] = (C[0] + 2 sum_{k>0} C[k] cos(pi (2e+1) k / 2n)) / n.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <type_traits>
#include "bk_dct.cuh"
#include "bk_dct2.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

static int ilog2i(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

__global__ void k_fill(double* p, long long n, unsigned long long seed) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned long long z = (i + 1) * 0x9E3779B97F4A7C15ull + seed;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31;
    p[i] = (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
  }
}
// naive forward DCT of line `line` (stride es, start base) -> out[k]
__global__ void k_naive_fwd(const double* in, long long base, long long es, int n, double* out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  double acc = 0;
  for (int e = 0; e < n; ++e) acc += in[base + e * es] * cospi((2.0 * e + 1.0) * k / (2.0 * n));
  out[k] = acc;
}
__global__ void k_symbol(double* a, int nx, int ny, const double* lx, const double* ly, double shift) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)nx * ny; i += (long long)gridDim.x * blockDim.x) {
    double t = 1.0 + lx[i % nx] + ly[i / nx];
    a[i] /= (t * t + shift);
  }
}

struct Tab { double2 *tw, *wn, *dtw; double* lam; };
static Tab make_tables(int n, double inv_h2) {
  const long double PI = 3.14159265358979323846264338327950288L;
  const int M = n / 2;
  std::vector<double2> tw(M / 2), wn(M + 1), dtw(M + 1);
  std::vector<double> lam(n);
  for (int k = 0; k < n; ++k) lam[k] = (double)(2.0L * cosl(PI * k / n) - 2.0L) * inv_h2;
  for (int k = 0; k < M / 2; ++k) tw[k] = make_double2((double)cosl(-2.0L * PI * k / M), (double)sinl(-2.0L * PI * k / M));
  for (int k = 0; k <= M; ++k) {
    wn[k] = make_double2((double)cosl(-2.0L * PI * k / n), (double)sinl(-2.0L * PI * k / n));
    dtw[k] = make_double2((double)cosl(-PI * k / (2.0L * n)), (double)sinl(-PI * k / (2.0L * n)));
  }
  Tab t;
  CK(cudaMalloc(&t.tw, 16 * (M / 2))); CK(cudaMalloc(&t.wn, 16 * (M + 1))); CK(cudaMalloc(&t.dtw, 16 * (M + 1))); CK(cudaMalloc(&t.lam, 8 * n));
  CK(cudaMemcpy(t.tw, tw.data(), 16 * (M / 2), cudaMemcpyHostToDevice)); CK(cudaMemcpy(t.wn, wn.data(), 16 * (M + 1), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(t.dtw, dtw.data(), 16 * (M + 1), cudaMemcpyHostToDevice)); CK(cudaMemcpy(t.lam, lam.data(), 8 * n, cudaMemcpyHostToDevice));
  return t;
}

static double maxdiff(const double* a, const double* b, long long n) {
  std::vector<double> ha(n), hb(n);
  CK(cudaMemcpy(ha.data(), a, 8 * n, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hb.data(), b, 8 * n, cudaMemcpyDeviceToHost));
  double m = 0, s = 0;
  for (long long i = 0; i < n; ++i) { m = fmax(m, fabs(ha[i] - hb[i])); s = fmax(s, fabs(hb[i])); }
  return m / (s > 0 ? s : 1);
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 1024;
  int W = argc > 2 ? atoi(argv[2]) : 4096 / n;
  const int nthr = argc > 3 ? atoi(argv[3]) : 512, reps = argc > 4 ? atoi(argv[4]) : 50;
  if (W < 1) W = 1;
  const int nx = n, ny = n, logM = ilog2i(n / 2), logW = ilog2i(W);
  const long long N = (long long)nx * ny;
  double *x, *a, *b, *c2, *ref;
  CK(cudaMalloc(&x, 8 * N)); CK(cudaMalloc(&a, 8 * N)); CK(cudaMalloc(&b, 8 * N)); CK(cudaMalloc(&c2, 8 * N)); CK(cudaMalloc(&ref, 8 * n));
  k_fill<<<1024, 256>>>(x, N, 7);
  Tab t = make_tables(n, 1.0);
  DctTables tb{t.tw, t.wn, t.dtw};
  SymbolArgs none{nullptr, nullptr, nullptr, 0.0, nullptr, nullptr, 0};
  LineGeom gx{n, 1, 1, nx, ny}, gy{n, nx, nx, N, 1};
  const size_t sm = sizeof(double2) * (size_t)DCT_PADDED(n / 2) * W, sm2 = sm + sizeof(double) * (size_t)n * W;
  const int mx = 160 * 1024;
  CK(cudaFuncSetAttribute(k_dct2<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
  CK(cudaFuncSetAttribute(k_dct2<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
  CK(cudaFuncSetAttribute(k_dct2<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
  CK(cudaFuncSetAttribute(k_dct2<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
  CK(cudaFuncSetAttribute(k_dct2<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
  const dim3 grid_x((ny + W - 1) / W), grid_y((nx + W - 1) / W, 1);
  printf("n=%d W=%d threads=%d smem=%zu (fused %zu) grid_x=%d grid_y=%d\n", n, W, nthr, sm, sm2, grid_x.x, grid_y.x);
  auto fx = [&](const double* in, double* out) { return bk_launch_pdl(k_dct2<false, 0>, grid_x, dim3(nthr), sm, 0, in, out, gx, logM, W, logW, tb, none); };
  auto ix = [&](const double* in, double* out) { return bk_launch_pdl(k_dct2<false, 1>, grid_x, dim3(nthr), sm, 0, in, out, gx, logM, W, logW, tb, none); };
  auto fy = [&](const double* in, double* out) { return bk_launch_pdl(k_dct2<true, 0>, grid_y, dim3(nthr), sm, 0, in, out, gy, logM, W, logW, tb, none); };
  auto iy = [&](const double* in, double* out) { return bk_launch_pdl(k_dct2<true, 1>, grid_y, dim3(nthr), sm, 0, in, out, gy, logM, W, logW, tb, none); };
  SymbolArgs sy{t.lam, t.lam, nullptr, 1.0, nullptr, nullptr, 0};
  auto fused_y = [&](const double* in, double* out) { return bk_launch_pdl(k_dct2<true, 2>, grid_y, dim3(nthr), sm2, 0, in, out, gy, logM, W, logW, tb, sy); };
  // ---- correctness
  CK(fx(x, a)); CK(cudaDeviceSynchronize());
  for (int line : {0, 1, ny / 2, ny - 1}) {
    k_naive_fwd<<<(n + 127) / 128, 128>>>(x, (long long)line * nx, 1, n, ref);
    printf("  x-forward line %4d  rel err %.2e\n", line, maxdiff(a + (long long)line * nx, ref, n));
  }
  CK(ix(a, b)); printf("  x round trip        rel err %.2e\n", maxdiff(b, x, N));
  CK(fy(x, a)); CK(iy(a, b)); printf("  y round trip        rel err %.2e\n", maxdiff(b, x, N));
  // full preconditioner: fx -> fused_y  -> ix   vs   fx -> fy -> symbol -> iy -> ix
  CK(fx(x, a)); CK(fused_y(a, b)); CK(ix(b, c2));
  CK(fx(x, a)); CK(fy(a, b)); k_symbol<<<1024, 256>>>(b, nx, ny, t.lam, t.lam, 1.0); CK(iy(b, a)); CK(ix(a, b));
  printf("  fused y-pass vs unfused chain rel err %.2e\n", maxdiff(c2, b, N));
  // ---- timing
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char* name, auto&& f, double bytes) {
    for (int r = 0; r < 3; ++r) f();
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) f();
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    printf("  %-34s %7.1f us  %6.0f GB/s\n", name, 1e3 * ms / reps, bytes / (1e6 * ms / reps));
  };
  timeit("x forward  k_dct2<0,0>", [&] { fx(x, a); }, 16.0 * N);
  timeit("x inverse  k_dct2<0,1>", [&] { ix(a, b); }, 16.0 * N);
  timeit("y forward  k_dct2<1,0>", [&] { fy(x, a); }, 16.0 * N);
  timeit("y fused    k_dct2<1,2>", [&] { fused_y(a, b); }, 16.0 * N);
  timeit("preconditioner (3 launches)", [&] { fx(x, a); fused_y(a, b); ix(b, c2); }, 48.0 * N);
  // ---- version 2 (bk_dct2.cuh), n = 1024 / W = 4 and n = 512 / W = 8 instantiated here
  auto v2 = [&](auto LM, auto LW) {
    constexpr int LOGM = decltype(LM)::value, LOGW = decltype(LW)::value;
    if ((2 << LOGM) != n || (1 << LOGW) != W) return;
    const size_t s1 = 16 * ((size_t)Dct2Cfg<LOGM>::MP * W + Dct2Cfg<LOGM>::TWN), s2 = s1 + 8 * (size_t)n * W;
    CK(cudaFuncSetAttribute(k_dct2v2<LOGM, LOGW, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    CK(cudaFuncSetAttribute(k_dct2v2<LOGM, LOGW, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    CK(cudaFuncSetAttribute(k_dct2v2<LOGM, LOGW, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    CK(cudaFuncSetAttribute(k_dct2v2<LOGM, LOGW, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    CK(cudaFuncSetAttribute(k_dct2v2<LOGM, LOGW, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    auto fx2 = [&](const double* in, double* out) { return bk_launch_pdl(k_dct2v2<LOGM, LOGW, false, 0>, grid_x, dim3(nthr), s1, 0, in, out, gx, tb, none); };
    auto ix2 = [&](const double* in, double* out) { return bk_launch_pdl(k_dct2v2<LOGM, LOGW, false, 1>, grid_x, dim3(nthr), s1, 0, in, out, gx, tb, none); };
    auto fy2 = [&](const double* in, double* out) { return bk_launch_pdl(k_dct2v2<LOGM, LOGW, true, 0>, grid_y, dim3(nthr), s1, 0, in, out, gy, tb, none); };
    auto iy2 = [&](const double* in, double* out) { return bk_launch_pdl(k_dct2v2<LOGM, LOGW, true, 1>, grid_y, dim3(nthr), s1, 0, in, out, gy, tb, none); };
    auto fu2 = [&](const double* in, double* out) { return bk_launch_pdl(k_dct2v2<LOGM, LOGW, true, 2>, grid_y, dim3(nthr), s2, 0, in, out, gy, tb, sy); };
    double *r1, *r2; CK(cudaMalloc(&r1, 8 * N)); CK(cudaMalloc(&r2, 8 * N));
    CK(fx(x, r1)); CK(fx2(x, a)); printf("  v2 x forward vs v1   rel err %.2e\n", maxdiff(a, r1, N));
    CK(ix(r1, r2)); CK(ix2(r1, a)); printf("  v2 x inverse vs v1   rel err %.2e\n", maxdiff(a, r2, N));
    CK(fy(x, r1)); CK(fy2(x, a)); printf("  v2 y forward vs v1   rel err %.2e\n", maxdiff(a, r1, N));
    CK(iy(r1, r2)); CK(iy2(r1, a)); printf("  v2 y inverse vs v1   rel err %.2e\n", maxdiff(a, r2, N));
    CK(fused_y(x, r1)); CK(fu2(x, a)); printf("  v2 y fused vs v1     rel err %.2e\n", maxdiff(a, r1, N));
    timeit("v2 x forward", [&] { fx2(x, a); }, 16.0 * N);
    timeit("v2 x inverse", [&] { ix2(a, b); }, 16.0 * N);
    timeit("v2 y forward", [&] { fy2(x, a); }, 16.0 * N);
    timeit("v2 y fused", [&] { fu2(a, b); }, 16.0 * N);
    timeit("v2 preconditioner (3 launches)", [&] { fx2(x, a); fu2(a, b); ix2(b, c2); }, 48.0 * N);
  };
  v2(std::integral_constant<int, 9>{}, std::integral_constant<int, 2>{});
  v2(std::integral_constant<int, 9>{}, std::integral_constant<int, 1>{});
  v2(std::integral_constant<int, 9>{}, std::integral_constant<int, 3>{});
  v2(std::integral_constant<int, 8>{}, std::integral_constant<int, 3>{});
  v2(std::integral_constant<int, 8>{}, std::integral_constant<int, 2>{});
  return 0;
}
