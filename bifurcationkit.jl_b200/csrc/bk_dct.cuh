// bk_dct.cuh -- fast shared-memory DCT-II / inverse for power-of-two line lengths (K6).
//
// A length-n DCT-II of a real line costs ONE complex FFT of length M = n/2:
//   Makhoul reordering  v[m] = x[2m], v[n-1-m] = x[2m+1]           (DCT-II -> DFT of a real sequence)
//   real-FFT packing    z[q] = v[2q] + i v[2q+1],  Z = FFT_M(z)    (length-n real DFT from a length-n/2 complex FFT)
//   unpack              V[k] = Ev[k] + e^{-2 pi i k/n} Od[k],  Ev/Od from Z[k], conj(Z[M-k])
//   twiddle             A[k] = e^{-i pi k/2n} V[k] = C[k] - i C[n-k]   (one complex value -> two outputs)
// and the inverse runs the same steps backwards.  The FFT is radix-4 (two radix-2 stages per pass in
// registers), in place in shared memory after a bit-reversed load; W lines are transformed per CTA.
// MODE 0: forward, 1: inverse, 2: forward + divide by the operator symbol + inverse (used for the last
// dimension, so an SH preconditioner application is 3 kernels in 2-D and 5 in 3-D).
#pragma once
#include "bk_common.cuh"

struct LineGeom {
  int n;         // line length
  long long es;  // element stride along the line
  int nx;        // extent of the contiguous (batch) index; 1 for x-lines
  long long os;  // stride of the outer index
  int nouter;    // number of outer indices
};

struct DctTables {
  const double2* tw;   // exp(-2 pi i k / M), k < M/2          (M = n/2)
  const double2* wn;   // exp(-2 pi i k / n), k <= n/2
  const double2* dtw;  // exp(-i pi k / 2n),  k <= n/2
};

struct SymbolArgs {
  const double* lam_e;  // eigenvalues along the transformed dimension
  const double* lam_x;  // eigenvalues along x (index = global x of the line), may be NULL
  const double* lam_o;  // eigenvalues along the outer index, may be NULL
  double shift;         // symbol = (1 + lam_e + lam_x + lam_o)^2 + shift
  // optional pass-through of trailing entries (bordered vectors: the preconditioner is the identity on the border);
  // done by one thread of the kernel instead of a separate 8-byte memcpy on the copy engine
  const double* tail_src;
  double* tail_dst;
  int tail_n;
};

__device__ __forceinline__ double2 dct_cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 dct_conj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ int dct_bitrev(int v, int logm) { return (int)(__brev((unsigned)v) >> (32 - logm)); }
// shared-memory index padding: one extra complex slot per 8 (power-of-two strides would otherwise map the
// radix-4 butterflies of the early passes onto the same banks: 16-way conflicts measured as ~20 us kernels)
__device__ __forceinline__ int dct_pad(int i) { return i + (i >> 3); }
#define DCT_PADDED(M) ((M) + ((M) >> 3) + 1)

// in-place DIT FFT of W lines of M complex points held bit-reversed in shared memory
template <bool STRIDED>
__device__ __forceinline__ void dct_fft(double2* s, int M, int logM, int W, int logW, const double2* __restrict__ tw,
                                        bool inverse) {
  int st = 0;
  if (logM & 1) {
    for (int b = threadIdx.x; b < (M >> 1) * W; b += blockDim.x) {
      int line = STRIDED ? (b & (W - 1)) : (b >> (logM - 1));
      int bf = STRIDED ? (b >> logW) : (b & ((M >> 1) - 1));
      const int MP = DCT_PADDED(M);
      double2* p0 = STRIDED ? s + (long long)dct_pad(2 * bf) * W + line : s + (long long)line * MP + dct_pad(2 * bf);
      double2* p1 = STRIDED ? s + (long long)dct_pad(2 * bf + 1) * W + line : s + (long long)line * MP + dct_pad(2 * bf + 1);
      double2 a = *p0, c = *p1;
      *p0 = make_double2(a.x + c.x, a.y + c.y);
      *p1 = make_double2(a.x - c.x, a.y - c.y);
    }
    __syncthreads();
    st = 1;
  }
  for (; st < logM; st += 2) {
    const int half = 1 << st;
    const int s1 = M >> (st + 1), s2 = M >> (st + 2);
    for (int g = threadIdx.x; g < (M >> 2) * W; g += blockDim.x) {
      int line = STRIDED ? (g & (W - 1)) : (g >> (logM - 2));
      int gi = STRIDED ? (g >> logW) : (g & ((M >> 2) - 1));
      int grp = gi >> st, pos = gi & (half - 1);
      int i = (grp << (st + 2)) + pos;
      double2 w1 = __ldg(tw + pos * s1), w2 = __ldg(tw + pos * s2);
      if (inverse) {
        w1.y = -w1.y;
        w2.y = -w2.y;
      }
      const double2 w3 = inverse ? make_double2(-w2.y, w2.x) : make_double2(w2.y, -w2.x);  // w2 * (+-i)
      const int MP = DCT_PADDED(M);
      double2* pa = STRIDED ? s + (long long)dct_pad(i) * W + line : s + (long long)line * MP + dct_pad(i);
      double2* pb = STRIDED ? s + (long long)dct_pad(i + half) * W + line : s + (long long)line * MP + dct_pad(i + half);
      double2* pc = STRIDED ? s + (long long)dct_pad(i + 2 * half) * W + line : s + (long long)line * MP + dct_pad(i + 2 * half);
      double2* pd = STRIDED ? s + (long long)dct_pad(i + 3 * half) * W + line : s + (long long)line * MP + dct_pad(i + 3 * half);
      double2 a = *pa, b = *pb, c = *pc, d = *pd;
      double2 t = dct_cmul(w1, b);
      double2 a1 = make_double2(a.x + t.x, a.y + t.y), b1 = make_double2(a.x - t.x, a.y - t.y);
      t = dct_cmul(w1, d);
      double2 c1 = make_double2(c.x + t.x, c.y + t.y), d1 = make_double2(c.x - t.x, c.y - t.y);
      t = dct_cmul(w2, c1);
      *pa = make_double2(a1.x + t.x, a1.y + t.y);
      *pc = make_double2(a1.x - t.x, a1.y - t.y);
      t = dct_cmul(w3, d1);
      *pb = make_double2(b1.x + t.x, b1.y + t.y);
      *pd = make_double2(b1.x - t.x, b1.y - t.y);
    }
    __syncthreads();
  }
}

// shared memory: s[M*W] complex, and for MODE 2 additionally cb[n*W] real
template <bool STRIDED, int MODE>
static __global__ void __launch_bounds__(1024) k_dct2(const double* __restrict__ in, double* __restrict__ out, LineGeom g, int logM,
                                                      int W, int logW, DctTables tb, SymbolArgs sy) {
  bk_pdl_sync();
  extern __shared__ __align__(16) double2 sdct[];
  const int n = g.n, M = n >> 1;
  double2* s = sdct;
  double* sd = reinterpret_cast<double*>(sdct);
  const int MP = DCT_PADDED(M);
  double* cb = reinterpret_cast<double*>(sdct + (size_t)MP * W);
  long long base, lstride;
  int nl, x0 = 0, o = 0;
  if (STRIDED) {
    x0 = blockIdx.x * W;
    o = blockIdx.y;
    nl = min(W, g.nx - x0);
    base = x0 + (long long)o * g.os;
    lstride = 1;
  } else {
    long long l0 = (long long)blockIdx.x * W;
    nl = (int)min((long long)W, (long long)g.nouter - l0);
    base = l0 * g.os;
    lstride = g.os;
  }
  const double inv_m = 1.0 / M;
  if (sy.tail_n > 0 && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < sy.tail_n) sy.tail_dst[threadIdx.x] = sy.tail_src[threadIdx.x];

  // ---------------- forward half (MODE 0, 2)
  if (MODE != 1) {
    for (int q = threadIdx.x; q < n * W; q += blockDim.x) {
      int line = STRIDED ? (q & (W - 1)) : (q >> (logM + 1));
      int e = STRIDED ? (q >> logW) : (q & (n - 1));
      double xv = (line < nl) ? in[base + line * lstride + (long long)e * g.es] : 0.0;
      int m = (e & 1) ? (n - 1 - (e >> 1)) : (e >> 1);
      int p = dct_bitrev(m >> 1, logM);
      long long ci = STRIDED ? (long long)dct_pad(p) * W + line : (long long)line * MP + dct_pad(p);
      sd[2 * ci + (m & 1)] = xv;
    }
    __syncthreads();
    dct_fft<STRIDED>(s, M, logM, W, logW, tb.tw, false);
    for (int q = threadIdx.x; q < (M + 1) * W; q += blockDim.x) {
      // items 0 .. M*W-1 cover k < M (shift/mask indexing); the last W items are k = M of every line
      int line, k;
      if (q < M * W) {
        line = STRIDED ? (q & (W - 1)) : (q >> logM);
        k = STRIDED ? (q >> logW) : (q & (M - 1));
      } else {
        line = q - M * W;
        k = M;
      }
      if (line >= nl) continue;
      int k0 = k & (M - 1), k1 = (M - k) & (M - 1);
      double2 zk = STRIDED ? s[(long long)dct_pad(k0) * W + line] : s[(long long)line * MP + dct_pad(k0)];
      double2 zc = dct_conj(STRIDED ? s[(long long)dct_pad(k1) * W + line] : s[(long long)line * MP + dct_pad(k1)]);
      double2 ev = make_double2(0.5 * (zk.x + zc.x), 0.5 * (zk.y + zc.y));
      double2 df = make_double2(zk.x - zc.x, zk.y - zc.y);
      double2 od = make_double2(0.5 * df.y, -0.5 * df.x);  // -i (zk - zc) / 2
      double2 t = dct_cmul(__ldg(tb.wn + k), od);
      double2 v = make_double2(ev.x + t.x, ev.y + t.y);
      double2 a = dct_cmul(__ldg(tb.dtw + k), v);  // C[k] - i C[n-k]
      if (MODE == 0) {
        long long gb = base + line * lstride;
        out[gb + (long long)k * g.es] = a.x;
        if (k >= 1 && k < M) out[gb + (long long)(n - k) * g.es] = -a.y;
      } else {
        // divide by the symbol and park C in shared memory
        double lx = sy.lam_x ? __ldg(sy.lam_x + (STRIDED ? x0 + line : 0)) : 0.0;
        double lo = sy.lam_o ? __ldg(sy.lam_o + o) : 0.0;
        double t1 = 1.0 + lx + lo + __ldg(sy.lam_e + k);
        cb[(long long)k * W + line] = a.x / (t1 * t1 + sy.shift);
        if (k >= 1 && k < M) {
          double t2 = 1.0 + lx + lo + __ldg(sy.lam_e + n - k);
          cb[(long long)(n - k) * W + line] = -a.y / (t2 * t2 + sy.shift);
        }
      }
    }
    __syncthreads();
  }
  // ---------------- inverse half (MODE 1, 2)
  if (MODE != 0) {
    for (int q = threadIdx.x; q < M * W; q += blockDim.x) {
      int line = STRIDED ? (q & (W - 1)) : (q >> logM);
      int k = STRIDED ? (q >> logW) : (q & (M - 1));
      double2 z = make_double2(0.0, 0.0);
      if (line < nl) {
        // V[j] = conj(dtw[j]) (C[j] - i C[n-j]),  C[n] = 0;  need j = k and j = M - k
        const int j2 = M - k;
        double ck, cnk, cj, cnj;
        if (MODE == 1) {
          long long gb = base + line * lstride;
          ck = in[gb + (long long)k * g.es];
          cnk = k > 0 ? in[gb + (long long)(n - k) * g.es] : 0.0;
          cj = in[gb + (long long)j2 * g.es];
          cnj = in[gb + (long long)(n - j2) * g.es];  // j2 >= 1 always (k < M)
        } else {
          ck = cb[(long long)k * W + line];
          cnk = k > 0 ? cb[(long long)(n - k) * W + line] : 0.0;
          cj = cb[(long long)j2 * W + line];
          cnj = cb[(long long)(n - j2) * W + line];
        }
        double2 vk = dct_cmul(dct_conj(__ldg(tb.dtw + k)), make_double2(ck, -cnk));
        double2 vj = dct_cmul(dct_conj(__ldg(tb.dtw + j2)), make_double2(cj, -cnj));
        double2 vjc = dct_conj(vj);
        double2 ev = make_double2(0.5 * (vk.x + vjc.x), 0.5 * (vk.y + vjc.y));
        double2 od = dct_cmul(dct_conj(__ldg(tb.wn + k)), make_double2(0.5 * (vk.x - vjc.x), 0.5 * (vk.y - vjc.y)));
        z = make_double2(ev.x - od.y, ev.y + od.x);  // Ev + i Od
      }
      int p = dct_bitrev(k, logM);
      if (STRIDED)
        s[(long long)dct_pad(p) * W + line] = z;
      else
        s[(long long)line * MP + dct_pad(p)] = z;
    }
    __syncthreads();
    dct_fft<STRIDED>(s, M, logM, W, logW, tb.tw, true);
    for (int q = threadIdx.x; q < n * W; q += blockDim.x) {
      int line = STRIDED ? (q & (W - 1)) : (q >> (logM + 1));
      int e = STRIDED ? (q >> logW) : (q & (n - 1));
      if (line >= nl) continue;
      int m = (e & 1) ? (n - 1 - (e >> 1)) : (e >> 1);
      long long ci = STRIDED ? (long long)dct_pad(m >> 1) * W + line : (long long)line * MP + dct_pad(m >> 1);
      out[base + line * lstride + (long long)e * g.es] = sd[2 * ci + (m & 1)] * inv_m;
    }
  }
}
