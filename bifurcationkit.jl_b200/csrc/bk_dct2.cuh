// bk_dct2.cuh -- second version of the shared-memory DCT-II kernels of bk_dct.cuh (same algorithm, same tables, same
// LineGeom / SymbolArgs interface, same results to rounding) with the integer work taken out of the inner loops.
//
// Why: ncu on k_dct2 (profiles/r01c_ncu_k_dct2.csv) shows 187 lane-instructions per point of which 15 % are fp64; the
// rest is address arithmetic with run-time shifts and 64-bit indices, twiddle loads from global memory inside every FFT
// pass (long-scoreboard stalls) and a post-processing step that visits Z[k] and Z[M-k] twice.  Here
//   * the line length is a template parameter: every shift / mask is an immediate, the FFT passes are unrolled;
//   * shared-memory indices are 32-bit; for passes with half >= 8 the four butterfly addresses are base + c * const;
//   * the FFT twiddles are staged once per CTA in shared memory;
//   * post- and pre-processing handle the pair (k, M-k) together: V[M-k] = conj(ev - wn[k] od) reuses everything;
//   * the contiguous direction moves two values per thread and access (x[2j], x[2j+1] are neighbours in memory);
//   * shared-memory bank conflicts (v1: 1.59 wavefronts per ideal wavefront in ncu) are designed out with the model in
//     tools/dctcheck/bank_model.py: index padding i + i/8 + i/64 (the bit-reversed scatter of the load phase strides by
//     M/16, which i + i/8 alone maps onto one bank: 8-way), per-pass twiddle tables that are contiguous in the butterfly
//     position (a strided read of one table is 8-way), and for the passes with half < 8 a lane order in which the group index
//     runs fastest.  Model, n = 1024 / W = 4: 2.42 -> 1.06 (x kernels), 1.33 -> 1.24 (y kernels).
// STATUS: opt-in (BK_DCT_V2=1, see bk_precond.cu).  Checked against a naive DCT on the host by compiling this header with
// BK_DCT_HOST_EMU (tools/dctcheck/dct2_host_emu.cpp: one emulated thread per CTA runs every phase to completion, which is
// exact because the items of a phase are independent); NOT yet run or timed on a GPU -- tools/dctcheck/dct_check.cu does both.
#pragma once
#ifndef BK_DCT_HOST_EMU
#include "bk_dct.cuh"
#endif

// per-pass twiddle tables: pass ST (= P0, P0 + 2, ...) holds w1[pos], w2[pos], pos < 2^ST, at offset 2 (2^ST - 2^P0) / 3
__host__ __device__ constexpr int d2_tw_off(int st, int p0) { return 2 * ((1 << st) - (1 << p0)) / 3; }

template <int LOGM>
struct Dct2Cfg {
  static constexpr int M = 1 << LOGM, N = 2 * M, P0 = LOGM & 1;
  static constexpr int MP = (M - 1) + ((M - 1) >> 3) + ((M - 1) >> 6) + 1;  // d2_pad(M - 1) + 1
  static constexpr int TWN = 2 * (M - (1 << P0)) / 3;
};
__device__ __forceinline__ int d2_pad(int i) { return i + (i >> 3) + (i >> 6); }
__device__ __forceinline__ double2 d2_cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 d2_conj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double2 d2_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 d2_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// complex slot of (line, padded index pi)
template <int LOGM, int LOGW, bool STRIDED>
__device__ __forceinline__ int d2_slot(int line, int pi) {
  return STRIDED ? (pi << LOGW) + line : line * Dct2Cfg<LOGM>::MP + pi;
}

// one radix-4 pass (two radix-2 stages) at stage ST, in place, DIT, bit-reversed input
template <int LOGM, int LOGW, bool STRIDED, int ST>
__device__ __forceinline__ void d2_pass4(double2* s, const double2* stw, bool inverse) {
  constexpr int M = 1 << LOGM, W = 1 << LOGW, HALF = 1 << ST;
  const double2* w1t = stw + d2_tw_off(ST, LOGM & 1);
  const double2* w2t = w1t + HALF;
  for (int g = threadIdx.x; g < (M >> 2) * W; g += blockDim.x) {
    const int line = STRIDED ? (g & (W - 1)) : (g >> (LOGM - 2));
    int gi = STRIDED ? (g >> LOGW) : (g & ((M >> 2) - 1));
    if (!STRIDED && ST > 0 && ST < 3 && (M >> 2) >= 8 * HALF) {
      // line-major layout, half < 8: within a block of 8 * HALF butterflies let the group index run fastest over the lanes
      // (a quarter-warp then touches 8 different groups = 8 different bank groups instead of 8 / HALF)
      const int r = gi & (8 * HALF - 1);
      gi = (gi & ~(8 * HALF - 1)) + ((r & 7) << ST) + (r >> 3);
    }
    const int pos = gi & (HALF - 1);
    const int i = ((gi >> ST) << (ST + 2)) + pos;
    double2 w1 = w1t[pos], w2 = w2t[pos];
    if (inverse) {
      w1.y = -w1.y;
      w2.y = -w2.y;
    }
    const double2 w3 = inverse ? make_double2(-w2.y, w2.x) : make_double2(w2.y, -w2.x);  // w2 * (+-i)
    int ia, ib, ic, id;
    if (HALF >= 64) {  // (c * HALF) is a multiple of 64: pad(i + c HALF) = pad(i) + c (HALF + HALF / 8 + HALF / 64)
      constexpr int STEP = HALF + (HALF >> 3) + (HALF >> 6);
      ia = d2_pad(i);
      ib = ia + STEP;
      ic = ia + 2 * STEP;
      id = ia + 3 * STEP;
    } else {
      ia = d2_pad(i);
      ib = d2_pad(i + HALF);
      ic = d2_pad(i + 2 * HALF);
      id = d2_pad(i + 3 * HALF);
    }
    double2* pa = s + d2_slot<LOGM, LOGW, STRIDED>(line, ia);
    double2* pb = s + d2_slot<LOGM, LOGW, STRIDED>(line, ib);
    double2* pc = s + d2_slot<LOGM, LOGW, STRIDED>(line, ic);
    double2* pd = s + d2_slot<LOGM, LOGW, STRIDED>(line, id);
    const double2 a = *pa, b = *pb, c = *pc, d = *pd;
    double2 t = d2_cmul(w1, b);
    const double2 a1 = d2_add(a, t), b1 = d2_sub(a, t);
    t = d2_cmul(w1, d);
    const double2 c1 = d2_add(c, t), d1 = d2_sub(c, t);
    t = d2_cmul(w2, c1);
    *pa = d2_add(a1, t);
    *pc = d2_sub(a1, t);
    t = d2_cmul(w3, d1);
    *pb = d2_add(b1, t);
    *pd = d2_sub(b1, t);
  }
  __syncthreads();
}

template <int LOGM, int LOGW, bool STRIDED, int ST>
struct D2Passes {
  static __device__ __forceinline__ void run(double2* s, const double2* stw, bool inverse) {
    if constexpr (ST < LOGM) {
      d2_pass4<LOGM, LOGW, STRIDED, ST>(s, stw, inverse);
      D2Passes<LOGM, LOGW, STRIDED, ST + 2>::run(s, stw, inverse);
    }
  }
};

template <int LOGM, int LOGW, bool STRIDED>
__device__ __forceinline__ void d2_fft(double2* s, const double2* stw, bool inverse) {
  constexpr int M = 1 << LOGM, W = 1 << LOGW;
  if constexpr (LOGM & 1) {  // leading radix-2 stage (twiddle 1)
    for (int b = threadIdx.x; b < (M >> 1) * W; b += blockDim.x) {
      const int line = STRIDED ? (b & (W - 1)) : (b >> (LOGM - 1));
      const int bf = STRIDED ? (b >> LOGW) : (b & ((M >> 1) - 1));
      const int p = d2_pad(2 * bf);  // 2 bf is even, so pad(2 bf + 1) = pad(2 bf) + 1
      double2* p0 = s + d2_slot<LOGM, LOGW, STRIDED>(line, p);
      double2* p1 = s + d2_slot<LOGM, LOGW, STRIDED>(line, p + 1);
      const double2 a = *p0, c = *p1;
      *p0 = d2_add(a, c);
      *p1 = d2_sub(a, c);
    }
    __syncthreads();
    D2Passes<LOGM, LOGW, STRIDED, 1>::run(s, stw, inverse);
  } else {
    D2Passes<LOGM, LOGW, STRIDED, 0>::run(s, stw, inverse);
  }
}

// MODE 0: forward, 1: inverse, 2: forward + divide by the operator symbol + inverse.
// shared memory: s[MP * W] complex | stw[TWN] complex (per-pass twiddle tables) | (MODE 2) cb[N * W] real
template <int LOGM, int LOGW, bool STRIDED, int MODE>
static __global__ void __launch_bounds__(512) k_dct2v2(const double* __restrict__ in, double* __restrict__ out, LineGeom g, DctTables tb,
                                                       SymbolArgs sy) {
  using C = Dct2Cfg<LOGM>;
  constexpr int M = C::M, n = C::N, MP = C::MP, W = 1 << LOGW;
#ifndef BK_DCT_HOST_EMU
  bk_pdl_sync();
  extern __shared__ __align__(16) double2 sdct2[];
#endif
  double2* s = sdct2;
  double* sd = reinterpret_cast<double*>(sdct2);
  double2* stw = sdct2 + MP * W;
  double* cb = reinterpret_cast<double*>(stw + C::TWN);
  long long base, lstride;
  int nl, x0 = 0, o = 0;
  if (STRIDED) {
    x0 = blockIdx.x * W;
    o = blockIdx.y;
    nl = min(W, g.nx - x0);
    base = x0 + (long long)o * g.os;
    lstride = 1;
  } else {
    const long long l0 = (long long)blockIdx.x * W;
    nl = (int)min((long long)W, (long long)g.nouter - l0);
    base = l0 * g.os;
    lstride = g.os;
  }
  const long long es = g.es;
  const double inv_m = 1.0 / M;
  if (sy.tail_n > 0 && blockIdx.x == 0 && blockIdx.y == 0 && (int)threadIdx.x < sy.tail_n) sy.tail_dst[threadIdx.x] = sy.tail_src[threadIdx.x];
  for (int st = C::P0; st < LOGM; st += 2) {  // per-pass twiddle tables: w1[pos] = tw[pos M / 2^(st+1)], w2[pos] = tw[pos M / 2^(st+2)]
    const int half = 1 << st;
    double2* t = stw + d2_tw_off(st, C::P0);
    for (int q = threadIdx.x; q < 2 * half; q += blockDim.x) {
      const int pos = q & (half - 1);
      t[q] = __ldg(tb.tw + (q < half ? (pos << (LOGM - st - 1)) : (pos << (LOGM - st - 2))));
    }
  }

  // ---------------------------------------------------------------- forward half (MODE 0, 2)
  if (MODE != 1) {
    // Makhoul reordering + real-FFT packing + bit reversal: x[2j] -> v[j], x[2j+1] -> v[n-1-j];  v[m] is component (m & 1)
    // of z[m >> 1];  (n-1-j) >> 1 = M-1-(j >> 1) and bitrev(M-1-a) = M-1-bitrev(a)
    for (int q = threadIdx.x; q < M * W; q += blockDim.x) {
      const int line = STRIDED ? (q & (W - 1)) : (q >> LOGM);
      const int j = STRIDED ? (q >> LOGW) : (q & (M - 1));
      double x0v = 0.0, x1v = 0.0;
      if (line < nl) {
        const double* p = in + base + line * lstride + (long long)(2 * j) * es;
        if (STRIDED) {
          x0v = p[0];
          x1v = p[es];
        } else {
          const double2 t = *reinterpret_cast<const double2*>(p);  // es == 1, lines start on 16-byte boundaries (n even)
          x0v = t.x;
          x1v = t.y;
        }
      }
      const int p0 = (int)(__brev((unsigned)(j >> 1)) >> (32 - LOGM));
      const int p1 = (M - 1) - p0;
      sd[2 * d2_slot<LOGM, LOGW, STRIDED>(line, d2_pad(p0)) + (j & 1)] = x0v;
      sd[2 * d2_slot<LOGM, LOGW, STRIDED>(line, d2_pad(p1)) + 1 - (j & 1)] = x1v;
    }
    __syncthreads();
    d2_fft<LOGM, LOGW, STRIDED>(s, stw, false);
    // unpack + DCT twiddle, pairs (k, M - k), k = 0 .. M/2;  k = 0 also produces C[M]
    for (int q = threadIdx.x; q < (M / 2 + 1) * W; q += blockDim.x) {
      int line, k;
      if (q < (M / 2) * W) {
        line = STRIDED ? (q & (W - 1)) : (q >> (LOGM - 1));
        k = STRIDED ? (q >> LOGW) : (q & (M / 2 - 1));
      } else {
        line = q - (M / 2) * W;
        k = M / 2;
      }
      if (line >= nl) continue;
      const int kc = (M - k) & (M - 1);
      const double2 zk = s[d2_slot<LOGM, LOGW, STRIDED>(line, d2_pad(k))];
      const double2 zc = d2_conj(s[d2_slot<LOGM, LOGW, STRIDED>(line, d2_pad(kc))]);
      const double2 ev = make_double2(0.5 * (zk.x + zc.x), 0.5 * (zk.y + zc.y));
      const double2 od = make_double2(0.5 * (zk.y - zc.y), -0.5 * (zk.x - zc.x));  // -i (zk - zc) / 2
      const double2 t = d2_cmul(__ldg(tb.wn + k), od);
      const double2 a = d2_cmul(__ldg(tb.dtw + k), d2_add(ev, t));                 // C[k] - i C[n-k]
      const int k2 = M - k;                                                        // partner (k = 0 -> M, k = M/2 -> itself)
      const double2 a2 = d2_cmul(__ldg(tb.dtw + k2), d2_conj(d2_sub(ev, t)));      // C[M-k] - i C[M+k]
      if (MODE == 0) {
        double* ob = out + base + line * lstride;
        ob[(long long)k * es] = a.x;
        if (k >= 1) ob[(long long)(n - k) * es] = -a.y;
        if (k2 != k) {
          ob[(long long)k2 * es] = a2.x;
          if (k2 < M) ob[(long long)(n - k2) * es] = -a2.y;
        }
      } else {
        const double lx = sy.lam_x ? __ldg(sy.lam_x + (STRIDED ? x0 + line : 0)) : 0.0;
        const double lo = sy.lam_o ? __ldg(sy.lam_o + o) : 0.0;
        const double c0 = 1.0 + lx + lo;
        double tt = c0 + __ldg(sy.lam_e + k);
        cb[k * W + line] = a.x / (tt * tt + sy.shift);
        if (k >= 1) {
          tt = c0 + __ldg(sy.lam_e + n - k);
          cb[(n - k) * W + line] = -a.y / (tt * tt + sy.shift);
        }
        if (k2 != k) {
          tt = c0 + __ldg(sy.lam_e + k2);
          cb[k2 * W + line] = a2.x / (tt * tt + sy.shift);
          if (k2 < M) {
            tt = c0 + __ldg(sy.lam_e + n - k2);
            cb[(n - k2) * W + line] = -a2.y / (tt * tt + sy.shift);
          }
        }
      }
    }
    __syncthreads();
  }
  // ---------------------------------------------------------------- inverse half (MODE 1, 2)
  if (MODE != 0) {
    // z[k] = Ev + i Od from V[k] = conj(dtw[k]) (C[k] - i C[n-k]) and V[M-k];  pair (k, M-k): z[M-k] = conj(Ev) + i conj(Od)
    for (int q = threadIdx.x; q < (M / 2 + 1) * W; q += blockDim.x) {
      int line, k;
      if (q < (M / 2) * W) {
        line = STRIDED ? (q & (W - 1)) : (q >> (LOGM - 1));
        k = STRIDED ? (q >> LOGW) : (q & (M / 2 - 1));
      } else {
        line = q - (M / 2) * W;
        k = M / 2;
      }
      const int j2 = M - k;  // k = 0 -> M (C[n] = 0 never read: cnk = 0)
      double2 z = make_double2(0.0, 0.0), z2 = make_double2(0.0, 0.0);
      if (line < nl) {
        double ck, cnk, cj, cnj;
        if (MODE == 1) {
          const double* ib = in + base + line * lstride;
          ck = ib[(long long)k * es];
          cnk = k > 0 ? ib[(long long)(n - k) * es] : 0.0;
          cj = ib[(long long)j2 * es];
          cnj = ib[(long long)(n - j2) * es];
        } else {
          ck = cb[k * W + line];
          cnk = k > 0 ? cb[(n - k) * W + line] : 0.0;
          cj = cb[j2 * W + line];
          cnj = cb[(n - j2) * W + line];
        }
        const double2 vk = d2_cmul(d2_conj(__ldg(tb.dtw + k)), make_double2(ck, -cnk));
        const double2 vjc = d2_conj(d2_cmul(d2_conj(__ldg(tb.dtw + j2)), make_double2(cj, -cnj)));
        const double2 ev = make_double2(0.5 * (vk.x + vjc.x), 0.5 * (vk.y + vjc.y));
        const double2 od = d2_cmul(d2_conj(__ldg(tb.wn + k)), make_double2(0.5 * (vk.x - vjc.x), 0.5 * (vk.y - vjc.y)));
        z = make_double2(ev.x - od.y, ev.y + od.x);    // Ev + i Od
        z2 = make_double2(ev.x + od.y, od.x - ev.y);   // conj(Ev) + i conj(Od)
      }
      const int p = (int)(__brev((unsigned)k) >> (32 - LOGM));
      s[d2_slot<LOGM, LOGW, STRIDED>(line, d2_pad(p))] = z;
      if (k >= 1 && k < M / 2) {
        const int p2 = (int)(__brev((unsigned)(M - k)) >> (32 - LOGM));
        s[d2_slot<LOGM, LOGW, STRIDED>(line, d2_pad(p2))] = z2;
      }
    }
    __syncthreads();
    d2_fft<LOGM, LOGW, STRIDED>(s, stw, true);
    for (int q = threadIdx.x; q < M * W; q += blockDim.x) {
      const int line = STRIDED ? (q & (W - 1)) : (q >> LOGM);
      const int j = STRIDED ? (q >> LOGW) : (q & (M - 1));
      if (line >= nl) continue;
      const int a0 = j >> 1, a1 = (M - 1) - a0;  // natural order after the DIT passes
      const double v0 = sd[2 * d2_slot<LOGM, LOGW, STRIDED>(line, d2_pad(a0)) + (j & 1)] * inv_m;
      const double v1 = sd[2 * d2_slot<LOGM, LOGW, STRIDED>(line, d2_pad(a1)) + 1 - (j & 1)] * inv_m;
      double* p = out + base + line * lstride + (long long)(2 * j) * es;
      if (STRIDED) {
        p[0] = v0;
        p[es] = v1;
      } else {
        *reinterpret_cast<double2*>(p) = make_double2(v0, v1);
      }
    }
  }
}
