"""The executable specification of the register-resident DCT kernels (tools/fftcheck/model.py: index maps, per-pass twiddle
tables, pair split, fused forward / symbol / inverse, thread by thread as csrc/bk_fft_fast.cuh runs them) against scipy.fft,
and the host-side table builder's digit reversal against the model's.  CPU-only; the device parity is tests/test_gpu_precond.py."""
import os
import sys

import numpy as np
import pytest
import scipy.fft as sf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "fftcheck"))
import model  # noqa: E402


@pytest.mark.parametrize("n,E", [(64, 4), (64, 32), (128, 8), (256, 16), (512, 4), (1024, 8), (1024, 32), (2048, 8)])
def test_pair_fft_model_matches_scipy(n, E):
    rng = np.random.default_rng(n + E)
    pl = model.Plan(n, E)
    assert int(np.prod(pl.rad)) == n and all(r <= E for r in pl.rad)
    assert sorted(pl.k_of_pos) == list(range(n))                     # the digit reversal is a permutation
    x1, x2 = rng.standard_normal(n), rng.standard_normal(n)
    C1, C2 = model.dct_pair_forward(pl, x1, x2)                       # the kernels return 2 C
    r1, r2 = sf.dct(x1, type=2), sf.dct(x2, type=2)                   # scipy's unnormalised DCT-II is 2 C as well
    assert max(np.abs(C1 - r1).max(), np.abs(C2 - r2).max()) < 1e-12 * n
    y1, y2 = model.dct_pair_inverse(pl, r1 / 2, r2 / 2)               # inverse kernels return n x
    assert max(np.abs(y1 / n - x1).max(), np.abs(y2 / n - x2).max()) < 1e-12
    s1, s2 = rng.uniform(0.5, 2.0, n), rng.uniform(0.5, 2.0, n)
    f1, f2 = model.fused_pair(pl, x1, x2, s1 / (2 * n), s2 / (2 * n))
    g1, g2 = sf.idct(sf.dct(x1, type=2) * s1, type=2), sf.idct(sf.dct(x2, type=2) * s2, type=2)
    assert max(np.abs(f1 - g1).max(), np.abs(f2 - g2).max()) < 1e-12


def test_shipped_padding_is_conflict_free_in_the_bank_model():
    """pad(i) = i + i/4 (+ i/2^PB) with the PB of csrc/bk_fft_fast.cuh::Cfg: one wavefront per 128 bytes in every pass"""
    for n, E, pb in [(1024, 8, 3), (512, 4, 0), (1024, 32, 5), (1024, 16, 4), (256, 4, 0)]:
        pl = model.Plan(n, E)
        pad = (lambda i, pb=pb: i + (i >> 2) + ((i >> pb) if pb else 0))
        PP = max(2, 64 // (n // E))
        ratios = model.bank_report_em(pl, pad, PP)
        assert max(ratios[:-1]) <= 1.0 + 1e-9, (n, E, ratios)   # every FFT pass: one wavefront per 128 bytes
        assert ratios[-1] <= 1.15, (n, E, ratios)                # the partner read Z[n-k]: at most 15 % extra wavefronts


def test_natural_order_staging_padding_is_conflict_free():
    """padn(k) = k + k/4 (+ k/2^NB) with the NB table of csrc/bk_fft_fast.cuh::Cfg: the digit-reversed scatter and the natural
    read of the contiguous kernels' staging array are both conflict-free in the bank model"""
    def nb(loge, logn):
        if loge == 2:
            return 0 if logn <= 6 else 4 if logn <= 8 else 6 if logn <= 10 else 8
        if loge == 3:
            return 0 if logn <= 6 else 3 if logn <= 9 else 6
        if loge == 4:
            return 0 if logn <= 8 else 4
        return 0 if logn <= 10 else 5
    for loge in (2, 3, 4, 5):
        for logn in range(max(6, loge + 1), 12):
            n, E = 1 << logn, 1 << loge
            pl = model.Plan(n, E)
            b = nb(loge, logn)
            padn = (lambda k, b=b: k + (k >> 2) + ((k >> b) if b else 0))
            PP = max(2, 64 // (n // E))
            w, r = model.bank_report_natural(pl, padn, PP)
            assert w <= 1.0 + 1e-9 and r <= 1.0 + 1e-9, (n, E, b, w, r)
