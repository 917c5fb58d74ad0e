"""GPU parity of the K6 transform kernels (bk_fft_fast.cuh: register-resident power-of-two DCT-II; bk_fft_gen.cuh:
mixed-radix DCT-II / DST-I of any length) against the oracle's scipy.fft restatement, through bk_precond_apply.
Tolerance 1e-11 relative (fp64; different summation order only)."""
import ctypes as C

import numpy as np
import pytest

import __graft_entry__ as g
from oracle import precond as oprecond

pytestmark = pytest.mark.gpu

LX, LY = 8 * np.pi, 4 * np.pi / np.sqrt(3)


@pytest.fixture(scope="module")
def bk():
    return g.load_package()


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


# fast kernels: every instantiated line length 64..2048 in both the contiguous (x) and the strided (y) role, ragged batch
# counts (odd number of line pairs per CTA), mixed fast / general dimensions, the reference example's own grid 151 x 100
@pytest.mark.parametrize("dims", [(64, 64), (128, 64), (64, 128), (256, 96), (96, 256), (512, 130), (130, 512), (1024, 1024),
                                  (2048, 64), (64, 2048), (1024, 100), (100, 1024), (151, 100), (96, 64), (48, 32), (30, 7),
                                  (62, 64), (64, 62)])
def test_sh_dct_preconditioner_2d(bk, dims):
    L = (LX, LY)
    ctx = bk.Context(bk.BK_SH2D, dims, L, krylov_m=2, params=(-0.1, 1.3))
    r = np.random.default_rng(3).standard_normal(ctx.N)
    for shift in (1.0, 0.25):
        ctx.precond_setup(bk.BK_PC_SH_DCT, shift)
        ref = oprecond.dct_precond(dims, L, shift)(r)
        assert _rel(ctx.precond_apply(r), ref) < 1e-11
        assert _rel(ctx.precond_apply(ctx.to_device(r)).numpy(), ref) < 1e-11


@pytest.mark.parametrize("dims", [(64, 64, 64), (128, 64, 32), (64, 128, 66), (128, 128, 128), (48, 32, 16), (64, 30, 64), (30, 64, 64)])
def test_sh_dct_preconditioner_3d(bk, dims):
    L = (np.pi, 1.3 * np.pi, 0.7 * np.pi)
    ctx = bk.Context(bk.BK_SH3D, dims, L, krylov_m=2, params=(0.1, 1.2))
    r = np.random.default_rng(4).standard_normal(ctx.N)
    ctx.precond_setup(bk.BK_PC_SH_DCT, 1.0)
    ref = oprecond.dct_precond(dims, L, 1.0)(r)
    assert _rel(ctx.precond_apply(ctx.to_device(r)).numpy(), ref) < 1e-11


def test_unaligned_device_vectors_take_the_general_path(bk):
    """16-byte vector accesses of the fast kernels need aligned vectors; a device pointer offset by one double must still
    give the same answer (general kernel), not a misaligned-address fault."""
    dims, L = (256, 128), (LX, LY)
    ctx = bk.Context(bk.BK_SH2D, dims, L, krylov_m=2, params=(-0.1, 1.3))
    ctx.precond_setup(bk.BK_PC_SH_DCT, 1.0)
    r = np.random.default_rng(5).standard_normal(ctx.N)
    ref = oprecond.dct_precond(dims, L, 1.0)(r)
    big_in, big_out = ctx.zeros(ctx.N + 2), ctx.zeros(ctx.N + 2)
    host = np.concatenate([[0.0], r, [0.0]])
    st = ctx.lib.bk_vec_upload(ctx.handle, big_in.dptr, host.ctypes.data, len(host))
    assert st == 0
    st = ctx.lib.bk_precond_apply(ctx.handle, C.c_void_p(big_in.dptr + 8), C.c_void_p(big_out.dptr + 8))
    assert st == 0, ctx.lib.bk_last_error(ctx.handle)
    assert _rel(big_out.numpy()[1:-1], ref) < 1e-11
    assert _rel(ctx.precond_apply(ctx.to_device(r)).numpy(), ref) < 1e-11


@pytest.mark.parametrize("dims", [(41, 21), (64, 32), (100, 37), (512, 512)])
def test_cgl_dst_helmholtz_preconditioner_sizes(bk, dims):
    """DST-I by the mixed-radix kernel (odd extension of length 2 (n + 1): 84 = 4 * 3 * 7, 44 = 4 * 11, 130 = 2 * 5 * 13,
    1026 = 2 * 3^3 * 19, ...) against scipy's dstn on both components."""
    Nx, Ny = dims
    L = (0.5 * np.pi, np.pi)
    ctx = bk.Context(bk.BK_CGL2D, dims, L, krylov_m=2, params=(1.2, 0.1, 1.0, -1.0, 1.0))
    rng = np.random.default_rng(6)
    r = rng.standard_normal(ctx.N)
    for a0, a1 in ((1.0, -0.05), (2.5, -1.0)):
        ctx.precond_setup(bk.BK_PC_CGL_DST, a0, a1)
        P = oprecond.dst_helmholtz_precond(Nx, Ny, *L, a0, a1)
        n = Nx * Ny
        ref = np.concatenate([P(r[:n]), P(r[n:])])
        assert _rel(ctx.precond_apply(ctx.to_device(r)).numpy(), ref) < 1e-11
