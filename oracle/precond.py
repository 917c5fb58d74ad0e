"""Preconditioners.  Test infrastructure only (see oracle/__init__.py).

The reference preconditions SH with a sparse factorisation of L1 + I
(examples/SH2d-fronts.jl:120-122 ``Pl = lu(par.L1 + I)``; examples/SH3d.jl:88
``cholesky(L1)``).  The Neumann-closure Laplacian of examples/SH2d-fronts.jl:13-29 is
diagonalised by the orthonormal DCT-II, so (L1 + shift I)^-1 is applied exactly by
DCT -> divide by ((1 + lx_i + ly_j (+ lz_k))^2 + shift) -> inverse DCT.
``sparse_lu_precond`` is the literal restatement; ``dct_precond`` the fast form;
tests check that they agree to rounding.
"""
import numpy as np
import scipy.fft as sfft
import scipy.sparse as sp
import scipy.sparse.linalg as spl


def neumann_eigs(n, h):
    return (2.0 * np.cos(np.pi * np.arange(n) / n) - 2.0) / h**2


def sh_symbol(dims, lengths, shift=1.0):
    """((1 + sum_d lambda_d)^2 + shift) with array shape dims[::-1] (x fastest)."""
    lam = 1.0
    nd = len(dims)
    for d, (n, L) in enumerate(zip(dims, lengths)):
        e = neumann_eigs(n, 2 * L / n)
        shape = [1] * nd
        shape[nd - 1 - d] = n
        lam = lam + e.reshape(shape)
    return lam**2 + shift


def dct_precond(dims, lengths, shift=1.0, workers=1):
    """Returns r -> (L1 + shift I)^-1 r."""
    sym = sh_symbol(dims, lengths, shift)
    shape = tuple(dims[::-1])

    def apply(r):
        R = sfft.dctn(r.reshape(shape), type=2, norm="ortho", workers=workers)
        R /= sym
        return sfft.idctn(R, type=2, norm="ortho", workers=workers).reshape(-1)

    return apply


def sparse_lu_precond(L1, shift=1.0):
    lu = spl.splu((L1 + shift * sp.identity(L1.shape[0])).tocsc())
    return lu.solve


def chan_lu_precond(n):
    """examples/chan.jl:108-111  Pl = lu(P)."""
    from .problems import chan_precond_matrix
    lu = spl.splu(chan_precond_matrix(n))
    return lu.solve


def dirichlet_eigs(n, h):
    return -(2.0 - 2.0 * np.cos(np.pi * np.arange(1, n + 1) / (n + 1))) / h**2


def dst_helmholtz_precond(Nx, Ny, lx, ly, a0, a1, workers=1):
    """r -> (a0 I + a1 Lap_dirichlet)^-1 r on one n=Nx*Ny component via DST-I
    (the Dirichlet Laplacian of examples/cGL2d.jl:6-22 is diagonalised by DST-I)."""
    ex = dirichlet_eigs(Nx, 2 * lx / Nx)
    ey = dirichlet_eigs(Ny, 2 * ly / Ny)
    sym = a0 + a1 * (ex[None, :] + ey[:, None])

    def apply(r):
        R = sfft.dstn(r.reshape(Ny, Nx), type=1, norm="ortho", workers=workers)
        R /= sym
        return sfft.idstn(R, type=1, norm="ortho", workers=workers).reshape(-1)

    return apply


def potrap_circulant_precond(Nx, Ny, lx, ly, M, T, r, nu, workers=1):
    """Preconditioner for the Trapeze periodic-orbit Jacobian of cGL (stand-in for the ILU of the assembled sparse PO
    Jacobian, examples/cGL2d.jl:209-213): the PO Jacobian linearised at the trivial state,
        rows i = 1..M-1:  A x_i - B x_{i-1},  A = I - h/2 J0, B = I + h/2 J0, x_0 == x_{M-1};   row M: x_M - x_1,
    J0 = Lap_dirichlet + r + nu*R, is block-circulant in time and diagonal in the DST-I basis, so it is inverted exactly by
    DST-I in space, the change of variables u1 +- i u2 (diagonalises R), a length-(M-1) DFT in time and a scalar division.
    Identity on the period unknown (and on a PALC border entry if present)."""
    n, Ns, K = Nx * Ny, 2 * Nx * Ny, M - 1
    hx, hy = 2 * lx / Nx, 2 * ly / Ny
    lam = dirichlet_eigs(Nx, hx)[None, :] + dirichlet_eigs(Ny, hy)[:, None]
    h = T / M
    gam = np.exp(-2j * np.pi * np.arange(K) / K)
    sp = (1 - gam)[:, None, None] - (h / 2) * (1 + gam)[:, None, None] * (lam[None] + r + 1j * nu)
    sm = (1 - gam)[:, None, None] - (h / 2) * (1 + gam)[:, None, None] * (lam[None] + r - 1j * nu)
    NM = Ns * M

    def apply(v):
        out = np.array(v, dtype=float)
        X = v[:NM].reshape(M, 2, Ny, Nx)
        W = sfft.dstn(X[:K], type=1, norm="ortho", axes=(-2, -1), workers=workers)
        wp, wm = W[:, 0] + 1j * W[:, 1], W[:, 0] - 1j * W[:, 1]
        yp = np.fft.ifft(np.fft.fft(wp, axis=0) / sp, axis=0)
        ym = np.fft.ifft(np.fft.fft(wm, axis=0) / sm, axis=0)
        Y = np.empty_like(W)
        Y[:, 0] = ((yp + ym) / 2).real
        Y[:, 1] = ((yp - ym) / 2j).real
        Z = sfft.idstn(Y, type=1, norm="ortho", axes=(-2, -1), workers=workers)
        O = out[:NM].reshape(M, 2, Ny, Nx)
        O[:K] = Z
        O[M - 1] = X[M - 1] + Z[0]
        return out  # entries beyond NM (period, border) pass through

    return apply
