"""Problem definitions P1-P4 exactly as the reference examples build them
(sparse Kronecker operators, fp64).  Test infrastructure only (see __init__).

Layout: Julia column-major, x fastest: flat index = i + j*Nx (+ k*Nx*Ny).
"""
import numpy as np
import scipy.sparse as sp

SQRT_EPS = float(np.sqrt(np.finfo(np.float64).eps))  # src/Problems.jl:69 (_getprecision)


# --------------------------------------------------------------------------- P1 chan
def chan_Nl(x, a=0.5, b=0.01):
    """examples/chan.jl:5"""
    return 1.0 + (x + a * x**2) / (1.0 + b * x**2)


def chan_dNl(x, a=0.5, b=0.01):
    """examples/chan.jl:6"""
    return (1.0 - b * x**2 + 2.0 * a * x) / (1.0 + b * x**2) ** 2


def chan_F(x, alpha, beta):
    """examples/chan.jl:8-19 F_chan"""
    n = len(x)
    f = np.empty_like(x)
    f[0] = x[0] - beta
    f[-1] = x[-1] - beta
    f[1:-1] = (x[:-2] - 2.0 * x[1:-1] + x[2:]) * (n - 1) ** 2 + alpha * chan_Nl(x[1:-1], b=beta)
    return f


def chan_dF(x, dx, alpha, beta):
    """examples/chan.jl:85-95 dF_chan"""
    n = len(x)
    out = np.empty_like(x)
    out[0] = dx[0]
    out[-1] = dx[-1]
    out[1:-1] = (dx[:-2] - 2.0 * dx[1:-1] + dx[2:]) * (n - 1) ** 2 + alpha * chan_dNl(x[1:-1], b=beta) * dx[1:-1]
    return out


def chan_sol0(n):
    """examples/chan.jl:23 (1-based i)"""
    i = np.arange(1, n + 1, dtype=np.float64)
    return (i - 1) * (n - i) / n**2 + 0.1


def chan_precond_matrix(n):
    """examples/chan.jl:108-109: tridiagonal Laplacian with identity boundary rows."""
    s = float((n - 1) ** 2)
    P = sp.diags([s * np.ones(n - 1), -2 * s * np.ones(n), s * np.ones(n - 1)], [-1, 0, 1], format="lil")
    P[0, 0] = 1.0
    P[0, 1] = 0.0
    P[n - 1, n - 2] = 0.0
    P[n - 1, n - 1] = 1.0
    return P.tocsc()


# --------------------------------------------------------------------------- Laplacians
def _d2(n, h, corner):
    d = -2.0 * np.ones(n)
    d[0] = corner
    d[-1] = corner
    return sp.diags([np.ones(n - 1), d, np.ones(n - 1)], [-1, 0, 1], format="csr") / h**2


def laplacian2d(Nx, Ny, lx, ly, bc="neumann"):
    """examples/SH2d-fronts.jl:13-29 (Neumann closure, corner diag -1/h^2) and
    examples/cGL2d.jl:6-22 (Dirichlet, diag -2/h^2 everywhere)."""
    hx = 2 * lx / Nx
    hy = 2 * ly / Ny
    c = -1.0 if bc == "neumann" else -2.0
    D2x = _d2(Nx, hx, c)
    D2y = _d2(Ny, hy, c)
    A = sp.kron(sp.identity(Ny), D2x) + sp.kron(D2y, sp.identity(Nx))
    return A.tocsr()


def laplacian3d(Nx, Ny, Nz, lx, ly, lz):
    """examples/SH3d.jl:16-41 (Neumann closure)."""
    hx, hy, hz = 2 * lx / Nx, 2 * ly / Ny, 2 * lz / Nz
    D2x, D2y, D2z = _d2(Nx, hx, -1.0), _d2(Ny, hy, -1.0), _d2(Nz, hz, -1.0)
    A2 = sp.kron(sp.identity(Ny), D2x) + sp.kron(D2y, sp.identity(Nx))
    A = sp.kron(sp.identity(Nz), A2) + sp.kron(sp.kron(D2z, sp.identity(Ny)), sp.identity(Nx))
    return A.tocsr()


# --------------------------------------------------------------------------- P2/P3 SH
class SwiftHohenberg:
    """F = -L1 u + l u + nu u^2 - u^3,  L1 = (I + Lap)^2.
    examples/SH2d-fronts.jl:31-34,55,124-127; examples/SH3d.jl:44-53,85."""

    def __init__(self, dims, lengths, l=-0.1, nu=1.3):
        self.dims = tuple(dims)
        self.lengths = tuple(lengths)
        if len(dims) == 2:
            lap = laplacian2d(dims[0], dims[1], lengths[0], lengths[1], "neumann")
        else:
            lap = laplacian3d(*dims, *lengths)
        n = lap.shape[0]
        IpL = (sp.identity(n) + lap).tocsr()
        self.L1 = (IpL @ IpL).tocsr()
        self.N = n
        self.l = l
        self.nu = nu

    def F(self, u, l=None):
        l = self.l if l is None else l
        return -(self.L1 @ u) + (l * u + self.nu * u**2 - u**3)

    def dF(self, u, du, l=None):
        l = self.l if l is None else l
        return -(self.L1 @ du) + (l + 2.0 * self.nu * u - 3.0 * u**2) * du

    def jac_sparse(self, u, l=None):
        l = self.l if l is None else l
        return (-self.L1 + sp.diags(l + 2.0 * self.nu * u - 3.0 * u**2)).tocsc()

    def grid(self):
        return [-L + 2 * L / n * np.arange(n) for n, L in zip(self.dims, self.lengths)]


def sh2d_sol0(Nx, Ny, lx, ly):
    """examples/SH2d-fronts.jl:44-51"""
    X = -lx + 2 * lx / Nx * np.arange(Nx)
    Y = -ly + 2 * ly / Ny * np.arange(Ny)
    s = np.cos(X)[None, :] + np.cos(X / 2)[None, :] * np.cos(np.sqrt(3.0) * Y / 2)[:, None]  # [j, i]
    s = s - s.min()
    s = s / s.max()
    s = s - 0.25
    s = s * 1.7
    return s.reshape(-1)


def sh2d_front_guess(u_hexa, Nx, Ny, lx, ly):
    """examples/SH2d-fronts.jl:75: 0.4 u_hexa exp(-(x+lx)^2/25)"""
    X = -lx + 2 * lx / Nx * np.arange(Nx)
    env = np.exp(-((X + lx) ** 2) / 25.0)
    return 0.4 * u_hexa * np.tile(env, Ny)


def sh3d_sol0(Nx, Ny, Nz, lx, ly, lz):
    """examples/SH3d.jl:77-80"""
    X = -lx + 2 * lx / Nx * np.arange(Nx)
    Y = -ly + 2 * ly / Ny * np.arange(Ny)
    s = np.cos(X)[None, None, :] * np.cos(Y)[None, :, None] * np.ones(Nz)[:, None, None]
    s = s - s.min()
    s = s / s.max()
    s = s * 1.2
    return s.reshape(-1)


# --------------------------------------------------------------------------- P4 cGL
class GinzburgLandau2D:
    """examples/cGL2d.jl:262-318 (NL!, dNL!, Fcgl!, dFcgl!), Dirichlet Laplacian :6-22,
    state [u1; u2] of length 2n."""

    def __init__(self, Nx, Ny, lx, ly, r=0.5, mu=0.1, nu=1.0, c3=-1.0, c5=1.0):
        self.Nx, self.Ny, self.lx, self.ly = Nx, Ny, lx, ly
        self.n = Nx * Ny
        self.N = 2 * self.n
        self.lap = laplacian2d(Nx, Ny, lx, ly, "dirichlet")
        self.Delta = sp.block_diag([self.lap, self.lap]).tocsr()
        self.r, self.mu, self.nu, self.c3, self.c5 = r, mu, nu, c3, c5

    def NL(self, u, r=None):
        r = self.r if r is None else r
        n = self.n
        u1, u2 = u[:n], u[n:]
        ua = u1**2 + u2**2
        f = np.empty_like(u)
        f[:n] = r * u1 - self.nu * u2 - ua * (self.c3 * u1 - self.mu * u2) - self.c5 * ua**2 * u1
        f[n:] = r * u2 + self.nu * u1 - ua * (self.c3 * u2 + self.mu * u1) - self.c5 * ua**2 * u2
        return f

    def dNL(self, u, du, r=None):
        r = self.r if r is None else r
        n = self.n
        u1, u2 = u[:n], u[n:]
        d1, d2 = du[:n], du[n:]
        mu, nu, c3, c5 = self.mu, self.nu, self.c3, self.c5
        f = np.empty_like(u)
        f[:n] = (-5 * c5 * u1**4 + (-6 * c5 * u2**2 - 3 * c3) * u1**2 + 2 * mu * u1 * u2 - c5 * u2**4 - c3 * u2**2 + r) * d1 + (
            -4 * c5 * u2 * u1**3 + mu * u1**2 + (-4 * c5 * u2**3 - 2 * c3 * u2) * u1 + 3 * u2**2 * mu - nu
        ) * d2
        f[n:] = (-4 * c5 * u2 * u1**3 - 3 * mu * u1**2 + (-4 * c5 * u2**3 - 2 * c3 * u2) * u1 - u2**2 * mu + nu) * d1 + (
            -c5 * u1**4 + (-6 * c5 * u2**2 - c3) * u1**2 - 2 * mu * u1 * u2 - 5 * c5 * u2**4 - 3 * c3 * u2**2 + r
        ) * d2
        return f

    def F(self, u, r=None):
        return self.NL(u, r) + self.Delta @ u

    def dF(self, u, du, r=None):
        return self.dNL(u, du, r) + self.Delta @ du

    def r_hopf(self):
        """Analytic Hopf point of the trivial state: r = -lambda_1(Delta)."""
        hx, hy = 2 * self.lx / self.Nx, 2 * self.ly / self.Ny
        lam = -(2 - 2 * np.cos(np.pi / (self.Nx + 1))) / hx**2 - (2 - 2 * np.cos(np.pi / (self.Ny + 1))) / hy**2
        return -lam

    def phi11(self):
        i = np.arange(1, self.Nx + 1)
        j = np.arange(1, self.Ny + 1)
        return (np.sin(np.pi * i / (self.Nx + 1))[None, :] * np.sin(np.pi * j / (self.Ny + 1))[:, None]).reshape(-1)
