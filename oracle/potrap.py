"""Trapezoid periodic-orbit functional (P5).  Test infrastructure only.

Unknown x = [x_1; ...; x_M; T] (slices contiguous, src/periodicorbit/PeriodicOrbitTrapeze.jl:141),
uniform mesh step 1/M (src/TimeMesh.jl:20-21).

po_residual  : PeriodicOrbitTrapeze.jl:249-287 (+ potrap_scheme! :209-242)
po_jvp       : PeriodicOrbitTrapeze.jl:294-330 (+ Jc :362-386)
functional_ref / dfunctional_ref: the *independent* restatement used by the reference's own
test (test/periodic_orbits_function_fd/test_potrap.jl:90-157), for cross-checking.
"""
import numpy as np


class Trapeze:
    def __init__(self, F, dF, phi, xpi, M, N):
        """F(u) -> vector field; dF(u, du) -> J(u) du  (parameters bound by the caller)."""
        self.F, self.dF, self.phi, self.xpi, self.M, self.N = F, dF, phi, xpi, M, N

    def slices(self, x):
        return x[:-1].reshape(self.M, self.N)

    def residual(self, x):
        M, N = self.M, self.N
        T = x[-1]
        u = self.slices(x)
        out = np.empty_like(x)
        o = out[:-1].reshape(M, N)
        h = T / M
        # rows 1..M-1 (0-based i=0..M-2): (u_i - u_{i-1}) - h/2 (F(u_i) + F(u_{i-1})), u_{-1} = u_{M-2}
        Fprev = self.F(u[M - 2])
        for i in range(M - 1):
            Fi = self.F(u[i])
            prev = u[i - 1] if i > 0 else u[M - 2]
            o[i] = (u[i] - prev) - (h / 2) * (Fi + Fprev)
            Fprev = Fi
        o[M - 1] = u[M - 1] - u[0]
        out[-1] = np.dot(x[:-1], self.phi) - np.dot(self.xpi, self.phi)
        return out

    def jvp(self, x, dx):
        M, N = self.M, self.N
        T, dT = x[-1], dx[-1]
        u, du = self.slices(x), self.slices(dx)
        out = np.empty_like(x)
        o = out[:-1].reshape(M, N)
        h, dh = T / M, dT / M
        Jprev = self.dF(u[M - 2], du[M - 2])
        Fprev = self.F(u[M - 2])
        for i in range(M - 1):
            Ji = self.dF(u[i], du[i])
            Fi = self.F(u[i])
            dprev = du[i - 1] if i > 0 else du[M - 2]
            o[i] = (du[i] - dprev) - (h / 2) * (Ji + Jprev) - (dh / 2) * (Fi + Fprev)
            Jprev, Fprev = Ji, Fi
        o[M - 1] = du[M - 1] - du[0]
        out[-1] = np.dot(dx[:-1], self.phi)
        return out


def functional_ref(F, x, M, N, phi, xpi, mesh=None):
    """test/periodic_orbits_function_fd/test_potrap.jl:90-120 `_functional` (mesh = time steps,
    length M-1... the test passes dt_i; uniform = 1/M)."""
    T = x[-1]
    u = x[:-1].reshape(M, N)
    out = np.empty_like(x)
    o = out[:-1].reshape(M, N)
    dts = np.full(M, 1.0 / M) if mesh is None else np.asarray(mesh)
    for i in range(1, M - 1):
        h = T * dts[i]
        o[i] = (u[i] - u[i - 1]) - h / 2 * (F(u[i]) + F(u[i - 1]))
    h = T * dts[0]
    o[0] = (u[0] - u[M - 2]) - h / 2 * (F(u[0]) + F(u[M - 2]))
    o[M - 1] = u[M - 1] - u[0]
    out[-1] = np.dot(x[:-1], phi) - np.dot(xpi, phi)
    return out


def dfunctional_ref(F, dF, x, dx, M, N, phi):
    """test/periodic_orbits_function_fd/test_potrap.jl:113-150 `_dfunctional` (uniform mesh)."""
    T, dT = x[-1], dx[-1]
    u, du = x[:-1].reshape(M, N), dx[:-1].reshape(M, N)
    out = np.empty_like(x)
    o = out[:-1].reshape(M, N)
    h, dh = T / M, dT / M
    o[0] = (du[0] - du[M - 2]) - h / 2 * (dF(u[0], du[0]) + dF(u[M - 2], du[M - 2]))
    for i in range(1, M - 1):
        o[i] = (du[i] - du[i - 1]) - h / 2 * (dF(u[i], du[i]) + dF(u[i - 1], du[i - 1]))
    o[0] -= dh / 2 * (F(u[0]) + F(u[M - 2]))
    for i in range(1, M - 1):
        o[i] -= dh / 2 * (F(u[i]) + F(u[i - 1]))
    o[M - 1] = du[M - 1] - du[0]
    out[-1] = np.dot(dx[:-1], phi)
    return out
