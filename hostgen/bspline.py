"""Cubic B-spline on a uniform grid, restating what the reference gets from
Interpolations.jl via `spline(f, x_grid) = scale(interpolate(f, BSpline(Cubic(Line(OnGrid())))), x_grid)`
(src/util.jl:11-13).

Coefficients are padded by one on each side (n+2 for n samples).  Interior rows are
(1/6, 2/3, 1/6); the `Line(OnGrid())` boundary rows say the second difference of the
coefficients vanishes at the first/last sample, which makes c[1] = y[0], c[n] = y[n-1].
Evaluation uses the four-tap cubic basis; gradient/hessian are the analytic derivatives of
the same basis (used by `spline_dx`, `spline_dx2`, util.jl:12-13, sampled at the knots).
"""
import numpy as np
from scipy.linalg import solve_banded


def prefilter(y):
    y = np.asarray(y, dtype=np.float64)
    n = y.shape[0]
    c = np.empty((n + 2,) + y.shape[1:], dtype=np.float64)
    c[1] = y[0]
    c[n] = y[n - 1]
    m = n - 2
    if m > 0:
        ab = np.zeros((3, m))
        ab[0, 1:] = 1.0 / 6.0
        ab[1, :] = 2.0 / 3.0
        ab[2, :-1] = 1.0 / 6.0
        rhs = np.array(y[1:n - 1], dtype=np.float64, copy=True)
        rhs[0] = rhs[0] - y[0] / 6.0
        rhs[-1] = rhs[-1] - y[n - 1] / 6.0
        c[2:n] = solve_banded((1, 1), ab, rhs)
    c[0] = 2.0 * c[1] - c[2]
    c[n + 1] = 2.0 * c[n] - c[n - 1]
    return c


class CubicBSpline:
    """y sampled at x0 + i*dx, i = 0..n-1."""

    def __init__(self, y, x0, dx, coefs=None):
        self.x0 = float(x0)
        self.dx = float(dx)
        self.coefs = prefilter(y) if coefs is None else np.asarray(coefs, dtype=np.float64)
        self.n = self.coefs.shape[0] - 2

    def _locate(self, x):
        t = (np.asarray(x, dtype=np.float64) - self.x0) / self.dx
        i = np.floor(t).astype(np.int64)
        i = np.clip(i, 0, self.n - 2)
        d = t - i
        return i, d

    def __call__(self, x):
        i, d = self._locate(x)
        c = self.coefs
        e = 1.0 - d
        w0 = e * e * e / 6.0
        w1 = 2.0 / 3.0 - d * d + d * d * d / 2.0
        w2 = 2.0 / 3.0 - e * e + e * e * e / 2.0
        w3 = d * d * d / 6.0
        return c[i] * w0 + c[i + 1] * w1 + c[i + 2] * w2 + c[i + 3] * w3

    def gradient(self, x):
        i, d = self._locate(x)
        c = self.coefs
        e = 1.0 - d
        w0 = -e * e / 2.0
        w1 = -2.0 * d + 1.5 * d * d
        w2 = 2.0 * e - 1.5 * e * e
        w3 = d * d / 2.0
        return (c[i] * w0 + c[i + 1] * w1 + c[i + 2] * w2 + c[i + 3] * w3) / self.dx

    def hessian(self, x):
        i, d = self._locate(x)
        c = self.coefs
        e = 1.0 - d
        w0 = e
        w1 = -2.0 + 3.0 * d
        w2 = -2.0 + 3.0 * e
        w3 = d
        return (c[i] * w0 + c[i + 1] * w1 + c[i + 2] * w2 + c[i + 3] * w3) / self.dx**2

    def knots(self):
        return self.x0 + self.dx * np.arange(self.n)


def spline(y, x0, dx):
    return CubicBSpline(y, x0, dx)


def spline_dx(f, x_grid):
    """util.jl:12 — spline of the parent spline's analytic gradient sampled on the grid."""
    return CubicBSpline(f.gradient(x_grid), f.x0, f.dx)


def spline_dx2(f, x_grid):
    """util.jl:13 — spline of the parent spline's analytic hessian sampled on the grid."""
    return CubicBSpline(f.hessian(x_grid), f.x0, f.dx)
