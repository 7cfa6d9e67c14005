"""TEST / BENCH HARNESS (not product): host-side 1-D background, the input-table generator.

Restates src/Bolt.jl:56-66 (CosmoParams) and src/background.jl:5-128.  All quantities are in
the reference's Mpc units.  x = ln(a).
"""
from dataclasses import dataclass, replace, fields
import numpy as np

from . import constants as C
from .bspline import CubicBSpline, spline_dx, spline_dx2


from bolt_b200.params import (CosmoParams, H0, rho_crit, T_nu, f0, dlnf0dlnq, to_ui, from_ui, dxdq, xq2q, q_grid)   # noqa: F401


def Omega_Lambda(par):                         # background.jl:7-18
    Tγ = (15.0 / np.pi ** 2 * rho_crit(par) * par.Ω_r) ** 0.25
    νfac = (90.0 * C.ZETA3 / (11.0 * np.pi ** 4)) * (par.Ω_r * par.h ** 2 / Tγ) * ((par.N_ν / 3.0) ** 0.75)
    Ω_ν = par.Σm_ν * νfac / par.h ** 2
    return 1.0 - (par.Ω_r * (1.0 + (2.0 / 3.0) * (7.0 * par.N_ν / 8.0) * (4.0 / 11.0) ** (4.0 / 3.0))
                  + par.Ω_b + par.Ω_c + Ω_ν)


def rhoP_0(a, par, quad_pts, quad_wts):        # background.jl:33-48
    a = np.asarray(a, dtype=np.float64)
    q, lqmi, lqma = q_grid(par, quad_pts)
    m = par.Σm_ν
    aa = a[..., None]
    eps = np.sqrt(q ** 2 + (aa * m) ** 2)
    w = f0(q, par) / dxdq(q, lqmi, lqma) * quad_wts
    Irho = q ** 2 * eps * w
    IP = q ** 2 * (q ** 2 / eps) * w
    rho = 4.0 * np.pi * a ** (-4.0) * Irho.sum(-1)
    P = 4.0 * np.pi / 3.0 * a ** (-4.0) * IP.sum(-1)
    return rho, P


def H_a(a, par, quad_pts, quad_wts):           # background.jl:58-64
    a = np.asarray(a, dtype=np.float64)
    rho_nu, _ = rhoP_0(a, par, quad_pts, quad_wts)
    return H0(par) * np.sqrt((par.Ω_c + par.Ω_b) * a ** (-3.0)
                             + rho_nu / rho_crit(par)
                             + par.Ω_r * a ** (-4.0) * (1.0 + (2.0 / 3.0) * (7.0 * par.N_ν / 8.0) * (4.0 / 11.0) ** (4.0 / 3.0))
                             + Omega_Lambda(par))


def calH_a(a, par, quad_pts, quad_wts):        # background.jl:66
    return a * H_a(a, par, quad_pts, quad_wts)


def eta(x, par, quad_pts, quad_wts):           # background.jl:73-77
    x = np.asarray(x, dtype=np.float64)
    logamin = -13.75
    logamax = np.log10(np.exp(x))[..., None]
    ap = xq2q(quad_pts, logamin, logamax)
    I = 1.0 / (ap * calH_a(ap, par, quad_pts, quad_wts)) / dxdq(ap, logamin, logamax)
    return (I * quad_wts).sum(-1)


def make_x_grid(x0=-20.0, dx=0.01, n=2001):
    """Julia's -20.0:0.01:0.0 is a twice-precision range whose elements are the correctly
    rounded decimals; reproduce that."""
    return np.round(x0 + dx * np.arange(n), 10)


class Background:
    """background.jl:85-128."""

    def __init__(self, par, x_grid=None, nq=15):
        if x_grid is None:
            x_grid = make_x_grid()
        self.par = par
        self.x_grid = np.asarray(x_grid, dtype=np.float64)
        self.x0 = float(self.x_grid[0])
        self.dx = float((self.x_grid[-1] - self.x_grid[0]) / (len(self.x_grid) - 1))
        self.nq = nq
        pts, wts = np.polynomial.legendre.leggauss(nq)   # FastGaussQuadrature.gausslegendre(nq)
        self.quad_pts, self.quad_wts = pts, wts
        self.H0 = H0(par)
        self.η0 = float(eta(0.0, par, pts, wts))
        self.ρ_crit = rho_crit(par)
        self.Ω_Λ = Omega_Lambda(par)
        a = np.exp(self.x_grid)
        self.ρ0M = CubicBSpline(rhoP_0(a, par, pts, wts)[0], self.x0, self.dx)
        self.H = CubicBSpline(calH_a(a, par, pts, wts), self.x0, self.dx)        # conformal ℋ(x)
        self.η = CubicBSpline(eta(self.x_grid, par, pts, wts), self.x0, self.dx)
        self.Hp = spline_dx(self.H, self.x_grid)
        self.Hpp = spline_dx2(self.H, self.x_grid)
        self.ηp = spline_dx(self.η, self.x_grid)
        self.ηpp = spline_dx2(self.η, self.x_grid)
