"""Host-side ionization history: RECFAST + tanh reionization + optical depth / visibility tables.

Restates src/ionization/recfast.jl:22-121 (constants), :156-310 (ion_recfast), :313-323,
:346-475 (piecewise Saha / evolution and recfastsolve), :478-536 (tanh reionization),
:674-726 (IonizationHistory) and src/ionization/ionization.jl:107-137 (τ, τ′, g̃).

The reference integrates with OrdinaryDiffEq's Tsit5 (reltol = tol) and finds the switch
redshifts with a Falsi root-find; those packages are not vendored, so this module uses
scipy's explicit RK (DOP853) at tighter tolerances and Brent's method -- equal to the
reference up to its own integration tolerance.  Pinned by the reference's Fortran-RECFAST
golden file (test/runtests.jl:38-48, 1e-4 absolute on Xe).
"""
import math
import numpy as np
from scipy.integrate import solve_ivp
from scipy.optimize import brentq

from . import constants as K
from .bspline import CubicBSpline, spline_dx, spline_dx2


class RECFAST:
    def __init__(self, bg, Yp=0.24, OmegaB=0.046, OmegaG=5.0469e-5, Tnow=None, tol=1e-8,
                 Hswitch=1, Heswitch=6):
        s = self
        s.bg = bg
        s.C = 2.99792458e8; s.k_B = 1.380658e-23; s.h_P = 6.6260755e-34
        s.m_e = 9.1093897e-31; s.m_H = 1.673575e-27; s.not4 = 3.9715e0
        s.sigma = 6.6524616e-29; s.a = 7.565914e-16; s.G = 6.6742e-11
        s.Lambda = 8.2245809e0; s.Lambda_He = 51.3e0
        s.L_H_ion = 1.096787737e7; s.L_H_alpha = 8.225916453e6
        s.L_He1_ion = 1.98310772e7; s.L_He2_ion = 4.389088863e7
        s.L_He_2s = 1.66277434e7; s.L_He_2p = 1.71134891e7
        s.A2P_s = 1.798287e9; s.A2P_t = 177.58e0
        s.L_He_2Pt = 1.690871466e7; s.L_He_2St = 1.5985597526e7; s.L_He2St_ion = 3.8454693845e6
        s.sigma_He_2Ps = 1.436289e-22; s.sigma_He_2Pt = 1.484872e-22
        s.AGauss1 = -0.14e0; s.AGauss2 = 0.079e0; s.zGauss1 = 7.28e0; s.zGauss2 = 6.73e0
        s.wGauss1 = 0.18e0; s.wGauss2 = 0.33e0
        s.a_PPB = 4.309; s.b_PPB = -0.6166; s.c_PPB = 0.6703; s.d_PPB = 0.5300
        s.a_VF = 10 ** (-16.744); s.b_VF = 0.711; s.T_0 = 10 ** 0.477121; s.T_1 = 10 ** 5.114
        s.a_trip = 10 ** (-16.306); s.b_trip = 0.761
        s.Lalpha = 1 / s.L_H_alpha; s.Lalpha_He = 1 / s.L_He_2p
        s.DeltaB = s.h_P * s.C * (s.L_H_ion - s.L_H_alpha); s.CDB = s.DeltaB / s.k_B
        s.DeltaB_He = s.h_P * s.C * (s.L_He1_ion - s.L_He_2s); s.CDB_He = s.DeltaB_He / s.k_B
        s.CB1 = s.h_P * s.C * s.L_H_ion / s.k_B
        s.CB1_He1 = s.h_P * s.C * s.L_He1_ion / s.k_B
        s.CB1_He2 = s.h_P * s.C * s.L_He2_ion / s.k_B
        s.CR = 2 * math.pi * (s.m_e / s.h_P) * (s.k_B / s.h_P)
        s.CK = s.Lalpha ** 3 / (8 * math.pi); s.CK_He = s.Lalpha_He ** 3 / (8 * math.pi)
        s.CL = s.C * s.h_P / (s.k_B * s.Lalpha); s.CL_He = s.C * s.h_P / (s.k_B / s.L_He_2s)
        s.CT = (8 / 3) * (s.sigma / (s.m_e * s.C)) * s.a
        s.Bfact = s.h_P * s.C * (s.L_He_2p - s.L_He_2s) / s.k_B
        s.H_frac = 1e-3
        s.Hswitch = Hswitch; s.Heswitch = Heswitch
        s.Yp = Yp; s.OmegaB = OmegaB; s.OmegaG = OmegaG
        s.HO = bg.H0 / K.H0_natural_unit_conversion
        s.Tnow = ((15 / math.pi ** 2 * bg.ρ_crit * OmegaG) ** 0.25 * K.Kelvin_natural_unit_conversion
                  if Tnow is None else Tnow)
        s.mu_H = 1 / (1 - Yp); s.mu_T = s.not4 / (s.not4 - (s.not4 - 1) * Yp)
        s.fHe = Yp / (s.not4 * (1 - Yp))
        s.Nnow = 3 * s.HO * s.HO * OmegaB / (8 * math.pi * s.G * s.mu_H * s.m_H)
        s.fu = 1.14 if Hswitch == 0 else 1.125
        s.b_He = 0.86
        s.tol = tol
        # plain-float copies of the ℋ, ℋ′ spline for fast scalar evaluation inside the ODE RHS
        s._cH = bg.H.coefs; s._cHp = bg.Hp.coefs; s._x0 = bg.x0; s._dx = bg.dx; s._n = bg.H.n

    def _spl(self, c, x):
        t = (x - self._x0) / self._dx
        i = int(math.floor(t))
        i = min(max(i, 0), self._n - 2)
        d = t - i; e = 1.0 - d
        return (c[i] * e * e * e / 6.0 + c[i + 1] * (2.0 / 3.0 - d * d + d * d * d / 2.0)
                + c[i + 2] * (2.0 / 3.0 - e * e + e * e * e / 2.0) + c[i + 3] * d * d * d / 6.0)

    def Hz_dHdz(self, z):
        a = 1.0 / (1.0 + z)
        x_a = math.log(a)
        Hc = self._spl(self._cH, x_a)
        Hz = Hc / a / K.H0_natural_unit_conversion
        dHdz = (-self._spl(self._cHp, x_a) + Hc) / K.H0_natural_unit_conversion
        return Hz, dHdz


def ion_recfast(y, r, z):
    """recfast.jl:156-310."""
    exp, sqrt, log, pi = math.exp, math.sqrt, math.log, math.pi
    x_H, x_He, Tmat = y[0], y[1], y[2]
    if not (Tmat > 0.0) or not (x_H + r.fHe * x_He > 0.0):
        # an over-long trial step of the explicit integrator: report NaN so the step is rejected
        return (math.nan, math.nan, math.nan)
    x = x_H + r.fHe * x_He
    n = r.Nnow * (1 + z) ** 3
    n_He = r.fHe * r.Nnow * (1 + z) ** 3
    Trad = r.Tnow * (1 + z)
    Hz, dHdz = r.Hz_dHdz(z)

    Rdown = 1e-19 * r.a_PPB * (Tmat / 1e4) ** r.b_PPB / (1. + r.c_PPB * (Tmat / 1e4) ** r.d_PPB)
    Rup = Rdown * (r.CR * Tmat) ** 1.5 * exp(-r.CDB / Tmat)

    sq_0 = sqrt(Tmat / r.T_0)
    sq_1 = sqrt(Tmat / r.T_1)
    Rdown_He = r.a_VF / (sq_0 * (1 + sq_0) ** (1 - r.b_VF))
    Rdown_He = Rdown_He / (1 + sq_1) ** (1 + r.b_VF)
    Rup_He = Rdown_He * (r.CR * Tmat) ** 1.5 * exp(-r.CDB_He / Tmat)
    Rup_He = 4. * Rup_He
    if (r.Bfact / Tmat) > 680.:
        He_Boltz = exp(680.)
    else:
        He_Boltz = exp(r.Bfact / Tmat)

    if r.Hswitch == 0:
        Kc = r.CK / Hz
    else:
        Kc = r.CK / Hz * (1.0
                          + r.AGauss1 * exp(-((log(1 + z) - r.zGauss1) / r.wGauss1) ** 2)
                          + r.AGauss2 * exp(-((log(1 + z) - r.zGauss2) / r.wGauss2) ** 2))

    Rdown_trip = r.a_trip / (sq_0 * (1 + sq_0) ** (1 - r.b_trip))
    Rdown_trip = Rdown_trip / ((1 + sq_1) ** (1 + r.b_trip))
    Rup_trip = Rdown_trip * exp(-r.h_P * r.C * r.L_He2St_ion / (r.k_B * Tmat))
    Rup_trip = Rup_trip * ((r.CR * Tmat) ** 1.5) * (4 / 3)

    if (x_He < 5.e-9) or (x_He > 0.980):
        Heflag = 0
    else:
        Heflag = r.Heswitch
    CfHe_t = 0.0
    if Heflag == 0:
        K_He = r.CK_He / Hz
    else:
        tauHe_s = r.A2P_s * r.CK_He * 3 * n_He * (1 - x_He) / Hz
        pHe_s = (1 - exp(-tauHe_s)) / tauHe_s
        K_He = 1 / (r.A2P_s * pHe_s * 3 * n_He * (1 - x_He))
        if ((Heflag == 2) or (Heflag >= 5)) and (x_H < 0.9999999):
            Doppler = 2 * r.k_B * Tmat / (r.m_H * r.not4 * r.C * r.C)
            Doppler = r.C * r.L_He_2p * sqrt(Doppler)
            gamma_2Ps = 3 * r.A2P_s * r.fHe * (1 - x_He) * r.C * r.C / (
                sqrt(pi) * r.sigma_He_2Ps * 8 * pi * Doppler * (1 - x_H)) / ((r.C * r.L_He_2p) ** 2)
            pb = 0.36
            qb = r.b_He
            AHcon = r.A2P_s / (1 + pb * (gamma_2Ps ** qb))
            K_He = 1 / ((r.A2P_s * pHe_s + AHcon) * 3 * n_He * (1 - x_He))
        if Heflag >= 3:
            tauHe_t = r.A2P_t * n_He * (1. - x_He) * 3
            tauHe_t = tauHe_t / (8 * pi * Hz * r.L_He_2Pt ** 3)
            pHe_t = (1 - exp(-tauHe_t)) / tauHe_t
            CL_PSt = r.h_P * r.C * (r.L_He_2Pt - r.L_He_2St) / r.k_B
            if (Heflag == 3) or (Heflag == 5) or (x_H > 0.99999):
                CfHe_t = r.A2P_t * pHe_t * exp(-CL_PSt / Tmat)
                CfHe_t = CfHe_t / (Rup_trip + CfHe_t)
            else:
                Doppler = 2 * r.k_B * Tmat / (r.m_H * r.not4 * r.C * r.C)
                Doppler = r.C * r.L_He_2Pt * sqrt(Doppler)
                gamma_2Pt = (3 * r.A2P_t * r.fHe * (1 - x_He) * r.C * r.C
                             / (sqrt(pi) * r.sigma_He_2Pt * 8 * pi * Doppler * (1 - x_H))
                             / ((r.C * r.L_He_2Pt) ** 2))
                pb = 0.66
                qb = 0.9
                AHcon = r.A2P_t / (1 + pb * gamma_2Pt ** qb) / 3
                CfHe_t = (r.A2P_t * pHe_t + AHcon) * exp(-CL_PSt / Tmat)
                CfHe_t = CfHe_t / (Rup_trip + CfHe_t)

    timeTh = (1 / (r.CT * Trad ** 4)) * (1 + x + r.fHe) / x
    timeH = 2 / (3 * r.HO * (1 + z) ** 1.5)

    if x_H > 0.99:
        f1 = 0.
    elif x_H > 0.985:
        f1 = (x * x_H * n * Rdown - Rup * (1 - x_H) * exp(-r.CL / Tmat)) / (Hz * (1 + z))
    else:
        f1 = (((x * x_H * n * Rdown - Rup * (1.0 - x_H) * exp(-r.CL / Tmat))
               * (1.0 + Kc * r.Lambda * n * (1.0 - x_H)))
              / (Hz * (1.0 + z) * (1.0 / r.fu + Kc * r.Lambda * n * (1.0 - x_H) / r.fu
                                   + Kc * Rup * n * (1.0 - x_H))))
    if x_He < 1e-15:
        f2 = 0.
    else:
        f2 = (((x * x_He * n * Rdown_He - Rup_He * (1 - x_He) * exp(-r.CL_He / Tmat))
               * (1 + K_He * r.Lambda_He * n_He * (1 - x_He) * He_Boltz))
              / (Hz * (1 + z)
                 * (1 + K_He * (r.Lambda_He + Rup_He) * n_He * (1 - x_He) * He_Boltz)))
        if Heflag >= 3:
            f2 = f2 + (x * x_He * n * Rdown_trip
                       - (1 - x_He) * 3 * Rup_trip * exp(-r.h_P * r.C * r.L_He_2St / (r.k_B * Tmat))
                       ) * CfHe_t / (Hz * (1 + z))

    if timeTh < r.H_frac * timeH:
        epsilon = Hz * (1 + x + r.fHe) / (r.CT * Trad ** 3 * x)
        f3 = r.Tnow + epsilon * ((1 + r.fHe) / (1 + r.fHe + x)) * (
            (f1 + r.fHe * f2) / x) - epsilon * dHdz / Hz + 3 * epsilon / (1 + z)
    else:
        f3 = r.CT * (Trad ** 4) * x / (1 + x + r.fHe) * (Tmat - Trad) / (Hz * (1 + z)) + 2 * Tmat / (1 + z)
    return (f1, f2, f3)


def _saha_rhs(r, z, CB):
    return math.exp(1.5 * math.log(r.CR * r.Tnow / (1 + z)) - CB / (r.Tnow * (1 + z))) / r.Nnow


def x_H0_H_Saha(r, z):                          # recfast.jl:380-384
    rhs = _saha_rhs(r, z, r.CB1)
    return 0.5 * (math.sqrt(rhs ** 2 + 4 * rhs) - rhs)


def _x_He_saha(r, z):
    rhs = 4 * _saha_rhs(r, z, r.CB1_He1)
    return 0.5 * (math.sqrt((rhs - 1) ** 2 + 4 * (1 + r.fHe) * rhs) - (rhs - 1))


def end_of_saha_condition(z, r):                # recfast.jl:350-357
    return (_x_He_saha(r, z) - 1) / r.fHe - 0.99


def end_He_evo_condition(z, r):                 # recfast.jl:386
    return x_H0_H_Saha(r, z) - 0.985


class RecfastHistory:
    """recfast.jl:446-475 (recfastsolve) plus the piecewise accessors :392-425."""

    def __init__(self, r, zinitial=10000., zfinal=0.):
        self.r = r
        self.zinitial, self.zfinal = zinitial, zfinal
        z_begin = min(zinitial, 3500.)
        self.z_He_evo_start = brentq(end_of_saha_condition, zfinal, z_begin, args=(r,), xtol=1e-10, rtol=1e-12)
        self.z_H_He_evo_start = brentq(end_He_evo_condition, zfinal, self.z_He_evo_start, args=(r,),
                                       xtol=1e-10, rtol=1e-12)
        rtol, atol = min(r.tol, 1e-8), 1e-12

        def rhs_He(z, u):                        # ion_recfast_H_Saha, recfast.jl:360-366
            du = ion_recfast((x_H0_H_Saha(r, z), u[0], u[1]), r, z)
            return (du[1], du[2])

        z0 = self.z_He_evo_start
        y2 = ((_x_He_saha(r, z0) - 1) / r.fHe, r.Tnow * (1 + z0))     # init_He_evolution :368-376
        self.sol_He = solve_ivp(rhs_He, (z0, self.z_H_He_evo_start), y2, method="DOP853",
                                rtol=rtol, atol=atol, dense_output=True).sol
        z3 = self.z_H_He_evo_start
        u3 = self.sol_He(z3)
        y3 = (x_H0_H_Saha(r, z3), u3[0], u3[1])
        self.sol_H_He = solve_ivp(lambda z, y: ion_recfast(y, r, z), (z3, zfinal), y3, method="DOP853",
                                  rtol=rtol, atol=atol, dense_output=True).sol

    def Xe(self, z):                             # Xe_RECFAST :392-412
        r = self.r
        if z > 8000.:
            return 1 + 2 * r.fHe
        elif z > 5000.:
            rhs = _saha_rhs(r, z, r.CB1_He2)
            return 0.5 * (math.sqrt((rhs - 1 - r.fHe) ** 2 + 4 * (1 + 2 * r.fHe) * rhs) - (rhs - 1 - r.fHe))
        elif z > 3500.:
            return 1 + r.fHe
        elif z > self.z_He_evo_start:
            return _x_He_saha(r, z)
        elif z > self.z_H_He_evo_start:
            return x_H0_H_Saha(r, z) + r.fHe * self.sol_He(z)[0]
        else:
            u = self.sol_H_He(z)
            return u[0] + r.fHe * u[1]

    def Tmat(self, z):                           # Tmat_RECFAST :414-423
        if z > self.z_He_evo_start:
            return self.r.Tnow * (1 + z)
        elif z > self.z_H_He_evo_start:
            return self.sol_He(z)[1]
        else:
            return self.sol_H_He(z)[2]


def reionization_Xe(rh, z):                      # recfast.jl:478-488 (zre hard-coded)
    r = rh.r
    X_fin = 1 + r.Yp / (r.not4 * (1 - r.Yp))
    zre, α, ΔH, zHe, ΔHe, fHe = 7.6711, 1.5, 0.5, 3.5, 0.5, X_fin - 1
    x_orig = rh.Xe(z)
    x_reio_H = (X_fin - x_orig) / 2 * (
        1 + math.tanh(((1 + zre) ** α - (1 + z) ** α) / (α * (1 + zre) ** (α - 1)) / ΔH)) + x_orig
    x_reio_He = fHe / 2 * (1 + math.tanh((zHe - z) / ΔHe))
    return x_reio_H + x_reio_He


class TanhReionizationHistory:                   # recfast.jl:506-536
    def __init__(self, rh, zre_ini=50.0):
        self.rh, self.zre_ini = rh, zre_ini
        r = rh.r

        def rhs(z, Tm):
            x_reio = reionization_Xe(rh, z)
            Hz, _ = r.Hz_dHdz(z)
            Trad = r.Tnow * (1 + z)
            return (r.CT * Trad ** 4 * x_reio / (1 + x_reio + r.fHe) * (Tm[0] - Trad) / (Hz * (1 + z))
                    + 2 * Tm[0] / (1 + z),)

        self.sol = solve_ivp(rhs, (zre_ini, rh.zfinal), (rh.Tmat(zre_ini),), method="DOP853",
                             rtol=min(r.tol, 1e-8), atol=1e-12, dense_output=True).sol

    def Xe(self, z):
        return self.rh.Xe(z) if z > self.zre_ini else reionization_Xe(self.rh, z)

    def Tmat(self, z):
        return self.rh.Tmat(z) if z > self.zre_ini else self.sol(z)[0]


class IonizationHistory:
    """recfast.jl:674-726 with τ, τ′, g̃ from ionization.jl:107-137."""

    def __init__(self, r, par, bg):
        x_grid = bg.x_grid
        x0, dx = bg.x0, bg.dx
        rhist = RecfastHistory(r)
        trhist = TanhReionizationHistory(rhist)
        self.rhist, self.trhist = rhist, trhist
        xinitial = math.log(1.0 / (1.0 + rhist.zinitial))
        Xe_initial = rhist.Xe(rhist.zinitial)
        x2z = lambda x: 1.0 / math.exp(x) - 1.0
        Xe = np.array([Xe_initial if x < xinitial else trhist.Xe(x2z(x)) for x in x_grid])
        Tmat = np.array([r.Tnow * (1 + x2z(x)) if x < xinitial else trhist.Tmat(x2z(x)) for x in x_grid])

        # ionization.jl:124-129 (τ′) and :107-117 (reverse cumulative trapezoid)
        a = np.exp(x_grid)
        n_H = par.Ω_b * bg.ρ_crit / (K.m_H * a ** 3) * (1 - par.Y_p)
        τp = -Xe * n_H * a * K.sigma_T / bg.H(x_grid)
        rx, ry = x_grid[::-1], τp[::-1]
        cum = np.concatenate([[0.0], np.cumsum((rx[1:] - rx[:-1]) * (ry[1:] + ry[:-1]) / 2.0)])
        τ = cum[::-1]
        g = -τp * np.exp(-τ)

        S = lambda y: CubicBSpline(y, x0, dx)
        self.Xe = S(Xe)
        self.τ = S(τ)
        self.g = S(g)
        self.Tmat = S(Tmat)
        csb2_pre = r.C ** -2 * r.k_B / r.m_H * (1 / r.mu_T + (1 - r.Yp) * self.Xe(x_grid))
        dTmat = spline_dx(self.Tmat, x_grid)
        self.csb2 = S(csb2_pre * (self.Tmat(x_grid) - (1.0 / 3.0) * dTmat(x_grid)))
        self.τ0 = float(τ[-1])
        self.τp = spline_dx(self.τ, x_grid)
        self.τpp = spline_dx2(self.τ, x_grid)
        self.gp = spline_dx(self.g, x_grid)
        self.gpp = spline_dx2(self.g, x_grid)
