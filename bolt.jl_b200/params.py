"""CosmoParams and the massive-neutrino momentum grid: the parameter container of the reference's API (src/Bolt.jl:56-66) and the
small helper functions of the hot path's scope table -- f0, dlnf0dlnq (src/background.jl:21-30), to_ui / from_ui / dxdq / xq2q
(src/util.jl:24-27), the q grid of hierarchy! (src/perturbations.jl:164-166).  The device evaluates the same expressions once per
cosmology at upload (bolt_cosmo_upload); the host copies here serve the API mirror's post-processing (rsa_perts, plin at x != 0).

The 1-D background and RECFAST generators are NOT part of the product: no Julia exists in the build image to produce the input
tables, so a restatement of those out-of-scope reference components lives in the `hostgen/` harness at the repository root.
"""
from dataclasses import dataclass, replace, fields
import math
import numpy as np


class C:
    """The three unit constants these helpers need, in the reference's "Mpc units" (src/Bolt.jl:49-51; CODATA 2018)."""
    _C_SI, _HBAR_SI, _EV_SI, _G_SI = 299792458.0, 6.62607015e-34 / (2.0 * math.pi), 1.602176634e-19, 6.67430e-11
    _MPC_SI = 1.0e6 * (149597870700.0 * 648000.0 / math.pi)
    km_s_Mpc_100 = 100.0e3 / _C_SI
    G_natural = _G_SI * _HBAR_SI / _C_SI**3 / _MPC_SI**2
    mass_natural = _EV_SI / (_HBAR_SI * _C_SI) * _MPC_SI


@dataclass(frozen=True)
class CosmoParams:
    """src/Bolt.jl:56-66 (same field names; Σm_ν is in Mpc^-1 like the reference)."""
    h: float = 0.7
    Ω_r: float = 5.0469e-5
    Ω_b: float = 0.046
    Ω_c: float = 0.224
    A: float = 2.097e-9
    n: float = 1.0
    Y_p: float = 0.24
    N_ν: float = 3.046
    Σm_ν: float = 0.06 * C.mass_natural

    def replace(self, **kw):
        return replace(self, **kw)

    @staticmethod
    def names():
        return [f.name for f in fields(CosmoParams)]


def H0(par):                                   # background.jl:5
    return par.h * C.km_s_Mpc_100


def rho_crit(par):                             # background.jl:6
    return (3.0 / (8.0 * np.pi)) * H0(par) ** 2 / C.G_natural


def T_nu(par):                                 # background.jl:22 (repeated all over the reference)
    return (par.N_ν / 3.0) ** 0.25 * (4.0 / 11.0) ** (1.0 / 3.0) * \
        (15.0 / np.pi ** 2 * rho_crit(par) * par.Ω_r) ** 0.25


def f0(q, par):                                # background.jl:21-25
    return 2.0 / (2.0 * np.pi) ** 3 / (np.exp(q / T_nu(par)) + 1.0)


def dlnf0dlnq(q, par):                         # background.jl:27-30
    Tν = T_nu(par)
    return -q / Tν / (1.0 + np.exp(-q / Tν))


# util.jl:24-27
def to_ui(lq, lqmi, lqma):
    return -1.0 + (1.0 - (-1.0)) / (lqma - lqmi) * (lq - lqmi)


def from_ui(x, lqmi, lqma):
    return lqmi + (lqma - lqmi) / (1.0 - (-1.0)) * (x - (-1.0))


def dxdq(q, lqmi, lqma):
    return (1.0 + to_ui(1.0 + lqmi, lqmi, lqma)) / (q * np.log(10.0))


def xq2q(x, lqmi, lqma):
    return 10.0 ** from_ui(x, lqmi, lqma)


def q_grid(par, quad_pts):
    """Momentum nodes q_i on [Tν/30, 30 Tν] (perturbations.jl:164-166)."""
    Tν = T_nu(par)
    lqmi, lqma = np.log10(Tν / 30.0), np.log10(Tν * 30.0)
    return xq2q(quad_pts, lqmi, lqma), lqmi, lqma


