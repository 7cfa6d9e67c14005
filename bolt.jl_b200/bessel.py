"""Host mirror of the reference's Bessel-moment / Filon interface (src/bessel/*.jl) over the CUDA library.

Same names, argument meaning and error behaviour as the reference; every number is computed on the device
(bolt_sph_j_moments, bolt_moment_table_*, bolt_filon_*: include/bolt_cuda.h).  There is no CPU fallback.

  sph_bessel_interpolator(ν, order, kη_min, kη_max, N; weniger_cut)  -> MomentTable       interpolator.jl:67-80
  MomentTable.__call__(x), getnu, getorder                                                 interpolator.jl:19-38
  sph_j_moment_weniger_1F2 / sph_j_moment_asymp / sph_j_moment_maclaurin_1F2               moments.jl:61-83
  J_moment_weniger_1F2 / J_moment_asymp (J_ν = sqrt(2t/π) j_{ν-1/2}: half-integer ν)       moments.jl:24-35
  integrate_sph_bessel_filon, _loop_integrate_sph_bessel_filon                             integrator.jl:7-38
  filon_chain: the loop form batched over k (one block per k on the device)
"""
import ctypes as C
import math

import numpy as np

from . import abi
from .capi import BoltError, lib

SMALL, ASYMP, MACLAURIN = 0, 1, 2


def _check(rc, what):
    if rc != 0:
        raise BoltError(f"{what} failed with {rc}: {lib().bolt_moments_last_error().decode()}")


def _moments(nu, powers, method, x, device=0):
    x = np.atleast_1d(np.ascontiguousarray(x, dtype=np.float64))
    powers = np.ascontiguousarray(powers, dtype=np.float64)
    if int(nu) != nu:
        raise BoltError(f"ν = {nu}: spherical moments need an integer order")
    out = np.zeros((len(x), len(powers)))
    _check(lib().bolt_sph_j_moments(int(device), int(nu), len(powers), abi.ptr(powers), int(method), abi.ptr(x), len(x), abi.ptr(out)),
           "bolt_sph_j_moments")
    return out


class WenigerCache1F2:
    """Placeholder for the reference's scratch object (weniger.jl:5-35): the device evaluator needs none."""

    def __init__(self, T=float):
        self.T = T


def sph_j_moment_weniger_1F2(x, nu, m, cache=None, device=0):
    """∫₀ˣ tᵐ j_ν(t) dt for small and moderate x (moments.jl:72-76)."""
    return float(_moments(nu, [m], SMALL, [x], device)[0, 0])


def sph_j_moment_asymp(x, nu, m, prefactor=None, device=0):
    """Lommel asymptotic form for large x (moments.jl:61-70).  `prefactor` is accepted for signature parity and recomputed."""
    return float(_moments(nu, [m], ASYMP, [x], device)[0, 0])


def sph_j_moment_asymp_prefactor(nu, m):
    """∫₀^∞ tᵐ j_ν(t) dt (moments.jl:57-58)."""
    return math.sqrt(math.pi) * 2.0 ** (m - 1) * math.gamma((nu + m + 1) / 2) / math.gamma((nu - m + 2) / 2)


def sph_j_moment_asymp_nu_2(x, m, prefactor=None, device=0):
    return sph_j_moment_asymp(x, 2, m, device=device)


def sph_j_moment_asymp_nu_3(x, m, prefactor=None, device=0):
    return sph_j_moment_asymp(x, 3, m, device=device)


def sph_j_moment_maclaurin_1F2(x, nu, m, device=0):
    """Maclaurin series of the ₁F₂ for vanishing x (moments.jl:79-83)."""
    return float(_moments(nu, [m], MACLAURIN, [x], device)[0, 0])


def _cyl(nu, alpha):
    if int(nu - 0.5) != nu - 0.5:
        raise BoltError(f"J_moment: ν = {nu} must be half-integer (J_ν = sqrt(2t/π) j_(ν-1/2))")
    return int(nu - 0.5), alpha + 0.5


def J_moment_weniger_1F2(x, nu, alpha, cache=None, device=0):
    """∫₀ˣ t^α J_ν(t) dt (moments.jl:24-27) through the spherical moment of order ν-1/2 and power α+1/2."""
    n, p = _cyl(nu, alpha)
    return math.sqrt(2 / math.pi) * float(_moments(n, [p], SMALL, [x], device)[0, 0])


def J_moment_asymp(x, nu, alpha_minus_half, device=0):
    """moments.jl:30-35 (takes α - 1/2 like the reference)."""
    n, p = _cyl(nu, alpha_minus_half + 0.5)
    return math.sqrt(2 / math.pi) * float(_moments(n, [p], ASYMP, [x], device)[0, 0])


def J_moment_asymp_nu_five_halves(x, alpha_minus_half, prefactor=None, device=0):
    return J_moment_asymp(x, 2.5, alpha_minus_half, device=device)


class MomentTable:
    """The reference's MomentTable (interpolator.jl:19-38): callable, returns the `order` moments at x."""

    def __init__(self, nu, order, keta_min, keta_max, N, weniger_cut=50.0, device=0):
        self._h = C.c_void_p()
        self.nu, self.order, self.keta_min, self.keta_max, self.N = int(nu), int(order), float(keta_min), float(keta_max), int(N)
        _check(lib().bolt_moment_table_create(int(device), int(nu), int(order), float(keta_min), float(keta_max), int(N), float(weniger_cut),
                                              C.byref(self._h)), "bolt_moment_table_create")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().bolt_moment_table_free(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def many(self, x):
        """[len(x)][order] moments (one device launch)."""
        x = np.atleast_1d(np.ascontiguousarray(x, dtype=np.float64))
        out = np.zeros((len(x), self.order))
        _check(lib().bolt_moment_table_eval(self._h, abi.ptr(x), len(x), abi.ptr(out)), "bolt_moment_table_eval")
        return out

    def __call__(self, x):
        return self.many([x])[0]

    def __repr__(self):
        return f"MomentTable(device, Float64 output, ν={self.nu}) of order {self.order} with kη over ({self.keta_min}, {self.keta_max})"


def sph_bessel_interpolator(nu, order, keta_min, keta_max, N, weniger_cut=50.0, device=0):
    return MomentTable(nu, order, keta_min, keta_max, N, weniger_cut, device)


def getnu(itp):
    return itp.nu


def getorder(itp):
    return itp.order


def sph_j_moment_maclaurin_all_orders(itp, x, device=0):
    return _moments(itp.nu, np.arange(itp.order), MACLAURIN, [x], device)[0]


def _arr(v):
    return np.atleast_1d(np.ascontiguousarray(v, dtype=np.float64))


def integrate_sph_bessel_filon(f, f1, f2, k, a, b, itp):
    """∫_a^b (f + f′(x-a) + f″(x-a)²/2) j_ν(kx) dx (integrator.jl:7-20).  Scalars or equal-length arrays."""
    arrs = np.broadcast_arrays(*[_arr(v) for v in (f, f1, f2, k, a, b)])
    arrs = [np.ascontiguousarray(v) for v in arrs]
    out = np.zeros(len(arrs[0]))
    _check(lib().bolt_filon_pieces(itp._h, len(out), *[abi.ptr(v) for v in arrs], abi.ptr(out)), "bolt_filon_pieces")
    return out if np.ndim(f) or np.ndim(k) or np.ndim(a) or np.ndim(b) else float(out[0])


def _loop_integrate_sph_bessel_filon(f, f1, f2, k, a, b, itp, itp_ka):
    """integrator.jl:25-38.  The re-use of I(ka) matters on a CPU; the device evaluates a whole chain at once (filon_chain),
    so this scalar form simply returns the piece and I(kb)."""
    return integrate_sph_bessel_filon(f, f1, f2, k, a, b, itp), itp(k * b)


def filon_chain(nodes, f, f1, f2, k, itp, timing=False):
    """For every k: Σ_i ∫_{nodes[i]}^{nodes[i+1]} (quadratic through f, f′, f″ at nodes[i]) j_ν(kx) dx.
    f, f1, f2: [len(k)][len(nodes)].  Returns [len(k)] (and the kernel's device time in ms when timing=True)."""
    nodes, k = _arr(nodes), _arr(k)
    f, f1, f2 = [np.ascontiguousarray(v, dtype=np.float64).reshape(len(k), len(nodes)) for v in (f, f1, f2)]
    out = np.zeros(len(k))
    ms = C.c_float(0.0)
    _check(lib().bolt_filon_chain(itp._h, len(k), len(nodes), abi.ptr(nodes), abi.ptr(f), abi.ptr(f1), abi.ptr(f2), abi.ptr(k), abi.ptr(out),
                                  C.byref(ms) if timing else None), "bolt_filon_chain")
    return (out, float(ms.value)) if timing else out
