"""Host-side mirror of the reference's exported function set for the hot path (src/Bolt.jl:8-9):
Hierarchy / boltsolve / boltsolve_rsa (src/perturbations.jl:7-33, 86-111), source_grid / source_grid_P,
quadratic_k / log10_k, cltt / clte / clee, plin (src/spectra.jl).  Same names, argument order, defaults
and return shapes as the Julia functions; the arithmetic runs in libbolt_cuda.so (no CPU fallback).
This file is the Python twin of julia/BoltCUDA.jl (the Julia shim cannot be executed in this image).
"""
import threading
import numpy as np

from . import abi, capi
from .params import CosmoParams


class BasicNewtonian:
    """src/perturbations.jl:4"""


_tls = threading.local()


def default_context(device=0):
    """One bolt_ctx per host thread (the reference calls plin/cltt from many threads)."""
    ctxs = getattr(_tls, "ctxs", None)
    if ctxs is None:
        ctxs = _tls.ctxs = {}
    if device not in ctxs:
        ctxs[device] = capi.Context(device)
    return ctxs[device]


def device_cosmo(par, bg, ih, ctx=None):
    """Upload (once per (bg, ih) pair and context) what Background and IonizationHistory computed on the host."""
    ctx = ctx or default_context()
    cache = ih.__dict__.setdefault("_bolt_dev", {})
    key = (id(ctx), id(bg))
    if key not in cache:
        cache[key] = capi.DeviceCosmo(ctx, abi.HostCosmo.from_host(par, bg, ih))
    return cache[key]


def _check_status(status, where):
    """The reference never checks the solver's retcode (perturbations.jl:29-32); the shim warns instead of staying silent
    (SURVEY 8b).  0 = ok, 4 = the RSA switch fired inside the solve (reported, results still returned)."""
    bad = np.asarray(status)
    bad = bad[(bad != abi.K_OK) & (bad != abi.K_RSA_TRIGGERED)]
    if bad.size:
        import warnings
        warnings.warn(f"{where}: {bad.size} k-mode(s) did not finish (status codes {sorted(set(bad.tolist()))}: 1 max steps, "
                      "2 step underflow, 3 non-finite); their remaining rows are zero", RuntimeWarning, stacklevel=3)


class Hierarchy:
    """src/perturbations.jl:7-21"""

    def __init__(self, integrator, par, bg, ih, k, ℓᵧ=8, ℓ_ν=8, ℓ_mν=10, nq=15):
        if nq != bg.nq:
            raise ValueError("nq must match the background's quadrature (bg.quad_pts)")
        self.integrator, self.par, self.bg, self.ih, self.k = integrator, par, bg, ih, float(k)
        self.ℓᵧ, self.ℓ_ν, self.ℓ_mν, self.nq = ℓᵧ, ℓ_ν, ℓ_mν, nq

    @property
    def n(self):
        return abi.state_dim(self.ℓᵧ, self.ℓ_ν, self.ℓ_mν, self.nq)


class Solution:
    """What boltsolve returns: callable like the reference's ODESolution, `sol(x) -> Vector(n)`.
    The solution is held on bg.x_grid (the device's Hermite dense output sampled at every grid point: exact there);
    between grid points a local four-point cubic through the neighbouring rows is used (O(dx⁴) like the
    reference's Hermite interpolant between its own steps, which the device does not keep)."""

    def __init__(self, x_grid, u, status, nsteps):
        self.t, self.u, self.retcode, self.nsteps = x_grid, u, int(status), int(nsteps)

    def __call__(self, x):
        nx = len(self.t)
        t = (x - self.t[0]) / ((self.t[-1] - self.t[0]) / (nx - 1))
        i = int(np.clip(np.floor(t), 0, nx - 2))
        w = t - i
        if w == 0.0:
            return self.u[i].copy()
        j = int(np.clip(i - 1, 0, nx - 4))          # rows j..j+3, Lagrange weights at s = t - j
        s_ = t - j
        wts = [-(s_ - 1) * (s_ - 2) * (s_ - 3) / 6, s_ * (s_ - 2) * (s_ - 3) / 2, -s_ * (s_ - 1) * (s_ - 3) / 2, s_ * (s_ - 1) * (s_ - 2) / 6]
        return sum(wt * self.u[j + m] for m, wt in enumerate(wts))


def _opts(h_or_trunc, reltol, abstol, **kw):
    ℓᵧ, ℓ_ν, ℓ_mν = h_or_trunc
    return abi.make_opts(ℓᵧ, ℓ_ν, ℓ_mν, reltol=reltol, abstol=abstol, **kw)


def boltsolve(hierarchy, ode_alg=None, reltol=1e-6, abstol=1e-6, ctx=None):
    """src/perturbations.jl:25-33.  `ode_alg` is accepted for signature compatibility; the device
    runs KenCarp4's ESDIRK tableau."""
    h = hierarchy
    dc = device_cosmo(h.par, h.bg, h.ih, ctx)
    out = dc.solve(np.array([h.k]), _opts((h.ℓᵧ, h.ℓ_ν, h.ℓ_mν), reltol, abstol), want=("u_hist",))
    _check_status(out["status"], "boltsolve")
    return Solution(h.bg.x_grid, out["u_hist"][0], out["status"][0], out["nsteps"][0])


def rsa_perts(u, hierarchy, x):
    """rsa_perts! (src/perturbations.jl:35-84): overwrite Θ₀..₂, 𝒩₀..₂ with the RSA expressions, zero the rest."""
    h = hierarchy
    k, ℓᵧ, ℓ_ν, nq, par, bg, ih = h.k, h.ℓᵧ, h.ℓ_ν, h.nq, h.par, h.bg, h.ih
    H02 = bg.H0 ** 2
    Hx, ηx, τp, τpp = bg.H(x), bg.η(x), ih.τp(x), ih.τpp(x)
    a = np.exp(x)
    Ω_ν = 7 * (2 / 3) * par.N_ν / 8 * (4 / 11) ** (4 / 3) * par.Ω_r
    csb2 = ih.csb2(x)
    iN = 2 * (ℓᵧ + 1); iM = iN + ℓ_ν + 1; iS = iM + (h.ℓ_mν + 1) * nq
    Φ, δ, v, δ_b, v_b = u[iS:iS + 5]
    from .params import q_grid, f0, dxdq
    q, lqmi, lqma = q_grid(par, bg.quad_pts)
    eps = np.sqrt(q ** 2 + (a * par.Σm_ν) ** 2)
    w = f0(q, par) / dxdq(q, lqmi, lqma) * bg.quad_wts
    ρM = 4 * np.pi * np.sum(q ** 2 * eps * w * u[iM:iM + nq])
    σM = 4 * np.pi * np.sum(q ** 2 * (q ** 2 / eps) * w * u[iM + 2 * nq:iM + 3 * nq])
    Ψ = -Φ - 12 * H02 / k ** 2 / a ** 2 * (par.Ω_r * u[2] + Ω_ν * u[iN + 2] + σM / bg.ρ_crit / 4)
    Φp = Ψ - k ** 2 / (3 * Hx ** 2) * Φ + H02 / (2 * Hx ** 2) * (
        par.Ω_c / a * δ + par.Ω_b / a * δ_b + 4 * par.Ω_r / a ** 2 * u[0] + 4 * Ω_ν / a ** 2 * u[iN] + ρM / a ** 2 / bg.ρ_crit)
    u[0] = Φ - Hx / k * τp * v_b
    u[1] = Hx / k * (-2 * Φp + τp * (Φ - csb2 * δ_b) + Hx / k * (τpp - τp) * v_b)
    u[2] = 0
    u[iN] = Φ; u[iN + 1] = -2 * Hx / k * Φp; u[iN + 2] = 0
    # the reference's zeroing loop uses 1-based u[ℓ] for ℓ in 3:ℓᵧ (perturbations.jl:78-82), reproduced as written
    for ℓ in range(3, ℓᵧ + 1):
        u[ℓ - 1] = 0
        u[(ℓᵧ + 1) + ℓ - 1] = 0
    for ℓ in range(3, ℓ_ν + 1):
        u[2 * (ℓᵧ + 1) + ℓ - 1] = 0


def boltsolve_rsa(hierarchy, ode_alg=None, reltol=1e-6, abstol=1e-6, ctx=None):
    """src/perturbations.jl:86-111 -> Matrix(n, n_x) (returned as an array of shape (n, n_x))."""
    h = hierarchy
    sol = boltsolve(h, reltol=reltol, abstol=abstol, ctx=ctx)
    x_grid = h.bg.x_grid
    results = sol.u.T.copy()
    kη = h.k * h.bg.η(x_grid)
    od = -h.ih.τp(x_grid) * h.bg.H(x_grid) / h.bg.η(x_grid)
    hor = np.nonzero(kη > 240)[0]
    xrsa_hor = (hor[0] if len(hor) else len(x_grid) - 1)
    odi = np.nonzero(od > 100)[0]
    xrsa_od = (odi[0] if len(odi) else len(x_grid) - 1)   # (:99 tests the wrong variable in the reference; same result)
    switch = x_grid[max(xrsa_hor, xrsa_od)]
    for i in np.nonzero(x_grid > switch)[0]:
        col = results[:, i].copy()
        rsa_perts(col, h, x_grid[i])
        results[:, i] = col
    return results


class SourceInterpolant:
    """LinearInterpolation((x_grid, k_grid), grid, extrapolation_bc=Line()) (src/spectra.jl:21)."""

    def __init__(self, x_grid, k_grid, grid, other=None, meta=None):
        self.x_grid, self.k_grid, self.grid = x_grid, np.asarray(k_grid, dtype=np.float64), grid   # grid[i_x, i_k]
        self._other, self.meta = other, meta or {}

    def __call__(self, x, k):
        xg, kg = self.x_grid, self.k_grid
        tx = (np.asarray(x, dtype=np.float64) - xg[0]) / ((xg[-1] - xg[0]) / (len(xg) - 1))
        ix = np.clip(np.floor(tx).astype(int), 0, len(xg) - 2); wx = tx - ix
        k = np.asarray(k, dtype=np.float64)
        jk = np.clip(np.searchsorted(kg, k, side="right") - 1, 0, len(kg) - 2)
        wk = (k - kg[jk]) / (kg[jk + 1] - kg[jk])
        g = self.grid
        return ((1 - wx) * (1 - wk) * g[ix, jk] + wx * (1 - wk) * g[ix + 1, jk]
                + (1 - wx) * wk * g[ix, jk + 1] + wx * wk * g[ix + 1, jk + 1])


def _source_grids(par, bg, ih, k_grid, ℓᵧ, reltol, ctx):
    dc = device_cosmo(par, bg, ih, ctx)
    k_grid = np.ascontiguousarray(k_grid, dtype=np.float64)
    out = dc.solve(k_grid, _opts((ℓᵧ, 8, 10), reltol, 1e-6), want=("S_T", "S_P"))
    meta = dict(status=out["status"], nsteps=out["nsteps"], nreject=out["nreject"])
    _check_status(out["status"], "source_grid")
    sT = SourceInterpolant(bg.x_grid, k_grid, np.ascontiguousarray(out["S_T"].T), meta=meta)
    sP = SourceInterpolant(bg.x_grid, k_grid, np.ascontiguousarray(out["S_P"].T), meta=meta)
    sT._other, sP._other = sP, sT
    return sT, sP


# The sibling grid of the last source_grid / source_grid_P call (one entry).  The entry keeps `bg` and `ih` alive and is
# matched by identity (`is`), never by a bare id(): an id can be reused by a new object once the old one is collected.
_pair_cache = []


def _sibling(which, par, bg, ih, k_grid, ℓᵧ, reltol, ctx):
    kb = np.asarray(k_grid, dtype=np.float64).tobytes()
    if _pair_cache:
        e = _pair_cache[0]
        if e["bg"] is bg and e["ih"] is ih and e["k"] == kb and e["trunc"] == (ℓᵧ, reltol) and e["left"] == which:
            _pair_cache.clear()
            return e["pair"][which]
    pair = _source_grids(par, bg, ih, k_grid, ℓᵧ, reltol, ctx)
    _pair_cache.clear()
    _pair_cache.append(dict(bg=bg, ih=ih, k=kb, trunc=(ℓᵧ, reltol), pair=pair, left=1 - which))
    return pair[which]


def source_grid(par, bg, ih, k_grid, integrator, ℓᵧ=8, reltol=1e-11, ctx=None):
    """src/spectra.jl:6-23.  The device emits the temperature AND polarization source from the same
    solve, so the sibling grid is cached for a following source_grid_P call with the same arguments."""
    return _sibling(0, par, bg, ih, k_grid, ℓᵧ, reltol, ctx)


def source_grid_P(par, bg, ih, k_grid, integrator, ℓᵧ=8, reltol=1e-11, ctx=None):
    """src/spectra.jl:25-42"""
    return _sibling(1, par, bg, ih, k_grid, ℓᵧ, reltol, ctx)


def quadratic_k(kmin, kmax, nk):
    """src/spectra.jl:60-63"""
    i = np.arange(1, nk + 1)
    return kmin + (kmax - kmin) * (i / nk) ** 2


def log10_k(kmin, kmax, nk):
    """src/spectra.jl:65-68"""
    i = np.arange(1, nk + 1)
    return 10.0 ** (np.log10(kmin) + (np.log10(kmax / kmin)) * (i - 1) / (nk - 1))


def _ix_start(bg):
    return int(np.argmax(bg.x_grid > -8))       # findfirst(bg.x_grid .> -8), 0-based (src/spectra.jl:86)


def _is_quadratic(kgrid):
    n = len(kgrid)
    kmin = (kgrid[0] * n * n - kgrid[-1]) / (n * n - 1.0)
    return np.allclose(quadratic_k(kmin, kgrid[-1], n), kgrid, rtol=1e-13, atol=0), kmin


def _project(ells, sT, sP, kd_min, kd_max, n_kd, par, bg, ih, ctx):
    scalar = np.isscalar(ells)
    ells = np.atleast_1d(np.asarray(ells, dtype=np.int32))
    uniq, inv = np.unique(ells, return_inverse=True)      # the library wants strictly increasing multipoles; the caller may repeat
    ref = sT if sT is not None else sP
    dc = device_cosmo(par, bg, ih if ih is not None else ref.meta.get("ih"), ctx)
    tt, te, ee = dc.project(None if sT is None else np.ascontiguousarray(sT.grid.T),
                            None if sP is None else np.ascontiguousarray(sP.grid.T),
                            ref.k_grid, uniq, kd_min, kd_max, n_kd, _ix_start(bg))
    res = [None if a is None else (a[inv][0] if scalar else a[inv]) for a in (tt, te, ee)]
    return res


def _dense(args, nsrc):
    """Dispatch the three reference methods: (ℓ, s_itp..., kgrid, par, bg) or (ℓ | ℓ⃗, par, bg, ih, sf...)."""
    if isinstance(args[1], SourceInterpolant):
        ell, srcs, kgrid, par, bg = args[0], args[1:1 + nsrc], np.asarray(args[1 + nsrc]), args[2 + nsrc], args[3 + nsrc]
        ok, kmin = _is_quadratic(kgrid)
        if not ok:
            raise NotImplementedError("only quadratic_k dense grids are supported on the device")
        return ell, srcs, kmin, kgrid[-1], len(kgrid), par, bg, None
    ell, par, bg, ih, srcs = args[0], args[1], args[2], args[3], args[4:4 + nsrc]
    return ell, srcs, 0.01 * bg.H0, 1000 * bg.H0, 5000, par, bg, ih     # src/spectra.jl:133,138,143


def _ih_of(ih, src):
    if ih is not None:
        return ih
    raise ValueError("the (ℓ, s_itp, kgrid, par, bg) method needs the ionization history that made s_itp; "
                     "pass ih=... (the Julia method reads nothing from ih either, but the device tables are keyed on it)")


def cltt(*args, ih=None, ctx=None):
    """src/spectra.jl:84-97, 132-135, 147-150"""
    ell, (sf,), kmin, kmax, n, par, bg, ih2 = _dense(args, 1)
    return _project(ell, sf, None, kmin, kmax, n, par, bg, _ih_of(ih2 or ih, sf), ctx)[0]


def clte(*args, ih=None, ctx=None):
    """src/spectra.jl:99-114, 137-140, 152-155"""
    ell, (sf, sfP), kmin, kmax, n, par, bg, ih2 = _dense(args, 2)
    return _project(ell, sf, sfP, kmin, kmax, n, par, bg, _ih_of(ih2 or ih, sf), ctx)[1]


def clee(*args, ih=None, ctx=None):
    """src/spectra.jl:116-130, 142-145, 157-160"""
    ell, (sfP,), kmin, kmax, n, par, bg, ih2 = _dense(args, 1)
    return _project(ell, None, sfP, kmin, kmax, n, par, bg, _ih_of(ih2 or ih, sfP), ctx)[2]


def plin(k, par, bg, ih, n_q=15, ℓᵧ=50, ℓ_ν=50, ℓ_mν=20, x=0, reltol=1e-5, ctx=None):
    """src/spectra.jl:163-198.  Accepts a scalar k (like the reference) or a vector of k (one batched call)."""
    if n_q != bg.nq:
        raise ValueError("n_q must match the background's quadrature")
    dc = device_cosmo(par, bg, ih, ctx)
    ks = np.atleast_1d(np.asarray(k, dtype=np.float64))
    if x != 0:
        # perturb(x) at an interior abscissa: the device solves the modes and returns their histories on bg.x_grid;
        # the epilogue (spectra.jl:170-197, a few dozen flop per mode) runs here on the interpolated state
        out = dc.solve(ks, _opts((ℓᵧ, ℓ_ν, ℓ_mν), reltol, 1e-6), want=("u_hist",))
        _check_status(out["status"], "plin")
        pk = np.array([_plin_from_state(Solution(bg.x_grid, out["u_hist"][i], out["status"][i], out["nsteps"][i])(x),
                                        ks[i], par, bg, x, ℓᵧ, ℓ_ν, ℓ_mν) for i in range(len(ks))])
        return pk[0] if np.isscalar(k) else pk
    pk, status, _ = dc.plin(ks, _opts((ℓᵧ, ℓ_ν, ℓ_mν), reltol, 1e-6))
    _check_status(status, "plin")
    return pk[0] if np.isscalar(k) else pk


def _plin_from_state(u, k, par, bg, x, ℓᵧ, ℓ_ν, ℓ_mν):
    """The epilogue of plin (src/spectra.jl:170-197) on a state vector `u = perturb(x)`."""
    from .params import q_grid, f0, dxdq
    nq = bg.nq
    iM = 2 * (ℓᵧ + 1) + (ℓ_ν + 1); iS = iM + (ℓ_mν + 1) * nq
    q, lqmi, lqma = q_grid(par, bg.quad_pts)
    a = np.exp(x)
    eps = np.sqrt(q ** 2 + (a * par.Σm_ν) ** 2)
    w = f0(q, par) / dxdq(q, lqmi, lqma) * bg.quad_wts
    ρ0M, Hx = bg.ρ0M(x), bg.H(x)
    Mρ = 4 * np.pi * np.sum(q ** 2 * eps * w * u[iM:iM + nq]) / ρ0M                  # ρ_σ (perturbations.jl:127-145)
    Mθ = k * 4 * np.pi * np.sum(q ** 3 * w * u[iM + nq:iM + 2 * nq]) / ρ0M            # θ   (perturbations.jl:148-158)
    δcN, vcN, δbN, vbN = u[iS + 1], u[iS + 2], u[iS + 3], u[iS + 4]
    vmν = -Mθ / k
    Tγ = (15 / np.pi ** 2 * bg.ρ_crit * par.Ω_r) ** 0.25
    νfac = (90 * 1.2020569 / (11 * np.pi ** 4)) * (par.Ω_r * par.h ** 2 / Tγ) * (par.N_ν / 3) ** 0.75
    Ω_ν = par.Σm_ν * νfac / par.h ** 2
    Ωm = par.Ω_c + par.Ω_b + Ω_ν
    δc = δcN - 3 * Hx * vcN / k; δb = δbN - 3 * Hx * vbN / k
    δmν = Mρ - 3 * Hx * vmν / k
    δm = (par.Ω_c * δc + par.Ω_b * δb + Ω_ν * δmν) / Ωm
    return (2 * np.pi ** 2 / k ** 3) * δm ** 2 * par.A * (k / 0.05) ** (par.n - 1)


def spectra(ells, par, bg, ih, k_grid, ℓᵧ=8, reltol=1e-11, ctx=None):
    """Fused source_grid + source_grid_P + cltt/clte/clee with the source grids kept in HBM
    (bolt_spectra): the whole-path call the benchmark times."""
    dc = device_cosmo(par, bg, ih, ctx)
    ells = np.asarray(ells, dtype=np.int32)
    tt, te, ee, status, nsteps = dc.spectra(k_grid, _opts((ℓᵧ, 8, 10), reltol, 1e-6), ells, 0.01 * bg.H0, 1000 * bg.H0, 5000,
                                            _ix_start(bg))
    _check_status(status, "spectra")
    return tt, te, ee, dict(status=status, nsteps=nsteps, nreject=dc.last_nreject)
