"""Host input generator (background, RECFAST, optical depth) against the reference's fixtures."""
import numpy as np
import pytest

from conftest import load_golden


def test_recfast_matches_fortran_golden():
    """test/runtests.jl:38-48: |Xe_fortran - Xe_RECFAST| < 1e-4 with CosmoParams(Σm_ν=0, N_ν=3, Ω_r=5.042e-5), Tnow=2.725."""
    import bolt_b200 as B
    import hostgen as HG
    from hostgen.recfast import RecfastHistory
    g = load_golden("recfast_xe.npz")
    par = B.CosmoParams(Σm_ν=0.0, N_ν=3.0, Ω_r=5.042e-5)
    bg = HG.Background(par)
    r = HG.RECFAST(bg, OmegaB=par.Ω_b, Yp=par.Y_p, OmegaG=par.Ω_r, Tnow=2.725)
    rh = RecfastHistory(r)
    mine = np.array([rh.Xe(z) for z in g["z"]])
    assert np.all(np.abs(mine - g["Xe"]) < 1e-4)


def test_bspline_interpolates_and_has_natural_ends():
    from hostgen.bspline import CubicBSpline
    x = np.linspace(-3.0, 2.0, 51)
    y = np.sin(x) + 0.1 * x ** 2
    s = CubicBSpline(y, x[0], x[1] - x[0])
    assert np.allclose(s(x), y, rtol=0, atol=1e-14)
    # Line(OnGrid()) boundary: zero second derivative at both end samples
    assert abs(s.hessian(x[0])) < 1e-10 and abs(s.hessian(x[-1])) < 1e-10
    xm = 0.5 * (x[1:] + x[:-1])
    assert np.max(np.abs(s(xm)[8:-8] - (np.sin(xm) + 0.1 * xm ** 2)[8:-8])) < 1e-6
    # gradient / hessian are the analytic derivatives of the same basis
    h = 1e-5
    assert np.allclose(s.gradient(xm), (s(xm + h) - s(xm - h)) / (2 * h), rtol=1e-6, atol=1e-8)


def test_background_basics(cosmo):
    bg, par = cosmo.bg, cosmo.par
    assert len(bg.x_grid) == 2001 and bg.x_grid[1200] == -8.0 and cosmo.ix_start == 1201
    assert bg.H0 == pytest.approx(0.7 * 100 / 299792.458, rel=1e-12)
    assert 1000 * bg.H0 * bg.η0 == pytest.approx(3377.5, rel=2e-3)       # k_max η₀
    # radiation era: ℋ η -> 1
    assert bg.H(-20.0) * bg.η(-20.0) == pytest.approx(1.0, abs=1e-4)
    # flat universe closure today: H(a=1) = H0
    assert bg.H(0.0) == pytest.approx(bg.H0, rel=1e-6)


def test_ionization_history_sanity(cosmo):
    ih, bg = cosmo.ih, cosmo.bg
    assert 0.04 < ih.τ(-3.0) < 0.07                      # reionization optical depth (zre = 7.6711 hard-coded)
    g = ih.g(bg.x_grid)
    xpk = bg.x_grid[np.argmax(g)]
    assert 1050 < 1 / np.exp(xpk) - 1 < 1120             # visibility peak at recombination
    # ∫ g̃ dx = 1 - e^{-τ(x_min)} ≈ 1
    assert np.trapezoid(g, bg.x_grid) == pytest.approx(1.0, abs=2e-3)
    assert np.all(ih.τp(bg.x_grid[5:-5]) < 0)


def test_packed_descriptor_layout(cosmo):
    from bolt_b200 import abi
    hc = cosmo.hc
    assert hc.nd == 1 and hc.n_x == 2001 and hc.nq == 15
    assert hc.tables.shape == (abi.NTABLES, 2003, 1) and hc.scalars.shape == (abi.NSCALARS, 1)
    assert hc.scalar("η0") == cosmo.bg.η0 and hc.scalar("Σm_ν") == cosmo.par.Σm_ν
