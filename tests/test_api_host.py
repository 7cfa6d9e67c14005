"""Host-side logic of the reference-interface mirror (no GPU needed)."""
import numpy as np
import pytest


def test_k_grids_match_reference_formulas():
    import bolt_b200 as B
    kq = B.quadratic_k(0.1, 1000.0, 100)           # src/spectra.jl:60-63 (i = 1..n, so the first point is NOT kmin)
    assert len(kq) == 100 and kq[-1] == 1000.0
    assert kq[0] == pytest.approx(0.1 + 999.9 * 1e-4)
    kl = B.log10_k(10.0, 5000.0, 32)               # src/spectra.jl:65-68
    assert kl[0] == pytest.approx(10.0) and kl[-1] == pytest.approx(5000.0)
    assert np.allclose(np.diff(np.log10(kl)), np.log10(500.0) / 31)


def test_source_interpolant_bilinear_with_line_extrapolation():
    from bolt_b200.api import SourceInterpolant
    xg = np.linspace(-2.0, 0.0, 5); kg = np.array([1.0, 2.0, 4.0])
    f = lambda x, k: 3.0 * x + 2.0 * k - 0.5
    grid = f(xg[:, None], kg[None, :])
    itp = SourceInterpolant(xg, kg, grid)
    assert itp(-1.3, 3.1) == pytest.approx(f(-1.3, 3.1))
    assert itp(-0.7, 0.25) == pytest.approx(f(-0.7, 0.25))      # below the first k: linear extrapolation (Line())
    assert itp(-0.7, 9.0) == pytest.approx(f(-0.7, 9.0))
    assert np.allclose(itp(xg[:, None], kg[None, :]), grid)      # broadcastable like itp(xs, ks)


def test_state_dim_and_hierarchy_defaults(cosmo):
    import bolt_b200 as B
    h = B.Hierarchy(B.BasicNewtonian(), cosmo.par, cosmo.bg, cosmo.ih, 0.01)
    assert (h.ℓᵧ, h.ℓ_ν, h.ℓ_mν, h.nq) == (8, 8, 10, 15) and h.n == 197      # src/perturbations.jl:20-21
    h = B.Hierarchy(B.BasicNewtonian(), cosmo.par, cosmo.bg, cosmo.ih, 0.01, 50, 50, 20, 15)
    assert h.n == 473


def test_dense_grid_detection():
    from bolt_b200.api import _is_quadratic
    import bolt_b200 as B
    ok, kmin = _is_quadratic(B.quadratic_k(0.01, 1000.0, 5000))
    assert ok and kmin == pytest.approx(0.01, rel=1e-6)
    ok, _ = _is_quadratic(B.log10_k(0.01, 1000.0, 50))
    assert not ok
