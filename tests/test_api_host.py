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


def test_solution_interpolates_cubics_exactly():
    from bolt_b200.api import Solution
    xg = np.linspace(-20.0, 0.0, 2001)
    f = lambda x: np.stack([x ** 3 - 2 * x, 0.5 * x ** 2 + 1.0], axis=-1)
    sol = Solution(xg, f(xg), 0, 1)
    for x in (-20.0, -19.9973, -7.123456, -0.0031, 0.0):
        assert np.allclose(sol(x), f(np.array(x)), rtol=1e-11, atol=1e-9)


def test_plin_epilogue_on_host_matches_oracle(cosmo, oracle):
    """plin(x != 0) runs the epilogue of spectra.jl:170-197 on the host: at x = 0 it must reproduce the oracle's plin."""
    from bolt_b200 import abi
    from bolt_b200.api import _plin_from_state
    ks = np.array([20.0, 600.0]) * cosmo.bg.H0
    o = abi.make_opts(8, 8, 10, reltol=1e-5, abstol=1e-6)
    pk, st, _ = oracle.plin(ks, o)
    uf = oracle.solve(ks, o, want=("u_final",))["u_final"]
    mine = [_plin_from_state(uf[i], ks[i], cosmo.par, cosmo.bg, 0.0, 8, 8, 10) for i in range(2)]
    assert np.allclose(mine, pk, rtol=1e-10)


def test_sibling_cache_is_keyed_on_identity_not_id(cosmo, monkeypatch):
    """a new cosmology whose Background happens to reuse the id() of a collected one must not hit the cache."""
    from bolt_b200 import api
    calls = []

    def fake(par, bg, ih, k_grid, ℓᵧ, reltol, ctx):
        calls.append((bg, ih)); return ("T%d" % len(calls), "P%d" % len(calls))
    monkeypatch.setattr(api, "_source_grids", fake)
    api._pair_cache.clear()
    kg = np.array([1.0, 2.0])

    class Obj:
        pass
    bg1, ih1, bg2 = Obj(), Obj(), Obj()
    assert api.source_grid(None, bg1, ih1, kg, None) == "T1"
    assert api.source_grid_P(None, bg1, ih1, kg, None) == "P1" and len(calls) == 1        # sibling served from the cache
    assert api.source_grid_P(None, bg1, ih1, kg, None) == "P2"                            # consumed: solved again
    assert api.source_grid(None, bg2, ih1, kg, None) == "T3"                              # another background: never the cache
    assert api._pair_cache[0]["bg"] is bg2                                                 # and the entry keeps its objects alive
    api._pair_cache.clear()
