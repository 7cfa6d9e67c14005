"""K2 parity (through the C ABI): Bessel tables, LOS projection, k-integral, plin."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def golden_sources(cosmo):
    g = load_golden("oracle_c1.npz")
    ix0 = int(g["ix_start"]); n_x = cosmo.hc.n_x
    S_T = np.zeros((len(g["k"]), n_x)); S_P = np.zeros_like(S_T)
    S_T[:, ix0:] = g["S_T"]; S_P[:, ix0:] = g["S_P"]
    return g, S_T, S_P, ix0


def test_project_golden_sources_matches_oracle_cl(cosmo, dev):
    """Same source grids in, C_ℓ out: isolates K2 (j_ℓ tables, spline, LOS sum, midpoint rule)."""
    g, S_T, S_P, ix0 = golden_sources(cosmo)
    bg = cosmo.bg
    tt, te, ee = dev.project(S_T, S_P, g["k"], g["ell"], 0.01 * bg.H0, 1000 * bg.H0, 5000, ix0)
    assert np.abs(tt / g["tt"] - 1).max() < 1e-9
    assert np.abs(ee / g["ee"] - 1).max() < 1e-9
    assert np.abs(te - g["te"]).max() < 1e-9 * np.sqrt(g["tt"] * g["ee"]).max()
    assert np.all(te ** 2 <= tt * ee * (1 + 1e-12))


@pytest.mark.parametrize("ells", [[2], [2, 3, 4, 5, 6], [2500], [2, 17, 300, 301, 302, 999, 2500]])
def test_project_ragged_multipole_sets(cosmo, oracle, dev, ells):
    """Edge cases: single ℓ, group sizes that are not a multiple of the kernel's ℓ-tile, ℓ_min and ℓ_max."""
    g, S_T, S_P, ix0 = golden_sources(cosmo)
    bg = cosmo.bg
    ells = np.array(ells, dtype=np.int32)
    got = dev.project(S_T, S_P, g["k"], ells, 0.01 * bg.H0, 1000 * bg.H0, 5000, ix0)
    ref = oracle.project(S_T, S_P, g["k"], ells, 0.01 * bg.H0, 1000 * bg.H0, 5000, ix0)
    assert np.abs(got[0] / ref[0] - 1).max() < 1e-9 and np.abs(got[2] / ref[2] - 1).max() < 1e-9
    assert np.abs(got[1] - ref[1]).max() < 1e-9 * np.sqrt(ref[0] * ref[2]).max()


def test_project_single_source_and_small_dense_grid(cosmo, oracle, dev):
    g, S_T, S_P, ix0 = golden_sources(cosmo)
    bg = cosmo.bg
    ells = np.array([10, 100, 1000], dtype=np.int32)
    tt, te, ee = dev.project(S_T, None, g["k"], ells, 0.02 * bg.H0, 900 * bg.H0, 777, ix0 + 13)
    rtt, _, _ = oracle.project(S_T, None, g["k"], ells, 0.02 * bg.H0, 900 * bg.H0, 777, ix0 + 13)
    assert te is None and ee is None and np.abs(tt / rtt - 1).max() < 1e-9


def test_project_rejects_unsorted_multipoles(cosmo, dev):
    from bolt_b200 import capi
    g, S_T, S_P, ix0 = golden_sources(cosmo)
    with pytest.raises(capi.BoltError):
        dev.project(S_T, S_P, g["k"], np.array([10, 5], dtype=np.int32), 0.01, 1.0, 5000, ix0)


def test_fused_spectra_matches_golden_cl(cosmo, dev):
    """bolt_spectra (K1 -> K2 with source grids kept in HBM) vs the oracle's C_ℓ: north_star 1e-4 on TT/TE/EE."""
    from bolt_b200 import abi
    g = load_golden("oracle_c1.npz"); bg = cosmo.bg
    o = abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6)
    tt, te, ee, status, nsteps = dev.spectra(g["k"], o, g["ell"], 0.01 * bg.H0, 1000 * bg.H0, 5000, int(g["ix_start"]))
    assert np.all(status == 0)
    assert np.abs(tt / g["tt"] - 1).max() < 1e-4
    assert np.abs(ee / g["ee"] - 1).max() < 1e-4
    assert np.abs(te - g["te"]).max() < 1e-4 * np.sqrt(g["tt"] * g["ee"]).max()


def test_spectrum_scales_linearly_with_A(cosmo, gpu_ctx):
    """Size-independent property: C_ℓ ∝ A_s exactly (spectra.jl:92)."""
    from bolt_b200 import abi, capi
    g, S_T, S_P, ix0 = golden_sources(cosmo)
    bg = cosmo.bg
    hc2 = abi.HostCosmo(cosmo.hc.scalars.copy(), cosmo.hc.quad_pts, cosmo.hc.quad_wts, cosmo.hc.tables, cosmo.hc.x0, cosmo.hc.dx)
    hc2.scalars[abi.S["A"], 0] *= 2.0
    d1 = capi.DeviceCosmo(gpu_ctx, cosmo.hc); d2 = capi.DeviceCosmo(gpu_ctx, hc2)
    ells = np.arange(2, 2501, 41, dtype=np.int32)
    a = d1.project(S_T, S_P, g["k"], ells, 0.01 * bg.H0, 1000 * bg.H0, 5000, ix0)
    b = d2.project(S_T, S_P, g["k"], ells, 0.01 * bg.H0, 1000 * bg.H0, 5000, ix0)
    for x, y in zip(a, b):
        assert np.allclose(y, 2.0 * x, rtol=1e-13, atol=0)


def test_plin_matches_oracle(cosmo, oracle, dev):
    """plin defaults (ℓᵧ = ℓ_ν = 50, ℓ_mν = 20, reltol 1e-5; src/spectra.jl:163-164): north_star 1e-4 on P(k)."""
    from bolt_b200 import abi
    import bolt_b200 as B
    ks = B.log10_k(10 * cosmo.bg.H0, 5000 * cosmo.bg.H0, 4)
    o = abi.make_opts(50, 50, 20, reltol=1e-5, abstol=1e-6)
    pk, st, ns = dev.plin(ks, o)
    rk, rst, rns = oracle.plin(ks, o)
    assert np.all(st == 0) and np.all(pk > 0)
    assert np.abs(pk / rk - 1).max() < 1e-4


def test_full_size_c3_properties(cosmo, dev):
    """BASELINE config 3 at full size (2000 quadratic k-modes, ℓ = 2..2500): size-independent properties.
    The oracle would need hours here; parity at this size is carried by determinism, the Cauchy-Schwarz bound of the
    discrete k-sum, invariance under the choice of multipole subset, and agreement with the 100-mode golden spectrum."""
    import bolt_b200 as B
    from bolt_b200 import abi
    bg = cosmo.bg
    ks = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, 2000)
    ells = np.arange(2, 2501, dtype=np.int32)
    o = abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6)
    args = (0.01 * bg.H0, 1000 * bg.H0, 5000, cosmo.ix_start)
    tt, te, ee, st, ns = dev.spectra(ks, o, ells, *args)
    assert np.all(st == 0) and ns.min() > 100 and ns.max() < 5000
    assert np.all(np.isfinite(tt)) and np.all(tt > 0) and np.all(ee > 0)
    assert np.all(te ** 2 <= tt * ee * (1 + 1e-12))
    tt2, te2, ee2, _, ns2 = dev.spectra(ks, o, ells, *args)
    assert np.array_equal(tt, tt2) and np.array_equal(te, te2) and np.array_equal(ee, ee2) and np.array_equal(ns, ns2)   # bitwise
    sub = ells[7::97]
    stt, ste, see, _, _ = dev.spectra(ks, o, sub, *args)
    assert np.allclose(stt, tt[7::97], rtol=1e-13) and np.allclose(see, ee[7::97], rtol=1e-13)
    # the acoustic peaks are where they should be and the 100-mode spectrum is within its own k-interpolation error
    g = load_golden("oracle_c1.npz")
    sel = np.searchsorted(ells, g["ell"])
    assert np.abs(tt[sel] / g["tt"] - 1)[3:].max() < 0.08
    dl = ells * (ells + 1) * tt
    assert 180 <= ells[np.argmax(dl[:400])] <= 260


def test_c4_style_high_lgamma(cosmo, dev):
    """BASELINE config 4 in miniature: ℓᵧ = 50 through source_grid's interface (state n = 281, generic kernel path).
    Raising the photon truncation must not move the spectrum by more than the truncation error of ℓᵧ = 8."""
    import bolt_b200 as B
    from bolt_b200 import abi
    bg = cosmo.bg
    ks = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, 150)
    ells = np.array([10, 100, 220, 540, 800, 1500], dtype=np.int32)
    args = (0.01 * bg.H0, 1000 * bg.H0, 5000, cosmo.ix_start)
    lo = dev.spectra(ks, abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6), ells, *args)
    hi = dev.spectra(ks, abi.make_opts(50, 8, 10, reltol=1e-11, abstol=1e-6), ells, *args)
    assert np.all(hi[3] == 0) and abi.state_dim(50, 8, 10, 15) == 281
    assert np.abs(hi[0] / lo[0] - 1).max() < 0.03 and np.abs(hi[2] / lo[2] - 1).max() < 0.06


def test_spectra_batch_equals_per_cosmology_calls(cosmo, cosmo_nonu, gpu_ctx):
    """bolt_spectra_batch (one K1 launch over all cosmologies) returns exactly what bolt_spectra returns per cosmology:
    a k-mode's solve does not depend on what else is in the queue, and K2 runs per cosmology either way."""
    import bolt_b200 as B
    from bolt_b200 import abi, capi
    cs = [cosmo, cosmo_nonu, cosmo]
    dcs = [capi.DeviceCosmo(gpu_ctx, c.hc) for c in cs]
    nk = 48
    ks = np.stack([B.quadratic_k(0.1 * c.bg.H0, 1000 * c.bg.H0, nk) for c in cs])
    ells = np.array([2, 10, 50, 200, 600, 1200, 2000], dtype=np.int32)
    o = abi.make_opts(8, 8, 10, reltol=1e-7, abstol=1e-6)
    kmin = np.array([0.01 * c.bg.H0 for c in cs]); kmax = np.array([1000 * c.bg.H0 for c in cs])
    ix0 = int(np.argmax(cosmo.bg.x_grid > -8))
    tt, te, ee, st, ns = capi.spectra_batch(gpu_ctx, dcs, ks, o, ells, kmin, kmax, 800, ix0)
    assert tt.shape == (3, len(ells)) and st.shape == (3, nk)
    for i, dc in enumerate(dcs):
        t1, e1, p1, s1, n1 = dc.spectra(ks[i], o, ells, kmin[i], kmax[i], 800, ix0)
        assert np.array_equal(st[i], s1) and np.array_equal(ns[i], n1)
        assert np.array_equal(tt[i], t1) and np.array_equal(te[i], e1) and np.array_equal(ee[i], p1)
    assert not np.array_equal(tt[0], tt[1]) and np.array_equal(tt[0], tt[2])
    with pytest.raises(capi.BoltError):
        capi.spectra_batch(gpu_ctx, dcs * 6, np.tile(ks, (6, 1)), o, ells, np.tile(kmin, 6), np.tile(kmax, 6), 800, ix0)   # > BOLT_MAX_BATCH
