"""Batched input tables on the device (bolt_hostgen_batch) against the harness generator hostgen/ (which is pinned
by the reference's Fortran RECFAST fixture, tests/test_host_inputs.py), and end to end: spectra from device-made tables against
spectra from host-made tables."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def host_cosmo(par):
    import hostgen as HG
    from bolt_b200 import abi
    bg = HG.Background(par)
    ih = HG.IonizationHistory(HG.RECFAST(bg, OmegaB=par.Ω_b, Yp=par.Y_p, OmegaG=par.Ω_r), par, bg)
    return abi.HostCosmo.from_host(par, bg, ih), bg


def sampled(hc, t):
    """values of table t at the knots (from the coefficients: c[i]/6 + 2c[i+1]/3 + c[i+2]/6)"""
    c = hc.tables[t, :, 0]
    return c[:-2] / 6 + 2 * c[1:-1] / 3 + c[2:] / 6


def test_device_tables_match_the_host_generator(gpu_ctx):
    import bolt_b200 as B
    from bolt_b200 import abi, capi
    from hostgen import constants as K
    pars = [B.CosmoParams(), B.CosmoParams(h=0.62, Ω_b=0.052, Ω_c=0.29, Σm_ν=0.2 * K.mass_natural, Y_p=0.25)]
    dev, st = capi.hostgen_batch(pars)
    assert np.all(st == 0)
    for par, d in zip(pars, dev):
        h, bg = host_cosmo(par)
        assert np.allclose(d.scalars[:, 0], h.scalars[:, 0], rtol=1e-12)
        T = abi.T
        for nm, tol in (("H", 1e-12), ("Hp", 1e-9), ("Hpp", 1e-7), ("η", 1e-12), ("ρ0M", 1e-12)):
            a, b = sampled(d, T[nm]), sampled(h, T[nm])
            assert np.abs(a - b).max() <= tol * np.abs(b).max(), nm
        # ionization history: two different integrators (Dormand-Prince on the grid vs DOP853 + dense output), tolerance-level agreement;
        # the reference's own fixture tolerance for X_e is 1e-4 (test/runtests.jl:47)
        for nm, tol in (("τ", 2e-6), ("τp", 2e-6), ("g", 2e-5), ("csb2", 2e-6)):
            a, b = sampled(d, T[nm]), sampled(h, T[nm])
            assert np.abs(a - b).max() <= tol * np.abs(b).max(), (nm, np.abs(a - b).max() / np.abs(b).max())


def test_spectra_from_device_tables_match_host_tables(gpu_ctx):
    import bolt_b200 as B
    from bolt_b200 import abi, capi
    par = B.CosmoParams()
    (d,), st = capi.hostgen_batch([par])
    h, bg = host_cosmo(par)
    k = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, 200)
    ells = np.arange(2, 2001, 37, dtype=np.int32)
    o = abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6)
    ix0 = int(np.argmax(bg.x_grid > -8))
    a = capi.DeviceCosmo(gpu_ctx, d).spectra(k, o, ells, 0.01 * bg.H0, 1000 * bg.H0, 5000, ix0)
    b = capi.DeviceCosmo(gpu_ctx, h).spectra(k, o, ells, 0.01 * bg.H0, 1000 * bg.H0, 5000, ix0)
    assert np.all(a[3] == 0) and np.all(b[3] == 0)
    for i in (0, 2):
        assert np.abs(a[i] / b[i] - 1).max() < 1e-4           # the north-star C_l tolerance


def test_batch_is_independent_of_its_composition(gpu_ctx):
    import bolt_b200 as B
    from bolt_b200 import capi
    pars = [B.CosmoParams(h=0.6 + 0.02 * i, Ω_c=0.2 + 0.01 * i) for i in range(5)]
    all5, st = capi.hostgen_batch(pars)
    one, _ = capi.hostgen_batch([pars[3]])
    assert np.all(st == 0)
    assert np.array_equal(all5[3].tables, one[0].tables) and np.array_equal(all5[3].scalars, one[0].scalars)
