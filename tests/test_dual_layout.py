"""Host-side packing of forward-mode partials (nd > 1): layout across the ABI and central-difference construction."""
import numpy as np


def test_with_partials_layout_and_values(cosmo):
    from bolt_b200 import abi
    hc = cosmo.hc
    rng = np.random.default_rng(20261017)
    def perturbed(eps):
        return abi.HostCosmo(hc.scalars * (1 + eps), hc.quad_pts, hc.quad_wts, hc.tables * (1 + eps), hc.x0, hc.dx)
    d1, d2 = 1e-3, 5e-4
    dual = abi.HostCosmo.with_partials(hc, [(perturbed(d1), perturbed(-d1)), (perturbed(2 * d2), perturbed(-2 * d2))], [d1, d2])
    assert dual.nd == 3 and dual.desc.nd == 3
    assert dual.tables.shape == (abi.NTABLES, hc.n_x + 2, 3) and dual.scalars.shape == (abi.NSCALARS, 3)
    # value first, exactly the base values (bit-identical to a Vector{Dual}: value, partial_1, partial_2)
    assert np.array_equal(dual.tables[..., 0], hc.tables[..., 0]) and np.array_equal(dual.scalars[:, 0], hc.scalars[:, 0])
    assert np.allclose(dual.tables[..., 1], hc.tables[..., 0], rtol=1e-8, atol=1e-300)          # d/d(eps) of T(1+eps) = T
    assert np.allclose(dual.tables[..., 2], 2 * hc.tables[..., 0], rtol=1e-8, atol=1e-300)
    assert dual.tables.flags["C_CONTIGUOUS"] and dual.tables.strides[-1] == 8               # nd is the fastest axis


def test_amplitude_and_tilt_partials_need_no_host_rerun(cosmo, monkeypatch):
    import hostgen.partials as api
    from bolt_b200 import abi
    calls = []
    real = api.Background
    monkeypatch.setattr(api, "Background", lambda p, **kw: (calls.append(1), real(p, **kw))[1])
    dual, base, bg, ih, pm, steps = api.host_cosmo_with_partials(cosmo.par, ["A", "n"], rel_step=1e-3)
    assert len(calls) == 1                                   # only the base cosmology ran the host pipeline
    assert dual.scalars[abi.S["A"], 1] == 1.0 or abs(dual.scalars[abi.S["A"], 1] - 1.0) < 1e-9
    assert abs(dual.scalars[abi.S["n"], 2] - 1.0) < 1e-9 and dual.scalars[abi.S["A"], 2] == 0.0
    assert not dual.tables[..., 1:].any()                    # the tables do not depend on A or n


def test_partial_counts_without_an_instantiation_are_padded():
    """5 partials per call are served by the 6-partial kernels with an all-zero partial (HostCosmo.padded)."""
    import numpy as np
    from bolt_b200 import abi
    rng = np.random.default_rng(3)
    nd, n_x, nq = 6, 7, 4
    hc = abi.HostCosmo(rng.random((abi.NSCALARS, nd)), rng.random(nq), rng.random(nq), rng.random((abi.NTABLES, n_x + 2, nd)), -20.0, 0.01)
    p = hc.padded(7)
    assert p.nd == 7 and p.n_x == n_x and p.desc.nd == 7
    assert np.array_equal(p.scalars[:, :6], hc.scalars) and np.all(p.scalars[:, 6] == 0)
    assert np.array_equal(p.tables[:, :, :6], hc.tables) and np.all(p.tables[:, :, 6] == 0)
