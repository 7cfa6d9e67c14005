"""K1 parity (through the C ABI) against the CPU oracle and the committed oracle vectors."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def hist_err(a, b):
    """max |a-b| relative to each variable's own maximum over the history."""
    sc = np.abs(b).max(axis=0) + 1e-300
    return (np.abs(a - b) / sc).max()


@pytest.mark.parametrize("trunc,dt,kfac", [((8, 8, 10), 0.005, [0.3, 30.0, 400.0]),
                                           ((10, 8, 10), 0.01, [5.0]),
                                           ((12, 9, 7), 0.01, [60.0])])
def test_fixed_step_histories_match_oracle(cosmo, oracle, dev, trunc, dt, kfac):
    """north_star: in fixed-step mode per-k perturbation histories agree within 1e-8 relative."""
    from bolt_b200 import abi
    o = abi.make_opts(*trunc, fixed_dt=dt)
    ks = np.array(kfac) * cosmo.bg.H0
    g = dev.solve(ks, o, want=("S_T", "S_P", "u_hist", "u_final"))
    r = oracle.solve(ks, o, want=("S_T", "S_P", "u_hist", "u_final"))
    assert np.all(g["status"] == 0) and np.array_equal(g["nsteps"], r["nsteps"])
    for i in range(len(ks)):
        assert hist_err(g["u_hist"][i], r["u_hist"][i]) < 1e-8
        assert np.abs(g["u_final"][i] - r["u_final"][i]).max() < 1e-8 * np.abs(r["u_final"][i]).max()
        for key in ("S_T", "S_P"):
            a, b = g[key][i, :-1], r[key][i, :-1]          # last S_P row is ±Inf by construction
            assert np.abs(a - b).max() < 1e-8 * np.abs(b).max()


def test_fixed_step_plin_truncation(cosmo, oracle, dev):
    """n = 473 (ℓᵧ = ℓ_ν = 50, ℓ_mν = 20): the plin / CLASS-test state size."""
    from bolt_b200 import abi
    o = abi.make_opts(50, 50, 20, fixed_dt=0.02)
    ks = np.array([40.0]) * cosmo.bg.H0
    g = dev.solve(ks, o, want=("u_hist",)); r = oracle.solve(ks, o, want=("u_hist",))
    assert g["status"][0] == 0
    assert hist_err(g["u_hist"][0], r["u_hist"][0]) < 1e-8


def test_adaptive_sources_match_golden_c1(cosmo, dev):
    """BASELINE config 1 (100 quadratic k-modes, ℓᵧ = 8, reltol 1e-11): source grids vs the committed oracle output."""
    from bolt_b200 import abi
    g = load_golden("oracle_c1.npz")
    ix0 = int(g["ix_start"])
    o = abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6)
    out = dev.solve(g["k"], o, want=("S_T", "S_P"))
    assert np.all(out["status"] == 0)
    # same controller, same arithmetic up to rounding: step counts agree (allow a stray accept/reject flip)
    assert np.abs(out["nsteps"] - g["nsteps"]).max() <= 2
    for key in ("S_T", "S_P"):
        a, b = out[key][:, ix0:-1], g[key][:, :-1]
        err = np.abs(a - b).max(axis=1) / np.abs(b).max(axis=1)
        # Adaptive parity is tolerance-level (abstol 1e-6): identical step sequences agree to ~1e-9; a mode whose
        # accept/reject decision flips on a rounding difference shifts its step sequence and agrees to ~1e-4.
        assert np.median(err) < 1e-6 and np.sort(err)[-4] < 1e-5 and err.max() < 1e-3, (key, err.argmax(), err.max())
    assert not np.isfinite(out["S_P"][:, -1]).any()        # y = 0 at x = 0 (perturbations.jl:401-403)


def test_ix_first_skips_early_rows(cosmo, dev):
    from bolt_b200 import abi
    ks = np.array([10.0, 200.0]) * cosmo.bg.H0
    full = dev.solve(ks, abi.make_opts(8, 8, 10, reltol=1e-8, abstol=1e-6), want=("S_T",))
    part = dev.solve(ks, abi.make_opts(8, 8, 10, reltol=1e-8, abstol=1e-6, ix_first=1201), want=("S_T",))
    assert np.all(part["S_T"][:, :1201] == 0)
    assert np.array_equal(part["S_T"][:, 1201:], full["S_T"][:, 1201:])


def test_deterministic_and_order_independent(cosmo, dev):
    from bolt_b200 import abi
    o = abi.make_opts(8, 8, 10, reltol=1e-9, abstol=1e-6)
    ks = np.array([3.0, 700.0, 50.0, 120.0]) * cosmo.bg.H0
    a = dev.solve(ks, o, want=("S_T",)); b = dev.solve(ks[::-1].copy(), o, want=("S_T",))
    assert np.array_equal(a["S_T"], b["S_T"][::-1]) and np.array_equal(a["nsteps"], b["nsteps"][::-1])


def test_max_steps_and_bad_arguments(cosmo, dev):
    from bolt_b200 import abi, capi
    ks = np.array([500.0]) * cosmo.bg.H0
    out = dev.solve(ks, abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6, max_steps=50), want=("u_final",))
    assert out["status"][0] == abi.K_MAXSTEPS and out["nsteps"][0] + out["nreject"][0] == 50
    with pytest.raises(capi.BoltError):
        dev.solve(ks, abi.make_opts(2, 8, 10), want=("u_final",))           # ℓᵧ < 3: source_function needs Θ₃
    with pytest.raises(capi.BoltError):
        dev.solve(ks, abi.make_opts(8, 8, 10, reltol=0.0), want=("u_final",))


def test_rsa_flagged_for_out_of_envelope_k(cosmo, dev):
    """The in-RHS RSA switch (perturbations.jl:216) is outside the supported envelope of the implicit stages;
    the kernel reports it instead of silently integrating something else."""
    from bolt_b200 import abi
    out = dev.solve(np.array([200.0]), abi.make_opts(8, 8, 10, fixed_dt=0.5), want=("u_final",))
    assert out["status"][0] == abi.K_RSA_TRIGGERED


def test_other_momentum_grid_and_truncations(gpu_ctx):
    """Generality of the kernel beyond the reference defaults: nq = 8 momentum nodes (Background(par; nq=8)), uneven truncations,
    heavier neutrinos -- the generic (runtime-loop) path against the oracle."""
    import bolt_b200 as B
    import hostgen as HG
    from bolt_b200 import abi, capi
    from hostgen import constants as K
    from oracle.oracle import OracleCosmo
    par = B.CosmoParams(Σm_ν=0.3 * K.mass_natural, h=0.65, Ω_c=0.27)
    bg = HG.Background(par, nq=8)
    ih = HG.IonizationHistory(HG.RECFAST(bg, OmegaB=par.Ω_b, Yp=par.Y_p, OmegaG=par.Ω_r), par, bg)
    hc = abi.HostCosmo.from_host(par, bg, ih)
    assert hc.nq == 8
    dc = capi.DeviceCosmo(gpu_ctx, hc); oc = OracleCosmo(hc)
    ks = np.array([1.0, 90.0]) * bg.H0
    o = abi.make_opts(6, 4, 5, fixed_dt=0.01)
    g = dc.solve(ks, o, want=("S_T", "S_P", "u_hist")); r = oc.solve(ks, o, want=("S_T", "S_P", "u_hist"))
    assert g["u_hist"].shape[2] == abi.state_dim(6, 4, 5, 8) == 72
    for i in range(2):
        assert hist_err(g["u_hist"][i], r["u_hist"][i]) < 1e-8
        assert np.abs(g["S_T"][i] - r["S_T"][i]).max() < 1e-8 * np.abs(r["S_T"][i]).max()


def test_long_chain_path_matches_generic_kernel(cosmo, dev, monkeypatch):
    """The runtime-truncation path (plin 50/50/20, C4 50/8/10) against the first-generation generic kernel
    (BOLT_K1_GENERIC=1, read at every launch).  Fixed step: same step sequence by construction, results to rounding.
    Adaptive: the two kernels round the error norm differently (reciprocal vs division, summation order), so the step
    sequences may part at a borderline accept/reject; the solutions agree at tolerance level."""
    from bolt_b200 import abi
    ks = np.array([0.5, 20.0, 300.0, 900.0]) * cosmo.bg.H0

    def both(o):
        monkeypatch.delenv("BOLT_K1_GENERIC", raising=False)
        a = dev.solve(ks, o, want=("S_T", "S_P", "u_final"))
        monkeypatch.setenv("BOLT_K1_GENERIC", "1")
        b = dev.solve(ks, o, want=("S_T", "S_P", "u_final"))
        monkeypatch.delenv("BOLT_K1_GENERIC", raising=False)
        return a, b

    def worst(a, b):
        w = 0.0
        for i in range(len(ks)):
            w = max(w, np.abs(a["u_final"][i] - b["u_final"][i]).max() / np.abs(b["u_final"][i]).max())
            for key in ("S_T", "S_P"):
                x, y = a[key][i, 1201:-1], b[key][i, 1201:-1]
                w = max(w, np.abs(x - y).max() / np.abs(y).max())
        return w

    for trunc in ((50, 8, 10), (50, 50, 20)):
        a, b = both(abi.make_opts(*trunc, fixed_dt=0.01, ix_first=1201))
        assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["nsteps"], b["nsteps"])
        assert worst(a, b) < 1e-8
        a, b = both(abi.make_opts(*trunc, reltol=1e-9, abstol=1e-6, ix_first=1201))
        assert np.all(a["status"] == 0) and np.all(b["status"] == 0)
        assert np.abs(a["nsteps"] - b["nsteps"]).max() <= 0.02 * b["nsteps"].max()
        assert worst(a, b) < 1e-6


@pytest.mark.parametrize("trunc", [(3, 3, 3), (4, 3, 5), (50, 8, 10)])
def test_long_chain_path_edge_truncations_match_oracle(cosmo, oracle, dev, trunc):
    """Shortest chains the library accepts (l_max = 3, source_function needs Θ₃: the truncation row is the first row above the
    three bottom rows) and very uneven chain lengths, fixed step, against the oracle."""
    from bolt_b200 import abi
    o = abi.make_opts(*trunc, fixed_dt=0.01)
    ks = np.array([2.0, 150.0]) * cosmo.bg.H0
    g = dev.solve(ks, o, want=("u_hist",)); r = oracle.solve(ks, o, want=("u_hist",))
    assert np.all(g["status"] == r["status"])
    for i in range(len(ks)):
        assert hist_err(g["u_hist"][i], r["u_hist"][i]) < 1e-8
    from bolt_b200.capi import BoltError
    with pytest.raises(BoltError):            # l_max < 3 is refused, not silently mis-solved
        dev.solve(ks, abi.make_opts(3, 2, 3, fixed_dt=0.01), want=("u_hist",))


def test_partial_counts_without_an_instantiation(cosmo, gpu_ctx):
    """Five partials have no kernel instantiation: the binding serves them through the six-partial kernels with an all-zero
    partial and hands back five (tangent directions that are multiples of each other must give multiples).  More than six
    fail loudly instead of silently dropping partials."""
    from bolt_b200 import abi, capi
    hc = cosmo.hc
    sc = np.zeros((abi.NSCALARS, 6)); tb = np.zeros(hc.tables.shape[:2] + (6,))
    sc[:, 0] = hc.scalars[:, 0]; tb[..., 0] = hc.tables[..., 0]
    for j in range(5):
        tb[..., 1 + j] = 1e-3 * (j + 1) * hc.tables[..., 0]           # five non-trivial partials
    dc = capi.DeviceCosmo(gpu_ctx, abi.HostCosmo(sc, hc.quad_pts, hc.quad_wts, tb, hc.x0, hc.dx))
    assert dc.nd_user == 6 and dc.hc.nd == 7
    out = dc.solve(np.array([10.0]) * cosmo.bg.H0, abi.make_opts(8, 8, 10, fixed_dt=0.05), want=("S_T", "u_final"))
    assert out["status"][0] == 0 and out["S_T"].shape[-1] == 6 and out["u_final"].shape[-1] == 6
    ref = out["u_final"][..., 1]
    assert np.abs(ref).max() > 0
    for j in range(1, 5):
        assert np.abs(out["u_final"][..., 1 + j] - (j + 1) * ref).max() <= 1e-9 * np.abs(ref).max()
    sc8 = np.zeros((abi.NSCALARS, 8)); tb8 = np.zeros(hc.tables.shape[:2] + (8,))
    sc8[:, 0] = hc.scalars[:, 0]; tb8[..., 0] = hc.tables[..., 0]
    with pytest.raises(capi.BoltError, match="partials"):
        capi.DeviceCosmo(gpu_ctx, abi.HostCosmo(sc8, hc.quad_pts, hc.quad_wts, tb8, hc.x0, hc.dx))
