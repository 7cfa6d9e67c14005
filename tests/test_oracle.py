"""The CPU oracle against the reference's own golden vectors (SURVEY 8c pins) and its internal consistency."""
import numpy as np
import pytest
from scipy.interpolate import CubicSpline
from scipy.special import spherical_jn

from conftest import load_golden


def test_oracle_vs_class_phi_deltab(cosmo_nonu):
    """test/runtests.jl:83-147: Φ and δ_b within 1e-3 of CLASS at k = 0.03 h/Mpc,
    ℓᵧ = ℓ_ν = 50, ℓ_mν = 20, nq = 15, reltol = abstol = 1e-9."""
    from bolt_b200 import abi
    from oracle.oracle import OracleCosmo
    g = load_golden("class_px.npz")
    c = cosmo_nonu
    oc = OracleCosmo(c.hc)
    k = c.par.h * float(g["k_p03"])
    o = abi.make_opts(50, 50, 20, reltol=1e-9, abstol=1e-9)
    out = oc.solve(np.array([k]), o, want=("u_hist",))
    assert out["status"][0] == 0
    uh = out["u_hist"][0]; n = uh.shape[1]
    cx = g["x_p03"][::-1]
    phi = CubicSpline(c.bg.x_grid, uh[:, n - 5])(cx)
    d_b = CubicSpline(c.bg.x_grid, uh[:, n - 2])(cx)
    assert np.all(np.abs(phi / g["phi_p03"][::-1] - 1) < 1e-3)
    assert np.all(np.abs(-d_b / g["d_b_p03"][::-1] - 1) < 1e-3)


def _mnu_check(cosmo, uh):
    """Φ, δ_b, δ_cdm and the massive-neutrino density contrast against CLASS (ncdm, no fluid approximation, reionization)."""
    from bolt_b200.params import q_grid, f0, dxdq
    g = load_golden("class_px_mnu.npz")
    par, bg = cosmo.par, cosmo.bg
    xg = bg.x_grid; n = uh.shape[1]
    cx = g["x"][::-1]; keep = cx > -12
    def rel(mine, ref, sign=1.0):
        return np.abs(sign * CubicSpline(xg, mine)(cx)[keep] / ref[::-1][keep] - 1).max()
    assert rel(uh[:, n - 5], g["phi"]) < 1.5e-3
    assert rel(uh[:, n - 2], g["d_b"], -1.0) < 1.5e-3
    assert rel(uh[:, n - 4], g["d_cdm"], -1.0) < 1.5e-3
    q, lqmi, lqma = q_grid(par, bg.quad_pts); w = f0(q, par) / dxdq(q, lqmi, lqma) * bg.quad_wts
    eps = np.sqrt(q ** 2 + (np.exp(xg)[:, None] * par.Σm_ν) ** 2)
    iM = 2 * 51 + 51
    rho = 4 * np.pi * np.sum(q ** 2 * eps * w * uh[:, iM:iM + 15], axis=1)       # ρ_σ (perturbations.jl:127-145)
    d_ncdm = rho / bg.ρ0M(xg) * np.exp(-4 * xg)                                     # scripts/plot_perts_x.jl:70
    assert rel(d_ncdm, g["d_ncdm"], -1.0) < 3e-2


def test_oracle_vs_class_massive_neutrinos(cosmo, oracle):
    """Extra pin (no reference test consumes it; scripts/plot_perts_x.jl plots it): default CosmoParams (Σm_ν = 0.06 eV),
    k = 0.03 h/Mpc, ℓᵧ = ℓ_ν = 50, ℓ_mν = 20, reltol 1e-8."""
    from bolt_b200 import abi
    g = load_golden("class_px_mnu.npz")
    out = oracle.solve(np.array([cosmo.par.h * float(g["k"])]), abi.make_opts(50, 50, 20, reltol=1e-8, abstol=1e-8), want=("u_hist",))
    assert out["status"][0] == 0
    _mnu_check(cosmo, out["u_hist"][0])


def camb_pk_nodes(kmin_h=1e-3, kmax_h=0.5, every=1):
    """CAMB's own k nodes [h/Mpc] and P [(Mpc/h)³]: the table has only 50 log-spaced points, so interpolating it across the
    BAO wiggles costs up to 3 % -- the pin is taken AT the nodes."""
    g = load_golden("camb_pk.npz")
    m = (g["k_h"] >= kmin_h) & (g["k_h"] <= kmax_h)
    return g["k_h"][m][::every], g["pk_h3"][m][::every]


def test_oracle_plin_vs_camb(cosmo, oracle):
    """Extra pin: the reference has no automated test of plin; its data directory holds CAMB's z = 0 matter power
    spectrum (scripts/first_plin.jl plots the ratio).  plin defaults (ℓᵧ = ℓ_ν = 50, ℓ_mν = 20, reltol 1e-5), P in (Mpc/h)³.
    Measured: within 0.24 % at all 27 nodes in 1e-3 ≤ k ≤ 0.5 h/Mpc."""
    from bolt_b200 import abi
    kh, pk_camb = camb_pk_nodes(every=3)
    pk, st, _ = oracle.plin(kh * cosmo.par.h, abi.make_opts(50, 50, 20, reltol=1e-5, abstol=1e-6))
    assert np.all(st == 0)
    assert np.abs(pk * cosmo.par.h ** 3 / pk_camb - 1).max() < 4e-3


def test_oracle_cl_vs_camb(cosmo, oracle):
    """test/runtests.jl:149-185: D_ℓ^TT, D_ℓ^EE within 11 % of CAMB for ℓ = 10:10:2500 (100 quadratic k-modes).
    The source grids are the committed oracle outputs (tests/golden/make_golden.py); the projection is re-run."""
    g = load_golden("oracle_c1.npz"); camb = load_golden("camb_cl.npz")
    ix0 = int(g["ix_start"]); n_x = cosmo.hc.n_x
    S_T = np.zeros((len(g["k"]), n_x)); S_P = np.zeros_like(S_T)
    S_T[:, ix0:] = g["S_T"]; S_P[:, ix0:] = g["S_P"]
    ells = g["ell"]
    bg = cosmo.bg
    tt, te, ee = oracle.project(S_T, S_P, g["k"], ells, 0.01 * bg.H0, 1000 * bg.H0, 5000, ix0)
    assert np.allclose(tt, g["tt"], rtol=1e-12) and np.allclose(ee, g["ee"], rtol=1e-12) and np.allclose(te, g["te"], rtol=1e-10, atol=1e-30)
    lf = ells * (ells + 1) / (2 * np.pi)
    assert np.all(np.abs(lf * tt / camb["tt"] - 1) < 1.1e-1)
    assert np.all(np.abs(lf * ee / camb["ee"] - 1) < 1.1e-1)
    # Cauchy-Schwarz holds term by term for the discrete k-sum
    assert np.all(te ** 2 <= tt * ee * (1 + 1e-12))


def test_oracle_reproduces_golden_source_columns(cosmo, oracle):
    from bolt_b200 import abi
    g = load_golden("oracle_c1.npz")
    ix0 = int(g["ix_start"])
    sel = np.array([3, 40])
    o = abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6)
    out = oracle.solve(g["k"][sel], o, want=("S_T", "S_P"))
    assert np.array_equal(out["nsteps"], g["nsteps"][sel])
    for key in ("S_T", "S_P"):
        a, b = out[key][:, ix0:-1], g[key][sel][:, :-1]
        assert np.max(np.abs(a - b)) <= 1e-9 * np.max(np.abs(b))


def test_lu_modes_agree(oracle, cosmo):
    """Skipping structural zeros must not change the arithmetic."""
    from bolt_b200 import abi
    o = abi.make_opts(8, 8, 10, fixed_dt=0.05)
    k = np.array([20 * cosmo.bg.H0])
    a = oracle.solve(k, o, want=("u_final",), lu_mode=0)["u_final"]
    b = oracle.solve(k, o, want=("u_final",), lu_mode=1)["u_final"]
    assert np.max(np.abs(a - b)) <= 1e-13 * np.max(np.abs(a))


def test_stepper_converges(oracle, cosmo):
    """Fixed-step KenCarp4 converges under step halving.  The early epoch is extremely stiff (|τ′R| ~ 1e14), where
    an ESDIRK method shows its stage order (2) rather than its classical order (4): require at least ratio 4."""
    from bolt_b200 import abi
    k = np.array([0.5 * cosmo.bg.H0])
    sols = []
    for dt in (0.1, 0.05, 0.025):
        o = abi.make_opts(8, 8, 10, fixed_dt=dt)
        sols.append(oracle.solve(k, o, want=("u_final",))["u_final"][0, -5:])
    e1 = np.abs(sols[0] - sols[2]).max(); e2 = np.abs(sols[1] - sols[2]).max()
    assert e1 / e2 > 4.0 and e2 < 1e-4


def test_rhs_is_linear_and_ic_on_constraint(oracle, cosmo):
    from bolt_b200 import abi
    o = abi.make_opts(8, 8, 10)
    k = 50 * cosmo.bg.H0
    u0 = oracle.initial_conditions(k, o)
    rng = np.random.default_rng(20261017)
    v = rng.standard_normal(len(u0)) * np.abs(u0).max()
    x = -7.3
    f = lambda u: oracle.hierarchy(k, o, x, u)[0]
    lhs = f(2.0 * u0 + 3.0 * v); rhs = 2.0 * f(u0) + 3.0 * f(v)
    assert np.max(np.abs(lhs - rhs)) <= 1e-12 * np.max(np.abs(rhs))
    # adiabatic ICs: δ = δ_b = 3Θ₀, v = v_b (perturbations.jl:308-312)
    assert u0[-4] == u0[-2] == 3 * u0[0] and u0[-3] == u0[-1]


def test_rsa_switch_is_dead_for_configured_k(oracle, cosmo):
    """the in-RHS RSA branch never fires for k <= 5000 H0."""
    from bolt_b200 import abi
    o = abi.make_opts(8, 8, 10)
    k = 5000 * cosmo.bg.H0
    u0 = oracle.initial_conditions(k, o)
    for x in np.linspace(-20, 0, 41):
        assert oracle.hierarchy(k, o, x, u0)[2] is False
    # ... and does fire (mutating u, zeroing radiation derivatives) for a far larger k at early times
    kbig = 200.0
    du, u, rsa = oracle.hierarchy(kbig, o, -12.0, u0)
    assert rsa and np.all(du[:27] == 0) and u[2] == 0 and u[18] == u0[-5]


def test_oracle_bessel_accuracy():
    from oracle.oracle import lib
    L = lib()
    for l in (2, 10, 100, 700, 2500):
        for x in (0.3, 5.0, 50.0, 699.5, 1500.0, 3377.5):
            ref = spherical_jn(l, x)
            got = L.oracle_sph_bessel_j(l, x)
            assert abs(got - ref) <= 1e-11 * max(abs(ref), 1e-3 / max(x, 1.0)), (l, x, got, ref)
