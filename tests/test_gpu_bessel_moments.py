"""Row n2 on the device: the reference's own test file (test/testbessel.jl) replayed through the C ABI with the reference's
tolerances, then the device against the oracle on seeded inputs, then size-independent properties of the full 2,000,000-node
table."""
import os
import sys

import mpmath
import numpy as np
import pytest

from oracle import bessel_moments_oracle as O  # noqa: E402
from test_bessel_moments import (J_MODERATE, MAC_001, MAC_0001, OCT_200, QUAD_200, REFS_4TH_1, REFS_4TH_1000, TOL, big, rel)  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    import bolt_b200.bessel as B
    return B


@pytest.fixture(scope="module")
def full_tables(B):
    """Exactly the tables the reference's tests build: (0, 1.6e4) with 2,000,000 nodes."""
    return {(nu, order): B.sph_bessel_interpolator(nu, order, 0.0, 1.6e4, 2_000_000) for (nu, order) in ((2, 3), (3, 3), (3, 4))}


def test_J_moments_moderate_argument(B):                      # testbessel.jl:9-33
    for alpha, ref in zip((0, 1, 2), J_MODERATE):
        assert rel(B.J_moment_asymp(200.0, 2.5, alpha - 0.5), ref) < TOL
        assert rel(B.J_moment_asymp_nu_five_halves(200.0, alpha - 0.5), ref) < TOL
        assert rel(B.J_moment_weniger_1F2(200.0, 2.5, alpha), ref) < TOL


def test_J_moments_small_argument(B):                         # :36-50; the reference asks 1e-15 of its Double64 sum, the device returns f64
    assert rel(B.J_moment_weniger_1F2(10.0, 2.5, 1.5), big("-0.98904817846028826228408967797229")) < 2e-15
    assert rel(B.J_moment_weniger_1F2(0.1, 2.5, -0.5), big("0.00001772317062480308")) < TOL


@pytest.mark.parametrize("nu,refs", [(2, QUAD_200), (3, OCT_200)])
def test_sph_moments_large_argument(B, nu, refs):             # :52-77, :100-125
    closed = B.sph_j_moment_asymp_nu_2 if nu == 2 else B.sph_j_moment_asymp_nu_3
    for m, ref in zip((0, 1, 2), refs):
        assert rel(B.sph_j_moment_asymp(200.0, nu, m), ref) < TOL
        assert rel(closed(200.0, m, B.sph_j_moment_asymp_prefactor(nu, m)), ref) < TOL
        assert rel(B.sph_j_moment_weniger_1F2(200.0, nu, m, B.WenigerCache1F2(float)), ref) < TOL


def test_sph_moments_small_argument_quadrupole(B):            # :79-97
    assert rel(B.sph_j_moment_weniger_1F2(0.1, 2, 1), big("1.66587318119689113603548520948514e-6")) < TOL
    assert rel(B.sph_j_moment_weniger_1F2(0.1, 2, 0), big("0.0000222127003021204915961147418458393")) < TOL
    assert rel(B.sph_j_moment_weniger_1F2(0.01, 2, 2), big("1.33332653062694211665894244946280e-12")) < TOL


def test_maclaurin_vanishing_argument_octupole(B):            # :128-153
    for x, refs in ((0.01, MAC_001), (0.001, MAC_0001)):
        for m in range(3):
            assert rel(B.sph_j_moment_maclaurin_1F2(x, 3, m), refs[m]) < TOL


def test_interpolator_large_argument(full_tables):            # :157-184
    for nu, refs in ((2, QUAD_200), (3, OCT_200)):
        moms = full_tables[(nu, 3)](200.0)
        for i in range(3):
            assert rel(moms[i], refs[i]) < 1e-12


def test_interpolator_small_argument_octupole(B, full_tables):        # :186-198
    refs = (2.380070697034449e-7, 1.904006180460421e-8, 1.586640331877486e-9)
    itp = full_tables[(3, 3)]
    assert B.getorder(itp) == 3 and B.getnu(itp) == 3
    moms = itp(0.1)
    for i in range(3):
        assert abs(moms[i] - refs[i]) < 1e-12


def test_interpolator_fourth_order(B, full_tables):           # :201-239
    itp = full_tables[(3, 4)]
    assert B.getorder(itp) == 4
    moms = itp(1.0)
    for i in range(4):
        assert abs(moms[i] - REFS_4TH_1[i]) < 1e-12
    moms = B.sph_bessel_interpolator(3, 4, 2.0, 1.6e4, 20)(1.0)          # Maclaurin branch
    for i in range(4):
        assert abs(moms[i] - REFS_4TH_1[i]) < 1e-12
    moms = itp(1000.0)
    for i in range(4):
        assert rel(moms[i], REFS_4TH_1000[i]) < 1e-12
    moms = B.sph_bessel_interpolator(3, 4, 0.0, 500.0, 20)(1000.0)       # Lommel branch
    for i in range(4):
        assert rel(moms[i], REFS_4TH_1000[i]) < 1e-12


def test_interpolator_maclaurin_branch(B):                    # :242-271
    itp = B.sph_bessel_interpolator(3, 3, 1.0, 1.6e4, 2_000_000)
    for x, refs in ((0.01, MAC_001), (0.001, MAC_0001)):
        for moms in (B.sph_j_moment_maclaurin_all_orders(itp, x), itp(x)):
            for m in range(3):
                assert rel(moms[m], refs[m]) < TOL


def test_filon_third_order(B, full_tables):                   # :275-305
    itp = full_tables[(3, 3)]
    refs = [big("0.365287615501162668736682652658444"), big("0.219009396999160523658045931736310"),
            big("0.146278218502002145078636720922135"), big("0.000548758594228308158105260833682184")]
    F = B.integrate_sph_bessel_filon
    assert abs(F(4., -0.2, 6.0, 10.0, 0., 2., itp) - refs[0]) < TOL
    assert abs(F(4., -0.2, 6.0, 10.0, 0., 1., itp) - refs[1]) < TOL
    assert abs(F(6.8, 5.8, 6.0, 10.0, 1., 2., itp) - refs[2]) < TOL
    assert abs(F(15.6, 11.8, 6., 10.0, 2., 4., itp) - refs[3]) < TOL
    s, _ = B._loop_integrate_sph_bessel_filon(6.8, 5.8, 6.0, 10.0, 1., 2., itp, itp(10 * 1.))
    assert abs(s - refs[2]) < TOL

    itp = B.sph_bessel_interpolator(3, 3, 2.0, 50.0, 20)
    refs = [big("1.42547119945725017346111489855163e-6"), big("1.42843411313369992225631279186481e-10"),
            big("3.21068375645105591375937123607342"), big("-9.05287087052870811987403003972375"),
            big("61.7007662060909735341421714015349")]
    assert abs((F(3.9983, -0.14, 6., 10.0, 0.01, 0.02, itp) - refs[0]) / refs[0]) < TOL
    assert abs((F(3999803 / 1000000, -(97 / 500), 6., 10.0, 0.001, 0.002, itp) - refs[1]) / refs[1]) < TOL
    assert abs((F(7494., 299.8, 6., 10.0, 50.0, 100.0, itp) - refs[2]) / refs[2]) < TOL
    assert abs((F(119964., 1199.8, 6., 10., 200., 201., itp) - refs[3]) / refs[3]) < TOL
    assert abs((F(74999004, 149999 / 5, 6., 10., 5000., 5100., itp) - refs[4]) / refs[4]) < TOL


# ---------------------------------------------------------------------------------------------------------------------
# device vs oracle on seeded inputs, and properties at full size
# ---------------------------------------------------------------------------------------------------------------------
def test_small_argument_evaluator_matches_40_digit_sums(B):
    """The device's quadrature-on-a-prefix evaluator against the oracle's 40-digit ₁F₂ over the whole range the table fill
    uses it on (x < 50), every power of both multipoles."""
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.uniform(0.0, 50.0, 300), [1e-6, 1e-3, 3.999999, 4.0, 4.000001, 49.999]])
    for nu in (2, 3):
        got = B._moments(nu, [0, 1, 2, 3], B.SMALL, x)
        for i, xi in enumerate(x):
            for m in range(4):
                ref = O.sph_j_moment_1F2(xi, nu, m)
                # relative to the running magnitude of the integral (the moment itself crosses zero for m >= 2)
                scale = max(abs(ref), mpmath.mpf(xi) ** (m - 1) if xi > 4 else 0)
                assert abs(big(got[i, m]) - ref) < 2e-14 * scale, (nu, m, xi, got[i, m], ref)


def test_asymptotic_matches_oracle(B):
    rng = np.random.default_rng(8)
    x = rng.uniform(50.0, 6e4, 200)
    for nu in (2, 3):
        got = B._moments(nu, [0, 1, 2, 3], B.ASYMP, x)
        f = O.sph_j_moment_asymp_nu_2 if nu == 2 else O.sph_j_moment_asymp_nu_3
        for i, xi in enumerate(x):
            for m in range(4):
                ref = f(xi, m, O.sph_j_moment_asymp_prefactor(nu, m))
                assert abs(got[i, m] - ref) < 1e-13 * max(abs(ref), xi ** (m - 1)), (nu, m, xi)


def test_table_matches_oracle_table(B):
    """Same node spacing as the full table on a range the oracle can fill in seconds; compares the spline BETWEEN nodes, so the
    prefilter (interior convolution + closed ends) is what is being checked."""
    h = 1.6e4 / (2_000_000 - 1)
    n = 20_001
    xmax = h * (n - 1)
    rng = np.random.default_rng(9)
    x = np.concatenate([rng.uniform(0.0, xmax, 400), [0.0, h * 0.5, h * 39.5, h * 40.5, xmax - h * 40.5, xmax - h * 0.5, xmax]])
    for nu, order in ((2, 3), (3, 4)):
        ora = O.MomentTable(nu, order, 0.0, xmax, n)
        dev = B.sph_bessel_interpolator(nu, order, 0.0, xmax, n)
        got = dev.many(x)
        for i, xi in enumerate(x):
            ref = ora(xi)
            # relative to the moment or to the running size of its integrand's envelope x^(m-1), whichever is larger
            scale = np.maximum(np.abs(ref), max(xi, 1e-3) ** (np.arange(order) - 1.0))
            assert np.all(np.abs(got[i] - ref) < 1e-13 * scale), (nu, xi, got[i], ref)


def test_short_table_whole_solve_matches_oracle(B):
    for N in (20, 60, 82, 83, 200):
        ora = O.MomentTable(3, 3, 2.0, 500.0, N)
        dev = B.sph_bessel_interpolator(3, 3, 2.0, 500.0, N)
        x = np.linspace(2.0, 500.0, 57)
        got = dev.many(x)
        ref = np.array([ora(xi) for xi in x])
        assert np.allclose(got, ref, rtol=1e-11, atol=1e-11 * np.abs(ref).max()), N


def test_full_table_properties(B, full_tables):
    """Size-independent checks on the real 2,000,000-node tables: the table reproduces its nodes' defining values (small-argument
    evaluator below 50, Lommel form above), is continuous across both branch seams, and the 4th-order table's first three moments
    equal the 3rd-order table's."""
    t3, t4 = full_tables[(3, 3)], full_tables[(3, 4)]
    h = 1.6e4 / (2_000_000 - 1)
    idx = np.array([0, 1, 39, 40, 41, 1000, 6249, 6250, 6251, 123457, 1999958, 1999959, 1999960, 1999998, 1999999])
    xn = idx * h
    got = t4.many(xn)
    small = B._moments(3, [0, 1, 2, 3], B.SMALL, xn[xn < 50.0])
    big_ = B._moments(3, [0, 1, 2, 3], B.ASYMP, xn[xn >= 50.0])
    ref = np.vstack([small, big_])
    assert np.all(np.abs(got - ref) <= 2e-13 * np.maximum(np.abs(ref), xn[:, None] ** np.arange(4)[None, :] * 1e-3 + 1e-12))
    xs = np.random.default_rng(3).uniform(0, 1.6e4, 5000)
    assert np.allclose(t3.many(xs), t4.many(xs)[:, :3], rtol=0, atol=0)
    # seam at the top of the table: inside (spline) vs outside (asymptotic form)
    a, b = t4.many([1.6e4])[0], t4.many([np.nextafter(1.6e4, np.inf)])[0]
    assert np.all(np.abs(a - b) < 1e-10 * np.abs(b))


def test_filon_chain_equals_sum_of_pieces_and_a_known_integral(B, full_tables):
    """The batched loop form against (i) the piece-by-piece rule and (ii) a closed form: with f ≡ 1 the chain telescopes to
    (I₀(k b) - I₀(k a)) / k whatever the nodes are."""
    itp = full_tables[(2, 3)]
    rng = np.random.default_rng(11)
    nodes = np.sort(rng.uniform(0.0, 14000.0, 700))
    k = rng.uniform(0.01, 1.0, 37)
    c0, c1, c2 = rng.normal(size=(3, len(k), 1))
    f = c0 + c1 * nodes + c2 * nodes ** 2
    f1 = c1 + 2 * c2 * nodes
    f2 = 2 * c2 + 0 * nodes
    got = B.filon_chain(nodes, f, f1, f2, k, itp)
    for i in (0, 17, 36):
        pieces = B.integrate_sph_bessel_filon(f[i, :-1], f1[i, :-1], f2[i, :-1], np.full(len(nodes) - 1, k[i]), nodes[:-1], nodes[1:], itp)
        assert abs(got[i] - pieces.sum()) < 1e-12 * np.abs(pieces).sum()
    one = np.ones_like(f)
    tele = B.filon_chain(nodes, one, 0 * one, 0 * one, k, itp)
    ref = np.array([(itp(kk * nodes[-1])[0] - itp(kk * nodes[0])[0]) / kk for kk in k])
    assert np.allclose(tele, ref, rtol=1e-11, atol=1e-13)


def test_bad_arguments_fail_loudly(B):
    from bolt_b200.capi import BoltError
    with pytest.raises(BoltError):
        B.sph_bessel_interpolator(5, 3, 0.0, 100.0, 1000)        # the reference has no ν = 5 table either (MethodError)
    with pytest.raises(BoltError):
        B.sph_bessel_interpolator(3, 3, 10.0, 1.0, 1000)
    itp = B.sph_bessel_interpolator(3, 2, 0.0, 100.0, 1000)
    with pytest.raises(BoltError):
        B.integrate_sph_bessel_filon(1., 0., 0., 1.0, 0., 1., itp)   # order 2 table cannot carry quadratic pieces


def test_empty_inputs_are_no_ops(B):
    """n = 0 everywhere: nothing launched, nothing written, no error (the reference's functions map over empty vectors)."""
    itp = B.sph_bessel_interpolator(3, 3, 0.0, 100.0, 1000)
    assert B._moments(3, [0, 1, 2], B.SMALL, np.zeros(0)).shape == (0, 3)
    assert itp.many(np.zeros(0)).shape == (0, 3)
    assert B.integrate_sph_bessel_filon(np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0), itp).shape == (0,)
    assert B.filon_chain(np.array([0.0, 1.0]), np.zeros((0, 2)), np.zeros((0, 2)), np.zeros((0, 2)), np.zeros(0), itp).shape == (0,)
    itp.close(); itp.close()      # idempotent
