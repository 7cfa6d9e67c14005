"""Row n2 (Bessel-moment / Filon LOS integrator): the oracle against EVERY known answer the reference's own test file holds
(test/testbessel.jl), with the reference's tolerances.  CPU only."""
import os
import sys

import mpmath
import numpy as np
import pytest

from oracle import bessel_moments_oracle as O  # noqa: E402

TOL = 1e-13          # test/testbessel.jl:4
big = mpmath.mpf

J_MODERATE = (big("1.02810300641345785268082956068118"), big("8.16960159665387499951281402709868"),
              big("1148.91466190490173313233835015677"))                                  # testbessel.jl:10-13
QUAD_200 = (big("0.78787782588093298580386838516258"), big("2.50028713446521582908074317765179"),
            big("105.635871208275569897958010677779"))                                    # :53-56
OCT_200 = (big("0.662361624450783328243409501185093"), big("1.49770949285120051392753792923852"),
           big("-163.183648420458825380490472539234"))                                    # :101-103


def rel(a, ref):
    return abs(1 - big(a) / ref)


def test_J_moments_moderate_argument():                       # testbessel.jl:9-33
    for alpha, ref in zip((0, 1, 2), J_MODERATE):
        x, nu = 200.0, 2.5
        assert rel(O.J_moment_asymp(x, nu, alpha - 0.5), ref) < TOL
        assert rel(O.J_moment_asymp_nu_five_halves(x, alpha - 0.5, O.J_moment_asymp_prefactor(nu, alpha)), ref) < TOL
        assert rel(O.J_moment_1F2(x, nu, alpha), ref) < TOL


def test_J_moments_small_argument():                          # :36-50
    assert rel(O.J_moment_1F2(10.0, 2.5, 1.5), big("-0.98904817846028826228408967797229")) < 1e-15
    assert rel(O.J_moment_1F2(0.1, 2.5, -0.5), big("0.00001772317062480308")) < TOL


@pytest.mark.parametrize("nu,refs,closed", [(2, QUAD_200, O.sph_j_moment_asymp_nu_2), (3, OCT_200, O.sph_j_moment_asymp_nu_3)])
def test_sph_moments_large_argument(nu, refs, closed):        # :52-77, :100-125
    for m, ref in zip((0, 1, 2), refs):
        assert rel(O.sph_j_moment_asymp(200.0, nu, m), ref) < TOL
        assert rel(closed(200.0, m, O.sph_j_moment_asymp_prefactor(nu, m)), ref) < TOL
        assert rel(O.sph_j_moment_1F2(200.0, nu, m), ref) < TOL


def test_sph_moments_small_argument_quadrupole():             # :79-97
    assert rel(O.sph_j_moment_1F2(0.1, 2, 1), big("1.66587318119689113603548520948514e-6")) < TOL
    assert rel(O.sph_j_moment_1F2(0.1, 2, 0), big("0.0000222127003021204915961147418458393")) < TOL
    assert rel(O.sph_j_moment_1F2(0.01, 2, 2), big("1.33332653062694211665894244946280e-12")) < TOL


MAC_001 = [big("2.38094356262526052651053721655034e-11"), big("1.90475434619627872194611999598176e-13"),
           big("1.58729497355699854415133110569146e-15")]
MAC_0001 = [big("2.38095229276896093875259000259011e-15"), big("1.90476182917611622651303789471188e-18"),
            big("1.58730152116402236652235367513154e-21")]


def test_maclaurin_vanishing_argument_octupole():             # :128-153
    for x, refs in ((0.01, MAC_001), (0.001, MAC_0001)):
        for m in range(3):
            assert rel(O.sph_j_moment_maclaurin_1F2(x, 3, m), refs[m]) < TOL


@pytest.fixture(scope="module")
def tables():
    """The reference's tests build 2,000,000-node tables on (0, 1.6e4); the spline is local (the prefilter's memory is
    0.268^d), so the oracle builds the same node spacing on a shorter range: x ≤ 1200 covers every x the tests evaluate."""
    h = 1.6e4 / (2_000_000 - 1)
    n = 150_001
    return {(nu, order): O.MomentTable(nu, order, 0.0, h * (n - 1), n) for (nu, order) in ((2, 3), (3, 3), (3, 4))}


def test_interpolator_large_argument(tables):                 # :157-184
    for nu, refs in ((2, QUAD_200), (3, OCT_200)):
        moms = tables[(nu, 3)](200.0)
        for i in range(3):
            assert rel(moms[i], refs[i]) < 1e-12


def test_interpolator_small_argument_octupole(tables):        # :186-198
    refs = (2.380070697034449e-7, 1.904006180460421e-8, 1.586640331877486e-9)
    moms = tables[(3, 3)](0.1)
    for i in range(3):
        assert abs(moms[i] - refs[i]) < 1e-12


REFS_4TH_1 = (big("0.00229425677577922706134041736035369"), big("0.00183049831049510967778943596015052"),
              big("0.00152235376642692867152374927665023"), big("0.00130283667913981697274149102492971"))
REFS_4TH_1000 = (big("0.667496353968182420081865144062472"), big("3.18644086496078566989171654620754"),
                 big("838.803790872929501155031335384019"), big("831383.108409725479693899756952862"))


def test_interpolator_fourth_order(tables):                   # :201-239
    moms = tables[(3, 4)](1.0)
    for i in range(4):
        assert abs(moms[i] - REFS_4TH_1[i]) < 1e-12
    moms = O.MomentTable(3, 4, 2.0, 1.6e4, 20)(1.0)           # Maclaurin branch below the table
    for i in range(4):
        assert abs(moms[i] - REFS_4TH_1[i]) < 1e-12
    moms = tables[(3, 4)](1000.0)
    for i in range(4):
        assert rel(moms[i], REFS_4TH_1000[i]) < 1e-12
    moms = O.MomentTable(3, 4, 0.0, 500.0, 20)(1000.0)        # Lommel branch above the table
    for i in range(4):
        assert rel(moms[i], REFS_4TH_1000[i]) < 1e-12


def test_interpolator_maclaurin_branch():                     # :242-271
    itp = O.MomentTable(3, 3, 1.0, 1.6e4, 200)
    for x, refs in ((0.01, MAC_001), (0.001, MAC_0001)):
        moms = itp(x)
        for m in range(3):
            assert rel(moms[m], refs[m]) < TOL


def test_filon_third_order(tables):                           # :275-305
    itp = tables[(3, 3)]
    refs = [big("0.365287615501162668736682652658444"), big("0.219009396999160523658045931736310"),
            big("0.146278218502002145078636720922135"), big("0.000548758594228308158105260833682184")]
    F = O.integrate_sph_bessel_filon
    assert abs(F(4., -0.2, 6.0, 10.0, 0., 2., itp) - refs[0]) < TOL
    assert abs(F(4., -0.2, 6.0, 10.0, 0., 1., itp) - refs[1]) < TOL
    assert abs(F(6.8, 5.8, 6.0, 10.0, 1., 2., itp) - refs[2]) < TOL
    assert abs(F(15.6, 11.8, 6., 10.0, 2., 4., itp) - refs[3]) < TOL
    s, _ = O.loop_integrate_sph_bessel_filon(6.8, 5.8, 6.0, 10.0, 1., 2., itp, itp(10 * 1.))
    assert abs(s - refs[2]) < TOL

    itp = O.MomentTable(3, 3, 2.0, 50.0, 20)                  # Maclaurin and Lommel branches
    refs = [big("1.42547119945725017346111489855163e-6"), big("1.42843411313369992225631279186481e-10"),
            big("3.21068375645105591375937123607342"), big("-9.05287087052870811987403003972375"),
            big("61.7007662060909735341421714015349")]
    assert abs((F(3.9983, -0.14, 6., 10.0, 0.01, 0.02, itp) - refs[0]) / refs[0]) < TOL
    assert abs((F(3999803 / 1000000, -(97 / 500), 6., 10.0, 0.001, 0.002, itp) - refs[1]) / refs[1]) < TOL
    assert abs((F(7494., 299.8, 6., 10.0, 50.0, 100.0, itp) - refs[2]) / refs[2]) < TOL
    assert abs((F(119964., 1199.8, 6., 10., 200., 201., itp) - refs[3]) / refs[3]) < TOL
    assert abs((F(74999004, 149999 / 5, 6., 10., 5000., 5100., itp) - refs[4]) / refs[4]) < TOL


def test_restated_weniger_summation_matches_the_known_answers_and_hyp1f2():
    """The reference's own summation of the 1F2 (weniger.jl:50-235, restated on 32-digit numbers): every known answer the reference
    checks its Weniger sums against (testbessel.jl:27,40,46,70,83,89,94,118), and agreement with the independent 40-digit hyp1f2
    over the range the table fill uses it on."""
    for alpha, ref in zip((0, 1, 2), J_MODERATE):
        assert rel(O.J_moment_weniger_1F2(200.0, 2.5, alpha), ref) < TOL
    assert rel(O.J_moment_weniger_1F2(10.0, 2.5, 1.5), big("-0.98904817846028826228408967797229")) < 1e-15
    assert rel(O.J_moment_weniger_1F2(0.1, 2.5, -0.5), big("0.00001772317062480308")) < TOL
    for nu, refs in ((2, QUAD_200), (3, OCT_200)):
        for m, ref in zip((0, 1, 2), refs):
            assert rel(O.sph_j_moment_weniger_1F2(200.0, nu, m), ref) < TOL
    assert rel(O.sph_j_moment_weniger_1F2(0.1, 2, 1), big("1.66587318119689113603548520948514e-6")) < TOL
    assert rel(O.sph_j_moment_weniger_1F2(0.1, 2, 0), big("0.0000222127003021204915961147418458393")) < TOL
    assert rel(O.sph_j_moment_weniger_1F2(0.01, 2, 2), big("1.33332653062694211665894244946280e-12")) < TOL
    rng = np.random.default_rng(21)
    for x in np.concatenate([rng.uniform(0.01, 50.0, 12), [1e-3, 49.99]]):
        for nu in (2, 3):
            for m in range(4):
                a, b = O.sph_j_moment_weniger_1F2(x, nu, m), O.sph_j_moment_1F2(x, nu, m)
                assert abs(a - b) < 1e-20 * max(abs(b), big(x) ** (m - 1)), (x, nu, m)
