"""CPU oracle for the Bessel-moment / Filon line-of-sight integrator.

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing else).  The product path is the CUDA library.

Restates, function by function, the reference's
  src/bessel/moments.jl      (moments  ∫₀ˣ tᵐ j_ν(t) dt  and  ∫₀ˣ t^α J_ν(t) dt)
  src/bessel/interpolator.jl (MomentTable: cubic B-spline table of the moments + Maclaurin / Lommel branches outside it)
  src/bessel/integrator.jl   (Filon rule: quadratic source piece × j_ν(kx) integrated with the moments)
  src/bessel/weniger.jl      (weniger1F2: Weniger's sequence transformation of the ₁F₂, the reference's summation in Double64;
                              restated at the end of this file on 32-digit mpmath numbers)
The ₁F₂ of moments.jl:24-27,72-76 is evaluated two ways: by the restated weniger1F2 (the reference's method) and by mpmath's own
hyp1f2 at 40 digits (an independent evaluation of the same function; this is what the table fill uses, for speed).  The two agree
to 1e-23 or better, and both are pinned by the reference's own big-float known answers (test/testbessel.jl, every testset), see
tests/test_bessel_moments.py.
"""
import math

import mpmath
import numpy as np
from scipy.linalg import solve_banded

mpmath.mp.dps = 40


# ------------------------------------------------------------------------------------------------------------------
# moments.jl
# ------------------------------------------------------------------------------------------------------------------
def s2(t, alpha_minus_half, nu, jmax=20, tol=1e-16):
    """Lommel-function asymptotic series s⁽²⁾ (moments.jl:5-21), float64 like the reference's asymptotic branch."""
    s = 1.0
    sk = 1.0
    ti = 1.0 / t
    ti2 = ti * ti
    nu2 = nu * nu
    am1 = alpha_minus_half - 0.5
    for j in range(jmax + 1):
        sk *= (nu2 - (am1 - 2 * j) ** 2) * ti2
        s += sk
        if abs(sk) < tol * abs(s):
            break
    return s * t ** alpha_minus_half * math.sqrt(ti)


def _besselj(nu, x):
    return float(mpmath.besselj(nu, x))


def J_moment_1F2(x, nu, alpha):
    """∫₀ˣ t^α J_ν(t) dt from the hypergeometric ₁F₂ (moments.jl:24-27); 40-digit mpmath stands in for Double64 + Weniger."""
    x, nu, alpha = mpmath.mpf(x), mpmath.mpf(nu), mpmath.mpf(alpha)
    return (1 / (alpha + nu + 1)) * mpmath.exp((alpha + nu + 1) * mpmath.log(x) - nu * mpmath.log(2) - mpmath.loggamma(nu + 1)) * \
        mpmath.hyp1f2((1 + alpha + nu) / 2, (3 + alpha + nu) / 2, 1 + nu, -x * x / 4)


def J_moment_asymp_prefactor(nu, alpha):
    """moments.jl:37-38."""
    return math.exp(math.log(2.0) * alpha + math.lgamma((nu + alpha + 1) / 2) - math.lgamma((nu - alpha + 1) / 2))


def J_moment_asymp(x, nu, alpha_minus_half):
    """moments.jl:30-35."""
    alpha = alpha_minus_half + 0.5
    return J_moment_asymp_prefactor(nu, alpha) + x * (
        (alpha + nu - 1) * _besselj(nu, x) * s2(x, alpha_minus_half - 1, nu - 1) - _besselj(nu - 1, x) * s2(x, alpha_minus_half, nu))


def J_moment_asymp_nu_five_halves(x, alpha_minus_half, prefactor):
    """moments.jl:41-51 (closed-form J_{3/2}, J_{5/2})."""
    sx, cx = math.sin(x), math.cos(x)
    xi = 1.0 / x
    c1 = math.sqrt(2 / math.pi * xi)
    j32 = sx * xi - cx
    j52 = 3 * sx * xi * xi - sx - 3 * cx * xi
    return prefactor + c1 * x * ((alpha_minus_half + 2) * j52 * s2(x, alpha_minus_half - 1, 1.5) - j32 * s2(x, alpha_minus_half, 2.5))


def sph_j_moment_asymp_prefactor(nu, m):
    """moments.jl:57-58."""
    return J_moment_asymp_prefactor(nu + 0.5, m - 0.5) * math.sqrt(math.pi / 2)


def sph_j_moment_asymp(x, nu, m, prefactor=None):
    """moments.jl:61-70."""
    if prefactor is None:
        prefactor = sph_j_moment_asymp_prefactor(nu, m)
    nup, num = nu + 0.5, nu - 0.5
    return prefactor + x * math.sqrt(math.pi / 2) * (
        (m + nu - 1) * _besselj(nup, x) * s2(x, m - 2, num) - _besselj(num, x) * s2(x, m - 1, nup))


def sph_j_moment_asymp_nu_2(x, m, prefactor):
    """moments.jl:87-98."""
    amh = m - 1
    sx, cx = math.sin(x), math.cos(x)
    xi = 1.0 / x
    c1 = math.sqrt(xi)
    j32 = sx * xi - cx
    j52 = 3 * sx * xi * xi - sx - 3 * cx * xi
    return prefactor + c1 * x * ((amh + 2) * j52 * s2(x, amh - 1, 1.5) - j32 * s2(x, amh, 2.5))


def sph_j_moment_asymp_nu_3(x, m, prefactor):
    """moments.jl:100-112."""
    amh = m - 1
    sx, cx = math.sin(x), math.cos(x)
    xi = 1.0 / x
    xi2 = xi * xi
    xi3 = xi2 * xi
    c1 = math.sqrt(xi)
    j72 = 15 * sx * xi3 - 15 * cx * xi2 - 6 * sx * xi + cx
    j52 = 3 * sx * xi2 - sx - 3 * cx * xi
    return prefactor + c1 * x * ((amh + 3) * j72 * s2(x, amh - 1, 2.5) - j52 * s2(x, amh, 3.5))


def sph_j_moment_1F2(x, nu, m):
    """∫₀ˣ tᵐ j_ν(t) dt from ₁F₂ (moments.jl:72-76), 40 digits."""
    x = mpmath.mpf(x)
    nup = nu + mpmath.mpf(3) / 2
    return (1 / mpmath.mpf(m + nu + 1)) * mpmath.exp((m + nu + 1) * mpmath.log(x) - (nup - 1) * mpmath.log(2) - mpmath.loggamma(nup)) * \
        mpmath.hyp1f2(mpmath.mpf(1 + m + nu) / 2, mpmath.mpf(3 + m + nu) / 2, nup, -x * x / 4) * mpmath.sqrt(mpmath.pi / 2)


def maclaurin_1F2(a, b1, b2, z):
    """weniger.jl:238-249, float64."""
    eps10 = 10 * np.finfo(float).eps
    S0, S1, j = 1.0, 1.0 + a * z / (b1 * b2), 1
    while j <= 1 or abs(S0 - S1) > eps10 * max(abs(S0), abs(S1)):
        r = 1.0 / (j + 1.0)
        r *= a + j
        r /= (b1 + j) * (b2 + j)
        S0, S1 = S1, S1 + (S1 - S0) * r * z
        j += 1
    return S1


def sph_j_moment_maclaurin_1F2(x, nu, m):
    """moments.jl:79-83, float64."""
    nup = nu + 1.5
    return (1.0 / (m + nu + 1)) * math.exp((m + nu + 1) * math.log(x) - (nup - 1) * math.log(2.0) - math.lgamma(nup)) * \
        maclaurin_1F2((1 + m + nu) / 2, (3 + m + nu) / 2, nup, -x * x / 4) * math.sqrt(math.pi / 2)


# ------------------------------------------------------------------------------------------------------------------
# interpolator.jl
# ------------------------------------------------------------------------------------------------------------------
def _asymp_nu(nu):
    if nu == 2:
        return sph_j_moment_asymp_nu_2
    if nu == 3:
        return sph_j_moment_asymp_nu_3
    raise ValueError("the reference dispatches the asymptotic branch for ν = 2, 3 only (interpolator.jl:13-16)")


def _asymp_nodes(nu, m, x, pref):
    """Vectorised sph_j_moment_asymp_nu_{2,3} for the table fill (same expressions as above on numpy arrays)."""
    amh = m - 1
    sx, cx = np.sin(x), np.cos(x)
    xi = 1.0 / x

    def s2v(t, a, v):
        s = np.ones_like(t)
        sk = np.ones_like(t)
        ti2 = 1.0 / (t * t)
        am1 = a - 0.5
        live = np.ones(t.shape, bool)
        for j in range(21):
            sk = np.where(live, sk * (v * v - (am1 - 2 * j) ** 2) * ti2, sk)
            s = np.where(live, s + sk, s)
            live &= ~(np.abs(sk) < 1e-16 * np.abs(s))
        return s * t ** a * np.sqrt(1.0 / t)

    c1 = np.sqrt(xi)
    j52 = 3 * sx * xi * xi - sx - 3 * cx * xi
    if nu == 2:
        j32 = sx * xi - cx
        return pref + c1 * x * ((amh + 2) * j52 * s2v(x, amh - 1, 1.5) - j32 * s2v(x, amh, 2.5))
    j72 = 15 * sx * xi ** 3 - 15 * cx * xi * xi - 6 * sx * xi + cx
    return pref + c1 * x * ((amh + 3) * j72 * s2v(x, amh - 1, 2.5) - j52 * s2v(x, amh, 3.5))


def prefilter_cubic_line(y):
    """Cubic B-spline coefficients with Interpolations.jl's Line(OnGrid()) ends: one padding coefficient each side,
    (c[i-1] + 4 c[i] + c[i+1]) / 6 = y[i], and c[-1] - 2 c[0] + c[1] = 0 at both ends (interpolator.jl:109)."""
    n = len(y)
    ab = np.zeros((5, n + 2))          # solve_banded storage: ab[2 + i - j, j] = A[i, j]
    rhs = np.zeros(n + 2)

    def put(i, j, v):
        ab[2 + i - j, j] = v

    for i in range(1, n + 1):
        put(i, i - 1, 1 / 6)
        put(i, i, 4 / 6)
        put(i, i + 1, 1 / 6)
    rhs[1:n + 1] = y
    put(0, 0, 1.0), put(0, 1, -2.0), put(0, 2, 1.0)
    put(n + 1, n + 1, 1.0), put(n + 1, n, -2.0), put(n + 1, n - 1, 1.0)
    return solve_banded((2, 2), ab, rhs)


def eval_cubic(c, t):
    """Uniform cubic B-spline at fractional node position t (0-based node index), c carries the two padding coefficients."""
    i = min(max(int(math.floor(t)), 0), len(c) - 4)
    u = t - i
    w0 = (1 - u) ** 3 / 6
    w1 = (3 * u ** 3 - 6 * u ** 2 + 4) / 6
    w2 = (-3 * u ** 3 + 3 * u ** 2 + 3 * u + 1) / 6
    w3 = u ** 3 / 6
    return w0 * c[i] + w1 * c[i + 1] + w2 * c[i + 2] + w3 * c[i + 3]


class MomentTable:
    """sph_bessel_interpolator + the three-branch call (interpolator.jl:67-110)."""

    def __init__(self, nu, order, keta_min, keta_max, N, weniger_cut=50.0):
        self.nu, self.order, self.keta_min, self.keta_max, self.N = nu, order, float(keta_min), float(keta_max), N
        self.prefactors = [sph_j_moment_asymp_prefactor(nu, m) for m in range(order)]
        xs = np.linspace(self.keta_min, self.keta_max, N)
        small = xs < weniger_cut
        self.coef = []
        for m in range(order):
            y = np.empty(N)
            y[small] = [float(sph_j_moment_1F2(x, nu, m)) if x > 0 else 0.0 for x in xs[small]]
            if (~small).any():
                y[~small] = _asymp_nodes(nu, m, xs[~small], self.prefactors[m])
            self.coef.append(prefilter_cubic_line(y))

    def __call__(self, x):
        if self.keta_min <= x <= self.keta_max:
            t = (x - self.keta_min) / (self.keta_max - self.keta_min) * (self.N - 1)
            return np.array([eval_cubic(c, t) for c in self.coef])
        if x > self.keta_max:
            f = _asymp_nu(self.nu)
            return np.array([f(x, m, self.prefactors[m]) for m in range(self.order)])
        return np.array([sph_j_moment_maclaurin_1F2(x, self.nu, m) for m in range(self.order)])


def table_many(itp, x):
    """MomentTable call vectorised over x INSIDE the table (numpy): the CPU arm of the Filon benchmark."""
    t = (x - itp.keta_min) / (itp.keta_max - itp.keta_min) * (itp.N - 1)
    i = np.clip(np.floor(t).astype(np.int64), 0, itp.N - 2)
    u = t - i
    w = ((1 - u) ** 3 / 6, (3 * u ** 3 - 6 * u ** 2 + 4) / 6, (-3 * u ** 3 + 3 * u ** 2 + 3 * u + 1) / 6, u ** 3 / 6)
    return np.stack([sum(w[j] * c[i + j] for j in range(4)) for c in itp.coef], axis=-1)


def filon_chain(nodes, f, f1, f2, k, itp):
    """The loop form of integrator.jl:25-38 over consecutive nodes for every k (rows of f, f1, f2), numpy-vectorised."""
    out = np.empty(len(k))
    a = nodes[:-1]
    for ik, kk in enumerate(k):
        I = table_many(itp, kk * nodes)
        dI = I[1:] - I[:-1]
        c2 = f2[ik, :-1] / 2
        af2 = a * f2[ik, :-1]
        c1 = f1[ik, :-1] - af2
        c0 = f[ik, :-1] - a * (f1[ik, :-1] - af2 / 2)
        out[ik] = np.sum((c0 / kk) * dI[:, 0] + (c1 / kk ** 2) * dI[:, 1] + (c2 / kk ** 3) * dI[:, 2])
    return out


# ------------------------------------------------------------------------------------------------------------------
# integrator.jl
# ------------------------------------------------------------------------------------------------------------------
def integrate_sph_bessel_filon(f, f1, f2, k, a, b, itp):
    """∫_a^b (quadratic through f, f′, f″ at a) · j_ν(k x) dx with the moment table (integrator.jl:7-20)."""
    c2 = f2 / 2
    af2 = a * f2
    c1 = f1 - af2
    c0 = f - a * (f1 - af2 / 2)
    ki = 1 / k
    dI = itp(k * b) - itp(k * a)
    return (c0 * ki) * dI[0] + (c1 * ki * ki) * dI[1] + (c2 * ki ** 3) * dI[2]


def loop_integrate_sph_bessel_filon(f, f1, f2, k, a, b, itp, itp_ka):
    """integrator.jl:25-38: the piecewise form that re-uses I(ka) from the previous piece."""
    c2 = f2 / 2
    af2 = a * f2
    c1 = f1 - af2
    c0 = f - a * (f1 - af2 / 2)
    ki = 1 / k
    itp_kb = itp(k * b)
    dI = itp_kb - itp_ka
    return (c0 * ki) * dI[0] + (c1 * ki * ki) * dI[1] + (c2 * ki ** 3) * dI[2], itp_kb


# ------------------------------------------------------------------------------------------------------------------
# weniger.jl: the reference's own summation of the 1F2 (Weniger's sequence transformation, 1-based arrays as there)
# ------------------------------------------------------------------------------------------------------------------
def weniger1F2(alpha, beta, z, dps=32, kmax=100000):
    """weniger1F2(α, β::SVector{2}, z, cache) restated (weniger.jl:50-235) with mpmath numbers of `dps` digits standing in for
    Double64 (~32 digits).  Arrays are 1-based like the reference's (index 0 unused) so that every subscript can be compared with
    the source.  Used only to show that the reference's summation method and the oracle's 40-digit hyp1f2 agree."""
    with mpmath.workdps(dps):
        mpf, rf = mpmath.mpf, mpmath.rf
        eps = mpf(2) ** (-int(dps * 3.3219))
        a = mpf(alpha); b = [None, mpf(beta[0]), mpf(beta[1])]; z = mpf(z)
        if abs(z) < eps:
            return mpf(1)
        gam = mpf(3) / 2
        zeta = 1 / z
        p, q, r, rho = 1, 2, 5, 3

        def arr(n):
            return [mpf(0)] * (n + 1)
        C, Crho, C1, C2, C3, P = arr(r), arr(rho + 2), arr(rho + 1), arr(rho + 2), arr(rho + 2), arr(rho + 2)
        Q, N, dN, dNold, D, dD, dDold, R = arr(q + 2), arr(r + 1), arr(r + 1), arr(r + 1), arr(r + 1), arr(r + 1), arr(r + 1), arr(r + 1)
        C[1] = mpf(1)
        Crho[rho + 2] = mpf(1)
        for s in range(rho, -1, -1):
            Crho[s + 1] = -(s + 1) * Crho[s + 2] / (rho + 1 - s)
        C2[rho + 2] = 1 / rf(gam - rho - 2, rho + 2)
        C3[rho + 1] = 1 / rf(gam - rho - 1, rho + 2)
        C3[rho + 2] = 1 / rf(gam - rho, rho + 2)
        P[1] = gam * (a + 1)
        err = abs(gam) * (abs(a) + 1)
        Q[1] = 2 * (b[1] + 1) * (b[2] + 1)
        N[r + 1] = b[1] * b[2] * zeta / a / (gam - 1)
        dN[r] = N[r + 1] / rf(gam - rho - 1, rho)
        D[r + 1] = b[1] * b[2] * zeta / a / (gam - 1)
        dD[r] = D[r + 1] / rf(gam - rho - 1, rho)
        R[r + 1] = N[r + 1] / D[r + 1]

        def errcheck(x, y, tol):
            return mpmath.isfinite(x) and mpmath.isfinite(y) and abs(x - y) > max(abs(x), abs(y)) * tol
        k = 0
        while k < r or (k < kmax and errcheck(R[r], R[r + 1], 10 * eps)):
            for j in range(1, r + 1):
                N[j], D[j], R[j] = N[j + 1], D[j + 1], R[j + 1]
            t1 = mpf(0)
            for j in range(0, min(k, q + 1) + 1):
                t1 += C[j + 1] * Q[j + 1] * dN[r - j]
            if k <= rho:
                for j in range(0, k + 1):
                    t2 = (b[1] + j + 1) * (b[2] + j + 1) * (j + 2)
                    t1 += C[j + 1] * mpf(-1) ** (k - j) * rf(j + gam, k - rho - 1) * t2
            t2 = mpf(0)
            s2 = mpf(0)
            for s in range(max(0, rho + 1 - k), rho + 1):
                s2 += Crho[s + 1] * C1[s + 1] * (N[r - rho + s] + (gam + k - rho + s - 2) * N[r - rho + s - 1])
            s2 += (gam + k - 1) * N[r] / rf(gam + 2 * k - rho - 1, rho + 2)
            t2 += P[1] * s2
            s2 = mpf(0)
            for s in range(max(0, rho + 1 - k), rho + 2):
                s2 += Crho[s + 1] * C2[s + 1] * (gam + 2 * k - 2 * rho + 2 * s - 3) * N[r - rho + s - 1]
            dNold[r + 1] = s2
            for j in range(1, min(k, p + 1) + 1):
                t2 += C[j + 1] * P[j + 1] * (dNold[r + 2 - j] + dNold[r + 1 - j])
            N[r + 1] = zeta * t1 - t2
            t1 = mpf(0)
            for j in range(0, min(k, q + 1) + 1):
                t1 += C[j + 1] * Q[j + 1] * dD[r - j]
            t2 = mpf(0)
            s2 = mpf(0)
            for s in range(max(0, rho + 1 - k), rho + 1):
                s2 += Crho[s + 1] * C1[s + 1] * (D[r - rho + s] + (gam + k - rho + s - 2) * D[r - rho + s - 1])
            s2 += (gam + k - 1) * D[r] / rf(gam + 2 * k - rho - 1, rho + 2)
            t2 += P[1] * s2
            s2 = mpf(0)
            for s in range(max(0, rho + 1 - k), rho + 2):
                s2 += Crho[s + 1] * C2[s + 1] * (gam + 2 * k - 2 * rho + 2 * s - 3) * D[r - rho + s - 1]
            dDold[r + 1] = s2
            for j in range(1, min(k, p + 1) + 1):
                t2 += C[j + 1] * P[j + 1] * (dDold[r + 2 - j] + dDold[r + 1 - j])
            D[r + 1] = zeta * t1 - t2
            if abs(P[1]) < eps * err:
                return N[r + 1] / D[r + 1]
            sc = P[1] / rf(gam + 2 * k - rho - 1, rho + 2)
            N[r + 1] /= sc
            D[r + 1] /= sc
            R[r + 1] = N[r + 1] / D[r + 1]
            s1 = mpf(0)
            for s in range(max(0, rho - k), rho + 2):
                s1 += Crho[s + 1] * C3[s + 1] * (gam + 2 * k - 2 * rho + 2 * s - 1) * N[r - rho + s]
            dN[r + 1] = s1
            s1 = mpf(0)
            for s in range(max(0, rho - k), rho + 2):
                s1 += Crho[s + 1] * C3[s + 1] * (gam + 2 * k - 2 * rho + 2 * s - 1) * D[r - rho + s]
            dD[r + 1] = s1
            k += 1
            for j in range(min(k, p + 1), -1, -1):
                dNold[r - j] = (gam + 2 * k - rho - j - 4) * dNold[r - j + 1] + (k - j - 1) * dNold[r - j]
            for j in range(min(k, q + 1), -1, -1):
                dN[r - j] = (gam + 2 * k - rho - j - 2) * dN[r - j + 1] + (k - j) * dN[r - j]
            for j in range(min(k, p + 1), -1, -1):
                dDold[r - j] = (gam + 2 * k - rho - j - 4) * dDold[r - j + 1] + (k - j - 1) * dDold[r - j]
            for j in range(min(k, q + 1), -1, -1):
                dD[r - j] = (gam + 2 * k - rho - j - 2) * dD[r - j + 1] + (k - j) * dD[r - j]
            for j in range(min(k, rho), 0, -1):
                C[j + 1] += C[j]
            if k <= rho + 1:
                for s in range(max(0, rho + 1 - k), rho + 1):
                    C1[s + 1] = rf(k - rho + s, rho + 1 - s) / rf(gam + 2 * k - 2 * rho + s - 2, rho + 2)
                for s in range(max(0, rho + 1 - k), rho + 2):
                    C2[s + 1] = rf(k - rho + s, rho + 1 - s) / rf(gam + 2 * k - 2 * rho + s - 3, rho + 2)
            else:
                for s in range(0, rho + 1):
                    C1[s + 1] *= mpf(k) / (k - mpf(1) - rho + s) * (gam + 2 * k - 2 * rho + s - 4) / (gam + 2 * k - rho + s - 2) * \
                        (gam + 2 * k - 2 * rho + s - 3) / (gam + 2 * k - rho + s - 1)
                for s in range(0, rho + 2):
                    C2[s + 1] *= mpf(k) / (k - mpf(1) - rho + s) * (gam + 2 * k - 2 * rho + s - 5) / (gam + 2 * k - rho + s - 3) * \
                        (gam + 2 * k - 2 * rho + s - 4) / (gam + 2 * k - rho + s - 2)
            if k <= rho:
                for s in range(max(0, rho - k), rho + 2):
                    C3[s + 1] = rf(k - rho + s + 1, rho + 1 - s) / rf(gam + 2 * k - 2 * rho + s - 1, rho + 2)
            else:
                for s in range(max(0, rho - k), rho + 2):
                    C3[s + 1] *= (k + mpf(1)) / (k - rho + s) * (gam + 2 * k - 2 * rho + s - 3) / (gam + 2 * k - rho + s - 1) * \
                        (gam + 2 * k - 2 * rho + s - 2) / (gam + 2 * k - rho + s)
            t = (gam + k) * (a + k + 1)
            err = (abs(gam) + k) * (abs(a) + k + 1)
            for j in range(2, p + 3):
                s_ = t - P[j - 1]
                P[j - 1] = t
                t = s_
            P[p + 2] = t
            t = (b[1] + k + 1) * (b[2] + k + 1) * (k + 2)
            for j in range(2, q + 3):
                s_ = t - Q[j - 1]
                Q[j - 1] = t
                t = s_
            Q[q + 2] = t
        return R[r + 1] if mpmath.isfinite(R[r + 1]) else R[r]


def sph_j_moment_weniger_1F2(x, nu, m, dps=32):
    """sph_j_moment_weniger_₁F₂ (moments.jl:72-76) with the reference's own summation (weniger1F2 above)."""
    with mpmath.workdps(dps):
        x = mpmath.mpf(x)
        nup = nu + mpmath.mpf(3) / 2
        pre = (1 / mpmath.mpf(m + nu + 1)) * mpmath.exp((m + nu + 1) * mpmath.log(x) - (nup - 1) * mpmath.log(2) - mpmath.loggamma(nup))
        return pre * weniger1F2(mpmath.mpf(1 + m + nu) / 2, (mpmath.mpf(3 + m + nu) / 2, nup), -x * x / 4, dps) * mpmath.sqrt(mpmath.pi / 2)


def J_moment_weniger_1F2(x, nu, alpha, dps=32):
    """J_moment_weniger_₁F₂ (moments.jl:24-27) with the reference's own summation."""
    with mpmath.workdps(dps):
        x, nu, alpha = mpmath.mpf(x), mpmath.mpf(nu), mpmath.mpf(alpha)
        pre = (1 / (alpha + nu + 1)) * mpmath.exp((alpha + nu + 1) * mpmath.log(x) - nu * mpmath.log(2) - mpmath.loggamma(nu + 1))
        return pre * weniger1F2((1 + alpha + nu) / 2, ((3 + alpha + nu) / 2, 1 + nu), -x * x / 4, dps)
