// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
// CPU restatement (C++17, double precision) of the hot path of xzackli/Bolt.jl:
//   src/perturbations.jl  state layout, initial conditions, RHS, source functions, boltsolve
//   src/spectra.jl        source grids, k grids, j_l spline, line-of-sight sum, C_l, plin
//   src/util.jl           cubic B-spline evaluation, momentum-grid maps
//   src/background.jl     f0, dlnf0dlnq
// Each function cites the reference file:line it follows.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library, and only as the
// checker or the reported CPU baseline -- never as the thing shipped.
//
// PARITY PIN: the reference cannot be executed here (no Julia), and its stepper (OrdinaryDiffEq
// 6.20.0 KenCarp4, docs/Manifest.toml:937-941) is not vendored.  This oracle is pinned only by the
// reference's own fixtures: CLASS Phi/delta_b at 1e-3 (test/runtests.jl:83-147), CAMB TT/EE at 11%
// (test/runtests.jl:149-185), Fortran-RECFAST Xe at 1e-4 (test/runtests.jl:38-48, host generator).
// At the stepper level parity with OrdinaryDiffEq is UNPINNED.
//
// Stepper: the ESDIRK half of Kennedy & Carpenter's ARK4(3)6L[2]SA (what KenCarp4() runs on a
// non-split ODEProblem, src/perturbations.jl:28-31).  The hierarchy is linear in u, u' = A(x) u,
// so every implicit stage equation is a linear system; it is solved exactly with the Jacobian at
// the stage abscissa (the limit of OrdinaryDiffEq's Newton iteration at zero tolerance), using a
// dense n x n matrix and LU with partial pivoting like the reference's default linear solver.
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <limits>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/bolt_cuda.h"

namespace {

// ---------------------------------------------------------------------------------------------
// Forward-mode dual numbers (value + NP partials).  The reference obtains parameter gradients by running the
// whole stack on ForwardDiff.Dual{Tag,Float64,N} (examples/plot_deriv_cl.jl:28-33); every physics function below
// is therefore a template on the scalar type T, instantiated with double (the value path, bit-for-bit the
// arithmetic of the non-templated code it replaces) and with Du<NP> (the gradient oracle).
// ---------------------------------------------------------------------------------------------
template <int NP> struct Du {
  double v; double d[NP];
  Du() {}
  Du(double x) : v(x) { for (int i = 0; i < NP; i++) d[i] = 0.0; }
};
#define DU_FOR for (int i = 0; i < NP; i++)
template <int NP> inline Du<NP> operator-(const Du<NP>& a) { Du<NP> r; r.v = -a.v; DU_FOR r.d[i] = -a.d[i]; return r; }
template <int NP> inline Du<NP> operator+(const Du<NP>& a, const Du<NP>& b) { Du<NP> r; r.v = a.v + b.v; DU_FOR r.d[i] = a.d[i] + b.d[i]; return r; }
template <int NP> inline Du<NP> operator-(const Du<NP>& a, const Du<NP>& b) { Du<NP> r; r.v = a.v - b.v; DU_FOR r.d[i] = a.d[i] - b.d[i]; return r; }
template <int NP> inline Du<NP> operator*(const Du<NP>& a, const Du<NP>& b) { Du<NP> r; r.v = a.v * b.v; DU_FOR r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int NP> inline Du<NP> operator/(const Du<NP>& a, const Du<NP>& b) { Du<NP> r; r.v = a.v / b.v; DU_FOR r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v; return r; }
template <int NP> inline Du<NP> operator+(const Du<NP>& a, double b) { Du<NP> r = a; r.v += b; return r; }
template <int NP> inline Du<NP> operator+(double b, const Du<NP>& a) { Du<NP> r = a; r.v += b; return r; }
template <int NP> inline Du<NP> operator-(const Du<NP>& a, double b) { Du<NP> r = a; r.v -= b; return r; }
template <int NP> inline Du<NP> operator-(double b, const Du<NP>& a) { Du<NP> r; r.v = b - a.v; DU_FOR r.d[i] = -a.d[i]; return r; }
template <int NP> inline Du<NP> operator*(const Du<NP>& a, double b) { Du<NP> r; r.v = a.v * b; DU_FOR r.d[i] = a.d[i] * b; return r; }
template <int NP> inline Du<NP> operator*(double b, const Du<NP>& a) { Du<NP> r; r.v = a.v * b; DU_FOR r.d[i] = a.d[i] * b; return r; }
template <int NP> inline Du<NP> operator/(const Du<NP>& a, double b) { Du<NP> r; r.v = a.v / b; DU_FOR r.d[i] = a.d[i] / b; return r; }
template <int NP> inline Du<NP> operator/(double a, const Du<NP>& b) { Du<NP> r; r.v = a / b.v; DU_FOR r.d[i] = -r.v * b.d[i] / b.v; return r; }
template <int NP> inline Du<NP>& operator+=(Du<NP>& a, const Du<NP>& b) { a.v += b.v; DU_FOR a.d[i] += b.d[i]; return a; }
template <int NP> inline Du<NP>& operator*=(Du<NP>& a, const Du<NP>& b) { a = a * b; return a; }
inline double m_exp(double a) { return std::exp(a); }
inline double m_sqrt(double a) { return std::sqrt(a); }
inline double m_log10(double a) { return std::log10(a); }
inline double m_pow(double a, double b) { return std::pow(a, b); }
template <int NP> inline Du<NP> m_exp(const Du<NP>& a) { Du<NP> r; r.v = std::exp(a.v); DU_FOR r.d[i] = a.d[i] * r.v; return r; }
template <int NP> inline Du<NP> m_sqrt(const Du<NP>& a) { Du<NP> r; r.v = std::sqrt(a.v); DU_FOR r.d[i] = a.d[i] * 0.5 / r.v; return r; }
template <int NP> inline Du<NP> m_log10(const Du<NP>& a) { Du<NP> r; r.v = std::log10(a.v); DU_FOR r.d[i] = a.d[i] / (a.v * std::log(10.0)); return r; }
template <int NP> inline Du<NP> m_pow(const Du<NP>& a, double b) { Du<NP> r; r.v = std::pow(a.v, b); DU_FOR r.d[i] = a.d[i] * b * r.v / a.v; return r; }
template <int NP> inline Du<NP> m_pow(double a, const Du<NP>& b) { Du<NP> r; r.v = std::pow(a, b.v); DU_FOR r.d[i] = b.d[i] * r.v * std::log(a); return r; }
template <int NP> inline Du<NP> m_pow(const Du<NP>& a, const Du<NP>& b) {
  Du<NP> r; r.v = std::pow(a.v, b.v); DU_FOR r.d[i] = r.v * (b.d[i] * std::log(a.v) + b.v * a.d[i] / a.v); return r; }
inline double val(double a) { return a; }
template <int NP> inline double val(const Du<NP>& a) { return a.v; }
// read element i (nd doubles, value first) of a dual-capable array
inline void load_T(const double* p, int nd, double& out) { out = p[0]; }
template <int NP> inline void load_T(const double* p, int nd, Du<NP>& out) { out.v = p[0]; DU_FOR out.d[i] = (1 + i < nd) ? p[1 + i] : 0.0; }
inline void store_T(double* p, int nd, double a) { p[0] = a; }
template <int NP> inline void store_T(double* p, int nd, const Du<NP>& a) { p[0] = a.v; DU_FOR if (1 + i < nd) p[1 + i] = a.d[i]; }

// ---------------------------------------------------------------------------------------------
// host tables
// ---------------------------------------------------------------------------------------------
template <class T> struct Cosmo_ {
  int n_x = 0, nq = 0;
  double x0 = 0, dx = 0;
  T s[BOLT_NSCALARS];
  std::vector<T> tab[BOLT_NTABLES];
  std::vector<double> xq, wq;
  // per-cosmology constants of the momentum grid (perturbations.jl:164-166, background.jl:21-30)
  std::vector<T> q, f0w, dlnf0;  // q_i, f0(q_i)/dxdq(q_i)*w_i, dlnf0dlnq(q_i)
  T Tnu = T(0.0), Omega_nu = T(0.0);
};
typedef Cosmo_<double> Cosmo;

// util.jl:11 -- Interpolations' BSpline(Cubic(Line(OnGrid()))) evaluation on a uniform grid.
// X = double (abscissa without partials) or T (the Bessel argument k*chi carries partials in the gradient oracle).
template <class C, class X> inline auto spline_eval(const std::vector<C>& c, int n_x, double x0, double dx, X x) -> decltype(c[0] * x) {
  X t = (x - x0) / dx;
  int i = (int)std::floor(val(t));
  if (i < 0) i = 0;
  if (i > n_x - 2) i = n_x - 2;
  X d = t - (double)i, e = 1.0 - d;
  X w0 = e * e * e / 6.0;
  X w1 = 2.0 / 3.0 - d * d + d * d * d / 2.0;
  X w2 = 2.0 / 3.0 - e * e + e * e * e / 2.0;
  X w3 = d * d * d / 6.0;
  return c[i] * w0 + c[i + 1] * w1 + c[i + 2] * w2 + c[i + 3] * w3;
}
template <class T> inline T tab(const Cosmo_<T>& c, int which, double x) {
  return spline_eval(c.tab[which], c.n_x, c.x0, c.dx, x);
}

template <class T> void cosmo_from_desc(const bolt_cosmo_desc* d, Cosmo_<T>& c) {
  c.n_x = d->n_x; c.nq = d->nq; c.x0 = d->x0; c.dx = d->dx;
  const int nd = d->nd;
  for (int i = 0; i < BOLT_NSCALARS; i++) load_T(d->scalars + (size_t)i * nd, nd, c.s[i]);
  for (int t = 0; t < BOLT_NTABLES; t++) {
    c.tab[t].resize(d->n_x + 2);
    for (int i = 0; i < d->n_x + 2; i++) load_T(d->tables + ((size_t)t * (d->n_x + 2) + i) * nd, nd, c.tab[t][i]);
  }
  c.xq.assign(d->quad_pts, d->quad_pts + d->nq);
  c.wq.assign(d->quad_wts, d->quad_wts + d->nq);
  // Tν (perturbations.jl:164), q grid (:165-166 with util.jl:24-27), f0 and dlnf0dlnq (background.jl:21-30)
  const T N_nu = c.s[BOLT_S_N_nu], Om_r = c.s[BOLT_S_Omega_r], rho_crit = c.s[BOLT_S_rho_crit];
  c.Tnu = m_pow(N_nu / 3.0, 0.25) * std::pow(4.0 / 11.0, 1.0 / 3.0) *
          m_pow(15.0 / (M_PI * M_PI) * rho_crit * Om_r, 0.25);
  const T lqmi = m_log10(c.Tnu / 30.0), lqma = m_log10(c.Tnu * 30.0);
  c.q.resize(c.nq); c.f0w.resize(c.nq); c.dlnf0.resize(c.nq);
  for (int i = 0; i < c.nq; i++) {
    T lq = lqmi + (lqma - lqmi) / 2.0 * (c.xq[i] + 1.0);             // from_ui
    T q = m_pow(10.0, lq);                                           // xq2q
    T dxdq = (2.0 / (lqma - lqmi)) / (q * std::log(10.0));           // dxdq
    T f0 = 2.0 / std::pow(2.0 * M_PI, 3) / (m_exp(q / c.Tnu) + 1.0);
    c.q[i] = q;
    c.f0w[i] = f0 / dxdq * c.wq[i];
    c.dlnf0[i] = -q / c.Tnu / (1.0 + m_exp(-q / c.Tnu));
  }
  c.Omega_nu = 7.0 * (2.0 / 3.0) * N_nu / 8.0 * std::pow(4.0 / 11.0, 4.0 / 3.0) * Om_r;  // perturbations.jl:171
}

// ---------------------------------------------------------------------------------------------
// state layout (unpack, perturbations.jl:114-125), 0-based
// ---------------------------------------------------------------------------------------------
struct Layout {
  int L, Lnu, Lm, nq, n;
  int iT, iP, iN, iM, iS;  // offsets of Θ, Θᵖ, 𝒩, ℳ, (Φ,δ,v,δ_b,v_b)
  Layout(int L_, int Lnu_, int Lm_, int nq_) : L(L_), Lnu(Lnu_), Lm(Lm_), nq(nq_) {
    iT = 0; iP = L + 1; iN = 2 * (L + 1); iM = iN + Lnu + 1; iS = iM + (Lm + 1) * nq; n = iS + 5;
  }
};

template <class T> struct Mode_ {       // Hierarchy (perturbations.jl:7-21)
  const Cosmo_<T>* c;
  double k;                              // k grids carry no partials (spectra.jl:46-47,61)
  Layout lay;
};
typedef Mode_<double> Mode;

// ρ_σ (perturbations.jl:127-145): ρ = 4π Σ q² ε f0/dxdq ℳ0 w ;  σ = 4π Σ q² (q²/ε) f0/dxdq ℳ2 w
template <class T> inline void rho_sigma(const Cosmo_<T>& c, const T* M0, const T* M2, double a, T& rho, T& sigma) {
  const T m = c.s[BOLT_S_Sum_m_nu];
  T r = T(0.0), s = T(0.0);
  for (int i = 0; i < c.nq; i++) {
    T q = c.q[i], eps = m_sqrt(q * q + (a * m) * (a * m));
    r += q * q * eps * c.f0w[i] * M0[i];
    s += q * q * (q * q / eps) * c.f0w[i] * M2[i];
  }
  rho = 4.0 * M_PI * r; sigma = 4.0 * M_PI * s;
}
// θ (perturbations.jl:148-158)
template <class T> inline T theta_nu(const Cosmo_<T>& c, const T* M1) {
  T t = T(0.0);
  for (int i = 0; i < c.nq; i++) t += c.q[i] * c.q[i] * c.q[i] * c.f0w[i] * M1[i];
  return 4.0 * M_PI * t;
}

// hierarchy! (perturbations.jl:161-271).  `u` is non-const because the RSA branch writes into it
// (:219-227), exactly like the reference.  Returns whether the RSA branch was taken.
template <class T> bool hierarchy(const Mode_<T>& h, double x, T* u, T* du) {
  const Cosmo_<T>& c = *h.c; const Layout& l = h.lay;
  const double k = h.k; const int L = l.L, Lnu = l.Lnu, Lm = l.Lm, nq = l.nq;
  const T Om_r = c.s[BOLT_S_Omega_r], Om_b = c.s[BOLT_S_Omega_b], Om_c = c.s[BOLT_S_Omega_c];
  const T m_nu = c.s[BOLT_S_Sum_m_nu], H0 = c.s[BOLT_S_H0], H02 = H0 * H0, rho_crit = c.s[BOLT_S_rho_crit];
  const T Hx = tab(c, BOLT_T_H, x), eta = tab(c, BOLT_T_eta, x);
  const T taup = tab(c, BOLT_T_taup, x), taupp = tab(c, BOLT_T_taupp, x);
  const double a = std::exp(x);
  const T R = 4.0 * Om_r / (3.0 * Om_b * a);
  const T Om_nu = c.Omega_nu;
  const T csb2 = tab(c, BOLT_T_csb2, x);
  T *Th = u + l.iT, *P = u + l.iP, *N = u + l.iN, *M = u + l.iM;
  T *dT = du + l.iT, *dP = du + l.iP, *dN = du + l.iN, *dM = du + l.iM;
  const T Phi = u[l.iS], delta = u[l.iS + 1], v = u[l.iS + 2], delta_b = u[l.iS + 3], v_b = u[l.iS + 4];

  T rhoM, sigM;
  rho_sigma(c, M, M + 2 * nq, a, rhoM, sigM);                                   // :182
  const T Psi = -Phi - 12.0 * H02 / (k * k) / (a * a) *                         // :184-187
                (Om_r * Th[2] + Om_nu * N[2] + sigM / rho_crit / 4.0);
  const T dPhi = Psi - k * k / (3.0 * Hx * Hx) * Phi + H02 / (2.0 * Hx * Hx) * (   // :189-194
                 Om_c / a * delta + Om_b / a * delta_b + 4.0 * Om_r / (a * a) * Th[0] +
                 4.0 * Om_nu / (a * a) * N[0] + rhoM / (a * a) / rho_crit);
  const T ddelta = k / Hx * v - 3.0 * dPhi;                                     // :197-200
  const T dv = -v - k / Hx * Psi;
  const T ddelta_b = k / Hx * v_b - 3.0 * dPhi;
  const T dv_b = -v_b - k / Hx * (Psi + csb2 * delta_b) + taup * R * (3.0 * Th[1] + v_b);

  for (int iq = 0; iq < nq; iq++) {                                             // :203-213
    const T q = c.q[iq], eps = m_sqrt(q * q + (a * m_nu) * (a * m_nu)), df0 = c.dlnf0[iq];
    dM[0 * nq + iq] = -k / Hx * q / eps * M[1 * nq + iq] + dPhi * df0;
    dM[1 * nq + iq] = k / (3.0 * Hx) * (q / eps * (M[0 * nq + iq] - 2.0 * M[2 * nq + iq]) - eps / q * Psi * df0);
    for (int ell = 2; ell <= Lm - 1; ell++)
      dM[ell * nq + iq] = k / Hx * q / ((double)(2 * ell + 1) * eps) * ((double)ell * M[(ell - 1) * nq + iq] - (double)(ell + 1) * M[(ell + 1) * nq + iq]);
    dM[Lm * nq + iq] = q / eps * k / Hx * M[(Lm - 1) * nq + iq] - (double)(Lm + 1) / (Hx * eta) * M[Lm * nq + iq];
  }

  const bool rsa_on = (k * val(eta) > 240.0) && (-val(taup) * val(Hx) / val(eta) > 100.0);          // :216
  if (rsa_on) {                                                                 // :217-232
    Th[0] = Phi - Hx / k * taup * v_b;
    Th[1] = Hx / k * (-2.0 * dPhi + taup * (Phi - csb2 * delta_b) + Hx / k * (taupp - taup) * v_b);  // :221-222 overwrite :220
    Th[2] = T(0.0);
    N[0] = Phi; N[1] = -2.0 * Hx / k * dPhi; N[2] = T(0.0);
    for (int ell = 0; ell <= Lnu; ell++) dN[ell] = T(0.0);
    for (int ell = 0; ell <= L; ell++) { dT[ell] = T(0.0); dP[ell] = T(0.0); }
  } else {
    dN[0] = -k / Hx * N[1] - dPhi;                                              // :237-243
    dN[1] = k / (3.0 * Hx) * N[0] - 2.0 * k / (3.0 * Hx) * N[2] + k / (3.0 * Hx) * Psi;
    for (int ell = 2; ell <= Lnu - 1; ell++)
      dN[ell] = k / ((double)(2 * ell + 1) * Hx) * ((double)ell * N[ell - 1] - (double)(ell + 1) * N[ell + 1]);
    dN[Lnu] = k / Hx * N[Lnu - 1] - (double)(Lnu + 1) / (Hx * eta) * N[Lnu];
    const T Pi = Th[2] + P[2] + P[0];                                           // :247-253
    dT[0] = -k / Hx * Th[1] - dPhi;
    dT[1] = k / (3.0 * Hx) * Th[0] - 2.0 * k / (3.0 * Hx) * Th[2] + k / (3.0 * Hx) * Psi + taup * (Th[1] + v_b / 3.0);
    for (int ell = 2; ell <= L - 1; ell++)
      dT[ell] = (double)ell * k / ((double)(2 * ell + 1) * Hx) * Th[ell - 1] - (double)(ell + 1) * k / ((double)(2 * ell + 1) * Hx) * Th[ell + 1] +
                taup * (Th[ell] - Pi * (ell == 2 ? 1.0 : 0.0) / 10.0);
    dP[0] = -k / Hx * P[1] + taup * (P[0] - Pi / 2.0);                          // :256-260
    for (int ell = 1; ell <= L - 1; ell++)
      dP[ell] = (double)ell * k / ((double)(2 * ell + 1) * Hx) * P[ell - 1] - (double)(ell + 1) * k / ((double)(2 * ell + 1) * Hx) * P[ell + 1] +
                taup * (P[ell] - Pi * (ell == 2 ? 1.0 : 0.0) / 10.0);
    dT[L] = k / Hx * Th[L - 1] - ((double)(L + 1) / (Hx * eta) - taup) * Th[L];           // :263-264
    dP[L] = k / Hx * P[L - 1] - ((double)(L + 1) / (Hx * eta) - taup) * P[L];
  }
  du[l.iS] = dPhi; du[l.iS + 1] = ddelta; du[l.iS + 2] = dv; du[l.iS + 3] = ddelta_b; du[l.iS + 4] = dv_b;  // :269
  return rsa_on;
}

// initial_conditions (perturbations.jl:274-338)
template <class T> void initial_conditions(const Mode_<T>& h, double xi, T* u) {
  const Cosmo_<T>& c = *h.c; const Layout& l = h.lay;
  const double k = h.k; const int L = l.L, Lnu = l.Lnu, Lm = l.Lm, nq = l.nq;
  std::fill(u, u + l.n, T(0.0));
  const T Hx = tab(c, BOLT_T_H, xi), eta = tab(c, BOLT_T_eta, xi), taup = tab(c, BOLT_T_taup, xi);
  T *Th = u + l.iT, *P = u + l.iP, *N = u + l.iN, *M = u + l.iM;
  const double ai2 = std::exp(xi) * std::exp(xi), ai = std::sqrt(ai2);
  const T N_nu = c.s[BOLT_S_N_nu];
  const T f_nu = 1.0 / (1.0 + 1.0 / (7.0 * (3.0 / 3.0) * N_nu / 8.0 * std::pow(4.0 / 11.0, 4.0 / 3.0)));   // :288
  const double Rc = 1.0;
  const T Phi = (4.0 * f_nu + 10.0) / (4.0 * f_nu + 15.0) * Rc;                 // :292
  const T C = -((15.0 + 4.0 * f_nu) / (20.0 + 8.0 * f_nu)) * Phi;               // :294
  Th[0] = -40.0 * C / (15.0 + 4.0 * f_nu) / 4.0;                                // :297-302
  Th[1] = 10.0 * C / (15.0 + 4.0 * f_nu) * (k * k * eta) / (3.0 * k);
  Th[2] = -8.0 * k / (15.0 * Hx * taup) * Th[1];
  P[0] = (5.0 / 4.0) * Th[2];
  P[1] = -k / (4.0 * Hx * taup) * Th[2];
  P[2] = (1.0 / 4.0) * Th[2];
  for (int ell = 3; ell <= L; ell++) {                                          // :303-306
    Th[ell] = -(double)ell / (2 * ell + 1) * k / (Hx * taup) * Th[ell - 1];
    P[ell] = -(double)ell / (2 * ell + 1) * k / (Hx * taup) * P[ell - 1];
  }
  const T delta = 3.0 / 4.0 * (4.0 * Th[0]);                                    // :308-312
  const T delta_b = delta;
  const T v = -3.0 * k * Th[1];
  const T v_b = v;
  N[0] = Th[0];                                                                 // :316-321
  N[1] = Th[1];
  N[2] = -(k * k * eta * eta) / 15.0 * 1.0 / (1.0 + 2.0 / 5.0 * f_nu) * Phi / 2.0;
  for (int ell = 3; ell <= Lnu; ell++) N[ell] = k / ((double)(2 * ell + 1) * Hx) * N[ell - 1];
  const T m_nu = c.s[BOLT_S_Sum_m_nu];
  for (int iq = 0; iq < nq; iq++) {                                             // :325-334
    const T q = c.q[iq], eps = m_sqrt(q * q + (ai * m_nu) * (ai * m_nu)), df0 = c.dlnf0[iq];
    M[0 * nq + iq] = -N[0] * df0;
    M[1 * nq + iq] = -eps / q * N[1] * df0;
    M[2 * nq + iq] = -N[2] * df0;
    for (int ell = 3; ell <= Lm; ell++)
      M[ell * nq + iq] = q / eps * k / ((double)(2 * ell + 1) * Hx) * M[(ell - 1) * nq + iq];
  }
  u[l.iS] = Phi; u[l.iS + 1] = delta; u[l.iS + 2] = v; u[l.iS + 3] = delta_b; u[l.iS + 4] = v_b;   // :336
}

// source_function (perturbations.jl:343-383) and source_function_P (:386-404)
template <class T> void source_functions(const Mode_<T>& h, double x, const T* u, const T* du, T& S_T, T& S_P) {
  const Cosmo_<T>& c = *h.c; const Layout& l = h.lay;
  const double k = h.k; const int nq = l.nq;
  const T H0 = c.s[BOLT_S_H0], H02 = H0 * H0, Om_r = c.s[BOLT_S_Omega_r], rho_crit = c.s[BOLT_S_rho_crit];
  const T Hx = tab(c, BOLT_T_H, x), Hp = tab(c, BOLT_T_Hp, x), Hpp = tab(c, BOLT_T_Hpp, x);
  const T tau = tab(c, BOLT_T_tau, x), taup = tab(c, BOLT_T_taup, x), taupp = tab(c, BOLT_T_taupp, x);
  const T g = tab(c, BOLT_T_g, x), gp = tab(c, BOLT_T_gp, x), gpp = tab(c, BOLT_T_gpp, x);
  const double a = std::exp(x);
  const T Om_nu = c.Omega_nu;
  const T *Th = u + l.iT, *P = u + l.iP, *N = u + l.iN, *M = u + l.iM;
  const T *dT = du + l.iT, *dP = du + l.iP, *dN = du + l.iN, *dM = du + l.iM;
  const T Phi = u[l.iS], v_b = u[l.iS + 4];
  const T dPhi = du[l.iS], dv_b = du[l.iS + 4];
  T r_, sigM, sigMp;
  rho_sigma(c, M, M + 2 * nq, a, r_, sigM);                                     // :361
  rho_sigma(c, dM, dM + 2 * nq, a, r_, sigMp);                                  // :362
  const T Psi = -Phi - 12.0 * H02 / (k * k) / (a * a) * (Om_r * Th[2] + Om_nu * N[2] + sigM / rho_crit / 4.0);
  const T dPsi = -dPhi - 12.0 * H02 / (k * k) / (a * a) * (Om_r * (dT[2] - 2.0 * Th[2]) + Om_nu * (dN[2] - 2.0 * N[2]) +
                                                            (sigMp - 2.0 * sigM) / rho_crit / 4.0);   // :368-370
  const T Pi = Th[2] + P[2] + P[0];
  const T dPi = dT[2] + dP[2] + dP[0];
  const T term1 = g * (Th[0] + Psi + Pi / 4.0) + m_exp(-tau) * (dPsi - dPhi);               // :375
  const T term2 = (-1.0 / k) * (Hp * g * v_b + Hx * gp * v_b + Hx * g * dv_b);              // :376
  const T ddPi = 2.0 * k / (5.0 * Hx) * (-Hp / Hx * Th[1] + dT[1]) + (3.0 / 10.0) * (taupp * Pi + taup * dPi) -
                 3.0 * k / (5.0 * Hx) * (-Hp / Hx * (Th[3] + P[1] + P[3]) + (dT[3] + dP[1] + dP[3]));   // :377-378
  const T term3 = (3.0 / (4.0 * k * k)) * ((Hp * Hp + Hx * Hpp) * g * Pi + 3.0 * Hx * Hp * (gp * Pi + g * dPi) +
                                           Hx * Hx * (gpp * Pi + 2.0 * gp * dPi + g * ddPi));           // :379-381
  S_T = term1 + term2 + term3;
  const double x_end = c.x0 + c.dx * (c.n_x - 1);
  const T y = k * (tab(c, BOLT_T_eta, x_end) - tab(c, BOLT_T_eta, x));          // :401
  S_P = (3.0 / (4.0 * y * y)) * g * Pi;                                         // :403
}

// ---------------------------------------------------------------------------------------------
// dense LU with partial pivoting.  `skip_zeros` skips structurally-zero multipliers/columns;
// the arithmetic performed on non-zero entries is identical, it is only faster.
// ---------------------------------------------------------------------------------------------
struct DenseLU {
  int n = 0;
  std::vector<double> a;
  std::vector<int> piv, hi;
  void factor(std::vector<double>& A, int n_, bool skip_zeros) {
    n = n_; a.swap(A); piv.resize(n); hi.assign(n, n - 1);
    if (skip_zeros)
      for (int i = 0; i < n; i++) { int h = n - 1; while (h > i && a[(size_t)i * n + h] == 0.0) h--; hi[i] = h; }
    for (int k = 0; k < n; k++) {
      int p = k; double best = std::fabs(a[(size_t)k * n + k]);
      for (int i = k + 1; i < n; i++) { double v = std::fabs(a[(size_t)i * n + k]); if (v > best) { best = v; p = i; } }
      piv[k] = p;
      if (p != k) {
        for (int j = 0; j < n; j++) std::swap(a[(size_t)k * n + j], a[(size_t)p * n + j]);
        std::swap(hi[k], hi[p]);
      }
      const double pivot = a[(size_t)k * n + k];
      const int hk = hi[k];
      for (int i = k + 1; i < n; i++) {
        double m = a[(size_t)i * n + k];
        if (skip_zeros && m == 0.0) continue;
        m /= pivot; a[(size_t)i * n + k] = m;
        double* ri = &a[(size_t)i * n]; const double* rk = &a[(size_t)k * n];
        for (int j = k + 1; j <= hk; j++) ri[j] -= m * rk[j];
        if (hi[i] < hk) hi[i] = hk;
      }
    }
  }
  void solve(double* b) const {
    // rows were swapped in full during factorisation (LAPACK convention): permute b first
    for (int k = 0; k < n; k++) if (piv[k] != k) std::swap(b[k], b[piv[k]]);
    for (int k = 0; k < n; k++) {
      const double bk = b[k];
      if (bk != 0.0) for (int i = k + 1; i < n; i++) { double m = a[(size_t)i * n + k]; if (m != 0.0) b[i] -= m * bk; }
    }
    for (int i = n - 1; i >= 0; i--) {
      double s = b[i]; const double* ri = &a[(size_t)i * n];
      for (int j = i + 1; j <= hi[i]; j++) s -= ri[j] * b[j];
      b[i] = s / ri[i];
    }
  }
};

// ---------------------------------------------------------------------------------------------
// KenCarp4 (implicit tableau; Kennedy & Carpenter 2003 ARK4(3)6L[2]SA-ESDIRK), SURVEY 8c.
// ---------------------------------------------------------------------------------------------
const double KC_GAMMA = 0.25;
const double KC_C[6] = {0.0, 0.5, 83.0 / 250.0, 31.0 / 50.0, 17.0 / 20.0, 1.0};
const double KC_A[6][5] = {
    {0, 0, 0, 0, 0},
    {1.0 / 4.0, 0, 0, 0, 0},
    {8611.0 / 62500.0, -1743.0 / 31250.0, 0, 0, 0},
    {5012029.0 / 34652500.0, -654441.0 / 2922500.0, 174375.0 / 388108.0, 0, 0},
    {15267082809.0 / 155376265600.0, -71443401.0 / 120774400.0, 730878875.0 / 902184768.0, 2285395.0 / 8070912.0, 0},
    {82889.0 / 524892.0, 0.0, 15625.0 / 83664.0, 69875.0 / 102672.0, -2260.0 / 8211.0}};
const double KC_BHAT[6] = {4586570599.0 / 29645900160.0, 0.0, 178811875.0 / 945068544.0, 814220225.0 / 1159782912.0,
                           -3700637.0 / 11593932.0, 61727.0 / 225920.0};

struct SolveOut {
  double* S_T = nullptr;     // [n_x]
  double* S_P = nullptr;     // [n_x]
  double* u_hist = nullptr;  // [n_x][n]
  double* u_final = nullptr; // [n]
  int status = 0; int64_t nsteps = 0, nreject = 0, nfact = 0;
};

struct Stepper {
  const Mode& h; const int n; const bool skip_zeros;
  std::vector<double> Amat, tmp, col, e;
  DenseLU lu;
  bool rsa_seen = false;
  Stepper(const Mode& h_, bool sz) : h(h_), n(h_.lay.n), skip_zeros(sz), tmp(n), col(n), e(n) {}
  // W = I - hgam * A(x), A built column-by-column from the RHS (the hierarchy is linear in u):
  // the analogue of the reference's dense ForwardDiff Jacobian.
  void factor(double x, double hgam) {
    Amat.assign((size_t)n * n, 0.0);
    for (int j = 0; j < n; j++) {
      std::fill(e.begin(), e.end(), 0.0); e[j] = 1.0;
      rsa_seen |= hierarchy(h, x, e.data(), col.data());
      for (int i = 0; i < n; i++) if (col[i] != 0.0) Amat[(size_t)i * n + j] = -hgam * col[i];
    }
    for (int i = 0; i < n; i++) Amat[(size_t)i * n + i] += 1.0;
    lu.factor(Amat, n, skip_zeros);
  }
};

inline double rms_scaled(const double* err, const double* u0, const double* u1, int n, double abstol, double reltol) {
  double s = 0;
  for (int i = 0; i < n; i++) {
    double sc = abstol + reltol * std::max(std::fabs(u0[i]), std::fabs(u1[i]));
    double r = err[i] / sc; s += r * r;
  }
  return std::sqrt(s / n);
}

// boltsolve (perturbations.jl:25-33) + the sampling loop of source_grid (spectra.jl:13-18).
void solve_mode(const Mode& h, const bolt_opts& o, bool skip_zeros, SolveOut& out) {
  const Cosmo& c = *h.c; const Layout& l = h.lay; const int n = l.n;
  const double x_begin = c.x0, x_end = 0.0;
  const bool fixed = (o.mode == BOLT_MODE_FIXED);
  const double reltol = o.reltol, abstol = o.abstol;
  const int64_t max_steps = o.max_steps > 0 ? o.max_steps : 1000000;
  std::vector<double> u(n), unew(n), z[6], rhs(n), U(n), err(n), us(n), uh(n), dus(n), f0(n), u1(n), f1(n);
  for (auto& v : z) v.assign(n, 0.0);
  Stepper st(h, skip_zeros);

  initial_conditions(h, x_begin, u.data());
  { std::vector<double> ucopy(u); hierarchy(h, x_begin, ucopy.data(), f0.data()); }

  // sampling state: next grid row to emit
  int ix = 0;
  auto emit = [&](int i, double xs, const double* uu) {
    std::copy(uu, uu + n, us.begin());
    hierarchy(h, xs, us.data(), dus.data());     // spectra.jl:16 (may mutate us under RSA, like the reference)
    if (i >= o.ix_first) {
      double sT, sP; source_functions(h, xs, us.data(), dus.data(), sT, sP);   // spectra.jl:17,36
      if (out.S_T) out.S_T[i] = sT;
      if (out.S_P) out.S_P[i] = sP;
      if (out.u_hist) std::copy(uu, uu + n, out.u_hist + (size_t)i * n);
    }
  };
  emit(0, c.x0, u.data()); ix = 1;

  double x = x_begin, dt;
  if (fixed) {
    dt = o.fixed_dt;
  } else {
    // initial step: Hairer-Wanner as in OrdinaryDiffEq's ode_determine_initdt [dep-knowledge]
    double d0 = 0, d1 = 0;
    for (int i = 0; i < n; i++) { double sk = abstol + reltol * std::fabs(u[i]); d0 += (u[i] / sk) * (u[i] / sk); d1 += (f0[i] / sk) * (f0[i] / sk); }
    d0 = std::sqrt(d0 / n); d1 = std::sqrt(d1 / n);
    double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    dt0 = std::min(dt0, x_end - x_begin);
    for (int i = 0; i < n; i++) u1[i] = u[i] + dt0 * f0[i];
    hierarchy(h, x_begin + dt0, u1.data(), f1.data());
    double d2 = 0;
    for (int i = 0; i < n; i++) { double sk = abstol + reltol * std::fabs(u[i]); double r = (f1[i] - f0[i]) / sk; d2 += r * r; }
    d2 = std::sqrt(d2 / n) / dt0;
    double dm = std::max(d1, d2);
    double dt1 = (dm <= 1e-15) ? std::max(1e-6, dt0 * 1e-3) : std::pow(10.0, -(2.0 + std::log10(dm)) / 5.0);
    dt = std::min(100.0 * dt0, dt1);
  }
  for (int i = 0; i < n; i++) z[0][i] = dt * f0[i];   // z1 = dt f(u_n) (FSAL)

  // PI controller state (OrdinaryDiffEq defaults for KenCarp4 [dep-knowledge])
  const double beta1 = 7.0 / 40.0, beta2 = 2.0 / 20.0, safety = 0.9, qmin = 0.2, qmax = 10.0;
  const double qsteady_min = 1.0, qsteady_max = 1.2;
  double qold = 1e-4;
  int64_t fixed_left = fixed ? (int64_t)std::llround((x_end - x_begin) / o.fixed_dt) : 0;
  int64_t fixed_total = fixed_left;

  while (true) {
    bool clamped = false;
    if (fixed) { if (fixed_left == 0) break; }
    else {
      if (x >= x_end) break;
      if (x + dt >= x_end) { double dtn = x_end - x; for (int i = 0; i < n; i++) z[0][i] *= dtn / dt; dt = dtn; clamped = true; }
    }
    if (out.nsteps + out.nreject >= max_steps) { out.status = BOLT_K_MAXSTEPS; break; }
    // stages 2..6:  (I - γ dt A(x_n + c_i dt)) U_i = u_n + Σ_j a_ij z_j ;  z_i = (U_i - rhs)/γ
    for (int s = 1; s < 6; s++) {
      for (int i = 0; i < n; i++) { double r = u[i]; for (int j = 0; j < s; j++) r += KC_A[s][j] * z[j][i]; rhs[i] = r; }
      st.factor(x + KC_C[s] * dt, KC_GAMMA * dt); out.nfact++;
      U = rhs; st.lu.solve(U.data());
      for (int i = 0; i < n; i++) z[s][i] = (U[i] - rhs[i]) / KC_GAMMA;
    }
    unew = U;   // stiffly accurate: u_{n+1} = U_6
    bool accept = true; double EEst = 0, q11 = 0;
    if (!fixed) {
      for (int i = 0; i < n; i++) {
        double e = 0;
        for (int j = 0; j < 5; j++) e += (KC_A[5][j] - KC_BHAT[j]) * z[j][i];
        e += (KC_GAMMA - KC_BHAT[5]) * z[5][i];
        err[i] = e;
      }
      st.lu.solve(err.data());   // smooth_est: filter the estimate with W^{-1} of the last stage [dep-knowledge]
      EEst = rms_scaled(err.data(), u.data(), unew.data(), n, abstol, reltol);
      if (!std::isfinite(EEst)) { out.status = BOLT_K_NONFINITE; break; }
      // controller input floored at 1e-6 (rounding-noise guard, same rule as the device kernel; DESIGN.md "controller")
      q11 = std::pow(std::max(EEst, 1e-6), beta1);
      accept = EEst <= 1.0;
      if (getenv("ORACLE_DEBUG")) fprintf(stderr, "x=%.10g dt=%.4g EEst=%.17g acc=%d\n", x, dt, EEst, (int)accept);
    }
    if (accept) {
      // dense output on [x, x+dt]: cubic Hermite with f_n = z1/dt, f_{n+1} = z6/dt (KenCarp4 non-split
      // branch sets fsallast = z6/dt [dep-knowledge]); sample every grid row inside the step.
      const bool last = fixed ? (fixed_left == 1) : clamped;
      const double xn1 = last ? x_end : (fixed ? (x_begin + (double)(fixed_total - fixed_left + 1) * o.fixed_dt) : (x + dt));
      while (ix < c.n_x) {
        const double xs = c.x0 + c.dx * ix;
        if (!last && xs > xn1 + 1e-12) break;
        double th = (xs - x) / dt; if (th > 1.0) th = 1.0;
        for (int i = 0; i < n; i++) {
          double y0 = u[i], y1 = unew[i];
          uh[i] = (1.0 - th) * y0 + th * y1 + th * (th - 1.0) * ((1.0 - 2.0 * th) * (y1 - y0) + (th - 1.0) * z[0][i] + th * z[5][i]);
        }
        emit(ix, xs, uh.data());
        ix++;
      }
      x = xn1; u = unew; out.nsteps++;
      if (fixed) { fixed_left--; for (int i = 0; i < n; i++) z[0][i] = z[5][i]; }
      else {
        double q = q11 / std::pow(qold, beta2);
        q = std::max(1.0 / qmax, std::min(1.0 / qmin, q / safety));
        if (q <= qsteady_max && q >= qsteady_min) q = 1.0;
        qold = std::max(EEst, 1e-4);
        double dtn = dt / q;
        for (int i = 0; i < n; i++) z[0][i] = z[5][i] * (dtn / dt);
        dt = dtn;
      }
    } else {
      out.nreject++;
      double dtn = dt / std::min(1.0 / qmin, q11 / safety);
      for (int i = 0; i < n; i++) z[0][i] *= dtn / dt;
      dt = dtn;
      if (!(dt > 1e-14)) { out.status = BOLT_K_DT_UNDERFLOW; break; }
    }
  }
  if (st.rsa_seen && out.status == 0) out.status = BOLT_K_RSA_TRIGGERED;
  if (out.u_final) std::copy(u.begin(), u.end(), out.u_final);
}

// ---------------------------------------------------------------------------------------------
// Gradient oracle: boltsolve on a Dual-valued state (what the reference does when CosmoParams holds ForwardDiff.Dual,
// examples/plot_deriv_cl.jl:28-33).  The hierarchy is linear, u' = A(x;p) u, so in dual arithmetic every implicit stage
// (W + eps dW)(U + eps S_j) = r + eps r_j  is   W U = r,  W S_j = r_j + h*gamma * (dA/dp_j) U   with the SAME dense LU.
// (dA/dp_j) U is the partial part of hierarchy<Du>() evaluated on (U, zero partials): generic dual arithmetic on the
// restated right-hand side, nothing hand-derived.  Error control as DiffEqBase's ODE_DEFAULT_NORM sees a Dual state
// [dep-knowledge]: per element scale = abstol + reltol * max(|u_n,i|, |u_{n+1},i|) with |.| over value and partials,
// EEst = sqrt( sum over elements and components of (err/scale)^2 / (n * (1+NP)) ).  The error estimate of every component
// is smoothed with the value LU of the last stage (the dW term of a dual-valued smoothing solve is dropped, as on the device).
// ---------------------------------------------------------------------------------------------
template <int NP>
void solve_mode_sens(const Mode_<Du<NP>>& hd, const Mode& h, const bolt_opts& o, bool skip_zeros, int nd_out,
                     double* S_T, double* S_P, double* u_final, int& status_out, int64_t& nsteps_out, int64_t& nreject_out) {
  typedef Du<NP> T;
  const Cosmo& c = *h.c; const Layout& l = h.lay; const int n = l.n;
  const double x_begin = c.x0, x_end = 0.0;
  const bool fixed = (o.mode == BOLT_MODE_FIXED);
  const double reltol = o.reltol, abstol = o.abstol;
  const int64_t max_steps = o.max_steps > 0 ? o.max_steps : 1000000;
  std::vector<T> u(n), unew(n), z[6], rhs(n), U(n), err(n), us(n), uh(n), dus(n), f0(n), u1(n), f1(n), Uv(n), G(n);
  for (auto& v : z) v.assign(n, T(0.0));
  std::vector<double> col(n);
  Stepper st(h, skip_zeros);
  int status = 0; int64_t nsteps = 0, nreject = 0;

  initial_conditions(hd, x_begin, u.data());
  { std::vector<T> ucopy(u); hierarchy(hd, x_begin, ucopy.data(), f0.data()); }

  // solve W X = B for every component of the dual vector B (same LU)
  auto solve_all = [&](std::vector<T>& B) {
    for (int i = 0; i < n; i++) col[i] = B[i].v;
    st.lu.solve(col.data());
    for (int i = 0; i < n; i++) B[i].v = col[i];
    for (int j = 0; j < NP; j++) {
      for (int i = 0; i < n; i++) col[i] = B[i].d[j];
      st.lu.solve(col.data());
      for (int i = 0; i < n; i++) B[i].d[j] = col[i];
    }
  };
  auto norm1 = [&](const T& a) { double s = a.v * a.v; for (int j = 0; j < NP; j++) s += a.d[j] * a.d[j]; return std::sqrt(s); };
  auto rms = [&](const std::vector<T>& e, const std::vector<T>& a0, const std::vector<T>& a1) {
    double s = 0;
    for (int i = 0; i < n; i++) {
      const double sc = abstol + reltol * std::max(norm1(a0[i]), norm1(a1[i]));
      double r = e[i].v / sc; s += r * r;
      for (int j = 0; j < NP; j++) { r = e[i].d[j] / sc; s += r * r; }
    }
    return std::sqrt(s / ((double)n * nd_out));
  };

  int ix = 0;
  auto emit = [&](int i, double xs, const T* uu) {
    std::copy(uu, uu + n, us.begin());
    hierarchy(hd, xs, us.data(), dus.data());     // spectra.jl:16
    if (i >= o.ix_first) {
      T sT, sP; source_functions(hd, xs, us.data(), dus.data(), sT, sP);   // spectra.jl:17,36
      if (S_T) store_T(S_T + (size_t)i * nd_out, nd_out, sT);
      if (S_P) store_T(S_P + (size_t)i * nd_out, nd_out, sP);
    }
  };
  emit(0, c.x0, u.data()); ix = 1;

  double x = x_begin, dt;
  if (fixed) {
    dt = o.fixed_dt;
  } else {
    // initial step from the VALUE components (same rule as solve_mode; the device does the same)
    double d0 = 0, d1 = 0;
    for (int i = 0; i < n; i++) { double sk = abstol + reltol * std::fabs(u[i].v); d0 += (u[i].v / sk) * (u[i].v / sk); d1 += (f0[i].v / sk) * (f0[i].v / sk); }
    d0 = std::sqrt(d0 / n); d1 = std::sqrt(d1 / n);
    double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    dt0 = std::min(dt0, x_end - x_begin);
    std::vector<double> uv1(n), fv1(n);
    for (int i = 0; i < n; i++) uv1[i] = u[i].v + dt0 * f0[i].v;
    hierarchy(h, x_begin + dt0, uv1.data(), fv1.data());
    double d2 = 0;
    for (int i = 0; i < n; i++) { double sk = abstol + reltol * std::fabs(u[i].v); double r = (fv1[i] - f0[i].v) / sk; d2 += r * r; }
    d2 = std::sqrt(d2 / n) / dt0;
    double dm = std::max(d1, d2);
    double dt1 = (dm <= 1e-15) ? std::max(1e-6, dt0 * 1e-3) : std::pow(10.0, -(2.0 + std::log10(dm)) / 5.0);
    dt = std::min(100.0 * dt0, dt1);
  }
  for (int i = 0; i < n; i++) z[0][i] = dt * f0[i];

  const double beta1 = 7.0 / 40.0, beta2 = 2.0 / 20.0, safety = 0.9, qmin = 0.2, qmax = 10.0;
  double qold = 1e-4;
  int64_t fixed_left = fixed ? (int64_t)std::llround((x_end - x_begin) / o.fixed_dt) : 0;
  int64_t fixed_total = fixed_left;

  while (true) {
    bool clamped = false;
    if (fixed) { if (fixed_left == 0) break; }
    else {
      if (x >= x_end) break;
      if (x + dt >= x_end) { double dtn = x_end - x; for (int i = 0; i < n; i++) z[0][i] = z[0][i] * (dtn / dt); dt = dtn; clamped = true; }
    }
    if (nsteps + nreject >= max_steps) { status = BOLT_K_MAXSTEPS; break; }
    for (int s = 1; s < 6; s++) {
      for (int i = 0; i < n; i++) { T r = u[i]; for (int j = 0; j < s; j++) r += KC_A[s][j] * z[j][i]; rhs[i] = r; }
      const double xs = x + KC_C[s] * dt, hg = KC_GAMMA * dt;
      st.factor(xs, hg);
      // value first
      for (int i = 0; i < n; i++) col[i] = rhs[i].v;
      st.lu.solve(col.data());
      for (int i = 0; i < n; i++) { Uv[i] = T(col[i]); }
      hierarchy(hd, xs, Uv.data(), G.data());          // partial parts = (dA/dp_j) U   (Uv may be mutated under RSA: out of envelope)
      for (int i = 0; i < n; i++) U[i].v = col[i];
      for (int j = 0; j < NP; j++) {
        for (int i = 0; i < n; i++) col[i] = rhs[i].d[j] + hg * G[i].d[j];
        st.lu.solve(col.data());
        for (int i = 0; i < n; i++) U[i].d[j] = col[i];
      }
      for (int i = 0; i < n; i++) z[s][i] = (U[i] - rhs[i]) / KC_GAMMA;
    }
    unew = U;
    bool accept = true; double EEst = 0, q11 = 0;
    if (!fixed) {
      for (int i = 0; i < n; i++) {
        T e = T(0.0);
        for (int j = 0; j < 5; j++) e += (KC_A[5][j] - KC_BHAT[j]) * z[j][i];
        e += (KC_GAMMA - KC_BHAT[5]) * z[5][i];
        err[i] = e;
      }
      solve_all(err);
      EEst = rms(err, u, unew);
      if (!std::isfinite(EEst)) { status = BOLT_K_NONFINITE; break; }
      q11 = std::pow(std::max(EEst, 1e-6), beta1);
      accept = EEst <= 1.0;
    }
    if (accept) {
      const bool last = fixed ? (fixed_left == 1) : clamped;
      const double xn1 = last ? x_end : (fixed ? (x_begin + (double)(fixed_total - fixed_left + 1) * o.fixed_dt) : (x + dt));
      while (ix < c.n_x) {
        const double xs = c.x0 + c.dx * ix;
        if (!last && xs > xn1 + 1e-12) break;
        double th = (xs - x) / dt; if (th > 1.0) th = 1.0;
        if (ix >= o.ix_first) {
          for (int i = 0; i < n; i++) {
            const T y0 = u[i], y1 = unew[i];
            uh[i] = (1.0 - th) * y0 + th * y1 + th * (th - 1.0) * ((1.0 - 2.0 * th) * (y1 - y0) + (th - 1.0) * z[0][i] + th * z[5][i]);
          }
          emit(ix, xs, uh.data());
        }
        ix++;
      }
      x = xn1; u = unew; nsteps++;
      if (fixed) { fixed_left--; for (int i = 0; i < n; i++) z[0][i] = z[5][i]; }
      else {
        double q = q11 / std::pow(qold, beta2);
        q = std::max(1.0 / qmax, std::min(1.0 / qmin, q / safety));
        if (q <= 1.2 && q >= 1.0) q = 1.0;
        qold = std::max(EEst, 1e-4);
        double dtn = dt / q;
        for (int i = 0; i < n; i++) z[0][i] = z[5][i] * (dtn / dt);
        dt = dtn;
      }
    } else {
      nreject++;
      double dtn = dt / std::min(1.0 / qmin, q11 / safety);
      for (int i = 0; i < n; i++) z[0][i] = z[0][i] * (dtn / dt);
      dt = dtn;
      if (!(dt > 1e-14)) { status = BOLT_K_DT_UNDERFLOW; break; }
    }
  }
  if (st.rsa_seen && status == 0) status = BOLT_K_RSA_TRIGGERED;
  if (u_final) for (int i = 0; i < n; i++) store_T(u_final + (size_t)i * nd_out, nd_out, u[i]);
  status_out = status; nsteps_out = nsteps; nreject_out = nreject;
}

// ---------------------------------------------------------------------------------------------
// spectra.jl
// ---------------------------------------------------------------------------------------------
// spherical Bessel j_l(x) by Miller's downward recurrence (SpecialFunctions.sphericalbesselj is not
// vendored; any >= 1e-14-accurate j_l is equivalent).
double sph_bessel_j(int l, double x) {
  if (x == 0.0) return l == 0 ? 1.0 : 0.0;
  const double j0 = std::sin(x) / x;
  if (l == 0) return j0;
  const double j1 = std::sin(x) / (x * x) - std::cos(x) / x;
  if (l == 1) return j1;
  if (x > l + 0.5 + 10.0 * std::sqrt((double)l + 1.0)) {   // upward recurrence is stable for x >> l
    double jm = j0, jc = j1;
    for (int n = 1; n < l; n++) { double jn = (2 * n + 1) / x * jc - jm; jm = jc; jc = jn; }
    return jc;
  }
  int nstart = (int)(std::max((double)l, x) + 30.0 + 10.0 * std::sqrt(std::max((double)l, x)));
  double jp = 0.0, jc = 1e-280, jl = 0.0;
  for (int n = nstart; n >= 1; n--) {
    double jn = (2 * n + 1) / x * jc - jp;   // j_{n-1}
    jp = jc; jc = jn;
    if (n - 1 == l) jl = jc;
    if (std::fabs(jc) > 1e250) { jc *= 1e-250; jp *= 1e-250; jl *= 1e-250; }
  }
  // jc ~ j_0, jp ~ j_1 (unnormalised): normalise with the larger of the two
  if (std::fabs(j0) >= std::fabs(j1)) return jl * (j0 / jc);
  return jl * (j1 / jp);
}

// prefilter of spline(f, grid) for BSpline(Cubic(Line(OnGrid()))) (util.jl:11): c[1]=y[0], c[n]=y[n-1],
// interior (1/6, 2/3, 1/6), c[0] = 2c[1]-c[2], c[n+1] = 2c[n]-c[n-1].
void bspline_prefilter(const std::vector<double>& y, std::vector<double>& c) {
  const int n = (int)y.size();
  c.assign(n + 2, 0.0);
  c[1] = y[0]; c[n] = y[n - 1];
  const int m = n - 2;
  if (m > 0) {
    std::vector<double> cp(m), dp(m);
    const double a = 1.0 / 6.0, b = 2.0 / 3.0;
    for (int i = 0; i < m; i++) {
      double r = y[i + 1];
      if (i == 0) r -= a * y[0];
      if (i == m - 1) r -= a * y[n - 1];
      if (i == 0) { cp[0] = a / b; dp[0] = r / b; }
      else { double den = b - a * cp[i - 1]; cp[i] = a / den; dp[i] = (r - a * dp[i - 1]) / den; }
    }
    c[m + 1] = dp[m - 1];
    for (int i = m - 2; i >= 0; i--) c[i + 2] = dp[i] - cp[i] * c[i + 3];
  }
  c[0] = 2.0 * c[1] - c[2];
  c[n + 1] = 2.0 * c[n] - c[n - 1];
}

// cltt / clte / clee (spectra.jl:84-130) for one l, with Tl (:70-82), bessel_interpolator (:49-58)
// and the bilinear source interpolant with linear extrapolation in k (:21).
// T = double, or Du<NP>: sources [nk][n_x][nd], eta and eta0 (hence chi = eta0 - eta(x_i), the Bessel ARGUMENT), A and n carry
// partials; the table range kgrid[end]*eta0 does not (assume_nondual, spectra.jl:46-52) -- xmax_fixed > 0 overrides it.
template <class T>
void cl_one(const Cosmo_<T>& c, const double* S_T, const double* S_P, int nd, const double* kc, int nk, int ell,
            double kd_min, double kd_max, int n_kd, int ix_start, double xmax_fixed, T* tt, T* te, T* ee) {
  const T eta0 = c.s[BOLT_S_eta0], A = c.s[BOLT_S_A], ns = c.s[BOLT_S_n];
  const int n_x = c.n_x;
  // dense k grid: quadratic_k(kmin,kmax,n) (:60-63)
  std::vector<double> kd(n_kd);
  for (int i = 1; i <= n_kd; i++) { double r = (double)i / n_kd; kd[i - 1] = kd_min + (kd_max - kd_min) * (r * r); }
  // bessel_interpolator(l, kgrid[end]*η₀) (:49-58)
  const int NB = 5001;
  const double xmax = xmax_fixed > 0.0 ? xmax_fixed : kd[n_kd - 1] * val(eta0), dg = xmax / 5000.0;
  std::vector<double> y(NB), bc;
  for (int i = 0; i < NB; i++) y[i] = sph_bessel_j(ell, dg * i);
  bspline_prefilter(y, bc);
  // χ_i = η₀ - η(x_i), dx_i (:78-81)
  std::vector<T> chi(n_x); std::vector<double> dxs(n_x);
  for (int i = ix_start; i < n_x - 1; i++) {
    double xi = c.x0 + c.dx * i, xn = c.x0 + c.dx * (i + 1);
    chi[i] = eta0 - tab(c, BOLT_T_eta, xi);
    dxs[i] = xn - xi;
  }
  const double lfac = std::sqrt((double)(ell + 2) * (ell + 1) * ell * (ell - 1));   // :101,118
  T stt = T(0.0), ste = T(0.0), see = T(0.0);
  int jk = 0;
  for (int i = 0; i < n_kd - 1; i++) {                                           // :88-94
    const double k = (kd[i] + kd[i + 1]) / 2.0, dk = kd[i + 1] - kd[i];
    // bracket in the coarse grid; Line() extrapolation uses the end intervals
    while (jk < nk - 2 && kc[jk + 1] < k) jk++;
    while (jk > 0 && kc[jk] > k) jk--;
    const double w = (k - kc[jk]) / (kc[jk + 1] - kc[jk]);
    T th = T(0.0), ep = T(0.0);
    for (int ix = ix_start; ix < n_x - 1; ix++) {                                // Tl (:70-76)
      auto bes = spline_eval(bc, NB, 0.0, dg, k * chi[ix]);
      if (S_T) { T s0, s1; load_T(S_T + ((size_t)jk * n_x + ix) * nd, nd, s0); load_T(S_T + ((size_t)(jk + 1) * n_x + ix) * nd, nd, s1);
                 T s = (1.0 - w) * s0 + w * s1; th += bes * s * dxs[ix]; }
      if (S_P) { T s0, s1; load_T(S_P + ((size_t)jk * n_x + ix) * nd, nd, s0); load_T(S_P + ((size_t)(jk + 1) * n_x + ix) * nd, nd, s1);
                 T s = (1.0 - w) * s0 + w * s1; ep += bes * s * dxs[ix]; }
    }
    ep = ep * lfac;
    const T Pprim = A * m_pow(T(k / 0.05), ns - 1.0);                            // :92
    stt += th * th * Pprim * dk / k;
    ste += th * ep * Pprim * dk / k;
    see += ep * ep * Pprim * dk / k;
  }
  if (tt) *tt = 4.0 * M_PI * stt;
  if (te) *te = 4.0 * M_PI * ste;
  if (ee) *ee = 4.0 * M_PI * see;
}

// plin (spectra.jl:163-198) from the state at x = 0
template <class T> T plin_from_state(const Mode_<T>& h, const T* res) {
  const Cosmo_<T>& c = *h.c; const Layout& l = h.lay; const double k = h.k; const int nq = l.nq;
  const double x = 0.0;
  const T rho0M = tab(c, BOLT_T_rho0M, x), Hx = tab(c, BOLT_T_H, x);
  T rho, sig;
  rho_sigma(c, res + l.iM, res + l.iM + 2 * nq, std::exp(x), rho, sig);
  const T Mrho = rho / rho0M;                                                    // :170-172
  const T Mtheta = k * theta_nu(c, res + l.iM + nq) / rho0M;                     // :174-175
  const T dcN = res[l.iS + 1], dbN = res[l.iS + 3], vcN = res[l.iS + 2], vbN = res[l.iS + 4];
  const T vmnuN = -Mtheta / k;
  const T hh = c.s[BOLT_S_h], Om_r = c.s[BOLT_S_Omega_r], N_nu = c.s[BOLT_S_N_nu];
  const T Tg = m_pow(15.0 / (M_PI * M_PI) * c.s[BOLT_S_rho_crit] * Om_r, 0.25);
  const double zeta = 1.2020569;
  const T nufac = (90.0 * zeta / (11.0 * std::pow(M_PI, 4))) * (Om_r * hh * hh / Tg) * m_pow(N_nu / 3.0, 0.75);
  const T Om_nu = c.s[BOLT_S_Sum_m_nu] * nufac / (hh * hh);
  const T Om_c = c.s[BOLT_S_Omega_c], Om_b = c.s[BOLT_S_Omega_b];
  const T Om_m = Om_c + Om_b + Om_nu;
  const T dc = dcN - 3.0 * Hx * vcN / k, db = dbN - 3.0 * Hx * vbN / k;          // :189-190
  const T dmnu = Mrho - 3.0 * Hx * vmnuN / k;                                    // :192
  const T dm = (Om_c * dc + Om_b * db + Om_nu * dmnu) / Om_m;                    // :193
  const T Pprim = c.s[BOLT_S_A] * m_pow(T(k / 0.05), c.s[BOLT_S_n] - 1.0);
  return (2.0 * M_PI * M_PI / (k * k * k)) * dm * dm * Pprim;                    // :196
}

// the oracle handle: value tables always, the raw descriptor arrays kept for the dual instantiations
struct Handle {
  Cosmo c;
  int nd = 1;
  std::vector<double> scalars, tables, quad_pts, quad_wts;
  bolt_cosmo_desc desc;
};

template <int NP> void build_dual(const Handle& H, Cosmo_<Du<NP>>& cd) { cosmo_from_desc(&H.desc, cd); }

template <int NP>
int solve_sens_t(const Handle& H, const double* k, int nk, const bolt_opts* o, int lu_mode, double* S_T, double* S_P, double* u_final,
                 int32_t* status, int64_t* nsteps, int64_t* nreject) {
  Cosmo_<Du<NP>> cd; build_dual<NP>(H, cd);
  Layout lay(o->l_gamma, o->l_nu, o->l_mnu, H.c.nq);
  const int n = lay.n, n_x = H.c.n_x, nd = H.nd;
#pragma omp parallel for schedule(dynamic, 1)
  for (int j = 0; j < nk; j++) {
    const int ik = nk - 1 - j;
    Mode h{&H.c, k[ik], lay}; Mode_<Du<NP>> hd{&cd, k[ik], lay};
    int st = 0; int64_t ns = 0, nr = 0;
    solve_mode_sens<NP>(hd, h, *o, lu_mode != 0, nd, S_T ? S_T + (size_t)ik * n_x * nd : nullptr, S_P ? S_P + (size_t)ik * n_x * nd : nullptr,
                        u_final ? u_final + (size_t)ik * n * nd : nullptr, st, ns, nr);
    if (status) status[ik] = st;
    if (nsteps) nsteps[ik] = ns;
    if (nreject) nreject[ik] = nr;
  }
  return 0;
}

template <int NP>
int project_sens_t(const Handle& H, const double* S_T, const double* S_P, const double* k, int nk, const int32_t* ell, int nell,
                   double kd_min, double kd_max, int n_kd, int ix_start, double xmax_fixed, double* cl_tt, double* cl_te, double* cl_ee) {
  Cosmo_<Du<NP>> cd; build_dual<NP>(H, cd);
  const int nd = H.nd;
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < nell; i++) {
    Du<NP> tt, te, ee;
    cl_one(cd, S_T, S_P, nd, k, nk, ell[i], kd_min, kd_max, n_kd, ix_start, xmax_fixed, &tt, &te, &ee);
    if (cl_tt && S_T) store_T(cl_tt + (size_t)i * nd, nd, tt);
    if (cl_te && S_T && S_P) store_T(cl_te + (size_t)i * nd, nd, te);
    if (cl_ee && S_P) store_T(cl_ee + (size_t)i * nd, nd, ee);
  }
  return 0;
}

template <int NP>
int plin_sens_t(const Handle& H, const double* k, int nk, const bolt_opts* o, int lu_mode, double* pk, int32_t* status, int64_t* nsteps) {
  Cosmo_<Du<NP>> cd; build_dual<NP>(H, cd);
  Layout lay(o->l_gamma, o->l_nu, o->l_mnu, H.c.nq);
  const int nd = H.nd;
#pragma omp parallel for schedule(dynamic, 1)
  for (int j = 0; j < nk; j++) {
    const int ik = nk - 1 - j;
    Mode h{&H.c, k[ik], lay}; Mode_<Du<NP>> hd{&cd, k[ik], lay};
    std::vector<double> uf((size_t)lay.n * nd);
    bolt_opts oo = *o; oo.ix_first = H.c.n_x;
    int st = 0; int64_t ns = 0, nr = 0;
    solve_mode_sens<NP>(hd, h, oo, lu_mode != 0, nd, nullptr, nullptr, uf.data(), st, ns, nr);
    std::vector<Du<NP>> res(lay.n);
    for (int i = 0; i < lay.n; i++) load_T(uf.data() + (size_t)i * nd, nd, res[i]);
    store_T(pk + (size_t)ik * nd, nd, plin_from_state(hd, res.data()));
    if (status) status[ik] = st;
    if (nsteps) nsteps[ik] = ns;
  }
  return 0;
}

#define NP_DISPATCH(np, CALL)                                                                    \
  switch (np) {                                                                                  \
    case 1: return CALL(1); case 2: return CALL(2); case 3: return CALL(3); case 4: return CALL(4); \
    case 5: return CALL(5); case 6: return CALL(6); case 7: return CALL(7); case 8: return CALL(8); \
    default: return -1;                                                                          \
  }

}  // namespace

// ---------------------------------------------------------------------------------------------
// C entry points (ctypes).  Same descriptor structs as the product ABI so tests feed both alike.
// ---------------------------------------------------------------------------------------------
extern "C" {

void* oracle_cosmo_create(const bolt_cosmo_desc* d) {
  Handle* H = new Handle();
  H->nd = d->nd;
  const size_t nc = (size_t)d->n_x + 2;
  H->scalars.assign(d->scalars, d->scalars + (size_t)BOLT_NSCALARS * d->nd);
  H->tables.assign(d->tables, d->tables + (size_t)BOLT_NTABLES * nc * d->nd);
  H->quad_pts.assign(d->quad_pts, d->quad_pts + d->nq);
  H->quad_wts.assign(d->quad_wts, d->quad_wts + d->nq);
  H->desc = *d;
  H->desc.scalars = H->scalars.data(); H->desc.tables = H->tables.data();
  H->desc.quad_pts = H->quad_pts.data(); H->desc.quad_wts = H->quad_wts.data();
  cosmo_from_desc(&H->desc, H->c);
  return H;
}
void oracle_cosmo_free(void* c) { delete (Handle*)c; }
int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
// torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU-baseline legs of bench.py set the thread count explicitly
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
}

// RHS and IC at one (k, x) for unit tests of the kernels' building blocks
void oracle_initial_conditions(const void* cos, double k, const bolt_opts* o, double* u) {
  const Cosmo* c = &((const Handle*)cos)->c; Mode h{c, k, Layout(o->l_gamma, o->l_nu, o->l_mnu, c->nq)};
  initial_conditions(h, c->x0, u);
}
int oracle_hierarchy(const void* cos, double k, const bolt_opts* o, double x, double* u, double* du) {
  const Cosmo* c = &((const Handle*)cos)->c; Mode h{c, k, Layout(o->l_gamma, o->l_nu, o->l_mnu, c->nq)};
  return hierarchy(h, x, u, du) ? 1 : 0;
}
void oracle_source_functions(const void* cos, double k, const bolt_opts* o, double x, const double* u, const double* du, double* st, double* sp) {
  const Cosmo* c = &((const Handle*)cos)->c; Mode h{c, k, Layout(o->l_gamma, o->l_nu, o->l_mnu, c->nq)};
  source_functions(h, x, u, du, *st, *sp);
}
double oracle_spline_eval(const void* cos, int which, double x) { return tab(((const Handle*)cos)->c, which, x); }
double oracle_sph_bessel_j(int l, double x) { return sph_bessel_j(l, x); }

// bolt_solve analogue.  lu_mode: 0 = plain dense LU (the reference's choice), 1 = skip structural zeros.
int oracle_solve(const void* cos, const double* k, int nk, const bolt_opts* o, int lu_mode,
                 double* S_T, double* S_P, double* u_hist, double* u_final,
                 int32_t* status, int64_t* nsteps, int64_t* nreject) {
  const Cosmo* c = &((const Handle*)cos)->c;
  Layout lay(o->l_gamma, o->l_nu, o->l_mnu, c->nq);
  const int n = lay.n, n_x = c->n_x;
#pragma omp parallel for schedule(dynamic, 1)
  for (int j = 0; j < nk; j++) {
    const int ik = nk - 1 - j;   // largest k first (longest solves) for load balance
    Mode h{c, k[ik], lay};
    SolveOut out;
    if (S_T) out.S_T = S_T + (size_t)ik * n_x;
    if (S_P) out.S_P = S_P + (size_t)ik * n_x;
    if (u_hist) out.u_hist = u_hist + (size_t)ik * n_x * n;
    if (u_final) out.u_final = u_final + (size_t)ik * n;
    solve_mode(h, *o, lu_mode != 0, out);
    if (status) status[ik] = out.status;
    if (nsteps) nsteps[ik] = out.nsteps;
    if (nreject) nreject[ik] = out.nreject;
  }
  return 0;
}

// value + gradient in one pass (the cosmology must have been created with nd > 1): outputs carry nd doubles per element
//   S_T, S_P [nk][n_x][nd], u_final [nk][n][nd]
int oracle_solve_sens(const void* cos, const double* k, int nk, const bolt_opts* o, int lu_mode,
                      double* S_T, double* S_P, double* u_final, int32_t* status, int64_t* nsteps, int64_t* nreject) {
  const Handle& H = *(const Handle*)cos;
#define CALL(N) solve_sens_t<N>(H, k, nk, o, lu_mode, S_T, S_P, u_final, status, nsteps, nreject)
  NP_DISPATCH(H.nd - 1, CALL)
#undef CALL
}

int oracle_project(const void* cos, const double* S_T, const double* S_P, const double* k, int nk,
                   const int32_t* ell, int nell, double kd_min, double kd_max, int n_kd, int ix_start,
                   double* cl_tt, double* cl_te, double* cl_ee) {
  const Cosmo* c = &((const Handle*)cos)->c;
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < nell; i++)
    cl_one(*c, S_T, S_P, 1, k, nk, ell[i], kd_min, kd_max, n_kd, ix_start, 0.0,
           cl_tt ? cl_tt + i : nullptr, cl_te ? cl_te + i : nullptr, cl_ee ? cl_ee + i : nullptr);
  return 0;
}

// cltt/clte/clee with partials: S_T, S_P [nk][n_x][nd]; cl_* [nell][nd].  xmax_fixed > 0 fixes the j_l table range.
int oracle_project_sens(const void* cos, const double* S_T, const double* S_P, const double* k, int nk,
                        const int32_t* ell, int nell, double kd_min, double kd_max, int n_kd, int ix_start, double xmax_fixed,
                        double* cl_tt, double* cl_te, double* cl_ee) {
  const Handle& H = *(const Handle*)cos;
#define CALL(N) project_sens_t<N>(H, S_T, S_P, k, nk, ell, nell, kd_min, kd_max, n_kd, ix_start, xmax_fixed, cl_tt, cl_te, cl_ee)
  NP_DISPATCH(H.nd - 1, CALL)
#undef CALL
}

int oracle_plin(const void* cos, const double* k, int nk, const bolt_opts* o, int lu_mode, double* pk,
                int32_t* status, int64_t* nsteps) {
  const Cosmo* c = &((const Handle*)cos)->c;
  Layout lay(o->l_gamma, o->l_nu, o->l_mnu, c->nq);
#pragma omp parallel for schedule(dynamic, 1)
  for (int j = 0; j < nk; j++) {
    const int ik = nk - 1 - j;
    Mode h{c, k[ik], lay};
    std::vector<double> uf(lay.n);
    SolveOut out; out.u_final = uf.data();
    bolt_opts oo = *o; oo.ix_first = c->n_x;   // no source sampling needed
    solve_mode(h, oo, lu_mode != 0, out);
    pk[ik] = plin_from_state(h, uf.data());
    if (status) status[ik] = out.status;
    if (nsteps) nsteps[ik] = out.nsteps;
  }
  return 0;
}

// plin with partials: pk [nk][nd]
int oracle_plin_sens(const void* cos, const double* k, int nk, const bolt_opts* o, int lu_mode, double* pk,
                     int32_t* status, int64_t* nsteps) {
  const Handle& H = *(const Handle*)cos;
#define CALL(N) plin_sens_t<N>(H, k, nk, o, lu_mode, pk, status, nsteps)
  NP_DISPATCH(H.nd - 1, CALL)
#undef CALL
}

}  // extern "C"
