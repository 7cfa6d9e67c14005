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
// host tables
// ---------------------------------------------------------------------------------------------
struct Cosmo {
  int n_x = 0, nq = 0;
  double x0 = 0, dx = 0;
  double s[BOLT_NSCALARS];
  std::vector<double> tab[BOLT_NTABLES];
  std::vector<double> xq, wq;
  // per-cosmology constants of the momentum grid (perturbations.jl:164-166, background.jl:21-30)
  std::vector<double> q, f0w, dlnf0;  // q_i, f0(q_i)/dxdq(q_i)*w_i, dlnf0dlnq(q_i)
  double Tnu = 0, Omega_nu = 0;
};

// util.jl:11 -- Interpolations' BSpline(Cubic(Line(OnGrid()))) evaluation on a uniform grid.
inline double spline_eval(const std::vector<double>& c, int n_x, double x0, double dx, double x) {
  double t = (x - x0) / dx;
  int i = (int)std::floor(t);
  if (i < 0) i = 0;
  if (i > n_x - 2) i = n_x - 2;
  double d = t - i, e = 1.0 - d;
  double w0 = e * e * e / 6.0;
  double w1 = 2.0 / 3.0 - d * d + d * d * d / 2.0;
  double w2 = 2.0 / 3.0 - e * e + e * e * e / 2.0;
  double w3 = d * d * d / 6.0;
  return c[i] * w0 + c[i + 1] * w1 + c[i + 2] * w2 + c[i + 3] * w3;
}
inline double tab(const Cosmo& c, int which, double x) {
  return spline_eval(c.tab[which], c.n_x, c.x0, c.dx, x);
}

void cosmo_from_desc(const bolt_cosmo_desc* d, Cosmo& c) {
  c.n_x = d->n_x; c.nq = d->nq; c.x0 = d->x0; c.dx = d->dx;
  const int nd = d->nd;
  for (int i = 0; i < BOLT_NSCALARS; i++) c.s[i] = d->scalars[(size_t)i * nd];
  for (int t = 0; t < BOLT_NTABLES; t++) {
    c.tab[t].resize(d->n_x + 2);
    for (int i = 0; i < d->n_x + 2; i++) c.tab[t][i] = d->tables[((size_t)t * (d->n_x + 2) + i) * nd];
  }
  c.xq.assign(d->quad_pts, d->quad_pts + d->nq);
  c.wq.assign(d->quad_wts, d->quad_wts + d->nq);
  // Tν (perturbations.jl:164), q grid (:165-166 with util.jl:24-27), f0 and dlnf0dlnq (background.jl:21-30)
  const double N_nu = c.s[BOLT_S_N_nu], Om_r = c.s[BOLT_S_Omega_r], rho_crit = c.s[BOLT_S_rho_crit];
  c.Tnu = std::pow(N_nu / 3.0, 0.25) * std::pow(4.0 / 11.0, 1.0 / 3.0) *
          std::pow(15.0 / (M_PI * M_PI) * rho_crit * Om_r, 0.25);
  const double lqmi = std::log10(c.Tnu / 30.0), lqma = std::log10(c.Tnu * 30.0);
  c.q.resize(c.nq); c.f0w.resize(c.nq); c.dlnf0.resize(c.nq);
  for (int i = 0; i < c.nq; i++) {
    double lq = lqmi + (lqma - lqmi) / 2.0 * (c.xq[i] + 1.0);       // from_ui
    double q = std::pow(10.0, lq);                                   // xq2q
    double dxdq = (2.0 / (lqma - lqmi)) / (q * std::log(10.0));      // dxdq
    double f0 = 2.0 / std::pow(2.0 * M_PI, 3) / (std::exp(q / c.Tnu) + 1.0);
    c.q[i] = q;
    c.f0w[i] = f0 / dxdq * c.wq[i];
    c.dlnf0[i] = -q / c.Tnu / (1.0 + std::exp(-q / c.Tnu));
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

struct Mode {       // Hierarchy (perturbations.jl:7-21)
  const Cosmo* c;
  double k;
  Layout lay;
};

// ρ_σ (perturbations.jl:127-145): ρ = 4π Σ q² ε f0/dxdq ℳ0 w ;  σ = 4π Σ q² (q²/ε) f0/dxdq ℳ2 w
inline void rho_sigma(const Cosmo& c, const double* M0, const double* M2, double a, double& rho, double& sigma) {
  const double m = c.s[BOLT_S_Sum_m_nu];
  double r = 0, s = 0;
  for (int i = 0; i < c.nq; i++) {
    double q = c.q[i], eps = std::sqrt(q * q + (a * m) * (a * m));
    r += q * q * eps * c.f0w[i] * M0[i];
    s += q * q * (q * q / eps) * c.f0w[i] * M2[i];
  }
  rho = 4.0 * M_PI * r; sigma = 4.0 * M_PI * s;
}
// θ (perturbations.jl:148-158)
inline double theta_nu(const Cosmo& c, const double* M1) {
  double t = 0;
  for (int i = 0; i < c.nq; i++) t += c.q[i] * c.q[i] * c.q[i] * c.f0w[i] * M1[i];
  return 4.0 * M_PI * t;
}

// hierarchy! (perturbations.jl:161-271).  `u` is non-const because the RSA branch writes into it
// (:219-227), exactly like the reference.  Returns whether the RSA branch was taken.
bool hierarchy(const Mode& h, double x, double* u, double* du) {
  const Cosmo& c = *h.c; const Layout& l = h.lay;
  const double k = h.k; const int L = l.L, Lnu = l.Lnu, Lm = l.Lm, nq = l.nq;
  const double Om_r = c.s[BOLT_S_Omega_r], Om_b = c.s[BOLT_S_Omega_b], Om_c = c.s[BOLT_S_Omega_c];
  const double m_nu = c.s[BOLT_S_Sum_m_nu], H0 = c.s[BOLT_S_H0], H02 = H0 * H0, rho_crit = c.s[BOLT_S_rho_crit];
  const double Hx = tab(c, BOLT_T_H, x), eta = tab(c, BOLT_T_eta, x);
  const double taup = tab(c, BOLT_T_taup, x), taupp = tab(c, BOLT_T_taupp, x);
  const double a = std::exp(x);
  const double R = 4.0 * Om_r / (3.0 * Om_b * a);
  const double Om_nu = c.Omega_nu;
  const double csb2 = tab(c, BOLT_T_csb2, x);
  double *T = u + l.iT, *P = u + l.iP, *N = u + l.iN, *M = u + l.iM;
  double *dT = du + l.iT, *dP = du + l.iP, *dN = du + l.iN, *dM = du + l.iM;
  const double Phi = u[l.iS], delta = u[l.iS + 1], v = u[l.iS + 2], delta_b = u[l.iS + 3], v_b = u[l.iS + 4];

  double rhoM, sigM;
  rho_sigma(c, M, M + 2 * nq, a, rhoM, sigM);                                   // :182
  const double Psi = -Phi - 12.0 * H02 / (k * k) / (a * a) *                    // :184-187
                     (Om_r * T[2] + Om_nu * N[2] + sigM / rho_crit / 4.0);
  const double dPhi = Psi - k * k / (3.0 * Hx * Hx) * Phi + H02 / (2.0 * Hx * Hx) * (   // :189-194
                      Om_c / a * delta + Om_b / a * delta_b + 4.0 * Om_r / (a * a) * T[0] +
                      4.0 * Om_nu / (a * a) * N[0] + rhoM / (a * a) / rho_crit);
  const double ddelta = k / Hx * v - 3.0 * dPhi;                                // :197-200
  const double dv = -v - k / Hx * Psi;
  const double ddelta_b = k / Hx * v_b - 3.0 * dPhi;
  const double dv_b = -v_b - k / Hx * (Psi + csb2 * delta_b) + taup * R * (3.0 * T[1] + v_b);

  for (int iq = 0; iq < nq; iq++) {                                             // :203-213
    const double q = c.q[iq], eps = std::sqrt(q * q + (a * m_nu) * (a * m_nu)), df0 = c.dlnf0[iq];
    dM[0 * nq + iq] = -k / Hx * q / eps * M[1 * nq + iq] + dPhi * df0;
    dM[1 * nq + iq] = k / (3.0 * Hx) * (q / eps * (M[0 * nq + iq] - 2.0 * M[2 * nq + iq]) - eps / q * Psi * df0);
    for (int ell = 2; ell <= Lm - 1; ell++)
      dM[ell * nq + iq] = k / Hx * q / ((2 * ell + 1) * eps) * (ell * M[(ell - 1) * nq + iq] - (ell + 1) * M[(ell + 1) * nq + iq]);
    dM[Lm * nq + iq] = q / eps * k / Hx * M[(Lm - 1) * nq + iq] - (Lm + 1) / (Hx * eta) * M[Lm * nq + iq];
  }

  const bool rsa_on = (k * eta > 240.0) && (-taup * Hx / eta > 100.0);          // :216
  if (rsa_on) {                                                                 // :217-232
    T[0] = Phi - Hx / k * taup * v_b;
    T[1] = Hx / k * (-2.0 * dPhi + taup * (Phi - csb2 * delta_b) + Hx / k * (taupp - taup) * v_b);  // :221-222 overwrite :220
    T[2] = 0.0;
    N[0] = Phi; N[1] = -2.0 * Hx / k * dPhi; N[2] = 0.0;
    for (int ell = 0; ell <= Lnu; ell++) dN[ell] = 0.0;
    for (int ell = 0; ell <= L; ell++) { dT[ell] = 0.0; dP[ell] = 0.0; }
  } else {
    dN[0] = -k / Hx * N[1] - dPhi;                                              // :237-243
    dN[1] = k / (3.0 * Hx) * N[0] - 2.0 * k / (3.0 * Hx) * N[2] + k / (3.0 * Hx) * Psi;
    for (int ell = 2; ell <= Lnu - 1; ell++)
      dN[ell] = k / ((2 * ell + 1) * Hx) * (ell * N[ell - 1] - (ell + 1) * N[ell + 1]);
    dN[Lnu] = k / Hx * N[Lnu - 1] - (Lnu + 1) / (Hx * eta) * N[Lnu];
    const double Pi = T[2] + P[2] + P[0];                                       // :247-253
    dT[0] = -k / Hx * T[1] - dPhi;
    dT[1] = k / (3.0 * Hx) * T[0] - 2.0 * k / (3.0 * Hx) * T[2] + k / (3.0 * Hx) * Psi + taup * (T[1] + v_b / 3.0);
    for (int ell = 2; ell <= L - 1; ell++)
      dT[ell] = ell * k / ((2 * ell + 1) * Hx) * T[ell - 1] - (ell + 1) * k / ((2 * ell + 1) * Hx) * T[ell + 1] +
                taup * (T[ell] - Pi * (ell == 2 ? 1.0 : 0.0) / 10.0);
    dP[0] = -k / Hx * P[1] + taup * (P[0] - Pi / 2.0);                          // :256-260
    for (int ell = 1; ell <= L - 1; ell++)
      dP[ell] = ell * k / ((2 * ell + 1) * Hx) * P[ell - 1] - (ell + 1) * k / ((2 * ell + 1) * Hx) * P[ell + 1] +
                taup * (P[ell] - Pi * (ell == 2 ? 1.0 : 0.0) / 10.0);
    dT[L] = k / Hx * T[L - 1] - ((L + 1) / (Hx * eta) - taup) * T[L];           // :263-264
    dP[L] = k / Hx * P[L - 1] - ((L + 1) / (Hx * eta) - taup) * P[L];
  }
  du[l.iS] = dPhi; du[l.iS + 1] = ddelta; du[l.iS + 2] = dv; du[l.iS + 3] = ddelta_b; du[l.iS + 4] = dv_b;  // :269
  return rsa_on;
}

// initial_conditions (perturbations.jl:274-338)
void initial_conditions(const Mode& h, double xi, double* u) {
  const Cosmo& c = *h.c; const Layout& l = h.lay;
  const double k = h.k; const int L = l.L, Lnu = l.Lnu, Lm = l.Lm, nq = l.nq;
  std::fill(u, u + l.n, 0.0);
  const double Hx = tab(c, BOLT_T_H, xi), eta = tab(c, BOLT_T_eta, xi), taup = tab(c, BOLT_T_taup, xi);
  double *T = u + l.iT, *P = u + l.iP, *N = u + l.iN, *M = u + l.iM;
  const double ai2 = std::exp(xi) * std::exp(xi), ai = std::sqrt(ai2);
  const double N_nu = c.s[BOLT_S_N_nu];
  const double f_nu = 1.0 / (1.0 + 1.0 / (7.0 * (3.0 / 3.0) * N_nu / 8.0 * std::pow(4.0 / 11.0, 4.0 / 3.0)));   // :288
  const double Rc = 1.0;
  const double Phi = (4.0 * f_nu + 10.0) / (4.0 * f_nu + 15.0) * Rc;            // :292
  const double C = -((15.0 + 4.0 * f_nu) / (20.0 + 8.0 * f_nu)) * Phi;          // :294
  T[0] = -40.0 * C / (15.0 + 4.0 * f_nu) / 4.0;                                 // :297-302
  T[1] = 10.0 * C / (15.0 + 4.0 * f_nu) * (k * k * eta) / (3.0 * k);
  T[2] = -8.0 * k / (15.0 * Hx * taup) * T[1];
  P[0] = (5.0 / 4.0) * T[2];
  P[1] = -k / (4.0 * Hx * taup) * T[2];
  P[2] = (1.0 / 4.0) * T[2];
  for (int ell = 3; ell <= L; ell++) {                                          // :303-306
    T[ell] = -(double)ell / (2 * ell + 1) * k / (Hx * taup) * T[ell - 1];
    P[ell] = -(double)ell / (2 * ell + 1) * k / (Hx * taup) * P[ell - 1];
  }
  const double delta = 3.0 / 4.0 * (4.0 * T[0]);                                // :308-312
  const double delta_b = delta;
  const double v = -3.0 * k * T[1];
  const double v_b = v;
  N[0] = T[0];                                                                  // :316-321
  N[1] = T[1];
  N[2] = -(k * k * eta * eta) / 15.0 * 1.0 / (1.0 + 2.0 / 5.0 * f_nu) * Phi / 2.0;
  for (int ell = 3; ell <= Lnu; ell++) N[ell] = k / ((2 * ell + 1) * Hx) * N[ell - 1];
  const double m_nu = c.s[BOLT_S_Sum_m_nu];
  for (int iq = 0; iq < nq; iq++) {                                             // :325-334
    const double q = c.q[iq], eps = std::sqrt(q * q + (ai * m_nu) * (ai * m_nu)), df0 = c.dlnf0[iq];
    M[0 * nq + iq] = -N[0] * df0;
    M[1 * nq + iq] = -eps / q * N[1] * df0;
    M[2 * nq + iq] = -N[2] * df0;
    for (int ell = 3; ell <= Lm; ell++)
      M[ell * nq + iq] = q / eps * k / ((2 * ell + 1) * Hx) * M[(ell - 1) * nq + iq];
  }
  u[l.iS] = Phi; u[l.iS + 1] = delta; u[l.iS + 2] = v; u[l.iS + 3] = delta_b; u[l.iS + 4] = v_b;   // :336
}

// source_function (perturbations.jl:343-383) and source_function_P (:386-404)
void source_functions(const Mode& h, double x, const double* u, const double* du, double& S_T, double& S_P) {
  const Cosmo& c = *h.c; const Layout& l = h.lay;
  const double k = h.k; const int nq = l.nq;
  const double H0 = c.s[BOLT_S_H0], H02 = H0 * H0, Om_r = c.s[BOLT_S_Omega_r], rho_crit = c.s[BOLT_S_rho_crit];
  const double Hx = tab(c, BOLT_T_H, x), Hp = tab(c, BOLT_T_Hp, x), Hpp = tab(c, BOLT_T_Hpp, x);
  const double tau = tab(c, BOLT_T_tau, x), taup = tab(c, BOLT_T_taup, x), taupp = tab(c, BOLT_T_taupp, x);
  const double g = tab(c, BOLT_T_g, x), gp = tab(c, BOLT_T_gp, x), gpp = tab(c, BOLT_T_gpp, x);
  const double a = std::exp(x);
  const double Om_nu = c.Omega_nu;
  const double *T = u + l.iT, *P = u + l.iP, *N = u + l.iN, *M = u + l.iM;
  const double *dT = du + l.iT, *dP = du + l.iP, *dN = du + l.iN, *dM = du + l.iM;
  const double Phi = u[l.iS], v_b = u[l.iS + 4];
  const double dPhi = du[l.iS], dv_b = du[l.iS + 4];
  double r_, sigM, sigMp;
  rho_sigma(c, M, M + 2 * nq, a, r_, sigM);                                     // :361
  rho_sigma(c, dM, dM + 2 * nq, a, r_, sigMp);                                  // :362
  const double Psi = -Phi - 12.0 * H02 / (k * k) / (a * a) * (Om_r * T[2] + Om_nu * N[2] + sigM / rho_crit / 4.0);
  const double dPsi = -dPhi - 12.0 * H02 / (k * k) / (a * a) * (Om_r * (dT[2] - 2.0 * T[2]) + Om_nu * (dN[2] - 2.0 * N[2]) +
                                                                 (sigMp - 2.0 * sigM) / rho_crit / 4.0);   // :368-370
  const double Pi = T[2] + P[2] + P[0];
  const double dPi = dT[2] + dP[2] + dP[0];
  const double term1 = g * (T[0] + Psi + Pi / 4.0) + std::exp(-tau) * (dPsi - dPhi);        // :375
  const double term2 = (-1.0 / k) * (Hp * g * v_b + Hx * gp * v_b + Hx * g * dv_b);          // :376
  const double ddPi = 2.0 * k / (5.0 * Hx) * (-Hp / Hx * T[1] + dT[1]) + (3.0 / 10.0) * (taupp * Pi + taup * dPi) -
                      3.0 * k / (5.0 * Hx) * (-Hp / Hx * (T[3] + P[1] + P[3]) + (dT[3] + dP[1] + dP[3]));   // :377-378
  const double term3 = (3.0 / (4.0 * k * k)) * ((Hp * Hp + Hx * Hpp) * g * Pi + 3.0 * Hx * Hp * (gp * Pi + g * dPi) +
                                                Hx * Hx * (gpp * Pi + 2.0 * gp * dPi + g * ddPi));           // :379-381
  S_T = term1 + term2 + term3;
  const double x_end = c.x0 + c.dx * (c.n_x - 1);
  const double y = k * (tab(c, BOLT_T_eta, x_end) - tab(c, BOLT_T_eta, x));     // :401
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
void cl_one(const Cosmo& c, const double* S_T, const double* S_P, const double* kc, int nk, int ell,
            double kd_min, double kd_max, int n_kd, int ix_start, double* tt, double* te, double* ee) {
  const double eta0 = c.s[BOLT_S_eta0], A = c.s[BOLT_S_A], ns = c.s[BOLT_S_n];
  const int n_x = c.n_x;
  // dense k grid: quadratic_k(kmin,kmax,n) (:60-63)
  std::vector<double> kd(n_kd);
  for (int i = 1; i <= n_kd; i++) { double r = (double)i / n_kd; kd[i - 1] = kd_min + (kd_max - kd_min) * (r * r); }
  // bessel_interpolator(l, kgrid[end]*η₀) (:49-58)
  const int NB = 5001;
  const double xmax = kd[n_kd - 1] * eta0, dg = xmax / 5000.0;
  std::vector<double> y(NB), bc;
  for (int i = 0; i < NB; i++) y[i] = sph_bessel_j(ell, dg * i);
  bspline_prefilter(y, bc);
  // χ_i = η₀ - η(x_i), dx_i (:78-81)
  std::vector<double> chi(n_x), dxs(n_x);
  for (int i = ix_start; i < n_x - 1; i++) {
    double xi = c.x0 + c.dx * i, xn = c.x0 + c.dx * (i + 1);
    chi[i] = eta0 - tab(c, BOLT_T_eta, xi);
    dxs[i] = xn - xi;
  }
  const double lfac = std::sqrt((double)(ell + 2) * (ell + 1) * ell * (ell - 1));   // :101,118
  double stt = 0, ste = 0, see = 0;
  int jk = 0;
  for (int i = 0; i < n_kd - 1; i++) {                                           // :88-94
    const double k = (kd[i] + kd[i + 1]) / 2.0, dk = kd[i + 1] - kd[i];
    // bracket in the coarse grid; Line() extrapolation uses the end intervals
    while (jk < nk - 2 && kc[jk + 1] < k) jk++;
    while (jk > 0 && kc[jk] > k) jk--;
    const double w = (k - kc[jk]) / (kc[jk + 1] - kc[jk]);
    double th = 0, ep = 0;
    for (int ix = ix_start; ix < n_x - 1; ix++) {                                // Tl (:70-76)
      double bes = spline_eval(bc, NB, 0.0, dg, k * chi[ix]);
      if (S_T) { double s = (1.0 - w) * S_T[(size_t)jk * n_x + ix] + w * S_T[(size_t)(jk + 1) * n_x + ix]; th += bes * s * dxs[ix]; }
      if (S_P) { double s = (1.0 - w) * S_P[(size_t)jk * n_x + ix] + w * S_P[(size_t)(jk + 1) * n_x + ix]; ep += bes * s * dxs[ix]; }
    }
    ep *= lfac;
    const double Pprim = A * std::pow(k / 0.05, ns - 1.0);                       // :92
    stt += th * th * Pprim * dk / k;
    ste += th * ep * Pprim * dk / k;
    see += ep * ep * Pprim * dk / k;
  }
  if (tt) *tt = 4.0 * M_PI * stt;
  if (te) *te = 4.0 * M_PI * ste;
  if (ee) *ee = 4.0 * M_PI * see;
}

// plin (spectra.jl:163-198) from the state at x = 0
double plin_from_state(const Mode& h, const double* res) {
  const Cosmo& c = *h.c; const Layout& l = h.lay; const double k = h.k; const int nq = l.nq;
  const double x = 0.0;
  const double rho0M = tab(c, BOLT_T_rho0M, x), Hx = tab(c, BOLT_T_H, x);
  double rho, sig;
  rho_sigma(c, res + l.iM, res + l.iM + 2 * nq, std::exp(x), rho, sig);
  const double Mrho = rho / rho0M;                                               // :170-172
  const double Mtheta = k * theta_nu(c, res + l.iM + nq) / rho0M;                // :174-175
  const double dcN = res[l.iS + 1], dbN = res[l.iS + 3], vcN = res[l.iS + 2], vbN = res[l.iS + 4];
  const double vmnuN = -Mtheta / k;
  const double hh = c.s[BOLT_S_h], Om_r = c.s[BOLT_S_Omega_r], N_nu = c.s[BOLT_S_N_nu];
  const double Tg = std::pow(15.0 / (M_PI * M_PI) * c.s[BOLT_S_rho_crit] * Om_r, 0.25);
  const double zeta = 1.2020569;
  const double nufac = (90.0 * zeta / (11.0 * std::pow(M_PI, 4))) * (Om_r * hh * hh / Tg) * std::pow(N_nu / 3.0, 0.75);
  const double Om_nu = c.s[BOLT_S_Sum_m_nu] * nufac / (hh * hh);
  const double Om_c = c.s[BOLT_S_Omega_c], Om_b = c.s[BOLT_S_Omega_b];
  const double Om_m = Om_c + Om_b + Om_nu;
  const double dc = dcN - 3.0 * Hx * vcN / k, db = dbN - 3.0 * Hx * vbN / k;     // :189-190
  const double dmnu = Mrho - 3.0 * Hx * vmnuN / k;                               // :192
  const double dm = (Om_c * dc + Om_b * db + Om_nu * dmnu) / Om_m;               // :193
  const double Pprim = c.s[BOLT_S_A] * std::pow(k / 0.05, c.s[BOLT_S_n] - 1.0);
  return (2.0 * M_PI * M_PI / (k * k * k)) * dm * dm * Pprim;                    // :196
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C entry points (ctypes).  Same descriptor structs as the product ABI so tests feed both alike.
// ---------------------------------------------------------------------------------------------
extern "C" {

void* oracle_cosmo_create(const bolt_cosmo_desc* d) { Cosmo* c = new Cosmo(); cosmo_from_desc(d, *c); return c; }
void oracle_cosmo_free(void* c) { delete (Cosmo*)c; }
int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// RHS and IC at one (k, x) for unit tests of the kernels' building blocks
void oracle_initial_conditions(const void* cos, double k, const bolt_opts* o, double* u) {
  const Cosmo* c = (const Cosmo*)cos; Mode h{c, k, Layout(o->l_gamma, o->l_nu, o->l_mnu, c->nq)};
  initial_conditions(h, c->x0, u);
}
int oracle_hierarchy(const void* cos, double k, const bolt_opts* o, double x, double* u, double* du) {
  const Cosmo* c = (const Cosmo*)cos; Mode h{c, k, Layout(o->l_gamma, o->l_nu, o->l_mnu, c->nq)};
  return hierarchy(h, x, u, du) ? 1 : 0;
}
void oracle_source_functions(const void* cos, double k, const bolt_opts* o, double x, const double* u, const double* du, double* st, double* sp) {
  const Cosmo* c = (const Cosmo*)cos; Mode h{c, k, Layout(o->l_gamma, o->l_nu, o->l_mnu, c->nq)};
  source_functions(h, x, u, du, *st, *sp);
}
double oracle_spline_eval(const void* cos, int which, double x) { return tab(*(const Cosmo*)cos, which, x); }
double oracle_sph_bessel_j(int l, double x) { return sph_bessel_j(l, x); }

// bolt_solve analogue.  lu_mode: 0 = plain dense LU (the reference's choice), 1 = skip structural zeros.
int oracle_solve(const void* cos, const double* k, int nk, const bolt_opts* o, int lu_mode,
                 double* S_T, double* S_P, double* u_hist, double* u_final,
                 int32_t* status, int64_t* nsteps, int64_t* nreject) {
  const Cosmo* c = (const Cosmo*)cos;
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

int oracle_project(const void* cos, const double* S_T, const double* S_P, const double* k, int nk,
                   const int32_t* ell, int nell, double kd_min, double kd_max, int n_kd, int ix_start,
                   double* cl_tt, double* cl_te, double* cl_ee) {
  const Cosmo* c = (const Cosmo*)cos;
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < nell; i++)
    cl_one(*c, S_T, S_P, k, nk, ell[i], kd_min, kd_max, n_kd, ix_start,
           cl_tt ? cl_tt + i : nullptr, cl_te ? cl_te + i : nullptr, cl_ee ? cl_ee + i : nullptr);
  return 0;
}

int oracle_plin(const void* cos, const double* k, int nk, const bolt_opts* o, int lu_mode, double* pk,
                int32_t* status, int64_t* nsteps) {
  const Cosmo* c = (const Cosmo*)cos;
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

}  // extern "C"
