// K1 -- batched Einstein-Boltzmann hierarchy solve, one warp per k-mode (sm_100a, FP64).
//
// Replaces, for a whole batch of wavenumbers at once:
//   boltsolve                 src/perturbations.jl:25-33   (KenCarp4 ESDIRK, adaptive or fixed step)
//   initial_conditions        src/perturbations.jl:274-338
//   hierarchy!                src/perturbations.jl:161-271
//   the sampling loop of source_grid / source_grid_P (src/spectra.jl:13-18, 32-37) with
//   source_function / source_function_P (src/perturbations.jl:343-404), both from ONE solve.
//
// Layout.  The state of one mode lives in that warp's shared memory in the reference's own order
// (unpack, perturbations.jl:114-125).  Lanes own "chains": lane q < nq owns the massive-neutrino
// multipoles M[l*nq+q] (l = 0..l_mnu), lane nq owns Theta_l, lane nq+1 owns ThetaP_l, lane nq+2 owns
// N_l.  The l+-1 couplings of a chain are therefore a sequential sweep inside one lane; the couplings
// BETWEEN chains (Psi, Phi', Pi, v_b, the q-integrals rho_M, sigma_M) are warp shuffles.
//
// Implicit stages.  The hierarchy is linear, u' = A(x) u, so every ESDIRK stage is the linear system
// (I - gamma*dt*A(x_s)) U = rhs.  A is block-tridiagonal over the chains plus a rank-4 coupling through
// y = (Phi', Psi, Pi, v_b) and the five metric/matter scalars.  factor() runs a pivot-free downward
// elimination of every chain (stable: off-diagonals have opposite signs and the diagonal is >= 1), carries
// the y-dependence of rows l = 2,1,0 as 4-vectors, reduces them over the warp into a 4x4 system (partial
// pivoting, registers).  solve() repeats the sweep for a right-hand side, solves the 4x4 system and
// back-substitutes.  The factorisation costs O(n) like a solve, so it is redone at every stage abscissa
// and the stage equations are solved exactly (the oracle does the same with a dense LU).
#pragma once
#include "common.cuh"

namespace bolt {

struct SolveParams {
  const DevCosmo* const* cos_list;   // [ncos] device pointers; work item g belongs to cosmology g / nk_per
  int nk_per;
  const double* k;        // [nk] (nk = ncos * nk_per)
  const int* order;       // [nk] work order (largest k first)
  int nk;
  int L, Lnu, Lm, n;
  int mode;
  double reltol, abstol, fixed_dt;
  long long max_steps;
  int ix_first;
  double* S_T;            // [nk][n_x] or null
  double* S_P;            // [nk][n_x] or null
  double* u_hist;         // [nk][n_x][n] or null
  double* u_final;        // [nk][n] or null
  int* status;            // [nk]
  long long* nsteps;      // [nk]
  long long* nreject;     // [nk]
  int* counter;           // work queue head
  int out_nd;             // doubles per output element of S_T, S_P, u_final (1 + number of partials the CALLER carries)
  int comp_map[MAX_NP];   // dual kernels: output slot (1-based partial index) of the kernel's partial j
  double* dbg;            // optional step log of mode 0: [dbg_cap][4] = x, dt, EEst, accepted
  int dbg_cap;
  const DevCosmo* const* view_list;   // dual kernels, one cosmology: single-partial views [np] of the K1 cosmology (or null)
  int sync_mask;          // lockstep kernel: bit s set = the warps of a block meet before stage s (1..6)
};

static __constant__ double c_rl[MAX_L + 1];   // l/(2l+1)
static __constant__ double c_rl1[MAX_L + 1];  // 1 - l/(2l+1)

// KenCarp4 implicit tableau (Kennedy & Carpenter 2003, ARK4(3)6L[2]SA-ESDIRK); SURVEY 8c.
#define KC_GAMMA 0.25
static __device__ __constant__ double KC_C[6] = {0.0, 0.5, 83.0 / 250.0, 31.0 / 50.0, 17.0 / 20.0, 1.0};
static __device__ __constant__ double KC_A[6][5] = {
    {0, 0, 0, 0, 0},
    {1.0 / 4.0, 0, 0, 0, 0},
    {8611.0 / 62500.0, -1743.0 / 31250.0, 0, 0, 0},
    {5012029.0 / 34652500.0, -654441.0 / 2922500.0, 174375.0 / 388108.0, 0, 0},
    {15267082809.0 / 155376265600.0, -71443401.0 / 120774400.0, 730878875.0 / 902184768.0, 2285395.0 / 8070912.0, 0},
    {82889.0 / 524892.0, 0.0, 15625.0 / 83664.0, 69875.0 / 102672.0, -2260.0 / 8211.0}};
// b - bhat (b = last row of A plus gamma)
static __device__ __constant__ double KC_E[6] = {
    82889.0 / 524892.0 - 4586570599.0 / 29645900160.0, 0.0,
    15625.0 / 83664.0 - 178811875.0 / 945068544.0, 69875.0 / 102672.0 - 814220225.0 / 1159782912.0,
    -2260.0 / 8211.0 + 3700637.0 / 11593932.0, 0.25 - 61727.0 / 225920.0};

enum ChainKind { CH_M = 0, CH_T = 1, CH_P = 2, CH_N = 3, CH_IDLE = 4 };

// Everything a lane needs to know about the warp's mode and its own chain.
struct Lane {
  int lane, kind, base, stride, len;   // chain element l is at base + l*stride (internal shared-memory layout)
  int rbase, rstride, riS;             // the same element in the reference's unpack order (outputs)
  int nq, L, iS, n, maxlen;
  double k;
  double q, df0, wq;                   // massive-neutrino lanes only
};

// Background / ionization quantities at one abscissa (warp-uniform) plus the lane's q/eps.
struct Bg {
  double x, a, H, eta, taup, taupp, csb2, Hp;
  double kappa;     // k / H
  double qe, eq;    // q/eps, eps/q for M lanes (1 otherwise)
  double wPsi, wPhi;  // lane weights of chain rows 2 and 0 in the Psi and Phi' sums (see eval_bg)
  double cPsi, gPhi, k2;  // 12 H0^2/(k^2 a^2), H0^2/(2 H^2), k^2/(3 H^2)
  double R;
};

__device__ __forceinline__ void eval_bg(const DevCosmo& c, const Lane& ln, double x, Bg& b) {
  // lanes 0..5 evaluate one table each, then broadcast (perturbations.jl:168,172)
  const int which[6] = {BOLT_T_H, BOLT_T_eta, BOLT_T_taup, BOLT_T_taupp, BOLT_T_csb2, BOLT_T_Hp};
  double v = 0.0;
  if (ln.lane < 6) v = spline_eval(c.tab[which[ln.lane]], c.n_x, c.x0, c.dx, x);
  b.x = x;
  b.H = shfl_d(v, 0); b.eta = shfl_d(v, 1); b.taup = shfl_d(v, 2); b.taupp = shfl_d(v, 3);
  b.csb2 = shfl_d(v, 4); b.Hp = shfl_d(v, 5);
  b.a = exp(x);
  b.kappa = ln.k / b.H;
  const double Om_r = c.s[BOLT_S_Omega_r], Om_b = c.s[BOLT_S_Omega_b], rho_crit = c.s[BOLT_S_rho_crit];
  const double H0 = c.s[BOLT_S_H0];
  b.R = 4.0 * Om_r / (3.0 * Om_b * b.a);                       // :170
  b.cPsi = 12.0 * H0 * H0 / (ln.k * ln.k) / (b.a * b.a);       // :184
  b.gPhi = H0 * H0 / (2.0 * b.H * b.H);                        // :189
  b.k2 = ln.k * ln.k / (3.0 * b.H * b.H);
  b.qe = 1.0; b.eq = 1.0; b.wPsi = 0.0; b.wPhi = 0.0;
  const double ia2 = 1.0 / (b.a * b.a);
  if (ln.kind == CH_M) {
    const double am = b.a * c.s[BOLT_S_Sum_m_nu];
    const double eps = sqrt(ln.q * ln.q + am * am);            // :204
    b.qe = ln.q / eps; b.eq = eps / ln.q;
    // rho_M = sum wq*eps*M0, sigma_M = sum wq*(q^2/eps)*M2   (rho_sigma, :127-145)
    b.wPhi = ln.wq * eps * ia2 / rho_crit;                     // a^-2 rho_M / rho_crit        (:193)
    b.wPsi = ln.wq * (ln.q * ln.q / eps) / rho_crit * 0.25;    // sigma_M / rho_crit / 4       (:186)
  } else if (ln.kind == CH_T) {
    b.wPhi = 4.0 * Om_r * ia2; b.wPsi = Om_r;                  // :191, :184
  } else if (ln.kind == CH_N) {
    b.wPhi = 4.0 * c.Omega_nu * ia2; b.wPsi = c.Omega_nu;      // :192, :185
  }
}

// ---------------------------------------------------------------------------------------------------
// Right-hand side rows (hierarchy!, perturbations.jl:161-271) for the lane's chain, rows 0..nrows-1,
// reading chain values through `get(l)`.  Psi, Phi' and Pi are already known.
// ---------------------------------------------------------------------------------------------------
struct Metric { double Phi, delta, v, delta_b, v_b, Psi, dPhi, Pi; };

template <class Get>
__device__ __forceinline__ double rhs_row(const Lane& ln, const Bg& b, const Metric& m, int l, Get get) {
  const bool photon = (ln.kind == CH_T || ln.kind == CH_P);
  const double kq = b.kappa * b.qe;
  if (l == ln.len - 1) {   // truncation rows :212, :243, :263-264
    return kq * get(l - 1) - ((double)ln.len / (b.H * b.eta) - (photon ? b.taup : 0.0)) * get(l);
  }
  if (l == 0) {            // :207, :237, :248, :256
    double r = -kq * get(1);
    if (ln.kind == CH_M) r += m.dPhi * ln.df0;
    else if (ln.kind == CH_P) r += b.taup * (get(0) - m.Pi * 0.5);
    else r -= m.dPhi;
    return r;
  }
  const double rl = c_rl[l];
  double r = kq * (rl * get(l - 1) - (1.0 - rl) * get(l + 1));   // :208-210, :238-240, :249-252, :258-259
  if (l == 1) {
    if (ln.kind == CH_M) r -= b.kappa * (1.0 / 3.0) * b.eq * m.Psi * ln.df0;
    else if (ln.kind != CH_P) r += b.kappa * (1.0 / 3.0) * m.Psi;
    if (ln.kind == CH_T) r += b.taup * (m.v_b * (1.0 / 3.0));
  }
  if (photon) r += b.taup * (get(l) - (l == 2 ? m.Pi * 0.1 : 0.0));
  return r;
}

// Psi, Phi', Pi from row-0 and row-2 chain values of every lane (perturbations.jl:182-194, 247)
__device__ __forceinline__ void metric_from_chains(const Lane& ln, const Bg& b, double c0, double c2, Metric& m, const DevCosmo& c) {
  const double sPsi = warp_sum(b.wPsi * c2);
  const double sPhi = warp_sum(b.wPhi * c0);
  double pi = 0.0;
  if (ln.kind == CH_T) pi = c2; else if (ln.kind == CH_P) pi = c2 + c0;
  m.Pi = warp_sum(pi);
  m.Psi = -m.Phi - b.cPsi * sPsi;
  m.dPhi = m.Psi - b.k2 * m.Phi + b.gPhi * (c.s[BOLT_S_Omega_c] / b.a * m.delta + c.s[BOLT_S_Omega_b] / b.a * m.delta_b + sPhi);
}

// ---------------------------------------------------------------------------------------------------
// Stage matrix W = I - h*A(x):  factorisation state kept in registers between factor() and solve().
// ---------------------------------------------------------------------------------------------------
struct Factor {
  double h, hk, dtau;           // gamma*dt, h*kappa*q/eps, -h*tau' (photon lanes)
  double beta0[4], beta1[4], beta2[4];   // y-dependence of the lane's U_0, U_1, U_2
  double lo1, lo2;              // sub-diagonals of rows 1, 2
  double lu[4][4]; int perm[4]; // 4x4 border system, LU with partial pivoting
  // constants needed again to build the border right-hand side
  double hkap, vden;            // h*kappa, 1/(1+h)
  double e4c;                   // -3 h tau' R
};

__device__ __forceinline__ void chain_coefs(const Lane& ln, const Factor& f, const Bg& b, int l, double& bd, double& up, double& lo) {
  if (l == ln.len - 1) {
    bd = 1.0 + f.h * (double)ln.len / (b.H * b.eta) + f.dtau; up = 0.0; lo = -f.hk;
  } else {
    const double rl = c_rl[l];
    bd = 1.0 + ((l >= 1 || ln.kind == CH_P) ? f.dtau : 0.0);
    up = f.hk * (1.0 - rl); lo = -f.hk * rl;
  }
}

__device__ __forceinline__ void lu4_factor(Factor& f) {
#pragma unroll
  for (int i = 0; i < 4; i++) f.perm[i] = i;
#pragma unroll
  for (int kx = 0; kx < 4; kx++) {
    int p = kx; double best = fabs(f.lu[kx][kx]);
#pragma unroll
    for (int i = kx + 1; i < 4; i++) { double v = fabs(f.lu[i][kx]); if (v > best) { best = v; p = i; } }
    if (p != kx) {
#pragma unroll
      for (int i = kx + 1; i < 4; i++) if (i == p) {
#pragma unroll
        for (int j = 0; j < 4; j++) { double t = f.lu[kx][j]; f.lu[kx][j] = f.lu[i][j]; f.lu[i][j] = t; }
        int t = f.perm[kx]; f.perm[kx] = f.perm[i]; f.perm[i] = t;
      }
    }
    const double ip = 1.0 / f.lu[kx][kx];
#pragma unroll
    for (int i = kx + 1; i < 4; i++) {
      const double mlt = f.lu[i][kx] * ip; f.lu[i][kx] = mlt;
#pragma unroll
      for (int j = kx + 1; j < 4; j++) f.lu[i][j] -= mlt * f.lu[kx][j];
    }
  }
}
__device__ __forceinline__ void lu4_solve(const Factor& f, const double* rhs, double* y) {
  double t[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double v = 0.0;
#pragma unroll
    for (int j = 0; j < 4; j++) if (f.perm[i] == j) v = rhs[j];
    t[i] = v;
  }
#pragma unroll
  for (int i = 1; i < 4; i++)
#pragma unroll
    for (int j = 0; j < i; j++) t[i] -= f.lu[i][j] * t[j];
#pragma unroll
  for (int i = 3; i >= 0; i--) {
#pragma unroll
    for (int j = i + 1; j < 4; j++) t[i] -= f.lu[i][j] * t[j];
    t[i] /= f.lu[i][i];
  }
#pragma unroll
  for (int i = 0; i < 4; i++) y[i] = t[i];
}

// Factor W = I - h A(x_s).  Writes the inverse chain pivots to ib[] (shared memory).
__device__ __forceinline__ void factor(const DevCosmo& c, const Lane& ln, const Bg& b, double h, double* ib, Factor& f) {
  const bool photon = (ln.kind == CH_T || ln.kind == CH_P);
  f.h = h; f.hk = h * b.kappa * b.qe; f.dtau = photon ? -h * b.taup : 0.0;
  f.hkap = h * b.kappa; f.vden = 1.0 / (1.0 + h);
  f.e4c = -3.0 * h * b.taup * b.R;
  double ibn = 0.0, lo_next = 0.0;
  for (int l = ln.maxlen - 1; l >= 3; l--) {
    if (l < ln.len) {
      double bd, up, lo; chain_coefs(ln, f, b, l, bd, up, lo);
      const double bp = bd - (up * ibn) * lo_next;
      ibn = 1.0 / bp; lo_next = lo;
      ib[ln.base + l * ln.stride] = ibn;
    }
  }
  // rows 2, 1, 0 with the coupling to y = (Phi', Psi, Pi, v_b)
  double C0[4] = {0, 0, 0, 0}, C1[4] = {0, 0, 0, 0}, C2[4] = {0, 0, 0, 0};
  if (ln.kind == CH_M) { C0[0] = h * ln.df0; C1[1] = -f.hkap * (1.0 / 3.0) * b.eq * ln.df0; }
  else if (ln.kind == CH_T) { C0[0] = -h; C1[1] = f.hkap * (1.0 / 3.0); C1[3] = h * b.taup * (1.0 / 3.0); C2[2] = -h * b.taup * 0.1; }
  else if (ln.kind == CH_P) { C0[2] = -h * b.taup * 0.5; C2[2] = -h * b.taup * 0.1; }
  else if (ln.kind == CH_N) { C0[0] = -h; C1[1] = f.hkap * (1.0 / 3.0); }
  double ib2 = 0, ib1 = 0, ib0 = 0, m1 = 0, m0 = 0, lo1 = 0, lo2 = 0;
  if (ln.kind != CH_IDLE) {
    double bd, up, lo;
    chain_coefs(ln, f, b, 2, bd, up, lo);
    ib2 = 1.0 / (bd - (up * ibn) * lo_next); lo2 = lo;
    chain_coefs(ln, f, b, 1, bd, up, lo);
    m1 = up * ib2; ib1 = 1.0 / (bd - m1 * lo2); lo1 = lo;
    chain_coefs(ln, f, b, 0, bd, up, lo);
    m0 = up * ib1; ib0 = 1.0 / (bd - m0 * lo1);
    ib[ln.base + 2 * ln.stride] = ib2; ib[ln.base + ln.stride] = ib1; ib[ln.base] = ib0;
  }
  f.lo1 = lo1; f.lo2 = lo2;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const double V2 = C2[j], V1 = C1[j] - m1 * V2, V0 = C0[j] - m0 * V1;
    f.beta0[j] = V0 * ib0;
    f.beta1[j] = (V1 - lo1 * f.beta0[j]) * ib1;
    f.beta2[j] = (V2 - lo2 * f.beta1[j]) * ib2;
  }
  // reduce the y-dependence of the Psi, Phi', Pi sums and fetch Theta_1's from the Theta lane
  double sPsi[4], sPhi[4], sPi[4], t1[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    sPsi[j] = warp_sum(b.wPsi * f.beta2[j]);
    sPhi[j] = warp_sum(b.wPhi * f.beta0[j]);
    double pi = 0.0;
    if (ln.kind == CH_T) pi = f.beta2[j]; else if (ln.kind == CH_P) pi = f.beta2[j] + f.beta0[j];
    sPi[j] = warp_sum(pi);
    t1[j] = shfl_d(f.beta1[j], ln.nq);
  }
  // border equations E1..E4 in y (see DESIGN.md "stage system"); perturbations.jl:184-200
  const double Oc = c.s[BOLT_S_Omega_c] / b.a, Ob = c.s[BOLT_S_Omega_b] / b.a;
  const double hk = f.hkap;
  // Phi = rPhi + h y0 ; v = (rv - hk y1)/(1+h) ; delta = rdelta + hk v - 3h y0 ; delta_b = rdb - 3h y0 + hk y3
  const double dPhi_y[4] = {h, 0, 0, 0};
  const double dDel_y[4] = {-3.0 * h, -hk * hk * f.vden, 0, 0};
  const double dDb_y[4] = {-3.0 * h, 0, 0, hk};
#pragma unroll
  for (int j = 0; j < 4; j++) {
    f.lu[0][j] = dPhi_y[j] + b.cPsi * sPsi[j] + (j == 1 ? 1.0 : 0.0);                                   // E1: Psi + Phi + cPsi*SPsi = 0
    f.lu[1][j] = (j == 0 ? 1.0 : 0.0) - (j == 1 ? 1.0 : 0.0) + b.k2 * dPhi_y[j]
                 - b.gPhi * (Oc * dDel_y[j] + Ob * dDb_y[j] + sPhi[j]);                                // E2: Phi' definition
    f.lu[2][j] = (j == 2 ? 1.0 : 0.0) - sPi[j];                                                         // E3: Pi definition
    f.lu[3][j] = (j == 3 ? (1.0 + h - h * b.taup * b.R) : 0.0) + hk * ((j == 1 ? 1.0 : 0.0) + b.csb2 * dDb_y[j])
                 + f.e4c * t1[j];                                                                       // E4: v_b equation
  }
  lu4_factor(f);
}

// Solve W U = r for the vector in r[] (shared, overwritten by U).  If zout != null also writes
// zout = (U - rhs)/gamma with rhs recomputed by `rhs_of(idx)` (fused to avoid a second pass).
template <class RhsOf>
__device__ __forceinline__ void solve(const DevCosmo& c, const Lane& ln, const Bg& b, const Factor& f, const double* ib,
                                      double* r, double* zout, RhsOf rhs_of) {
  double ibn = 0.0, rn = 0.0;
  for (int l = ln.maxlen - 1; l >= 3; l--) {
    if (l < ln.len) {
      const int idx = ln.base + l * ln.stride;
      double v = r[idx];
      if (l < ln.len - 1) v -= (f.hk * (1.0 - c_rl[l]) * ibn) * rn;
      ibn = ib[idx]; rn = v; r[idx] = v;
    }
  }
  double a0 = 0, a1 = 0, a2 = 0, ib0 = 0, ib1 = 0, ib2 = 0;
  if (ln.kind != CH_IDLE) {
    const int i0 = ln.base, i1 = ln.base + ln.stride, i2 = ln.base + 2 * ln.stride;
    ib2 = ib[i2]; ib1 = ib[i1]; ib0 = ib[i0];
    const double r2 = r[i2] - (f.hk * (1.0 - c_rl[2]) * ibn) * rn;
    const double r1 = r[i1] - (f.hk * (1.0 - c_rl[1]) * ib2) * r2;
    const double r0 = r[i0] - (f.hk * ib1) * r1;
    a0 = r0 * ib0; a1 = (r1 - f.lo1 * a0) * ib1; a2 = (r2 - f.lo2 * a1) * ib2;
  }
  const double sPsi = warp_sum(b.wPsi * a2);
  const double sPhi = warp_sum(b.wPhi * a0);
  double pi = 0.0;
  if (ln.kind == CH_T) pi = a2; else if (ln.kind == CH_P) pi = a2 + a0;
  const double sPi = warp_sum(pi);
  const double t1 = shfl_d(a1, ln.nq);
  const int iS = ln.iS;
  const double rPhi = r[iS], rdel = r[iS + 1], rv = r[iS + 2], rdb = r[iS + 3], rvb = r[iS + 4];
  const double Oc = c.s[BOLT_S_Omega_c] / b.a, Ob = c.s[BOLT_S_Omega_b] / b.a;
  const double hk = f.hkap, h = f.h;
  const double vc = rv * f.vden, dc = rdel + hk * vc;
  double rhs[4], y[4];
  rhs[0] = -(rPhi + b.cPsi * sPsi);
  rhs[1] = -(b.k2 * rPhi - b.gPhi * (Oc * dc + Ob * rdb + sPhi));
  rhs[2] = sPi;
  rhs[3] = -(hk * b.csb2 * rdb + f.e4c * t1 - rvb);
  lu4_solve(f, rhs, y);
  __syncwarp();
  // scalars
  {
    const double Phi = rPhi + h * y[0];
    const double v = vc - hk * f.vden * y[1];
    const double del = rdel + hk * v - 3.0 * h * y[0];
    const double db = rdb - 3.0 * h * y[0] + hk * y[3];
    const double vb = y[3];
    if (ln.lane == 0) {
      if (zout) {
        zout[iS] = (Phi - rhs_of(iS)) * (1.0 / KC_GAMMA); zout[iS + 1] = (del - rhs_of(iS + 1)) * (1.0 / KC_GAMMA);
        zout[iS + 2] = (v - rhs_of(iS + 2)) * (1.0 / KC_GAMMA); zout[iS + 3] = (db - rhs_of(iS + 3)) * (1.0 / KC_GAMMA);
        zout[iS + 4] = (vb - rhs_of(iS + 4)) * (1.0 / KC_GAMMA);
      }
      r[iS] = Phi; r[iS + 1] = del; r[iS + 2] = v; r[iS + 3] = db; r[iS + 4] = vb;
    }
  }
  if (ln.kind != CH_IDLE) {
    double U0 = a0, U1 = a1, U2 = a2;
#pragma unroll
    for (int j = 0; j < 4; j++) { U0 += f.beta0[j] * y[j]; U1 += f.beta1[j] * y[j]; U2 += f.beta2[j] * y[j]; }
    const int i0 = ln.base, i1 = ln.base + ln.stride, i2 = ln.base + 2 * ln.stride;
    if (zout) {
      zout[i0] = (U0 - rhs_of(i0)) * (1.0 / KC_GAMMA); zout[i1] = (U1 - rhs_of(i1)) * (1.0 / KC_GAMMA);
      zout[i2] = (U2 - rhs_of(i2)) * (1.0 / KC_GAMMA);
    }
    r[i0] = U0; r[i1] = U1; r[i2] = U2;
    double Up = U2;
    for (int l = 3; l < ln.len; l++) {
      const int idx = ln.base + l * ln.stride;
      const double lo = (l == ln.len - 1) ? -f.hk : -f.hk * c_rl[l];
      const double U = (r[idx] - lo * Up) * ib[idx];
      if (zout) zout[idx] = (U - rhs_of(idx)) * (1.0 / KC_GAMMA);
      r[idx] = U; Up = U;
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------------
// Full right-hand side du = A(x) u for the arrays in shared memory (used for f(u0) and the initial-step
// heuristic only; the stepper itself never evaluates the RHS).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void rhs_full(const DevCosmo& c, const Lane& ln, const Bg& b, const double* u, double* du) {
  Metric m;
  const int iS = ln.iS;
  m.Phi = u[iS]; m.delta = u[iS + 1]; m.v = u[iS + 2]; m.delta_b = u[iS + 3]; m.v_b = u[iS + 4];
  double c0 = 0, c2 = 0;
  if (ln.kind != CH_IDLE) { c0 = u[ln.base]; c2 = u[ln.base + 2 * ln.stride]; }
  metric_from_chains(ln, b, c0, c2, m, c);
  const double T1 = shfl_d(ln.kind == CH_T ? u[ln.base + ln.stride] : 0.0, ln.nq);
  auto get = [&](int l) { return u[ln.base + l * ln.stride]; };
  #pragma unroll 1
  for (int l = 0; l < ln.len; l++) du[ln.base + l * ln.stride] = rhs_row(ln, b, m, l, get);
  if (ln.lane == 0) {   // :197-200
    du[iS] = m.dPhi;
    du[iS + 1] = b.kappa * m.v - 3.0 * m.dPhi;
    du[iS + 2] = -m.v - b.kappa * m.Psi;
    du[iS + 3] = b.kappa * m.v_b - 3.0 * m.dPhi;
    du[iS + 4] = -m.v_b - b.kappa * (m.Psi + b.csb2 * m.delta_b) + b.taup * b.R * (3.0 * T1 + m.v_b);
  }
  __syncwarp();
}

// initial_conditions (perturbations.jl:274-338) written into u[] (shared)
__device__ __forceinline__ void initial_conditions(const DevCosmo& c, const Lane& ln, const Bg& b, double* u) {
  const double k = ln.k, Hx = b.H, eta = b.eta, taup = b.taup;
  const double N_nu = c.s[BOLT_S_N_nu];
  const double f_nu = 1.0 / (1.0 + 1.0 / (7.0 * (3.0 / 3.0) * N_nu / 8.0 * pow(4.0 / 11.0, 4.0 / 3.0)));
  const double Phi = (4.0 * f_nu + 10.0) / (4.0 * f_nu + 15.0) * 1.0;
  const double C = -((15.0 + 4.0 * f_nu) / (20.0 + 8.0 * f_nu)) * Phi;
  const double T0 = -40.0 * C / (15.0 + 4.0 * f_nu) / 4.0;
  const double T1 = 10.0 * C / (15.0 + 4.0 * f_nu) * (k * k * eta) / (3.0 * k);
  const double T2 = -8.0 * k / (15.0 * Hx * taup) * T1;
  const double N2 = -(k * k * eta * eta) / 15.0 * 1.0 / (1.0 + 2.0 / 5.0 * f_nu) * Phi / 2.0;
  const int st = ln.stride;
  if (ln.kind == CH_T) {
    u[ln.base] = T0; u[ln.base + st] = T1; u[ln.base + 2 * st] = T2;
    double prev = T2;
    #pragma unroll 1
    for (int l = 3; l < ln.len; l++) { prev = -(double)l / (2 * l + 1) * k / (Hx * taup) * prev; u[ln.base + l * st] = prev; }
  } else if (ln.kind == CH_P) {
    u[ln.base] = (5.0 / 4.0) * T2; u[ln.base + st] = -k / (4.0 * Hx * taup) * T2;
    double prev = (1.0 / 4.0) * T2; u[ln.base + 2 * st] = prev;
    #pragma unroll 1
    for (int l = 3; l < ln.len; l++) { prev = -(double)l / (2 * l + 1) * k / (Hx * taup) * prev; u[ln.base + l * st] = prev; }
  } else if (ln.kind == CH_N) {
    u[ln.base] = T0; u[ln.base + st] = T1; u[ln.base + 2 * st] = N2;
    double prev = N2;
    #pragma unroll 1
    for (int l = 3; l < ln.len; l++) { prev = k / ((2 * l + 1) * Hx) * prev; u[ln.base + l * st] = prev; }
  } else if (ln.kind == CH_M) {
    const double df0 = ln.df0;
    u[ln.base] = -T0 * df0;
    u[ln.base + st] = -b.eq * T1 * df0;
    double prev = -N2 * df0; u[ln.base + 2 * st] = prev;
    #pragma unroll 1
    for (int l = 3; l < ln.len; l++) { prev = b.qe * k / ((2 * l + 1) * Hx) * prev; u[ln.base + l * st] = prev; }
  }
  if (ln.lane == 0) {
    const double delta = 3.0 / 4.0 * (4.0 * T0), v = -3.0 * k * T1;
    u[ln.iS] = Phi; u[ln.iS + 1] = delta; u[ln.iS + 2] = v; u[ln.iS + 3] = delta; u[ln.iS + 4] = v;
  }
  __syncwarp();
}

// Per-mode constants of the fast background evaluation
struct ModeConst {
  double k, k2_3, c12, H02h, R0, Oc, Ob, mnu;      // k, k^2/3, 12 H0^2/k^2, H0^2/2, 4 Om_r/(3 Om_b), Om_c, Om_b, sum m_nu
  double q2, iq, wPhi0, wPsi0;                       // lane: q^2, 1/q, weight prefactors (see eval_bg_fast)
};
__device__ __forceinline__ void mode_const(const DevCosmo& c, const Lane& ln, ModeConst& mc) {
  const double H0 = c.s[BOLT_S_H0], rho_crit = c.s[BOLT_S_rho_crit], Om_r = c.s[BOLT_S_Omega_r];
  mc.k = ln.k; mc.k2_3 = ln.k * ln.k / 3.0; mc.c12 = 12.0 * H0 * H0 / (ln.k * ln.k); mc.H02h = 0.5 * H0 * H0;
  mc.R0 = 4.0 * Om_r / (3.0 * c.s[BOLT_S_Omega_b]); mc.Oc = c.s[BOLT_S_Omega_c]; mc.Ob = c.s[BOLT_S_Omega_b];
  mc.mnu = c.s[BOLT_S_Sum_m_nu];
  mc.q2 = ln.q * ln.q; mc.iq = (ln.kind == CH_M) ? 1.0 / ln.q : 1.0;
  mc.wPhi0 = 0.0; mc.wPsi0 = 0.0;
  if (ln.kind == CH_M) { mc.wPhi0 = ln.wq / rho_crit; mc.wPsi0 = ln.wq * mc.q2 / rho_crit * 0.25; }
  else if (ln.kind == CH_T) { mc.wPhi0 = 4.0 * Om_r; mc.wPsi0 = Om_r; }
  else if (ln.kind == CH_N) { mc.wPhi0 = 4.0 * c.Omega_nu; mc.wPsi0 = c.Omega_nu; }
}

// ---------------------------------------------------------------------------------------------------
// One sample of the source grids at x_grid[ix] from the Hermite dense output of the current step
// (spectra.jl:13-18: u = perturb(x); hierarchy!(du,u,h,x); source_function(du,u,h,x)).
// Only chain rows l <= 3 enter the sources, so only u_l, l <= 4 is interpolated (all of u if u_hist).
// ---------------------------------------------------------------------------------------------------
struct Hermite { double th, c0, c1, d0, d1; };   // u(th) = c0*u0 + c1*u1 + d0*(dt f0) + d1*(dt f1)
__device__ __forceinline__ Hermite hermite_weights(double th) {
  // (1-th) y0 + th y1 + th(th-1)[(1-2th)(y1-y0) + (th-1) dt f0 + th dt f1]   [OrdinaryDiffEq hermite_interpolant]
  Hermite hm; hm.th = th;
  const double w = th * (th - 1.0);
  hm.c0 = (1.0 - th) - w * (1.0 - 2.0 * th);
  hm.c1 = th + w * (1.0 - 2.0 * th);
  hm.d0 = w * (th - 1.0);
  hm.d1 = w * th;
  return hm;
}

__device__ __forceinline__ void sample_sources(const DevCosmo& c, const Lane& ln, const SolveParams& p, int ik, int ix, double xs,
                                               const Hermite& hm, const double* u0, const double* u1, const double* z1, double s1,
                                               const double* z6, bool& rsa_flag, const ModeConst* mc = nullptr) {
  auto herm = [&](int idx) { return hm.c0 * u0[idx] + hm.c1 * u1[idx] + hm.d0 * (s1 * z1[idx]) + hm.d1 * z6[idx]; };
  if (p.u_hist) {
    double* out = p.u_hist + ((size_t)ik * c.n_x + ix) * ln.n;
    #pragma unroll 1
    for (int l = 0; l < ln.len; l++) out[ln.rbase + l * ln.rstride] = herm(ln.base + l * ln.stride);
    if (ln.lane < 5) out[ln.riS + ln.lane] = herm(ln.iS + ln.lane);
  }
  if (!p.S_T && !p.S_P) return;
  Bg b; double Hpp, tau, g, gp, gpp;
  if (mc) {
    // compact variant (register kernel): all 11 tables of hierarchy! + source_function (:168,172,347-349) from ONE spline
    // site, lanes 0..10 evaluating one table each; divisions hoisted as in eval_bg_fast
    const int which[11] = {BOLT_T_H, BOLT_T_eta, BOLT_T_taup, BOLT_T_taupp, BOLT_T_csb2, BOLT_T_Hp, BOLT_T_Hpp, BOLT_T_tau, BOLT_T_g, BOLT_T_gp, BOLT_T_gpp};
    double tv = 0.0;
    if (ln.lane < 11) tv = spline_eval(c.tab[which[ln.lane]], c.n_x, c.x0, c.dx, xs);
    b.x = xs; b.H = shfl_d(tv, 0); b.eta = shfl_d(tv, 1); b.taup = shfl_d(tv, 2); b.taupp = shfl_d(tv, 3); b.csb2 = shfl_d(tv, 4);
    b.Hp = shfl_d(tv, 5); Hpp = shfl_d(tv, 6); tau = shfl_d(tv, 7); g = shfl_d(tv, 8); gp = shfl_d(tv, 9); gpp = shfl_d(tv, 10);
    b.a = exp(xs);
    const double ia = 1.0 / b.a, ia2 = ia * ia, iH = 1.0 / b.H, iH2 = iH * iH;
    b.kappa = mc->k * iH; b.R = mc->R0 * ia; b.cPsi = mc->c12 * ia2; b.gPhi = mc->H02h * iH2; b.k2 = mc->k2_3 * iH2;
    b.qe = 1.0; b.eq = 1.0;
    if (ln.kind == CH_M) {
      const double am = b.a * mc->mnu;
      const double eps = sqrt(mc->q2 + am * am), ieps = 1.0 / eps;
      b.qe = ln.q * ieps; b.eq = eps * mc->iq; b.wPhi = mc->wPhi0 * eps * ia2; b.wPsi = mc->wPsi0 * ieps;
    } else { b.wPhi = mc->wPhi0 * ia2; b.wPsi = mc->wPsi0; }
  } else {
    eval_bg(c, ln, xs, b);
    // remaining tables of source_function (:347-349): lanes evaluate one each
    const int which[5] = {BOLT_T_Hpp, BOLT_T_tau, BOLT_T_g, BOLT_T_gp, BOLT_T_gpp};
    double tv = 0.0;
    if (ln.lane < 5) tv = spline_eval(c.tab[which[ln.lane]], c.n_x, c.x0, c.dx, xs);
    Hpp = shfl_d(tv, 0); tau = shfl_d(tv, 1); g = shfl_d(tv, 2); gp = shfl_d(tv, 3); gpp = shfl_d(tv, 4);
  }
  double uL[5] = {0, 0, 0, 0, 0};
  if (ln.kind != CH_IDLE) {
#pragma unroll
    for (int l = 0; l < 5; l++) if (l < ln.len) uL[l] = herm(ln.base + l * ln.stride);
  }
  Metric m;
  m.Phi = herm(ln.iS); m.delta = herm(ln.iS + 1); m.v = herm(ln.iS + 2); m.delta_b = herm(ln.iS + 3); m.v_b = herm(ln.iS + 4);
  metric_from_chains(ln, b, uL[0], uL[2], m, c);
  auto get = [&](int l) { double v = uL[0]; v = (l == 1) ? uL[1] : v; v = (l == 2) ? uL[2] : v; v = (l == 3) ? uL[3] : v; v = (l == 4) ? uL[4] : v; return v; };
  double d[4] = {0, 0, 0, 0};
  if (ln.kind != CH_IDLE) {
#pragma unroll
    for (int l = 0; l < 4; l++) d[l] = rhs_row(ln, b, m, l, get);
  }
  const double T1 = shfl_d(uL[1], ln.nq);
  const double dvb = -m.v_b - b.kappa * (m.Psi + b.csb2 * m.delta_b) + b.taup * b.R * (3.0 * T1 + m.v_b);   // :200
  // RSA switch (:216-232): overwrite Theta_0..2, N_0..2 in the sampled state and zero the radiation derivatives
  const bool rsa_on = (ln.k * b.eta > 240.0) && (-b.taup * b.H / b.eta > 100.0);
  if (rsa_on) {
    rsa_flag = true;
    if (ln.kind == CH_T) {
      uL[0] = m.Phi - b.H / ln.k * b.taup * m.v_b;
      uL[1] = b.H / ln.k * (-2.0 * m.dPhi + b.taup * (m.Phi - b.csb2 * m.delta_b) + b.H / ln.k * (b.taupp - b.taup) * m.v_b);
      uL[2] = 0.0;
    } else if (ln.kind == CH_N) {
      uL[0] = m.Phi; uL[1] = -2.0 * b.H / ln.k * m.dPhi; uL[2] = 0.0;
    }
    if (ln.kind == CH_T || ln.kind == CH_P || ln.kind == CH_N) { d[0] = d[1] = d[2] = d[3] = 0.0; }
  }
  // sigma_M and its derivative (:361-362)
  const double wS = (ln.kind == CH_M) ? b.wPsi * 4.0 * c.s[BOLT_S_rho_crit] : 0.0;   // wq q^2/eps
  const double sigM = warp_sum(wS * uL[2]);
  const double sigMp = warp_sum(wS * d[2]);
  const int lT = ln.nq, lP = ln.nq + 1, lN = ln.nq + 2;
  const double Th0 = shfl_d(uL[0], lT), Th1 = shfl_d(uL[1], lT), Th2 = shfl_d(uL[2], lT), Th3 = shfl_d(uL[3], lT);
  const double dTh1 = shfl_d(d[1], lT), dTh2 = shfl_d(d[2], lT), dTh3 = shfl_d(d[3], lT);
  const double P0 = shfl_d(uL[0], lP), P1 = shfl_d(uL[1], lP), P2 = shfl_d(uL[2], lP), P3 = shfl_d(uL[3], lP);
  const double dP0 = shfl_d(d[0], lP), dP1 = shfl_d(d[1], lP), dP2 = shfl_d(d[2], lP), dP3 = shfl_d(d[3], lP);
  const double N2 = shfl_d(uL[2], lN), dN2 = shfl_d(d[2], lN);
  const double Om_r = c.s[BOLT_S_Omega_r], rho_crit = c.s[BOLT_S_rho_crit];
  const double k = ln.k, Hx = b.H, Hp = b.Hp;
  const double Psi = -m.Phi - b.cPsi * (Om_r * Th2 + c.Omega_nu * N2 + sigM / rho_crit / 4.0);                 // :363-365
  const double dPsi = -m.dPhi - b.cPsi * (Om_r * (dTh2 - 2.0 * Th2) + c.Omega_nu * (dN2 - 2.0 * N2)
                                          + (sigMp - 2.0 * sigM) / rho_crit / 4.0);                           // :368-370
  const double Pi = Th2 + P2 + P0, dPi = dTh2 + dP2 + dP0;                                                     // :372-373
  const double term1 = g * (Th0 + Psi + Pi / 4.0) + exp(-tau) * (dPsi - m.dPhi);                                // :375
  const double term2 = (-1.0 / k) * (Hp * g * m.v_b + Hx * gp * m.v_b + Hx * g * dvb);                          // :376
  const double ddPi = 2.0 * k / (5.0 * Hx) * (-Hp / Hx * Th1 + dTh1) + (3.0 / 10.0) * (b.taupp * Pi + b.taup * dPi)
                      - 3.0 * k / (5.0 * Hx) * (-Hp / Hx * (Th3 + P1 + P3) + (dTh3 + dP1 + dP3));               // :377-378
  const double term3 = (3.0 / (4.0 * k * k)) * ((Hp * Hp + Hx * Hpp) * g * Pi + 3.0 * Hx * Hp * (gp * Pi + g * dPi)
                                                + Hx * Hx * (gpp * Pi + 2.0 * gp * dPi + g * ddPi));            // :379-381
  const double y = k * (c.eta_end - b.eta);                                                                    // :401
  if (ln.lane == 0) {
    if (p.S_T) p.S_T[((size_t)ik * c.n_x + ix) * p.out_nd] = term1 + term2 + term3;
    if (p.S_P) p.S_P[((size_t)ik * c.n_x + ix) * p.out_nd] = (3.0 / (4.0 * y * y)) * g * Pi;                                 // :403
  }
}

// ---------------------------------------------------------------------------------------------------
// The kernel: persistent warps pull k-modes from a queue.
// Shared memory per warp: 9 arrays of n doubles: u, z1..z6, work r, inverse pivots ib.
// ---------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------
// Register-resident stage solver, specialised at compile time on the truncations TR = Trunc<l_gamma, l_nu, l_mnu, nq>.
// Same algebra as factor()/solve() above, but: the lane's chain lives in a register array through the whole
// stage (sweeps fully unrolled, no index arithmetic; "is this the truncation row of my chain" folds to a
// compile-time constant except at the two or three rows where some chain ends); the shared-memory arrays are
// interleaved [l][chain] so every access is conflict-free at a constant offset; only the (Phi', Psi)
// components are reduced over the warp (the Pi and v_b components live on the photon lanes and are broadcast);
// the 4x4 border system is kept as a matrix and solved per right-hand side by Gaussian elimination with
// partial pivoting on the augmented 4x5 system (cheaper than a stored factorisation with permutation logic).
// ---------------------------------------------------------------------------------------------------
template <int LG_, int LNU_, int LMNU_, int NQ_, int NCH_ = 32>
struct Trunc {
  static constexpr int LG = LG_, LNU = LNU_, LMNU = LMNU_, NQ = NQ_;
  static constexpr int MAXL = (LG_ > LNU_ ? (LG_ > LMNU_ ? LG_ : LMNU_) : (LNU_ > LMNU_ ? LNU_ : LMNU_));
  static constexpr int MAXLEN = (NQ_ > 0) ? MAXL + 1 : 0;     // 0 = generic (runtime) kernel
  // row stride of the interleaved shared-memory layout: one private column per LANE (not per chain), so idle lanes and the
  // padded rows of short chains can run the unguarded, branch-free code on zeros
  static constexpr int NCH = NCH_;     // NQ+4: compact + one shared all-zero column for the idle lanes (both register kernels); 32: private column per lane
  static constexpr bool RT = false;
  // row l is the truncation row / an existing row of the lane's chain
  static __device__ __forceinline__ bool top(int kind, int l) {
    return (l == LG_ && (kind == CH_T || kind == CH_P)) || (l == LNU_ && kind == CH_N) || (l == LMNU_ && kind == CH_M);
  }
  static __device__ __forceinline__ bool act(int kind, int l) {
    return (l <= LG_ && (kind == CH_T || kind == CH_P)) || (l <= LNU_ && kind == CH_N) || (l <= LMNU_ && kind == CH_M);
  }
};

__host__ __device__ constexpr double RLc(int l) { return (double)l / (double)(2 * l + 1); }

// 1/x for normal x: MUFU.RCP64H seed (~20 bits) + 2 Newton steps (full double accuracy; not correctly rounded)
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0); y = fma(y, e, y);
  e = fma(-x, y, 1.0); y = fma(y, e, y);
  return y;
}

// Stage-time background: same quantities as eval_bg with the divisions hoisted (4 reciprocals, 1 exp, 1 sqrt).
struct BgS {
  double H, eta, taup, csb2, a;
  double kappa, qe, eq, wPsi, wPhi, cPsi, gPhi, k2, R, Oc_a, Ob_a, iHeta;
};
// ... from the four table values H, eta, tau', c_sb^2 at x (already in b)
__device__ __forceinline__ void bg_from_tables(const Lane& ln, const ModeConst& mc, double x, BgS& b);
__device__ __forceinline__ void eval_bg_fast(const DevCosmo& c, const Lane& ln, const ModeConst& mc, double x, BgS& b) {
  // every lane evaluates one of the four tables (lane & 3): no divergent branch, the extra loads are broadcasts
  static_assert(BOLT_T_H == 0 && BOLT_T_eta == 3 && BOLT_T_taup == 6 && BOLT_T_csb2 == 11, "table order");
  const int wl = ln.lane & 3;
  const int tab = 3 * wl + ((wl == 3) ? 2 : 0);       // = which[wl] without a select chain
  const double v = spline_eval(c.tab[tab], c.n_x, c.x0, c.dx, x);
  b.H = shfl_d(v, 0); b.eta = shfl_d(v, 1); b.taup = shfl_d(v, 2); b.csb2 = shfl_d(v, 3);
  bg_from_tables(ln, mc, x, b);
}
__device__ __forceinline__ void bg_from_tables(const Lane& ln, const ModeConst& mc, double x, BgS& b) {
  b.a = exp(x);
  const double ia = fast_rcp(b.a), ia2 = ia * ia, iH = fast_rcp(b.H), iH2 = iH * iH;
  b.kappa = mc.k * iH; b.R = mc.R0 * ia; b.cPsi = mc.c12 * ia2; b.gPhi = mc.H02h * iH2; b.k2 = mc.k2_3 * iH2;
  b.Oc_a = mc.Oc * ia; b.Ob_a = mc.Ob * ia; b.iHeta = iH * fast_rcp(b.eta);
  b.qe = 1.0; b.eq = 1.0;
  if (ln.kind == CH_M) {
    const double am = b.a * mc.mnu;
    const double eps = sqrt(mc.q2 + am * am), ieps = fast_rcp(eps);
    b.qe = ln.q * ieps; b.eq = eps * mc.iq;
    b.wPhi = mc.wPhi0 * eps * ia2; b.wPsi = mc.wPsi0 * ieps;
  } else { b.wPhi = mc.wPhi0 * ia2; b.wPsi = mc.wPsi0; }
}

// Factorisation state of one stage.  The inverse pivots and the beta vectors (23 doubles per lane) live in a lane-private
// column of a shared-memory scratch (fs: row r of the lane at fs[r*FS_STRIDE]); only the border matrix and a few scalars stay
// in registers.
template <class TR>
struct RegFactor {
  double* fs;
  double M[4][4];
  double h, hk, hkap, vden, e4c, lo1, lo2;
  static constexpr int FS_STRIDE = 32, FS_ROWS = (TR::MAXLEN > 0 ? TR::MAXLEN : 1) + 12;
  __device__ __forceinline__ double& ibv(int l) const { return fs[l * FS_STRIDE]; }
  __device__ __forceinline__ double& beta(int row, int j) const { return fs[(TR::MAXLEN + 4 * row + j) * FS_STRIDE]; }
};

// Solve M y = rhs by Gaussian elimination with partial pivoting on the augmented system (registers, select-based swaps).
__device__ __forceinline__ void ge4(const double (&Min)[4][4], const double (&rhs)[4], double (&y)[4]) {
  double a[4][5];
#pragma unroll
  for (int i = 0; i < 4; i++) {
#pragma unroll
    for (int j = 0; j < 4; j++) a[i][j] = Min[i][j];
    a[i][4] = rhs[i];
  }
#pragma unroll
  for (int kx = 0; kx < 3; kx++) {
#pragma unroll
    for (int i = kx + 1; i < 4; i++) {
      const bool sw = fabs(a[i][kx]) > fabs(a[kx][kx]);
#pragma unroll
      for (int j = kx; j < 5; j++) { const double t = a[kx][j]; a[kx][j] = sw ? a[i][j] : t; a[i][j] = sw ? t : a[i][j]; }
    }
    const double ip = fast_rcp(a[kx][kx]);
#pragma unroll
    for (int i = kx + 1; i < 4; i++) {
      const double m = a[i][kx] * ip;
#pragma unroll
      for (int j = kx + 1; j < 5; j++) a[i][j] -= m * a[kx][j];
    }
  }
  y[3] = a[3][4] * fast_rcp(a[3][3]);
  y[2] = (a[2][4] - a[2][3] * y[3]) * fast_rcp(a[2][2]);
  y[1] = (a[1][4] - a[1][2] * y[2] - a[1][3] * y[3]) * fast_rcp(a[1][1]);
  y[0] = (a[0][4] - a[0][1] * y[1] - a[0][2] * y[2] - a[0][3] * y[3]) * fast_rcp(a[0][0]);
}

// l/(2l+1) and 1 - l/(2l+1) for a compile-time l: as constant-bank operands (CB: no UMOV pair per use, -4 % in the value kernel)
// or as immediates (the kernel with partials: the constant-bank form costs it registers and spills)
template <bool CB> __device__ __forceinline__ double rl_of(int l) { return CB ? c_rl[l] : RLc(l); }
template <bool CB> __device__ __forceinline__ double rl1_of(int l) { return CB ? c_rl1[l] : 1.0 - RLc(l); }

template <class TR, bool CB = true, class FT>      // FT: RegFactor<TR> (lane-private scratch) or SlotFactor<TR> (stage_slot.cuh)
__device__ __forceinline__ void factor_reg(const Lane& ln, const BgS& b, double h, FT& f) {
  constexpr int MAXLEN = TR::MAXLEN;
  const int kind = ln.kind;
  const bool photon = (kind == CH_T || kind == CH_P);
  f.h = h; f.hk = h * b.kappa * b.qe; f.hkap = h * b.kappa; f.vden = fast_rcp(1.0 + h);
  f.e4c = -3.0 * h * b.taup * b.R;
  const double dtau = photon ? -h * b.taup : 0.0;
  const double btr = 1.0 + h * (double)ln.len * b.iHeta + dtau;
  double ibn = 0.0, lo_next = 0.0;
#pragma unroll
  for (int l = MAXLEN - 1; l >= 3; l--) {
    const bool act = TR::act(kind, l), top = TR::top(kind, l);
    const double bd = top ? btr : 1.0 + dtau;
    const double up = top ? 0.0 : f.hk * rl1_of<CB>(l);
    const double lo = top ? -f.hk : -f.hk * rl_of<CB>(l);
    const double rc = fast_rcp(bd - (up * ibn) * lo_next);      // unconditional (>= 1 on every lane): no branch around it
    const double ibl = act ? rc : 0.0;
    f.ibv(l) = ibl; ibn = ibl; lo_next = act ? lo : 0.0;
  }
  const bool live = kind != CH_IDLE;
  const double up2 = f.hk * rl1_of<CB>(2), up1 = f.hk * rl1_of<CB>(1), up0 = f.hk;
  const double lo2 = -f.hk * rl_of<CB>(2), lo1 = -f.hk * rl_of<CB>(1);
  const double rc2 = fast_rcp((1.0 + dtau) - (up2 * ibn) * lo_next);
  const double ib2 = live ? rc2 : 0.0;
  const double m1 = up1 * ib2;
  const double rc1 = fast_rcp((1.0 + dtau) - m1 * lo2);
  const double ib1 = live ? rc1 : 0.0;
  const double m0 = up0 * ib1;
  const double rc0 = fast_rcp((1.0 + (kind == CH_P ? dtau : 0.0)) - m0 * lo1);
  const double ib0 = live ? rc0 : 0.0;
  f.ibv(2) = ib2; f.ibv(1) = ib1; f.ibv(0) = ib0; f.lo1 = lo1; f.lo2 = lo2;
  // coupling of rows 0..2 to y = (Phi', Psi, Pi, v_b), by selects (no divergent branches)
  double C0[4] = {0, 0, 0, 0}, C1[4] = {0, 0, 0, 0}, C2[4] = {0, 0, 0, 0};
  {
    const bool isM = kind == CH_M, isT = kind == CH_T, isP = kind == CH_P, isTN = isT || kind == CH_N;
    const double htp = h * b.taup;
    C0[0] = isM ? h * ln.df0 : (isTN ? -h : 0.0);
    C1[1] = isM ? -f.hkap * (1.0 / 3.0) * b.eq * ln.df0 : (isTN ? f.hkap * (1.0 / 3.0) : 0.0);
    C1[3] = isT ? htp * (1.0 / 3.0) : 0.0;
    C2[2] = (isT || isP) ? -htp * 0.1 : 0.0;
    C0[2] = isP ? -htp * 0.5 : 0.0;
  }
  double be0[4], be1[4], be2[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const double V2 = C2[j], V1 = C1[j] - m1 * V2, V0 = C0[j] - m0 * V1;
    be0[j] = V0 * ib0;
    be1[j] = (V1 - lo1 * be0[j]) * ib1;
    be2[j] = (V2 - lo2 * be1[j]) * ib2;
    f.beta(0, j) = be0[j]; f.beta(1, j) = be1[j]; f.beta(2, j) = be2[j];
  }
  // y-dependence of the Psi / Phi' / Pi sums.  Components (Pi, v_b) are non-zero on the Theta lane only (and Pi on
  // the ThetaP lane), so only the (Phi', Psi) components need a warp reduction.
  const int lT = ln.nq, lP = ln.nq + 1;
  double sPsi[4], sPhi[4], sPi[4], t1[4];
#pragma unroll
  for (int j = 0; j < 2; j++) { sPsi[j] = warp_sum(b.wPsi * be2[j]); sPhi[j] = warp_sum(b.wPhi * be0[j]); }
  const double wPsiT = shfl_d(b.wPsi, lT), wPhiT = shfl_d(b.wPhi, lT);
#pragma unroll
  for (int j = 2; j < 4; j++) { sPsi[j] = wPsiT * shfl_d(be2[j], lT); sPhi[j] = wPhiT * shfl_d(be0[j], lT); }
#pragma unroll
  for (int j = 0; j < 4; j++) { sPi[j] = shfl_d(be2[j], lT); t1[j] = shfl_d(be1[j], lT); }
  sPi[2] += shfl_d(be2[2] + be0[2], lP);
  const double Oc = b.Oc_a, Ob = b.Ob_a;
  const double hk = f.hkap;
  const double dPhi_y[4] = {h, 0, 0, 0};
  const double dDel_y[4] = {-3.0 * h, -hk * hk * f.vden, 0, 0};
  const double dDb_y[4] = {-3.0 * h, 0, 0, hk};
#pragma unroll
  for (int j = 0; j < 4; j++) {
    f.M[0][j] = dPhi_y[j] + b.cPsi * sPsi[j] + (j == 1 ? 1.0 : 0.0);
    f.M[1][j] = (j == 0 ? 1.0 : 0.0) - (j == 1 ? 1.0 : 0.0) + b.k2 * dPhi_y[j] - b.gPhi * (Oc * dDel_y[j] + Ob * dDb_y[j] + sPhi[j]);
    f.M[2][j] = (j == 2 ? 1.0 : 0.0) - sPi[j];
    f.M[3][j] = (j == 3 ? (1.0 + h - h * b.taup * b.R) : 0.0) + hk * ((j == 1 ? 1.0 : 0.0) + b.csb2 * dDb_y[j]) + f.e4c * t1[j];
  }
}

// Solve W U = r for the lane's chain in rr[] (registers, overwritten by U) and the 5 scalars in r5[] (every lane
// holds the same copy).
template <class TR, bool CB = true>
__device__ __forceinline__ void solve_reg(const Lane& ln, const BgS& b, const RegFactor<TR>& f,
                                          double (&rr)[TR::MAXLEN > 0 ? TR::MAXLEN : 1], double (&r5)[5]) {
  constexpr int MAXLEN = TR::MAXLEN;
  const int kind = ln.kind;
  double ibn = 0.0, rn = 0.0;
#pragma unroll
  for (int l = MAXLEN - 1; l >= 3; l--) {
    const double up = TR::top(kind, l) ? 0.0 : f.hk * rl1_of<CB>(l);
    const double v = rr[l] - (up * ibn) * rn;
    rr[l] = v; rn = v; ibn = f.ibv(l);
  }
  const double r2 = rr[2] - (f.hk * rl1_of<CB>(2) * ibn) * rn;
  const double ib0 = f.ibv(0), ib1 = f.ibv(1), ib2 = f.ibv(2);
  const double r1 = rr[1] - (f.hk * rl1_of<CB>(1) * ib2) * r2;
  const double r0 = rr[0] - (f.hk * ib1) * r1;
  const double a0 = r0 * ib0, a1 = (r1 - f.lo1 * a0) * ib1, a2 = (r2 - f.lo2 * a1) * ib2;
  const int lT = ln.nq, lP = ln.nq + 1;
  const double sPsi = warp_sum(b.wPsi * a2);
  const double sPhi = warp_sum(b.wPhi * a0);
  const double sPi = shfl_d(a2, lT) + shfl_d(a2 + a0, lP);
  const double t1 = shfl_d(a1, lT);
  const double rPhi = r5[0], rdel = r5[1], rv = r5[2], rdb = r5[3], rvb = r5[4];
  const double hk = f.hkap, h = f.h;
  const double vc = rv * f.vden, dc = rdel + hk * vc;
  double rhs[4], y[4];
  rhs[0] = -(rPhi + b.cPsi * sPsi);
  rhs[1] = -(b.k2 * rPhi - b.gPhi * (b.Oc_a * dc + b.Ob_a * rdb + sPhi));
  rhs[2] = sPi;
  rhs[3] = -(hk * b.csb2 * rdb + f.e4c * t1 - rvb);
  ge4(f.M, rhs, y);
  r5[0] = rPhi + h * y[0];
  const double v = vc - hk * f.vden * y[1];
  r5[1] = rdel + hk * v - 3.0 * h * y[0];
  r5[2] = v;
  r5[3] = rdb - 3.0 * h * y[0] + hk * y[3];
  r5[4] = y[3];
  double U0 = a0, U1 = a1, U2 = a2;
#pragma unroll
  for (int j = 0; j < 4; j++) { U0 += f.beta(0, j) * y[j]; U1 += f.beta(1, j) * y[j]; U2 += f.beta(2, j) * y[j]; }
  rr[0] = U0; rr[1] = U1; rr[2] = U2;
  double Up = U2;
#pragma unroll
  for (int l = 3; l < MAXLEN; l++) {
    const double lo = TR::top(kind, l) ? -f.hk : -f.hk * rl_of<CB>(l);
    const double U = (rr[l] - lo * Up) * f.ibv(l);
    rr[l] = U; Up = U;
  }
}

// ---------------------------------------------------------------------------------------------------
// Runtime-truncation stage solver ("long-chain path"): any l_gamma / l_nu / l_mnu >= 2 (plin: 50/50/20, the C4 sweep: 50/8/10).
// Same algebra and the same hoisted background / reciprocal / per-solve elimination as the register-resident solver, but the
// chain rows stay in the warp's shared memory in the reference's own order (n doubles per array, nothing padded) and the
// sweeps are runtime loops over l.  Factorisation and the downward elimination of the right-hand side share one pass (both
// need up_l * ib_{l+1}).
// ---------------------------------------------------------------------------------------------------
struct TruncRT {
  static constexpr int LG = 0, LNU = 0, LMNU = 0, NQ = 0, MAXL = 0, MAXLEN = 0, NCH = 32;
  static constexpr bool RT = true;
};

struct RtFactor {
  double* bs;                 // lane-private column of the beta scratch: beta_row[j] at bs[(4*row + j) * 32]
  double M[4][4];
  double h, hk, hkap, vden, e4c, lo1, lo2, dtau, btr;
  __device__ __forceinline__ double& beta(int row, int j) const { return bs[(4 * row + j) * 32]; }
};

// Rows len-1 .. 3 of the lane's chain: (FACT) inverse pivots into ib[], and the downward elimination of r[].
// Returns the state the bottom rows continue from.
template <bool FACT>
__device__ __forceinline__ void rt_down(const Lane& ln, const RtFactor& f, double* ib, double* r, double& ibn, double& lo_next, double& rn) {
  ibn = 0.0; lo_next = 0.0; rn = 0.0;
  const int len = ln.len;
  if (len > 3) {                 // truncation row (l = len-1 >= 3) peeled: no coupling upward, r unchanged
    const int idx = ln.base + (len - 1) * ln.stride;
    if (FACT) { ibn = fast_rcp(f.btr); ib[idx] = ibn; } else ibn = ib[idx];
    lo_next = -f.hk; rn = r[idx];
  }
  const double bd = 1.0 + f.dtau;
  // Uniform loop over l (c_rl[l] is a constant-bank operand), branch-free: a lane whose chain does not reach row l works on
  // its own element 0 (always a valid address), keeps its carried state by selects and has its stores predicated off.
#pragma unroll 4
  for (int l = ln.maxlen - 2; l >= 3; l--) {
    const bool act = l < len - 1;
    const int idx = ln.base + (act ? l : 0) * ln.stride;
    const double rl = c_rl[l];
    const double m = (f.hk * (1.0 - rl)) * ibn;
    double ibl;
    if (FACT) { ibl = fast_rcp(bd - m * lo_next); if (act) ib[idx] = ibl; lo_next = act ? -f.hk * rl : lo_next; }
    else ibl = ib[idx];
    const double v = r[idx] - m * rn;
    if (act) r[idx] = v;
    rn = act ? v : rn; ibn = act ? ibl : ibn;
  }
}

// Factor W = I - h A(x_s) and eliminate r[] downward in the same pass.
__device__ __forceinline__ void rt_factor_down(const Lane& ln, const BgS& b, double h, RtFactor& f, double* ib, double* r,
                                               double& r0, double& r1, double& r2) {
  const int kind = ln.kind;
  const bool photon = (kind == CH_T || kind == CH_P);
  const bool live = kind != CH_IDLE;
  f.h = h; f.hk = h * b.kappa * b.qe; f.hkap = h * b.kappa; f.vden = fast_rcp(1.0 + h);
  f.e4c = -3.0 * h * b.taup * b.R;
  f.dtau = photon ? -h * b.taup : 0.0;
  f.btr = 1.0 + h * (double)ln.len * b.iHeta + f.dtau;
  double ibn, lo_next, rn;
  rt_down<true>(ln, f, ib, r, ibn, lo_next, rn);
  const bool top2 = (ln.len == 3);      // l = 2 is the truncation row of a chain with l_max = 2
  const double dtau = f.dtau;
  const double up2 = top2 ? 0.0 : f.hk * (1.0 - RLc(2)), up1 = f.hk * (1.0 - RLc(1)), up0 = f.hk;
  const double lo2 = top2 ? -f.hk : -f.hk * RLc(2), lo1 = -f.hk * RLc(1);
  const double mm2 = up2 * ibn;
  const double rc2 = fast_rcp((top2 ? f.btr : 1.0 + dtau) - mm2 * lo_next);
  const double ib2 = live ? rc2 : 0.0;
  const double m1 = up1 * ib2;
  const double rc1 = fast_rcp((1.0 + dtau) - m1 * lo2);
  const double ib1 = live ? rc1 : 0.0;
  const double m0 = up0 * ib1;
  const double rc0 = fast_rcp((1.0 + (kind == CH_P ? dtau : 0.0)) - m0 * lo1);
  const double ib0 = live ? rc0 : 0.0;
  f.lo1 = lo1; f.lo2 = lo2;
  r2 = 0.0; r1 = 0.0; r0 = 0.0;
  if (live) {
    const int i0 = ln.base, i1 = i0 + ln.stride, i2 = i1 + ln.stride;
    ib[i2] = ib2; ib[i1] = ib1; ib[i0] = ib0;
    r2 = r[i2] - mm2 * rn; r1 = r[i1] - m1 * r2; r0 = r[i0] - m0 * r1;
  }
  // coupling of rows 0..2 to y = (Phi', Psi, Pi, v_b), by selects (no divergent branches)
  double C0[4] = {0, 0, 0, 0}, C1[4] = {0, 0, 0, 0}, C2[4] = {0, 0, 0, 0};
  {
    const bool isM = kind == CH_M, isT = kind == CH_T, isP = kind == CH_P, isTN = isT || kind == CH_N;
    const double htp = h * b.taup;
    C0[0] = isM ? h * ln.df0 : (isTN ? -h : 0.0);
    C1[1] = isM ? -f.hkap * (1.0 / 3.0) * b.eq * ln.df0 : (isTN ? f.hkap * (1.0 / 3.0) : 0.0);
    C1[3] = isT ? htp * (1.0 / 3.0) : 0.0;
    C2[2] = (isT || isP) ? -htp * 0.1 : 0.0;
    C0[2] = isP ? -htp * 0.5 : 0.0;
  }
  double be0[4], be1[4], be2[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const double V2 = C2[j], V1 = C1[j] - m1 * V2, V0 = C0[j] - m0 * V1;
    be0[j] = V0 * ib0;
    be1[j] = (V1 - lo1 * be0[j]) * ib1;
    be2[j] = (V2 - lo2 * be1[j]) * ib2;
    f.beta(0, j) = be0[j]; f.beta(1, j) = be1[j]; f.beta(2, j) = be2[j];
  }
  const int lT = ln.nq, lP = ln.nq + 1;
  double sPsi[4], sPhi[4], sPi[4], t1[4];
#pragma unroll
  for (int j = 0; j < 2; j++) { sPsi[j] = warp_sum(b.wPsi * be2[j]); sPhi[j] = warp_sum(b.wPhi * be0[j]); }
  const double wPsiT = shfl_d(b.wPsi, lT), wPhiT = shfl_d(b.wPhi, lT);
#pragma unroll
  for (int j = 2; j < 4; j++) { sPsi[j] = wPsiT * shfl_d(be2[j], lT); sPhi[j] = wPhiT * shfl_d(be0[j], lT); }
#pragma unroll
  for (int j = 0; j < 4; j++) { sPi[j] = shfl_d(be2[j], lT); t1[j] = shfl_d(be1[j], lT); }
  sPi[2] += shfl_d(be2[2] + be0[2], lP);
  const double Oc = b.Oc_a, Ob = b.Ob_a;
  const double hk = f.hkap;
  const double dPhi_y[4] = {h, 0, 0, 0};
  const double dDel_y[4] = {-3.0 * h, -hk * hk * f.vden, 0, 0};
  const double dDb_y[4] = {-3.0 * h, 0, 0, hk};
#pragma unroll
  for (int j = 0; j < 4; j++) {
    f.M[0][j] = dPhi_y[j] + b.cPsi * sPsi[j] + (j == 1 ? 1.0 : 0.0);
    f.M[1][j] = (j == 0 ? 1.0 : 0.0) - (j == 1 ? 1.0 : 0.0) + b.k2 * dPhi_y[j] - b.gPhi * (Oc * dDel_y[j] + Ob * dDb_y[j] + sPhi[j]);
    f.M[2][j] = (j == 2 ? 1.0 : 0.0) - sPi[j];
    f.M[3][j] = (j == 3 ? (1.0 + h - h * b.taup * b.R) : 0.0) + hk * ((j == 1 ? 1.0 : 0.0) + b.csb2 * dDb_y[j]) + f.e4c * t1[j];
  }
}

// Downward elimination only (the smoothing solve of the error estimate re-uses the last stage's factorisation).
__device__ __forceinline__ void rt_down_only(const Lane& ln, const RtFactor& f, double* ib, double* r, double& r0, double& r1, double& r2) {
  double ibn, lo_next, rn;
  rt_down<false>(ln, f, ib, r, ibn, lo_next, rn);
  r2 = 0.0; r1 = 0.0; r0 = 0.0;
  if (ln.kind != CH_IDLE) {
    const int i0 = ln.base, i1 = i0 + ln.stride, i2 = i1 + ln.stride;
    const bool top2 = (ln.len == 3);
    const double up2 = top2 ? 0.0 : f.hk * (1.0 - RLc(2));
    r2 = r[i2] - (up2 * ibn) * rn;
    r1 = r[i1] - (f.hk * (1.0 - RLc(1)) * ib[i2]) * r2;
    r0 = r[i0] - (f.hk * ib[i1]) * r1;
  }
}

// Border solve and upward sweep.  ZMODE: zout[idx] holds the stage's right-hand side and receives z = (U - rhs)/gamma;
// otherwise U is written to r[] (error smoothing).  r5[] are the five metric/matter scalars (same copy on every lane).
template <bool ZMODE>
__device__ __forceinline__ void rt_finish(const Lane& ln, const BgS& b, const RtFactor& f, const double* ib, double* r, double* zout,
                                          double r0, double r1, double r2, double (&r5)[5]) {
  const bool live = ln.kind != CH_IDLE;
  double ib0 = 0.0, ib1 = 0.0, ib2 = 0.0;
  const int i0 = ln.base, i1 = i0 + ln.stride, i2 = i1 + ln.stride;
  if (live) { ib0 = ib[i0]; ib1 = ib[i1]; ib2 = ib[i2]; }
  const double a0 = r0 * ib0, a1 = (r1 - f.lo1 * a0) * ib1, a2 = (r2 - f.lo2 * a1) * ib2;
  const int lT = ln.nq, lP = ln.nq + 1;
  const double sPsi = warp_sum(b.wPsi * a2);
  const double sPhi = warp_sum(b.wPhi * a0);
  const double sPi = shfl_d(a2, lT) + shfl_d(a2 + a0, lP);
  const double t1 = shfl_d(a1, lT);
  const double rPhi = r5[0], rdel = r5[1], rv = r5[2], rdb = r5[3], rvb = r5[4];
  const double hk = f.hkap, h = f.h;
  const double vc = rv * f.vden, dc = rdel + hk * vc;
  double rhs[4], y[4];
  rhs[0] = -(rPhi + b.cPsi * sPsi);
  rhs[1] = -(b.k2 * rPhi - b.gPhi * (b.Oc_a * dc + b.Ob_a * rdb + sPhi));
  rhs[2] = sPi;
  rhs[3] = -(hk * b.csb2 * rdb + f.e4c * t1 - rvb);
  ge4(f.M, rhs, y);
  r5[0] = rPhi + h * y[0];
  const double v = vc - hk * f.vden * y[1];
  r5[1] = rdel + hk * v - 3.0 * h * y[0];
  r5[2] = v;
  r5[3] = rdb - 3.0 * h * y[0] + hk * y[3];
  r5[4] = y[3];
  double U2s = 0.0;
  if (live) {
    double U0 = a0, U1 = a1, U2 = a2;
#pragma unroll
    for (int j = 0; j < 4; j++) { U0 += f.beta(0, j) * y[j]; U1 += f.beta(1, j) * y[j]; U2 += f.beta(2, j) * y[j]; }
    if (ZMODE) {
      zout[i0] = (U0 - zout[i0]) * (1.0 / KC_GAMMA); zout[i1] = (U1 - zout[i1]) * (1.0 / KC_GAMMA); zout[i2] = (U2 - zout[i2]) * (1.0 / KC_GAMMA);
    } else { r[i0] = U0; r[i1] = U1; r[i2] = U2; }
    U2s = U2;
  }
  // upward sweep: uniform and branch-free like the downward one (idle lanes: base = stride = len = 0, nothing stored)
  {
    double Up = live ? U2s : 0.0;
    const int len = ln.len;
#pragma unroll 4
    for (int l = 3; l < ln.maxlen - 1; l++) {
      const bool act = l < len - 1;
      const int idx = ln.base + (act ? l : 0) * ln.stride;
      const double U = (r[idx] + (f.hk * c_rl[l]) * Up) * ib[idx];
      if (act) { if (ZMODE) zout[idx] = (U - zout[idx]) * (1.0 / KC_GAMMA); else r[idx] = U; }
      Up = act ? U : Up;
    }
    if (len > 3) {               // truncation row
      const int idx = ln.base + (len - 1) * ln.stride;
      const double U = (r[idx] + f.hk * Up) * ib[idx];
      if (ZMODE) zout[idx] = (U - zout[idx]) * (1.0 / KC_GAMMA); else r[idx] = U;
    }
  }
}

// Chain ownership of a lane for cosmology c.  MAXLEN == 0: generic layout = the reference's unpack order.
// MAXLEN > 0: interleaved layout [l][chain] (+5 scalars after MAXLEN*NCH) for the register-resident solver.
template <class TR>
__device__ __forceinline__ void lane_setup(const DevCosmo& c, const SolveParams& p, Lane& ln, int lane = threadIdx.x & 31) {
  ln.lane = lane; ln.nq = c.nq; ln.L = p.L; ln.n = p.n;
  ln.riS = 2 * (p.L + 1) + (p.Lnu + 1) + (p.Lm + 1) * c.nq;
  ln.maxlen = max(p.L, max(p.Lnu, p.Lm)) + 1;
  ln.q = 0; ln.df0 = 0; ln.wq = 0;
  if (ln.lane < c.nq) { ln.kind = CH_M; ln.rbase = 2 * (p.L + 1) + (p.Lnu + 1) + ln.lane; ln.rstride = c.nq; ln.len = p.Lm + 1;
                        ln.q = c.q[ln.lane]; ln.df0 = c.df0[ln.lane]; ln.wq = c.wq[ln.lane]; }
  else if (ln.lane == c.nq) { ln.kind = CH_T; ln.rbase = 0; ln.rstride = 1; ln.len = p.L + 1; }
  else if (ln.lane == c.nq + 1) { ln.kind = CH_P; ln.rbase = p.L + 1; ln.rstride = 1; ln.len = p.L + 1; }
  else if (ln.lane == c.nq + 2) { ln.kind = CH_N; ln.rbase = 2 * (p.L + 1); ln.rstride = 1; ln.len = p.Lnu + 1; }
  else { ln.kind = CH_IDLE; ln.rbase = 0; ln.rstride = 0; ln.len = 0; }
  if constexpr (TR::MAXLEN > 0) {
    const bool own = (TR::NCH == 32) || ln.kind != CH_IDLE;     // compact layout: idle lanes own no column ...
    ln.base = own ? ln.lane : 0; ln.stride = own ? TR::NCH : 0; ln.iS = TR::MAXLEN * TR::NCH;
    // ... or (NCH = NQ + 4) share one dummy column that only ever holds zeros, so that the kernel needs no idle-lane guards
    if (TR::NCH == TR::NQ + 4 && !own) { ln.base = TR::NCH - 1; ln.stride = TR::NCH; }
  }
  else { ln.base = ln.rbase; ln.stride = ln.rstride; ln.iS = ln.riS; }
}

// Shared-memory doubles per warp.
template <class TR>
__host__ __device__ constexpr int k1_array_len(int n) { return TR::MAXLEN > 0 ? TR::MAXLEN * TR::NCH + 8 : n; }
template <class TR>
__host__ __device__ constexpr int k1_num_arrays() { return TR::MAXLEN > 0 ? 7 : 9; }
// extra doubles per warp: the lane-private factor scratch of the register path
template <class TR>
__host__ __device__ constexpr int k1_extra_doubles() { return TR::MAXLEN > 0 ? (TR::MAXLEN + 12) * 32 : (TR::RT ? 12 * 32 : 0); }

// K1_MINBLOCKS: resident warps per SM the register allocator must allow.  8 (254 registers, no spills) measured faster than
// 12 (168 registers: 9-12 warps all land there because registers are allocated per SM sub-partition; ~450 B of spills even
// with the factor state in shared memory and the compact NCH=18 layout) in both the latency- and the throughput-bound regime:
// 296 / 2000 / 8000 modes 39.8 / 56.9 / 185.6 ms at 8 warps vs 55.4 / 85.1 / 209.1 ms at 12 (profiles/r1_k1_hierarchy.md).
#ifndef K1_MINBLOCKS
#define K1_MINBLOCKS 8
#endif
// WPB > 1 (register-resident truncations only): WPB warps per block, each still owning its own k-mode, kept in loose LOCKSTEP by a
// named barrier before the stages selected by p.sync_mask (default: once per step, before stage 1 -- measured as good as before
// every stage, profiles/r2/k1_warp_lockstep_sweep.log, and cheaper in barrier waits).  The kernel is instruction-fetch bound when its warps roam
// independently through the 61 KB step loop (ncu, profiles/r2/k1_warp_occupancy_sweep.md: stall_no_inst 2 % with one warp per
// SM, 10 % with four, 18 % with eight): every stage executes the same 25 KB of code, so warps that enter it together share the
// instruction-cache fills.  A warp that runs out of work keeps arriving at the barrier until every warp of the block has.
template <class TR, int WPB = 1>
__global__ void __launch_bounds__(32 * WPB, (K1_MINBLOCKS / WPB > 0 ? K1_MINBLOCKS / WPB : 1)) hierarchy_kernel_t(SolveParams p) {
  extern __shared__ double sm_all[];
  __shared__ int nghost;
  constexpr int NCH = TR::NCH, MAXLEN = TR::MAXLEN;
  const int n = p.n;
  const int na = k1_array_len<TR>(n);
  double* const sm = sm_all + (size_t)(threadIdx.x >> 5) * ((size_t)k1_num_arrays<TR>() * na + k1_extra_doubles<TR>());
  Lane ln;
  const bool fixed = (p.mode == BOLT_MODE_FIXED);
  const double reltol = p.reltol, abstol = p.abstol;
  if constexpr (WPB > 1) { if (threadIdx.x == 0) nghost = 0; __syncthreads(); }

  while (true) {
    int w = 0;
    if ((threadIdx.x & 31) == 0) w = atomicAdd(p.counter, 1);
    w = __shfl_sync(FULL, w, 0);
    if (w >= p.nk) break;
    const int ik = p.order[w];
    const DevCosmo& c = *p.cos_list[ik / p.nk_per];
    lane_setup<TR>(c, p, ln);
    ln.k = p.k[ik];
    const double x_begin = c.x0, x_end = 0.0;

    // Array slots.  u/u_{n+1} and z1/z6 swap roles on every accepted step (no copies): physical slots 0/2 hold u and
    // u_{n+1} (alias of the z2 slot, free once the error estimate is formed), slots 1/6 hold z1 and z6; z3..z5 fixed.
    bool flipU = false, flipZ = false;
    double* const Z2 = sm + (size_t)3 * na;
    double* const Z3 = sm + (size_t)4 * na;
    double* const Z4 = sm + (size_t)5 * na;
    double* r = (MAXLEN > 0) ? Z2 : sm + (size_t)7 * na;          // scratch (generic: work vector of the solver)
    double* ib = (MAXLEN > 0) ? nullptr : sm + (size_t)8 * na;    // generic: inverse pivots
#define SLOT_U  (sm + (flipU ? (size_t)2 * na : (size_t)0))
#define SLOT_Z1 (sm + (flipU ? (size_t)0 : (size_t)2 * na))
#define SLOT_Z0 (sm + (flipZ ? (size_t)6 * na : (size_t)na))
#define SLOT_Z5 (sm + (flipZ ? (size_t)na : (size_t)6 * na))
    double* U = SLOT_U; double* Z0 = SLOT_Z0; double* Z1 = SLOT_Z1; double* Z5 = SLOT_Z5;
    if constexpr (MAXLEN > 0 || TR::RT) {   // zero-coefficient terms are still loaded by the branch-free stage assembly
      for (int i = ln.lane; i < 7 * na; i += 32) sm[i] = 0.0;
      __syncwarp();
    }

    Bg b;
    eval_bg(c, ln, x_begin, b);
    initial_conditions(c, ln, b, U);
    rhs_full(c, ln, b, U, Z5);          // f(u0) in the z6 slot (plays the role of z6/dt of a previous step)
    bool rsa_flag = (ln.k * b.eta > 240.0) && (-b.taup * b.H / b.eta > 100.0);

    int ix = 0;     // next x_grid row to sample; row 0 (sol(x0) = u0) is emitted by the first accepted step with theta = 0
    int status = BOLT_K_OK;
    long long nsteps = 0, nreject = 0;

    double x = x_begin, dt;
    auto sumsq_scaled = [&](const double* num, const double* a0, const double* a1) {
      double s = 0.0;
      #pragma unroll 1
      for (int l = 0; l < ln.len; l++) {
        const int idx = ln.base + l * ln.stride;
        const double sc = abstol + reltol * fmax(fabs(a0[idx]), fabs(a1[idx]));
        const double q = num[idx] / sc; s += q * q;
      }
      if (ln.lane < 5) {
        const int idx = ln.iS + ln.lane;
        const double sc = abstol + reltol * fmax(fabs(a0[idx]), fabs(a1[idx]));
        const double q = num[idx] / sc; s += q * q;
      }
      return warp_sum(s);
    };
    if (fixed) {
      dt = p.fixed_dt;
    } else {
      // initial step, Hairer-Wanner as in OrdinaryDiffEq's ode_determine_initdt (same as the oracle)
      const double d0 = sqrt(sumsq_scaled(U, U, U) / n), d1 = sqrt(sumsq_scaled(Z5, U, U) / n);
      double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
      dt0 = fmin(dt0, x_end - x_begin);
      #pragma unroll 1
      for (int l = 0; l < ln.len; l++) { const int idx = ln.base + l * ln.stride; r[idx] = U[idx] + dt0 * Z5[idx]; }
      if (ln.lane < 5) { const int idx = ln.iS + ln.lane; r[idx] = U[idx] + dt0 * Z5[idx]; }
      __syncwarp();
      Bg b1; eval_bg(c, ln, x_begin + dt0, b1);
      rhs_full(c, ln, b1, r, Z0);
      #pragma unroll 1
      for (int l = 0; l < ln.len; l++) { const int idx = ln.base + l * ln.stride; Z0[idx] -= Z5[idx]; }
      if (ln.lane < 5) { const int idx = ln.iS + ln.lane; Z0[idx] -= Z5[idx]; }
      __syncwarp();
      const double d2 = sqrt(sumsq_scaled(Z0, U, U) / n) / dt0;
      const double dm = fmax(d1, d2);
      const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(dm)) / 5.0);
      dt = fmin(100.0 * dt0, dt1);
    }
    // z1 slot <- f(u0): true z1 = s1 * Z0 with s1 = dt
    flipZ = !flipZ; Z0 = SLOT_Z0; Z5 = SLOT_Z5;
    double s1 = dt;

    const double beta1 = 7.0 / 40.0, beta2 = 2.0 / 20.0, safety = 0.9, qmin = 0.2, qmax = 10.0;
    double qold = 1e-4;
    const long long fixed_total = fixed ? llround((x_end - x_begin) / p.fixed_dt) : 0;
    long long fixed_left = fixed_total;
    const long long max_steps = p.max_steps > 0 ? p.max_steps : 1000000;
    ModeConst mc; mode_const(c, ln, mc);

    while (true) {
      bool clamped = false;
      if (fixed) { if (fixed_left == 0) break; }
      else {
        if (x >= x_end) break;
        if (x + dt >= x_end) { const double dtn = x_end - x; s1 *= dtn / dt; dt = dtn; clamped = true; }
      }
      if (nsteps + nreject >= max_steps) { status = BOLT_K_MAXSTEPS; break; }

      Bg bs;
      bool accept = true; double EEst = 0.0, q11 = 0.0;
      if constexpr (MAXLEN > 0) {
        // ---------------- register-resident stages ----------------
        RegFactor<TR> f;
        f.fs = sm + (size_t)7 * na + ln.lane;
        BgS bf;
        double rr[MAXLEN], r5[5];
        const int lo_ = ln.base;   // column of the lane inside a row of the interleaved layout
        // NCH = 32: a private column per lane; NCH = NQ + 4: compact layout + one shared all-zero column for the idle lanes.  Either
        // way no access needs an idle-lane guard.  The compact layout also lets the stage assembly run FLAT over the whole state
        // (7 lane-strided elements per lane instead of the lane's 11 rows + 5 scalars).
        constexpr bool FLAT = (NCH == TR::NQ + 4);
        const bool live = (NCH == 32) || FLAT || ln.kind != CH_IDLE;
        for (int s = 1; s <= 6; s++) {
          if constexpr (WPB > 1) { if ((p.sync_mask >> s) & 1) asm volatile("bar.sync 1, %0;" ::"r"(32 * WPB) : "memory"); }     // lockstep (see the kernel's header)
          // z slot of this stage: Z1 and Z5 swap physical slots with the step parity, Z2..Z4 sit at slots 3..5
          double* zout = sm + (size_t)((s == 1) ? (flipU ? 0 : 2) : (s >= 5) ? (flipZ ? 1 : 6) : s + 1) * na;
          if (s <= 5) {
            // branch-free assembly: coefficients of stages >= s are zero.  The right-hand side is parked in the stage's own
            // z slot (not yet written) so that z_s = (U - rhs)/gamma needs no register copy
            const double a0 = KC_A[s][0] * s1, a1 = KC_A[s][1], a2 = KC_A[s][2], a3 = KC_A[s][3], a4 = KC_A[s][4];
            if constexpr (FLAT) {
              constexpr int NFLAT = MAXLEN * NCH + 5;
#pragma unroll
              for (int t = 0; t < (NFLAT + 31) / 32; t++) {
                const int i = ln.lane + 32 * t;
                if (i < NFLAT) zout[i] = U[i] + a0 * Z0[i] + a1 * Z1[i] + a2 * Z2[i] + a3 * Z3[i] + a4 * Z4[i];
              }
              eval_bg_fast(c, ln, mc, x + KC_C[s] * dt, bf);
              __syncwarp();
#pragma unroll
              for (int l = 0; l < MAXLEN; l++) rr[l] = zout[lo_ + l * NCH];
#pragma unroll
              for (int j = 0; j < 5; j++) r5[j] = zout[ln.iS + j];
            } else {
#pragma unroll
              for (int l = 0; l < MAXLEN; l++) {
                const int idx = lo_ + l * NCH;      // padded rows / idle lanes hold zeros
                double v = 0.0;
                if (live) { v = U[idx] + a0 * Z0[idx] + a1 * Z1[idx] + a2 * Z2[idx] + a3 * Z3[idx] + a4 * Z4[idx]; zout[idx] = v; }
                rr[l] = v;
              }
#pragma unroll
              for (int j = 0; j < 5; j++) {
                const int idx = ln.iS + j;
                const double v = U[idx] + a0 * Z0[idx] + a1 * Z1[idx] + a2 * Z2[idx] + a3 * Z3[idx] + a4 * Z4[idx];
                r5[j] = v; zout[idx] = v;
              }
              eval_bg_fast(c, ln, mc, x + KC_C[s] * dt, bf);
            }
            rsa_flag |= (ln.k * bf.eta > 240.0) && (-bf.taup * bf.H > 100.0 * bf.eta);
            factor_reg<TR>(ln, bf, KC_GAMMA * dt, f);
          } else {
            // "stage 7": error estimate err = sum (b - bhat)_j z_j, smoothed below by W^{-1} of the last stage (smooth_est);
            // u_{n+1} = u_n + sum b_j z_j goes to the z2 slot, which is free once err is formed
            const double e0 = KC_E[0] * s1, b0 = KC_A[5][0] * s1;
#pragma unroll
            for (int l = 0; l < MAXLEN; l++) {
              const int idx = lo_ + l * NCH;
              double e = 0.0;
              if (live) {
                const double z0 = Z0[idx], z2 = Z2[idx], z3 = Z3[idx], z4 = Z4[idx], z5 = Z5[idx];
                e = e0 * z0 + KC_E[2] * z2 + KC_E[3] * z3 + KC_E[4] * z4 + KC_E[5] * z5;
                Z1[idx] = U[idx] + b0 * z0 + KC_A[5][2] * z2 + KC_A[5][3] * z3 + KC_A[5][4] * z4 + KC_GAMMA * z5;
              }
              rr[l] = e;
            }
#pragma unroll
            for (int j = 0; j < 5; j++) {
              const int idx = ln.iS + j;
              const double z0 = Z0[idx], z2 = Z2[idx], z3 = Z3[idx], z4 = Z4[idx], z5 = Z5[idx];
              r5[j] = e0 * z0 + KC_E[2] * z2 + KC_E[3] * z3 + KC_E[4] * z4 + KC_E[5] * z5;
              Z1[idx] = U[idx] + b0 * z0 + KC_A[5][2] * z2 + KC_A[5][3] * z3 + KC_A[5][4] * z4 + KC_GAMMA * z5;
            }
            if (fixed) break;
          }
          solve_reg<TR>(ln, bf, f, rr, r5);
          if (s <= 5) {
#pragma unroll
            for (int l = 0; l < MAXLEN; l++) if (live) { const int idx = lo_ + l * NCH; zout[idx] = (rr[l] - zout[idx]) * (1.0 / KC_GAMMA); }
            // every lane holds identical scalars and stores them itself (same value, same address): a lane later reads
            // back what it wrote, so no warp-level synchronisation is needed anywhere in the stage loop
            if constexpr (FLAT) {
              // the parked scalars were written by one lane each: every lane reads, the warp converges, every lane writes the same
              // value; the closing barrier also publishes the rows to the next stage's flat pass
              double zz[5];
#pragma unroll
              for (int j = 0; j < 5; j++) zz[j] = (r5[j] - zout[ln.iS + j]) * (1.0 / KC_GAMMA);
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 5; j++) zout[ln.iS + j] = zz[j];
              __syncwarp();
            } else {
#pragma unroll
              for (int j = 0; j < 5; j++) { const int idx = ln.iS + j; zout[idx] = (r5[j] - zout[idx]) * (1.0 / KC_GAMMA); }
            }
          }
        }
        __syncwarp();     // the sampling / rotation code below reads other lanes' data
        if (!fixed) {
          double ssum = 0.0;
#pragma unroll
          for (int l = 0; l < MAXLEN; l++) {
            const int idx = lo_ + l * NCH;
            if (live) {
              const double sc = abstol + reltol * fmax(fabs(U[idx]), fabs(Z1[idx]));
              const double q = rr[l] * fast_rcp(sc); ssum += q * q;
            }
          }
          if (ln.lane < 5) {
            double e = r5[0]; e = (ln.lane == 1) ? r5[1] : e; e = (ln.lane == 2) ? r5[2] : e; e = (ln.lane == 3) ? r5[3] : e; e = (ln.lane == 4) ? r5[4] : e;
            const int idx = ln.iS + ln.lane;
            const double sc = abstol + reltol * fmax(fabs(U[idx]), fabs(Z1[idx]));
            const double q = e * fast_rcp(sc); ssum += q * q;
          }
          EEst = sqrt(warp_sum(ssum) / n);
        }
      } else if constexpr (TR::RT) {
        // ---------------- runtime-truncation stages (long chains; rows in shared memory) ----------------
        // Elementwise passes (stage assembly, error/u_{n+1} pass, error norm) run flat over the n entries, lane-strided and
        // conflict-free; only the sweeps follow the chains.  __syncwarp() separates the two access patterns.
        RtFactor f;
        f.bs = sm + (size_t)9 * na + ln.lane;
        BgS bf;
        double r5[5], q5[5], r0, r1, r2;
        const int iS = ln.iS;
        for (int s = 1; s <= 6; s++) {
          if constexpr (WPB > 1) { if ((p.sync_mask >> s) & 1) asm volatile("bar.sync 1, %0;" ::"r"(32 * WPB) : "memory"); }     // lockstep
          double* zout = (s == 1) ? Z1 : (s == 2) ? Z2 : (s == 3) ? Z3 : (s == 4) ? Z4 : Z5;
          if (s <= 5) {
            const double a0 = KC_A[s][0] * s1, a1 = KC_A[s][1], a2 = KC_A[s][2], a3 = KC_A[s][3], a4 = KC_A[s][4];
#pragma unroll 4
            for (int i = ln.lane; i < n; i += 32) {
              const double v = U[i] + a0 * Z0[i] + a1 * Z1[i] + a2 * Z2[i] + a3 * Z3[i] + a4 * Z4[i];   // zero coefficients: branch-free
              r[i] = v; zout[i] = v;
            }
            eval_bg_fast(c, ln, mc, x + KC_C[s] * dt, bf);
            rsa_flag |= (ln.k * bf.eta > 240.0) && (-bf.taup * bf.H > 100.0 * bf.eta);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 5; j++) { r5[j] = r[iS + j]; q5[j] = r5[j]; }
            rt_factor_down(ln, bf, KC_GAMMA * dt, f, ib, r, r0, r1, r2);
            rt_finish<true>(ln, bf, f, ib, r, zout, r0, r1, r2, r5);
            // every lane holds the same scalars and stores them itself (same value, same address; no read-modify-write)
#pragma unroll
            for (int j = 0; j < 5; j++) zout[iS + j] = (r5[j] - q5[j]) * (1.0 / KC_GAMMA);
            __syncwarp();
          } else {
            const double e0 = KC_E[0] * s1, b0 = KC_A[5][0] * s1;
#pragma unroll 4
            for (int i = ln.lane; i < n; i += 32) {
              const double z0 = Z0[i], z2 = Z2[i], z3 = Z3[i], z4 = Z4[i], z5 = Z5[i];
              Z1[i] = U[i] + b0 * z0 + KC_A[5][2] * z2 + KC_A[5][3] * z3 + KC_A[5][4] * z4 + KC_GAMMA * z5;
              r[i] = e0 * z0 + KC_E[2] * z2 + KC_E[3] * z3 + KC_E[4] * z4 + KC_E[5] * z5;
            }
            __syncwarp();
            if (fixed) break;
#pragma unroll
            for (int j = 0; j < 5; j++) r5[j] = r[iS + j];
            rt_down_only(ln, f, ib, r, r0, r1, r2);
            rt_finish<false>(ln, bf, f, ib, r, nullptr, r0, r1, r2, r5);
#pragma unroll
            for (int j = 0; j < 5; j++) r[iS + j] = r5[j];
            __syncwarp();
          }
        }
        if (!fixed) {
          double ssum = 0.0;
#pragma unroll 4
          for (int i = ln.lane; i < n; i += 32) {
            const double sc = abstol + reltol * fmax(fabs(U[i]), fabs(Z1[i]));
            const double q = r[i] * fast_rcp(sc); ssum += q * q;
          }
          EEst = sqrt(warp_sum(ssum) / n);
        }
      } else {
        // ---------------- generic stages (runtime chain lengths, work vectors in shared memory) ----------------
        Factor f;
        for (int s = 1; s < 6; s++) {
          const double a0 = KC_A[s][0] * s1, a1 = KC_A[s][1], a2 = KC_A[s][2], a3 = KC_A[s][3], a4 = KC_A[s][4];
          const double* z0 = Z0; const double* z1p = Z1; const double* z2p = Z2; const double* z3p = Z3; const double* z4p = Z4;
          auto rhs_of = [&](int idx) {
            double v = U[idx] + a0 * z0[idx];
            if (s > 1) v += a1 * z1p[idx];
            if (s > 2) v += a2 * z2p[idx];
            if (s > 3) v += a3 * z3p[idx];
            if (s > 4) v += a4 * z4p[idx];
            return v;
          };
          for (int l = 0; l < ln.len; l++) { const int idx = ln.base + l * ln.stride; r[idx] = rhs_of(idx); }
          if (ln.lane < 5) { const int idx = ln.iS + ln.lane; r[idx] = rhs_of(idx); }
          __syncwarp();
          eval_bg(c, ln, x + KC_C[s] * dt, bs);
          rsa_flag |= (ln.k * bs.eta > 240.0) && (-bs.taup * bs.H / bs.eta > 100.0);
          factor(c, ln, bs, KC_GAMMA * dt, ib, f);
          double* zout = (s == 1) ? Z1 : (s == 2) ? Z2 : (s == 3) ? Z3 : (s == 4) ? Z4 : Z5;
          solve(c, ln, bs, f, ib, r, zout, rhs_of);
        }
        // error estimate err = sum (b - bhat)_j z_j, and u_{n+1} = u_n + sum b_j z_j (= U_6 up to rounding).
        // u_{n+1} goes to the z2 slot, which is free once err is formed.
        {
          const double e0 = KC_E[0] * s1, b0 = KC_A[5][0] * s1;
          auto pass = [&](int idx) {
            const double z0 = Z0[idx], z2 = Z2[idx], z3 = Z3[idx], z4 = Z4[idx], z5 = Z5[idx];
            r[idx] = e0 * z0 + KC_E[2] * z2 + KC_E[3] * z3 + KC_E[4] * z4 + KC_E[5] * z5;
            Z1[idx] = U[idx] + b0 * z0 + KC_A[5][2] * z2 + KC_A[5][3] * z3 + KC_A[5][4] * z4 + KC_GAMMA * z5;
          };
          for (int l = 0; l < ln.len; l++) pass(ln.base + l * ln.stride);
          if (ln.lane < 5) pass(ln.iS + ln.lane);
          __syncwarp();
        }
        if (!fixed) {
          auto none = [&](int) { return 0.0; };
          solve(c, ln, bs, f, ib, r, (double*)nullptr, none);      // smooth_est: W^{-1} err with the last stage's W
          EEst = sqrt(sumsq_scaled(r, U, Z1) / n);
        }
      }
      if (!fixed) {
        if (!isfinite(EEst)) { status = BOLT_K_NONFINITE; break; }
        // Controller input floored at 1e-6: below that the estimate is rounding noise of the stiff start-up phase and
        // would make the step sequence implementation-dependent (DESIGN.md "controller"); acceptance uses the raw value.
        q11 = exp(beta1 * log(fmax(EEst, 1e-6)));
        accept = EEst <= 1.0;
        if (p.dbg && ik == 0 && ln.lane == 0 && nsteps + nreject < p.dbg_cap) {
          double* d = p.dbg + 4 * (nsteps + nreject); d[0] = x; d[1] = dt; d[2] = EEst; d[3] = accept ? 1.0 : 0.0;
        }
      }
      if (accept) {
        const bool last = fixed ? (fixed_left == 1) : clamped;
        const double xn1 = last ? x_end : (fixed ? (x_begin + (double)(fixed_total - fixed_left + 1) * p.fixed_dt) : (x + dt));
        while (ix < c.n_x) {
          const double xs = c.x0 + c.dx * ix;
          if (!last && xs > xn1 + 1e-12) break;
          if (ix >= p.ix_first) {
            double th = (xs - x) / dt; if (th > 1.0) th = 1.0;
            Hermite hm = hermite_weights(th);
            sample_sources(c, ln, p, ik, ix, xs, hm, U, Z1, Z0, s1, Z5, rsa_flag, (MAXLEN > 0 || TR::RT) ? &mc : nullptr);
          }
          ix++;
        }
        x = xn1; nsteps++;
        // rotate: u <- u_{n+1}; z1 <- z6 (scaled by dt_new/dt through s1)
        flipU = !flipU; flipZ = !flipZ; U = SLOT_U; Z1 = SLOT_Z1; Z0 = SLOT_Z0; Z5 = SLOT_Z5;
        if (fixed) { fixed_left--; s1 = 1.0; }
        else {
          double q = q11 * exp(-beta2 * log(qold));
          q = fmax(1.0 / qmax, fmin(1.0 / qmin, q / safety));
          if (q <= 1.2 && q >= 1.0) q = 1.0;
          qold = fmax(EEst, 1e-4);
          const double dtn = dt / q;
          s1 = dtn / dt; dt = dtn;
        }
      } else {
        nreject++;
        const double dtn = dt / fmin(1.0 / qmin, q11 / safety);
        s1 *= dtn / dt; dt = dtn;
        if (!(dt > 1e-14)) { status = BOLT_K_DT_UNDERFLOW; break; }
      }
    }
    if (rsa_flag && status == BOLT_K_OK) status = BOLT_K_RSA_TRIGGERED;
    if (p.u_final) {
      double* out = p.u_final + (size_t)ik * n * p.out_nd;
#pragma unroll 1
      for (int l = 0; l < ln.len; l++) out[(size_t)(ln.rbase + l * ln.rstride) * p.out_nd] = U[ln.base + l * ln.stride];
      if (ln.lane < 5) out[(size_t)(ln.riS + ln.lane) * p.out_nd] = U[ln.iS + ln.lane];
    }
    if (ln.lane == 0) {
      if (p.status) p.status[ik] = status;
      if (p.nsteps) p.nsteps[ik] = nsteps;
      if (p.nreject) p.nreject[ik] = nreject;
    }
    __syncwarp();
  }
  if constexpr (WPB > 1) {
    // out of work: keep the block's barrier complete until every warp is
    if ((threadIdx.x & 31) == 0) atomicAdd(&nghost, 1);
    while (true) {
      asm volatile("bar.sync 1, %0;" ::"r"(32 * WPB) : "memory");
      if (*(volatile int*)&nghost == WPB) break;
    }
  }
}

}  // namespace bolt
