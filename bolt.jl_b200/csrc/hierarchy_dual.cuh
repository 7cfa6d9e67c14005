// K1 with forward-mode partials (nd = 1 + NP): value and parameter gradient of every k-mode from ONE pass.
//
// The reference gets gradients by running the whole stack on ForwardDiff.Dual numbers (examples/plot_deriv_cl.jl:28-33),
// i.e. the stiff solver integrates a Dual-valued state.  Because the hierarchy is linear, u' = A(x;p) u, the
// sensitivities S_j = du/dp_j obey  S_j' = A S_j + (dA/dp_j) u, and every implicit stage of the ESDIRK scheme is
//     W U = r ,            W S_j = r_j + h G_j ,   G_j = (dA/dp_j)(x_s) U ,   W = I - h A(x_s)
// -- the SAME matrix for the value and for every partial.  So the kernel factors W once per stage (the value-path
// factor()/solve() of hierarchy_kernel.cuh), obtains all G_j from one evaluation of the right-hand side in dual
// arithmetic on (U, zero partials), and back-solves NP more right-hand sides.  In adaptive mode the error norm runs over
// the value and all partials, as in the reference (a Dual-valued state makes OrdinaryDiffEq's norm see every component).
//
// Physics below is the same as in hierarchy_kernel.cuh (same reference line citations), written once for a scalar
// type T = Dual<NP>.  The state lives in shared memory component-major: array a, component j at (a*(1+NP)+j)*n.
#pragma once
#include "hierarchy_kernel.cuh"
#include "dual.cuh"

namespace bolt {

// ---- dual views of the cosmology -----------------------------------------------------------------------------
template <int NP> __device__ __forceinline__ Dual<NP> cs_d(const DevCosmo& c, int i) {
  Dual<NP> r; r.v = c.s[i];
#pragma unroll
  for (int j = 0; j < NP; j++) r.d[j] = c.ds[i][j];
  return r;
}
template <int NP> __device__ __forceinline__ Dual<NP> ctab_d(const DevCosmo& c, int t, double x) {
  double tt = (x - c.x0) / c.dx;
  int i = (int)floor(tt);
  i = max(0, min(i, c.n_x - 2));
  const double d = tt - (double)i, e = 1.0 - d;
  const double w0 = e * e * e * (1.0 / 6.0), w1 = 2.0 / 3.0 - d * d + d * d * d * 0.5;
  const double w2 = 2.0 / 3.0 - e * e + e * e * e * 0.5, w3 = d * d * d * (1.0 / 6.0);
  Dual<NP> r;
  const double* cv = c.tab[t] + i;
  r.v = cv[0] * w0 + cv[1] * w1 + cv[2] * w2 + cv[3] * w3;
#pragma unroll
  for (int j = 0; j < NP; j++) {
    const double* cp = c.dtab[t] + (size_t)j * (c.n_x + 2) + i;
    r.d[j] = cp[0] * w0 + cp[1] * w1 + cp[2] * w2 + cp[3] * w3;
  }
  return r;
}
template <int NP> __device__ __forceinline__ Dual<NP> shfl_T(const Dual<NP>& a, int src) {
  Dual<NP> r; r.v = shfl_d(a.v, src);
#pragma unroll
  for (int j = 0; j < NP; j++) r.d[j] = shfl_d(a.d[j], src);
  return r;
}
template <int NP> __device__ __forceinline__ Dual<NP> warp_sum_T(const Dual<NP>& a) {
  Dual<NP> r; r.v = warp_sum(a.v);
#pragma unroll
  for (int j = 0; j < NP; j++) r.d[j] = warp_sum(a.d[j]);
  return r;
}

// component-major state array: component j of element idx at p[j*n + idx]
template <int NP> struct DArr {
  double* p; int n;
  __device__ __forceinline__ Dual<NP> get(int idx) const {
    Dual<NP> r; r.v = p[idx];
#pragma unroll
    for (int j = 0; j < NP; j++) r.d[j] = p[(size_t)(1 + j) * n + idx];
    return r;
  }
  __device__ __forceinline__ Dual<NP> getv(int idx) const { return Dual<NP>(p[idx]); }   // value with zero partials
  __device__ __forceinline__ void put(int idx, const Dual<NP>& a) const {
    p[idx] = a.v;
#pragma unroll
    for (int j = 0; j < NP; j++) p[(size_t)(1 + j) * n + idx] = a.d[j];
  }
  __device__ __forceinline__ double* comp(int j) const { return p + (size_t)j * n; }
};

// ---- background at one abscissa in dual arithmetic (eval_bg; perturbations.jl:164-172) --------------------------
template <int NP> struct BgD {
  typedef Dual<NP> T;
  double x, a;
  T H, eta, taup, taupp, csb2, Hp, kappa, qe, eq, wPsi, wPhi, cPsi, gPhi, k2, R;
};
// everything derived from the table values b.H, b.eta, b.taup, b.csb2 (already set) at abscissa x
template <int NP> __device__ __forceinline__ void finish_bg_d(const DevCosmo& c, const Lane& ln, double x, BgD<NP>& b);
template <int NP> __device__ __forceinline__ void eval_bg_d(const DevCosmo& c, const Lane& ln, double x, BgD<NP>& b) {
  b.H = ctab_d<NP>(c, BOLT_T_H, x); b.eta = ctab_d<NP>(c, BOLT_T_eta, x); b.taup = ctab_d<NP>(c, BOLT_T_taup, x);
  b.taupp = ctab_d<NP>(c, BOLT_T_taupp, x); b.csb2 = ctab_d<NP>(c, BOLT_T_csb2, x); b.Hp = ctab_d<NP>(c, BOLT_T_Hp, x);
  finish_bg_d<NP>(c, ln, x, b);
}
// Stage variant: the (1+NP) components of the four tables a stage needs (H, eta, tau', c_s^2) are evaluated by ONE lane
// each -- lane t*(1+NP)+comp -- instead of by every lane; bg_stage_prefetch issues the loads (call it early: the value is
// not needed before the stage value is solved), eval_bg_d_stage broadcasts.  tau'' and H' are not used by a stage.
template <int NP> __device__ __forceinline__ double bg_stage_prefetch(const DevCosmo& c, const Lane& ln, double x) {
  constexpr int ND = 1 + NP;
  static_assert(4 * ND <= 32, "one (table, component) pair per lane");
  const int which[4] = {BOLT_T_H, BOLT_T_eta, BOLT_T_taup, BOLT_T_csb2};
  double tv = 0.0;
  if (ln.lane < 4 * ND) {
    const int t = ln.lane / ND, comp = ln.lane - t * ND;
    const int tab = (t == 0) ? which[0] : (t == 1) ? which[1] : (t == 2) ? which[2] : which[3];
    const double* cp = comp == 0 ? c.tab[tab] : c.dtab[tab] + (size_t)(comp - 1) * (c.n_x + 2);
    tv = spline_eval(cp, c.n_x, c.x0, c.dx, x);
  }
  return tv;
}
template <int NP> __device__ __forceinline__ void eval_bg_d_stage(const DevCosmo& c, const Lane& ln, double x, double tv, BgD<NP>& b) {
  constexpr int ND = 1 + NP;
  auto fetch = [&](int t) {
    Dual<NP> r; r.v = shfl_d(tv, t * ND);
#pragma unroll
    for (int j = 0; j < NP; j++) r.d[j] = shfl_d(tv, t * ND + 1 + j);
    return r;
  };
  b.H = fetch(0); b.eta = fetch(1); b.taup = fetch(2); b.csb2 = fetch(3);
  b.taupp = Dual<NP>(0.0); b.Hp = Dual<NP>(0.0);
  finish_bg_d<NP>(c, ln, x, b);
}
template <int NP> __device__ __forceinline__ void finish_bg_d(const DevCosmo& c, const Lane& ln, double x, BgD<NP>& b) {
  typedef Dual<NP> T;
  b.x = x; b.a = exp(x);
  b.kappa = ln.k / b.H;
  const T Om_r = cs_d<NP>(c, BOLT_S_Omega_r), Om_b = cs_d<NP>(c, BOLT_S_Omega_b), rho_crit = cs_d<NP>(c, BOLT_S_rho_crit);
  const T H0 = cs_d<NP>(c, BOLT_S_H0);
  b.R = 4.0 * Om_r / (3.0 * Om_b * b.a);
  b.cPsi = 12.0 * H0 * H0 / (ln.k * ln.k) / (b.a * b.a);
  b.gPhi = H0 * H0 / (2.0 * b.H * b.H);
  b.k2 = ln.k * ln.k / (3.0 * b.H * b.H);
  b.qe = T(1.0); b.eq = T(1.0); b.wPsi = T(0.0); b.wPhi = T(0.0);
  const double ia2 = 1.0 / (b.a * b.a);
  T Om_nu; Om_nu.v = c.Omega_nu;
#pragma unroll
  for (int j = 0; j < NP; j++) Om_nu.d[j] = c.dOmega_nu[j];
  if (ln.kind == CH_M) {
    T q, wq; q.v = c.q[ln.lane]; wq.v = c.wq[ln.lane];
#pragma unroll
    for (int j = 0; j < NP; j++) { q.d[j] = c.dq[ln.lane][j]; wq.d[j] = c.dwq[ln.lane][j]; }
    const T am = b.a * cs_d<NP>(c, BOLT_S_Sum_m_nu);
    const T eps = dsqrt(q * q + am * am);
    b.qe = q / eps; b.eq = eps / q;
    b.wPhi = wq * eps * ia2 / rho_crit;
    b.wPsi = wq * (q * q / eps) / rho_crit * 0.25;
  } else if (ln.kind == CH_T) {
    b.wPhi = 4.0 * Om_r * ia2; b.wPsi = Om_r;
  } else if (ln.kind == CH_N) {
    b.wPhi = 4.0 * Om_nu * ia2; b.wPsi = Om_nu;
  }
}

template <int NP> struct MetricD { Dual<NP> Phi, delta, v, delta_b, v_b, Psi, dPhi, Pi; };

// rhs_row in dual arithmetic (hierarchy!, perturbations.jl:203-264)
template <int NP, class Get>
__device__ __forceinline__ Dual<NP> rhs_row_d(const Lane& ln, const BgD<NP>& b, const MetricD<NP>& m, int l, Get get) {
  typedef Dual<NP> T;
  const bool photon = (ln.kind == CH_T || ln.kind == CH_P);
  const T kq = b.kappa * b.qe;
  if (l == ln.len - 1) {
    T damp = (double)ln.len / (b.H * b.eta);
    if (photon) damp = damp - b.taup;
    return kq * get(l - 1) - damp * get(l);
  }
  if (l == 0) {
    T r = -(kq * get(1));
    if (ln.kind == CH_M) r = r + m.dPhi * ln.df0;
    else if (ln.kind == CH_P) r = r + b.taup * (get(0) - m.Pi * 0.5);
    else r = r - m.dPhi;
    return r;
  }
  const double rl = c_rl[l];
  T r = kq * (rl * get(l - 1) - (1.0 - rl) * get(l + 1));
  if (l == 1) {
    if (ln.kind == CH_M) r = r - b.kappa * (1.0 / 3.0) * b.eq * m.Psi * ln.df0;
    else if (ln.kind != CH_P) r = r + b.kappa * (1.0 / 3.0) * m.Psi;
    if (ln.kind == CH_T) r = r + b.taup * (m.v_b * (1.0 / 3.0));
  }
  if (photon) r = r + b.taup * (get(l) - (l == 2 ? m.Pi * 0.1 : T(0.0)));
  return r;
}

template <int NP>
__device__ __forceinline__ void metric_from_chains_d(const Lane& ln, const BgD<NP>& b, const Dual<NP>& c0, const Dual<NP>& c2,
                                                     MetricD<NP>& m, const DevCosmo& c) {
  typedef Dual<NP> T;
  const T sPsi = warp_sum_T(b.wPsi * c2);
  const T sPhi = warp_sum_T(b.wPhi * c0);
  T pi(0.0);
  if (ln.kind == CH_T) pi = c2; else if (ln.kind == CH_P) pi = c2 + c0;
  m.Pi = warp_sum_T(pi);
  m.Psi = -m.Phi - b.cPsi * sPsi;
  m.dPhi = m.Psi - b.k2 * m.Phi + b.gPhi * (cs_d<NP>(c, BOLT_S_Omega_c) / b.a * m.delta + cs_d<NP>(c, BOLT_S_Omega_b) / b.a * m.delta_b + sPhi);
}

// G = (dA/dp) U for a plain (partial-free) stage value U: same rows as rhs_row_d, but every product is dual-coefficient x
// double-state (NP+1 multiplies instead of 2NP+1).  Only the partials of the result are used.
template <int NP> struct MetricG { double Phi, delta, v, delta_b, v_b, Pi; Dual<NP> Psi, dPhi; };

template <int NP, class Get>
__device__ __forceinline__ Dual<NP> rhs_row_g(const Lane& ln, const BgD<NP>& b, const MetricG<NP>& m, int l, Get get) {
  typedef Dual<NP> T;
  const bool photon = (ln.kind == CH_T || ln.kind == CH_P);
  const T kq = b.kappa * b.qe;
  if (l == ln.len - 1) {
    T damp = (double)ln.len / (b.H * b.eta);
    if (photon) damp = damp - b.taup;
    return kq * get(l - 1) - damp * get(l);
  }
  if (l == 0) {
    T r = -(kq * get(1));
    if (ln.kind == CH_M) r = r + m.dPhi * ln.df0;
    else if (ln.kind == CH_P) r = r + b.taup * (get(0) - m.Pi * 0.5);
    else r = r - m.dPhi;
    return r;
  }
  const double rl = c_rl[l];
  T r = kq * (rl * get(l - 1) - (1.0 - rl) * get(l + 1));
  if (l == 1) {
    if (ln.kind == CH_M) r = r - b.kappa * (1.0 / 3.0) * b.eq * m.Psi * ln.df0;
    else if (ln.kind != CH_P) r = r + b.kappa * (1.0 / 3.0) * m.Psi;
    if (ln.kind == CH_T) r = r + b.taup * (m.v_b * (1.0 / 3.0));
  }
  if (photon) r = r + b.taup * (get(l) - (l == 2 ? m.Pi * 0.1 : 0.0));
  return r;
}

// Solve W X_j = R_j for NR right-hand sides at once, in place in the NR component arrays W[j] (shared memory), with the
// factorisation of the current stage.  Same algebra as solve(); the NR dependency chains are interleaved so that they
// hide each other's latency (the dual kernel runs at 2 warps per SM, so ILP is the only latency hiding there is).
template <int NR>
__device__ __forceinline__ void solve_multi(const DevCosmo& c, const Lane& ln, const Bg& b, const Factor& f, const double* ib,
                                            double* const (&W)[NR]) {
  double ibn = 0.0, rn[NR];
#pragma unroll
  for (int j = 0; j < NR; j++) rn[j] = 0.0;
#pragma unroll 1
  for (int l = ln.maxlen - 1; l >= 3; l--) {
    if (l < ln.len) {
      const int idx = ln.base + l * ln.stride;
      const double m = (l < ln.len - 1) ? f.hk * (1.0 - c_rl[l]) * ibn : 0.0;
#pragma unroll
      for (int j = 0; j < NR; j++) { const double v = W[j][idx] - m * rn[j]; W[j][idx] = v; rn[j] = v; }
      ibn = ib[idx];
    }
  }
  double a0[NR], a1[NR], a2[NR];
  const int i0 = ln.base, i1 = ln.base + ln.stride, i2 = ln.base + 2 * ln.stride;
  if (ln.kind != CH_IDLE) {
    const double ib2 = ib[i2], ib1 = ib[i1], ib0 = ib[i0];
    const double m2 = f.hk * (1.0 - c_rl[2]) * ibn, m1 = f.hk * (1.0 - c_rl[1]) * ib2, m0 = f.hk * ib1;
#pragma unroll
    for (int j = 0; j < NR; j++) {
      const double r2 = W[j][i2] - m2 * rn[j];
      const double r1 = W[j][i1] - m1 * r2;
      const double r0 = W[j][i0] - m0 * r1;
      a0[j] = r0 * ib0; a1[j] = (r1 - f.lo1 * a0[j]) * ib1; a2[j] = (r2 - f.lo2 * a1[j]) * ib2;
    }
  } else {
#pragma unroll
    for (int j = 0; j < NR; j++) { a0[j] = 0.0; a1[j] = 0.0; a2[j] = 0.0; }
  }
  const int iS = ln.iS;
  const double Oc = c.s[BOLT_S_Omega_c] / b.a, Ob = c.s[BOLT_S_Omega_b] / b.a;
  const double hk = f.hkap, h = f.h;
  double y[NR][4];
#pragma unroll
  for (int j = 0; j < NR; j++) {
    const double sPsi = warp_sum(b.wPsi * a2[j]);
    const double sPhi = warp_sum(b.wPhi * a0[j]);
    double pi = 0.0;
    if (ln.kind == CH_T) pi = a2[j]; else if (ln.kind == CH_P) pi = a2[j] + a0[j];
    const double sPi = warp_sum(pi);
    const double t1 = shfl_d(a1[j], ln.nq);
    const double rPhi = W[j][iS], rdel = W[j][iS + 1], rv = W[j][iS + 2], rdb = W[j][iS + 3], rvb = W[j][iS + 4];
    const double vc = rv * f.vden, dc = rdel + hk * vc;
    double rhs[4];
    rhs[0] = -(rPhi + b.cPsi * sPsi);
    rhs[1] = -(b.k2 * rPhi - b.gPhi * (Oc * dc + Ob * rdb + sPhi));
    rhs[2] = sPi;
    rhs[3] = -(hk * b.csb2 * rdb + f.e4c * t1 - rvb);
    lu4_solve(f, rhs, y[j]);
    // scalars (every lane computes the same values; lane 0 stores after the warp has read the inputs)
    const double Phi = rPhi + h * y[j][0];
    const double v = vc - hk * f.vden * y[j][1];
    const double del = rdel + hk * v - 3.0 * h * y[j][0];
    const double db = rdb - 3.0 * h * y[j][0] + hk * y[j][3];
    __syncwarp();
    if (ln.lane == 0) { W[j][iS] = Phi; W[j][iS + 1] = del; W[j][iS + 2] = v; W[j][iS + 3] = db; W[j][iS + 4] = y[j][3]; }
  }
  if (ln.kind != CH_IDLE) {
    double Up[NR];
#pragma unroll
    for (int j = 0; j < NR; j++) {
      double U0 = a0[j], U1 = a1[j], U2 = a2[j];
#pragma unroll
      for (int q = 0; q < 4; q++) { U0 += f.beta0[q] * y[j][q]; U1 += f.beta1[q] * y[j][q]; U2 += f.beta2[q] * y[j][q]; }
      W[j][i0] = U0; W[j][i1] = U1; W[j][i2] = U2; Up[j] = U2;
    }
#pragma unroll 1
    for (int l = 3; l < ln.len; l++) {
      const int idx = ln.base + l * ln.stride;
      const double lo = (l == ln.len - 1) ? -f.hk : -f.hk * c_rl[l];
      const double ibl = ib[idx];
#pragma unroll
      for (int j = 0; j < NR; j++) { const double U = (W[j][idx] - lo * Up[j]) * ibl; W[j][idx] = U; Up[j] = U; }
    }
  }
  __syncwarp();
}

// du = A(x;p) u in dual arithmetic.  zero_state_partials = true evaluates on (u, 0): the partials of the result are then
// G_j = (dA/dp_j) u, the inhomogeneity of the sensitivity equations.
template <int NP>
__device__ __forceinline__ void rhs_full_d(const DevCosmo& c, const Lane& ln, const BgD<NP>& b, const DArr<NP>& u, const DArr<NP>& du,
                                           bool zero_state_partials) {
  typedef Dual<NP> T;
  auto ld = [&](int idx) { return zero_state_partials ? u.getv(idx) : u.get(idx); };
  MetricD<NP> m;
  const int iS = ln.iS;
  m.Phi = ld(iS); m.delta = ld(iS + 1); m.v = ld(iS + 2); m.delta_b = ld(iS + 3); m.v_b = ld(iS + 4);
  T c0(0.0), c2(0.0);
  if (ln.kind != CH_IDLE) { c0 = ld(ln.base); c2 = ld(ln.base + 2 * ln.stride); }
  metric_from_chains_d<NP>(ln, b, c0, c2, m, c);
  const T T1 = shfl_T(ln.kind == CH_T ? ld(ln.base + ln.stride) : T(0.0), ln.nq);
  auto get = [&](int l) { return ld(ln.base + l * ln.stride); };
#pragma unroll 1
  for (int l = 0; l < ln.len; l++) du.put(ln.base + l * ln.stride, rhs_row_d<NP>(ln, b, m, l, get));
  if (ln.lane == 0) {
    du.put(iS, m.dPhi);
    du.put(iS + 1, b.kappa * m.v - 3.0 * m.dPhi);
    du.put(iS + 2, -m.v - b.kappa * m.Psi);
    du.put(iS + 3, b.kappa * m.v_b - 3.0 * m.dPhi);
    du.put(iS + 4, -m.v_b - b.kappa * (m.Psi + b.csb2 * m.delta_b) + b.taup * b.R * (3.0 * T1 + m.v_b));
  }
  __syncwarp();
}

// initial_conditions (perturbations.jl:274-338) in dual arithmetic
template <int NP>
__device__ __forceinline__ void initial_conditions_d(const DevCosmo& c, const Lane& ln, const BgD<NP>& b, const DArr<NP>& u) {
  typedef Dual<NP> T;
  const double k = ln.k;
  const T Hx = b.H, eta = b.eta, taup = b.taup;
  const T N_nu = cs_d<NP>(c, BOLT_S_N_nu);
  const double c411 = pow(4.0 / 11.0, 4.0 / 3.0);
  const T f_nu = 1.0 / (1.0 + 1.0 / (7.0 * (3.0 / 3.0) * N_nu / 8.0 * c411));
  const T Phi = (4.0 * f_nu + 10.0) / (4.0 * f_nu + 15.0) * 1.0;
  const T C = -((15.0 + 4.0 * f_nu) / (20.0 + 8.0 * f_nu)) * Phi;
  const T T0 = -40.0 * C / (15.0 + 4.0 * f_nu) / 4.0;
  const T T1 = 10.0 * C / (15.0 + 4.0 * f_nu) * (k * k * eta) / (3.0 * k);
  const T T2 = -8.0 * k / (15.0 * Hx * taup) * T1;
  const T N2 = -(k * k * eta * eta) / 15.0 * 1.0 / (1.0 + 2.0 / 5.0 * f_nu) * Phi / 2.0;
  const int st = ln.stride;
  if (ln.kind == CH_T) {
    u.put(ln.base, T0); u.put(ln.base + st, T1); u.put(ln.base + 2 * st, T2);
    T prev = T2;
#pragma unroll 1
    for (int l = 3; l < ln.len; l++) { prev = -(double)l / (2 * l + 1) * k / (Hx * taup) * prev; u.put(ln.base + l * st, prev); }
  } else if (ln.kind == CH_P) {
    u.put(ln.base, (5.0 / 4.0) * T2); u.put(ln.base + st, -k / (4.0 * Hx * taup) * T2);
    T prev = (1.0 / 4.0) * T2; u.put(ln.base + 2 * st, prev);
#pragma unroll 1
    for (int l = 3; l < ln.len; l++) { prev = -(double)l / (2 * l + 1) * k / (Hx * taup) * prev; u.put(ln.base + l * st, prev); }
  } else if (ln.kind == CH_N) {
    u.put(ln.base, T0); u.put(ln.base + st, T1); u.put(ln.base + 2 * st, N2);
    T prev = N2;
#pragma unroll 1
    for (int l = 3; l < ln.len; l++) { prev = k / ((2 * l + 1) * Hx) * prev; u.put(ln.base + l * st, prev); }
  } else if (ln.kind == CH_M) {
    const double df0 = ln.df0;
    u.put(ln.base, -T0 * df0);
    u.put(ln.base + st, -b.eq * T1 * df0);
    T prev = -N2 * df0; u.put(ln.base + 2 * st, prev);
#pragma unroll 1
    for (int l = 3; l < ln.len; l++) { prev = b.qe * k / ((2 * l + 1) * Hx) * prev; u.put(ln.base + l * st, prev); }
  }
  if (ln.lane == 0) {
    const T delta = 3.0 / 4.0 * (4.0 * T0), v = -3.0 * k * T1;
    u.put(ln.iS, Phi); u.put(ln.iS + 1, delta); u.put(ln.iS + 2, v); u.put(ln.iS + 3, delta); u.put(ln.iS + 4, v);
  }
  __syncwarp();
}

// One sample of both source grids with partials (spectra.jl:13-18; perturbations.jl:343-404)
template <int NP>
__device__ __forceinline__ void sample_sources_d(const DevCosmo& c, const Lane& ln, const SolveParams& p, int ik, int ix, double xs,
                                                 const Hermite& hm, const DArr<NP>& u0, const DArr<NP>& u1, const DArr<NP>& z1, double s1,
                                                 const DArr<NP>& z6, bool& rsa_flag, bool write_value = true, const int* cmap = nullptr) {
  // write_value / cmap: the CTA-per-mode kernel samples every partial on the warp that owns it (NP = 1 on a single-partial view):
  // that warp writes only its partial, into the output slot cmap[0]
  typedef Dual<NP> T;
  constexpr int ND = 1 + NP;
  const int* const cm = cmap ? cmap : p.comp_map;
  auto herm = [&](int idx) { return hm.c0 * u0.get(idx) + hm.c1 * u1.get(idx) + (hm.d0 * s1) * z1.get(idx) + hm.d1 * z6.get(idx); };
  if (!p.S_T && !p.S_P) return;
  BgD<NP> b; eval_bg_d<NP>(c, ln, xs, b);
  const T Hpp = ctab_d<NP>(c, BOLT_T_Hpp, xs), tau = ctab_d<NP>(c, BOLT_T_tau, xs), g = ctab_d<NP>(c, BOLT_T_g, xs);
  const T gp = ctab_d<NP>(c, BOLT_T_gp, xs), gpp = ctab_d<NP>(c, BOLT_T_gpp, xs);
  T uL[5];
#pragma unroll
  for (int l = 0; l < 5; l++) uL[l] = (ln.kind != CH_IDLE && l < ln.len) ? herm(ln.base + l * ln.stride) : T(0.0);
  MetricD<NP> m;
  m.Phi = herm(ln.iS); m.delta = herm(ln.iS + 1); m.v = herm(ln.iS + 2); m.delta_b = herm(ln.iS + 3); m.v_b = herm(ln.iS + 4);
  metric_from_chains_d<NP>(ln, b, uL[0], uL[2], m, c);
  auto get = [&](int l) { T v = uL[0]; if (l == 1) v = uL[1]; if (l == 2) v = uL[2]; if (l == 3) v = uL[3]; if (l == 4) v = uL[4]; return v; };
  T d[4];
#pragma unroll
  for (int l = 0; l < 4; l++) d[l] = (ln.kind != CH_IDLE) ? rhs_row_d<NP>(ln, b, m, l, get) : T(0.0);
  const T T1 = shfl_T(uL[1], ln.nq);
  const T dvb = -m.v_b - b.kappa * (m.Psi + b.csb2 * m.delta_b) + b.taup * b.R * (3.0 * T1 + m.v_b);
  const bool rsa_on = (ln.k * b.eta.v > 240.0) && (-b.taup.v * b.H.v / b.eta.v > 100.0);
  if (rsa_on) {
    rsa_flag = true;
    if (ln.kind == CH_T) {
      uL[0] = m.Phi - b.H / ln.k * b.taup * m.v_b;
      uL[1] = b.H / ln.k * (-2.0 * m.dPhi + b.taup * (m.Phi - b.csb2 * m.delta_b) + b.H / ln.k * (b.taupp - b.taup) * m.v_b);
      uL[2] = T(0.0);
    } else if (ln.kind == CH_N) {
      uL[0] = m.Phi; uL[1] = -2.0 * b.H / ln.k * m.dPhi; uL[2] = T(0.0);
    }
    if (ln.kind == CH_T || ln.kind == CH_P || ln.kind == CH_N) { d[0] = T(0.0); d[1] = T(0.0); d[2] = T(0.0); d[3] = T(0.0); }
  }
  const T rho_crit = cs_d<NP>(c, BOLT_S_rho_crit), Om_r = cs_d<NP>(c, BOLT_S_Omega_r);
  const T wS = (ln.kind == CH_M) ? b.wPsi * 4.0 * rho_crit : T(0.0);
  const T sigM = warp_sum_T(wS * uL[2]);
  const T sigMp = warp_sum_T(wS * d[2]);
  const int lT = ln.nq, lP = ln.nq + 1, lN = ln.nq + 2;
  const T Th0 = shfl_T(uL[0], lT), Th1 = shfl_T(uL[1], lT), Th2 = shfl_T(uL[2], lT), Th3 = shfl_T(uL[3], lT);
  const T dTh1 = shfl_T(d[1], lT), dTh2 = shfl_T(d[2], lT), dTh3 = shfl_T(d[3], lT);
  const T P0 = shfl_T(uL[0], lP), P1 = shfl_T(uL[1], lP), P2 = shfl_T(uL[2], lP), P3 = shfl_T(uL[3], lP);
  const T dP0 = shfl_T(d[0], lP), dP1 = shfl_T(d[1], lP), dP2 = shfl_T(d[2], lP), dP3 = shfl_T(d[3], lP);
  const T N2 = shfl_T(uL[2], lN), dN2 = shfl_T(d[2], lN);
  T Om_nu; Om_nu.v = c.Omega_nu;
#pragma unroll
  for (int j = 0; j < NP; j++) Om_nu.d[j] = c.dOmega_nu[j];
  const double k = ln.k;
  const T Hx = b.H, Hp = b.Hp;
  const T Psi = -m.Phi - b.cPsi * (Om_r * Th2 + Om_nu * N2 + sigM / rho_crit / 4.0);
  const T dPsi = -m.dPhi - b.cPsi * (Om_r * (dTh2 - 2.0 * Th2) + Om_nu * (dN2 - 2.0 * N2) + (sigMp - 2.0 * sigM) / rho_crit / 4.0);
  const T Pi = Th2 + P2 + P0, dPi = dTh2 + dP2 + dP0;
  const T term1 = g * (Th0 + Psi + Pi / 4.0) + dexp(-tau) * (dPsi - m.dPhi);
  const T term2 = (-1.0 / k) * (Hp * g * m.v_b + Hx * gp * m.v_b + Hx * g * dvb);
  const T ddPi = 2.0 * k / (5.0 * Hx) * (-(Hp / Hx) * Th1 + dTh1) + (3.0 / 10.0) * (b.taupp * Pi + b.taup * dPi)
                 - 3.0 * k / (5.0 * Hx) * (-(Hp / Hx) * (Th3 + P1 + P3) + (dTh3 + dP1 + dP3));
  const T term3 = (3.0 / (4.0 * k * k)) * ((Hp * Hp + Hx * Hpp) * g * Pi + 3.0 * Hx * Hp * (gp * Pi + g * dPi)
                                           + Hx * Hx * (gpp * Pi + 2.0 * gp * dPi + g * ddPi));
  T eta_end; eta_end.v = c.eta_end;
#pragma unroll
  for (int j = 0; j < NP; j++) eta_end.d[j] = c.deta_end[j];
  const T y = k * (eta_end - b.eta);
  if (ln.lane == 0) {
    const T sT = term1 + term2 + term3;
    const T sP = (3.0 / (4.0 * y * y)) * g * Pi;
    if (p.S_T) { double* o = p.S_T + ((size_t)ik * c.n_x + ix) * p.out_nd; if (write_value) o[0] = sT.v;
#pragma unroll
      for (int j = 0; j < NP; j++) o[cm[j]] = sT.d[j]; }
    if (p.S_P) { double* o = p.S_P + ((size_t)ik * c.n_x + ix) * p.out_nd; if (write_value) o[0] = sP.v;
#pragma unroll
      for (int j = 0; j < NP; j++) o[cm[j]] = sP.d[j]; }
  }
}

// Shared memory: 7 state arrays x (1+NP) components + work r + inverse pivots ib, each n doubles.
template <int NP> __host__ __device__ constexpr size_t k1_dual_smem_doubles(int n) { return (size_t)(7 * (1 + NP) + 2) * n; }

template <int NP>
__global__ void __launch_bounds__(32) hierarchy_dual_kernel(SolveParams p) {
  extern __shared__ double sm[];
  typedef Dual<NP> T;
  constexpr int ND = 1 + NP;
  const int n = p.n;
  Lane ln;
  const bool fixed = (p.mode == BOLT_MODE_FIXED);
  const double reltol = p.reltol, abstol = p.abstol;
  const size_t astr = (size_t)ND * n;   // stride between state arrays

  while (true) {
    int w = 0;
    if (threadIdx.x == 0) w = atomicAdd(p.counter, 1);
    w = __shfl_sync(FULL, w, 0);
    if (w >= p.nk) break;
    const int ik = p.order[w];
    const DevCosmo& c = *p.cos_list[ik / p.nk_per];
    lane_setup<Trunc<0, 0, 0, 0>>(c, p, ln);
    ln.k = p.k[ik];
    const double x_begin = c.x0, x_end = 0.0;

    bool flipU = false, flipZ = false;
    const DArr<NP> Z2{sm + 3 * astr, n}, Z3{sm + 4 * astr, n}, Z4{sm + 5 * astr, n};
    double* r = sm + 7 * astr;
    double* ib = r + n;
#define DSLOT_U  DArr<NP>{sm + (flipU ? 2 * astr : 0), n}
#define DSLOT_Z1 DArr<NP>{sm + (flipU ? 0 : 2 * astr), n}
#define DSLOT_Z0 DArr<NP>{sm + (flipZ ? 6 * astr : astr), n}
#define DSLOT_Z5 DArr<NP>{sm + (flipZ ? astr : 6 * astr), n}
    DArr<NP> U = DSLOT_U, Z0 = DSLOT_Z0, Z1 = DSLOT_Z1, Z5 = DSLOT_Z5;

    BgD<NP> bd;
    eval_bg_d<NP>(c, ln, x_begin, bd);
    initial_conditions_d<NP>(c, ln, bd, U);
    rhs_full_d<NP>(c, ln, bd, U, Z5, false);       // f(u0) with its partials
    bool rsa_flag = (ln.k * bd.eta.v > 240.0) && (-bd.taup.v * bd.H.v / bd.eta.v > 100.0);

    int ix = 0;
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
      // initial step from the value part only (same numbers as the value-only kernel)
      const double d0 = sqrt(sumsq_scaled(U.p, U.p, U.p) / n), d1 = sqrt(sumsq_scaled(Z5.p, U.p, U.p) / n);
      double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
      dt0 = fmin(dt0, x_end - x_begin);
#pragma unroll 1
      for (int l = 0; l < ln.len; l++) { const int idx = ln.base + l * ln.stride; r[idx] = U.p[idx] + dt0 * Z5.p[idx]; }
      if (ln.lane < 5) { const int idx = ln.iS + ln.lane; r[idx] = U.p[idx] + dt0 * Z5.p[idx]; }
      __syncwarp();
      Bg b1; eval_bg(c, ln, x_begin + dt0, b1);
      rhs_full(c, ln, b1, r, Z0.p);
#pragma unroll 1
      for (int l = 0; l < ln.len; l++) { const int idx = ln.base + l * ln.stride; Z0.p[idx] -= Z5.p[idx]; }
      if (ln.lane < 5) { const int idx = ln.iS + ln.lane; Z0.p[idx] -= Z5.p[idx]; }
      __syncwarp();
      const double d2 = sqrt(sumsq_scaled(Z0.p, U.p, U.p) / n) / dt0;
      const double dm = fmax(d1, d2);
      const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(dm)) / 5.0);
      dt = fmin(100.0 * dt0, dt1);
    }
    flipZ = !flipZ; Z0 = DSLOT_Z0; Z5 = DSLOT_Z5;
    double s1 = dt;

    const double beta1 = 7.0 / 40.0, beta2 = 2.0 / 20.0, safety = 0.9, qmin = 0.2, qmax = 10.0;
    double qold = 1e-4;
    const long long fixed_total = fixed ? llround((x_end - x_begin) / p.fixed_dt) : 0;
    long long fixed_left = fixed_total;
    const long long max_steps = p.max_steps > 0 ? p.max_steps : 1000000;

    while (true) {
      bool clamped = false;
      if (fixed) { if (fixed_left == 0) break; }
      else {
        if (x >= x_end) break;
        if (x + dt >= x_end) { const double dtn = x_end - x; s1 *= dtn / dt; dt = dtn; clamped = true; }
      }
      if (nsteps + nreject >= max_steps) { status = BOLT_K_MAXSTEPS; break; }

      Bg bs; Factor f;
      bool accept = true; double EEst = 0.0, q11 = 0.0;
      for (int s = 1; s < 6; s++) {
        const double a0 = KC_A[s][0] * s1, a1 = KC_A[s][1], a2 = KC_A[s][2], a3 = KC_A[s][3], a4 = KC_A[s][4];
        const double h = KC_GAMMA * dt;
        const DArr<NP> zout = (s == 1) ? Z1 : (s == 2) ? Z2 : (s == 3) ? Z3 : (s == 4) ? Z4 : Z5;
        // component comp of rhs_i = u_n + sum_j a_ij z_j
        auto rhs_comp = [&](int comp, int idx) {
          const size_t o = (size_t)comp * n + idx;
          double v = U.p[o] + a0 * Z0.p[o];
          if (s > 1) v += a1 * Z1.p[o];
          if (s > 2) v += a2 * Z2.p[o];
          if (s > 3) v += a3 * Z3.p[o];
          if (s > 4) v += a4 * Z4.p[o];
          return v;
        };
        const double xs = x + KC_C[s] * dt;
        // ---- value: factor W and solve W U = r ----
        auto rhs0 = [&](int idx) { return rhs_comp(0, idx); };
#pragma unroll 1
        for (int l = 0; l < ln.len; l++) { const int idx = ln.base + l * ln.stride; r[idx] = rhs0(idx); }
        if (ln.lane < 5) { const int idx = ln.iS + ln.lane; r[idx] = rhs0(idx); }
        __syncwarp();
        eval_bg(c, ln, xs, bs);
        rsa_flag |= (ln.k * bs.eta > 240.0) && (-bs.taup * bs.H / bs.eta > 100.0);
        factor(c, ln, bs, h, ib, f);
        solve(c, ln, bs, f, ib, r, zout.p, rhs0);
        // ---- G_j = (dA/dp_j) U: rows of the right-hand side with dual coefficients on the plain stage value U (in r[]) ----
        eval_bg_d<NP>(c, ln, xs, bd);
        {
          MetricG<NP> m;
          const int iS = ln.iS;
          m.Phi = r[iS]; m.delta = r[iS + 1]; m.v = r[iS + 2]; m.delta_b = r[iS + 3]; m.v_b = r[iS + 4];
          double c0 = 0.0, c2 = 0.0;
          if (ln.kind != CH_IDLE) { c0 = r[ln.base]; c2 = r[ln.base + 2 * ln.stride]; }
          const T sPsi = warp_sum_T(bd.wPsi * c2), sPhi = warp_sum_T(bd.wPhi * c0);
          double pi = 0.0;
          if (ln.kind == CH_T) pi = c2; else if (ln.kind == CH_P) pi = c2 + c0;
          m.Pi = warp_sum(pi);
          m.Psi = -(bd.cPsi * sPsi) - m.Phi;
          m.dPhi = m.Psi - bd.k2 * m.Phi + bd.gPhi * (cs_d<NP>(c, BOLT_S_Omega_c) * (m.delta / bd.a) + cs_d<NP>(c, BOLT_S_Omega_b) * (m.delta_b / bd.a) + sPhi);
          const double T1 = shfl_d(ln.kind == CH_T ? r[ln.base + ln.stride] : 0.0, ln.nq);
          auto get = [&](int l) { return r[ln.base + l * ln.stride]; };
          // zout partial components <- r_j + h * G_j: the right-hand sides of the sensitivity systems, solved in place below
#pragma unroll 1
          for (int l = 0; l < ln.len; l++) {
            const int idx = ln.base + l * ln.stride;
            const T g = rhs_row_g<NP>(ln, bd, m, l, get);
#pragma unroll
            for (int j = 0; j < NP; j++) zout.p[(size_t)(1 + j) * n + idx] = rhs_comp(1 + j, idx) + h * g.d[j];
          }
          if (ln.lane == 0) {
            const T g0 = m.dPhi, g1 = bd.kappa * m.v - 3.0 * m.dPhi, g2 = -(bd.kappa * m.Psi) - m.v, g3 = bd.kappa * m.v_b - 3.0 * m.dPhi;
            const T g4 = -(bd.kappa * (m.Psi + bd.csb2 * m.delta_b)) + bd.taup * bd.R * (3.0 * T1 + m.v_b) - m.v_b;
#pragma unroll
            for (int j = 0; j < NP; j++) {
              double* zp = zout.p + (size_t)(1 + j) * n + iS;
              zp[0] = rhs_comp(1 + j, iS) + h * g0.d[j]; zp[1] = rhs_comp(1 + j, iS + 1) + h * g1.d[j]; zp[2] = rhs_comp(1 + j, iS + 2) + h * g2.d[j];
              zp[3] = rhs_comp(1 + j, iS + 3) + h * g3.d[j]; zp[4] = rhs_comp(1 + j, iS + 4) + h * g4.d[j];
            }
          }
          __syncwarp();
        }
        // ---- partials: W S_j = r_j + h G_j, all NP systems at once with the same factorisation; then z_{s,j} = (S_j - r_j)/gamma ----
        {
          double* Wp[NP];
#pragma unroll
          for (int j = 0; j < NP; j++) Wp[j] = zout.comp(1 + j);
          solve_multi<NP>(c, ln, bs, f, ib, Wp);
#pragma unroll 1
          for (int l = 0; l < ln.len; l++) {
            const int idx = ln.base + l * ln.stride;
#pragma unroll
            for (int j = 0; j < NP; j++) Wp[j][idx] = (Wp[j][idx] - rhs_comp(1 + j, idx)) * (1.0 / KC_GAMMA);
          }
          if (ln.lane < 5) {
            const int idx = ln.iS + ln.lane;
#pragma unroll
            for (int j = 0; j < NP; j++) Wp[j][idx] = (Wp[j][idx] - rhs_comp(1 + j, idx)) * (1.0 / KC_GAMMA);
          }
          __syncwarp();
        }
      }
      // u_{n+1} for every component (into the z2 slot) -- the error vectors are formed per component below
      {
        const double b0 = KC_A[5][0] * s1;
        auto pass = [&](int idx) {
#pragma unroll 1
          for (int j = 0; j < ND; j++) {
            const size_t o = (size_t)j * n + idx;
            Z1.p[o] = U.p[o] + b0 * Z0.p[o] + KC_A[5][2] * Z2.p[o] + KC_A[5][3] * Z3.p[o] + KC_A[5][4] * Z4.p[o] + KC_GAMMA * Z5.p[o];
          }
        };
#pragma unroll 1
        for (int l = 0; l < ln.len; l++) pass(ln.base + l * ln.stride);
        if (ln.lane < 5) pass(ln.iS + ln.lane);
        __syncwarp();
      }
      if (!fixed) {
        // Error control over the value AND the partials, like the reference: with a Dual-valued state OrdinaryDiffEq's norm
        // runs over all dual components [dep-knowledge: DiffEqBase ODE_DEFAULT_NORM on ForwardDiff.Dual], each element scaled
        // by abstol + reltol*max(|u_n|,|u_{n+1}|) with |.| taken over (value, partials).  Each component's estimate
        // sum (b-bhat)_j z_j is smoothed by W^{-1} of the last stage (smooth_est) with the factorisation in hand.
        const double e0 = KC_E[0] * s1;
        auto elem_scale = [&](int idx) {
          double n0 = 0.0, n1 = 0.0;
#pragma unroll
          for (int j = 0; j < ND; j++) { const double a = U.p[(size_t)j * n + idx], b2 = Z1.p[(size_t)j * n + idx]; n0 += a * a; n1 += b2 * b2; }
          return abstol + reltol * sqrt(fmax(n0, n1));
        };
        // the z3 slot (logical Z2) is recomputed by the next attempt whether this step is accepted or not: its ND component
        // arrays become the error vectors, smoothed in place all at once
        auto errpass = [&](int idx) {
#pragma unroll
          for (int j = 0; j < ND; j++) {
            const size_t o = (size_t)j * n + idx;
            Z2.p[o] = e0 * Z0.p[o] + KC_E[2] * Z2.p[o] + KC_E[3] * Z3.p[o] + KC_E[4] * Z4.p[o] + KC_E[5] * Z5.p[o];
          }
        };
#pragma unroll 1
        for (int l = 0; l < ln.len; l++) errpass(ln.base + l * ln.stride);
        if (ln.lane < 5) errpass(ln.iS + ln.lane);
        __syncwarp();
        double* We[ND];
#pragma unroll
        for (int j = 0; j < ND; j++) We[j] = Z2.comp(j);
        solve_multi<ND>(c, ln, bs, f, ib, We);
        double ssum = 0.0;
        auto acc = [&](int idx) {
          const double isc = 1.0 / elem_scale(idx);
#pragma unroll
          for (int j = 0; j < ND; j++) { const double q = We[j][idx] * isc; ssum += q * q; }
        };
#pragma unroll 1
        for (int l = 0; l < ln.len; l++) acc(ln.base + l * ln.stride);
        if (ln.lane < 5) acc(ln.iS + ln.lane);
        EEst = sqrt(warp_sum(ssum) / ((double)n * p.out_nd));     // DiffEqBase: sum of squares over value and partials / totallength of the CALLER's dual state (zero partials count)
        if (!isfinite(EEst)) { status = BOLT_K_NONFINITE; break; }
        q11 = exp(beta1 * log(fmax(EEst, 1e-6)));
        accept = EEst <= 1.0;
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
            sample_sources_d<NP>(c, ln, p, ik, ix, xs, hm, U, Z1, Z0, s1, Z5, rsa_flag);
          }
          ix++;
        }
        x = xn1; nsteps++;
        flipU = !flipU; flipZ = !flipZ; U = DSLOT_U; Z1 = DSLOT_Z1; Z0 = DSLOT_Z0; Z5 = DSLOT_Z5;
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
    if (p.u_final) {   // [nk][n][nd]
      double* out = p.u_final + (size_t)ik * n * p.out_nd;
#pragma unroll 1
      for (int l = 0; l < ln.len; l++)
        for (int j = 0; j < ND; j++) out[(size_t)(ln.rbase + l * ln.rstride) * p.out_nd + (j ? p.comp_map[j - 1] : 0)] = U.p[(size_t)j * n + ln.base + l * ln.stride];
      if (ln.lane < 5) for (int j = 0; j < ND; j++) out[(size_t)(ln.riS + ln.lane) * p.out_nd + (j ? p.comp_map[j - 1] : 0)] = U.p[(size_t)j * n + ln.iS + ln.lane];
    }
    if (ln.lane == 0) {
      if (p.status) p.status[ik] = status;
      if (p.nsteps) p.nsteps[ik] = nsteps;
      if (p.nreject) p.nreject[ik] = nreject;
    }
    __syncwarp();
  }
}

// plin (src/spectra.jl:163-198) from the dual state at x = 0; u_final is [nk][n][nd], pk is [nk][nd]
template <int NP>
__global__ void plin_kernel_d(const DevCosmo* cos, const double* __restrict__ kk, int nk, const double* __restrict__ u_final,
                              int L, int Lnu, int Lm, double* __restrict__ pk) {
  typedef Dual<NP> T;
  constexpr int ND = 1 + NP;
  const int ik = blockIdx.x * blockDim.x + threadIdx.x;
  if (ik >= nk) return;
  const DevCosmo& c = *cos;
  const int nq = c.nq;
  const int iM = 2 * (L + 1) + (Lnu + 1), iS = iM + (Lm + 1) * nq, n = iS + 5;
  const double* res = u_final + (size_t)ik * n * ND;
  auto ld = [&](int idx) { T r; r.v = res[(size_t)idx * ND];
#pragma unroll
    for (int j = 0; j < NP; j++) r.d[j] = res[(size_t)idx * ND + 1 + j];
    return r; };
  const double k = kk[ik], x = 0.0, a = 1.0;
  const T rho0M = ctab_d<NP>(c, BOLT_T_rho0M, x), Hx = ctab_d<NP>(c, BOLT_T_H, x);
  const T m = cs_d<NP>(c, BOLT_S_Sum_m_nu);
  T rho(0.0), th(0.0);
  for (int i = 0; i < nq; i++) {
    T q, wq; q.v = c.q[i]; wq.v = c.wq[i];
#pragma unroll
    for (int j = 0; j < NP; j++) { q.d[j] = c.dq[i][j]; wq.d[j] = c.dwq[i][j]; }
    const T eps = dsqrt(q * q + (a * m) * (a * m));
    rho = rho + wq * eps * ld(iM + i);
    th = th + wq * q * ld(iM + nq + i);
  }
  const T Mrho = rho / rho0M, Mtheta = k * th / rho0M;
  const T dcN = ld(iS + 1), dbN = ld(iS + 3), vcN = ld(iS + 2), vbN = ld(iS + 4);
  const T vmnuN = -(Mtheta / k);
  const T hh = cs_d<NP>(c, BOLT_S_h), Om_r = cs_d<NP>(c, BOLT_S_Omega_r), N_nu = cs_d<NP>(c, BOLT_S_N_nu);
  const T Tg = dpow(15.0 / (M_PI * M_PI) * cs_d<NP>(c, BOLT_S_rho_crit) * Om_r, 0.25);
  const double zeta = 1.2020569;
  const T nufac = (90.0 * zeta / (11.0 * pow(M_PI, 4.0))) * (Om_r * hh * hh / Tg) * dpow(N_nu / 3.0, 0.75);
  const T Om_nu = m * nufac / (hh * hh);
  const T Om_c = cs_d<NP>(c, BOLT_S_Omega_c), Om_b = cs_d<NP>(c, BOLT_S_Omega_b);
  const T Om_m = Om_c + Om_b + Om_nu;
  const T dc = dcN - 3.0 * Hx * vcN / k, db = dbN - 3.0 * Hx * vbN / k;
  const T dmnu = Mrho - 3.0 * Hx * vmnuN / k;
  const T dm = (Om_c * dc + Om_b * db + Om_nu * dmnu) / Om_m;
  const T Pprim = cs_d<NP>(c, BOLT_S_A) * dpow(k / 0.05, cs_d<NP>(c, BOLT_S_n) - 1.0);
  const T P = (2.0 * M_PI * M_PI / (k * k * k)) * dm * dm * Pprim;
  pk[(size_t)ik * ND] = P.v;
#pragma unroll
  for (int j = 0; j < NP; j++) pk[(size_t)ik * ND + 1 + j] = P.d[j];
}

}  // namespace bolt
