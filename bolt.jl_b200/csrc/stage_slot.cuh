// Stage factorisation in a SHARED-MEMORY SLOT (sm_100a, FP64) -- the register-resident truncations of source_grid.
//
// W_s = I - gamma dt A(x_n + c_s dt) depends on (x_n, dt) only, not on the solution, so one warp can factor it into a slot that
// any number of other warps back-solve with: factor_stage() writes the chain pivots, the beta vectors, the pivoted LU of the 4x4
// border system and the stage scalars; solve_slot() is solve_reg() of hierarchy_kernel.cuh reading them from the slot.  Used by
// K1 with partials, one CTA per mode (hierarchy_dual_cta.cuh: value warp factors, every sensitivity warp solves).  The first
// CTA-per-mode value kernel built on these (round 2, solver + two factoriser warps) was superseded by hierarchy_pipe.cuh.
#pragma once
#include "hierarchy_kernel.cuh"

namespace bolt {

// Shared-memory layout of one CTA (doubles).
template <class TR>
struct CtaLayout {
  static constexpr int NCH = TR::NCH, MAXLEN = TR::MAXLEN;
  static constexpr int NA = MAXLEN * NCH + 8;        // one state array: interleaved [l][chain] + 5 scalars (as hierarchy_kernel_t)
  static constexpr int NARR = 7;                     // u / u_{n+1}, z1..z6 (rotating, see the solver)
  // factor slot of one stage: per-lane rows (column = the lane's chain column) ...
  static constexpr int R_IB = 0, R_BETA = MAXLEN, R_HK = MAXLEN + 12, R_WPSI = MAXLEN + 13, R_WPHI = MAXLEN + 14, LROWS = MAXLEN + 15;
  // ... and a warp-uniform block: pivoted LU of the border system + the stage scalars the back-solve needs
  static constexpr int UNI = LROWS * NCH;
  static constexpr int U_L = 0, U_U = 6, U_ID = 12, U_PERM = 16, U_H = 17, U_HKAP = 18, U_VDEN = 19, U_E4C = 20, U_CPSI = 21,
                       U_K2 = 22, U_GPHI = 23, U_OCA = 24, U_OBA = 25, U_CSB2 = 26, U_RSA = 27, NUNI = 28;
  static constexpr int SLOT = ((UNI + NUNI + 1) / 2) * 2;
  static constexpr int NSLOT = 5;
  static constexpr int O_SLOTS = ((NARR * NA + 1) / 2) * 2;
  static constexpr int O_DESC = O_SLOTS + NSLOT * SLOT;        // x, dt, ik (as double), spare
  static constexpr int O_BARS = O_DESC + 4;                    // start, full[5]
  static constexpr int TOTAL = O_BARS + 8;
};

// Factorisation holder of factor_reg() writing into a slot column.
template <class TR>
struct SlotFactor {
  double* col;
  double M[4][4];
  double h, hk, hkap, vden, e4c, lo1, lo2;
  __device__ __forceinline__ double& ibv(int l) const { return col[l * TR::NCH]; }
  __device__ __forceinline__ double& beta(int row, int j) const { return col[(TR::MAXLEN + 4 * row + j) * TR::NCH]; }
};

// LU of the 4x4 border system with partial pivoting (registers, select-based row swaps: no divergent branch).
// P M = L U;  L: l10 l20 l21 l30 l31 l32;  U: u01 u02 u03 u12 u13 u23;  idg: 1/u_ii;  code: perm[i] in bits 2i..2i+1
// (row i of P M is row perm[i] of M).
__device__ __forceinline__ void lu4_pivot(const double (&Min)[4][4], double (&L)[6], double (&Uu)[6], double (&idg)[4], int& code) {
  double a[4][4]; int pr[4] = {0, 1, 2, 3};
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) a[i][j] = Min[i][j];
#pragma unroll
  for (int kx = 0; kx < 3; kx++) {
#pragma unroll
    for (int i = kx + 1; i < 4; i++) {
      const bool sw = fabs(a[i][kx]) > fabs(a[kx][kx]);
#pragma unroll
      for (int j = 0; j < 4; j++) { const double t = a[kx][j]; a[kx][j] = sw ? a[i][j] : t; a[i][j] = sw ? t : a[i][j]; }
      const int tp = pr[kx]; pr[kx] = sw ? pr[i] : tp; pr[i] = sw ? tp : pr[i];
    }
    const double ip = fast_rcp(a[kx][kx]);
    idg[kx] = ip;
#pragma unroll
    for (int i = kx + 1; i < 4; i++) {
      const double m = a[i][kx] * ip;
      a[i][kx] = m;
#pragma unroll
      for (int j = kx + 1; j < 4; j++) a[i][j] -= m * a[kx][j];
    }
  }
  idg[3] = fast_rcp(a[3][3]);
  L[0] = a[1][0]; L[1] = a[2][0]; L[2] = a[2][1]; L[3] = a[3][0]; L[4] = a[3][1]; L[5] = a[3][2];
  Uu[0] = a[0][1]; Uu[1] = a[0][2]; Uu[2] = a[0][3]; Uu[3] = a[1][2]; Uu[4] = a[1][3]; Uu[5] = a[2][3];
  code = pr[0] | (pr[1] << 2) | (pr[2] << 4) | (pr[3] << 6);
}

// One stage of the factoriser: everything of W_s = I - h A(x_s) the back-solve needs, into `slot`.
template <class TR>
__device__ __forceinline__ void factor_stage(const DevCosmo& c, const Lane& ln, const ModeConst& mc, double xs, double h, double* slot,
                                             long long* tstamp = nullptr) {
  typedef CtaLayout<TR> LY;
  BgS bf;
  eval_bg_fast(c, ln, mc, xs, bf);
  if (tstamp) tstamp[0] = clock64();
  SlotFactor<TR> f;
  f.col = slot + ln.base;
  factor_reg<TR>(ln, bf, h, f);
  f.col[LY::R_HK * LY::NCH] = f.hk; f.col[LY::R_WPSI * LY::NCH] = bf.wPsi; f.col[LY::R_WPHI * LY::NCH] = bf.wPhi;
  if (tstamp) tstamp[1] = clock64();
  double L[6], Uu[6], idg[4]; int code;
  lu4_pivot(f.M, L, Uu, idg, code);
  if (ln.lane == 0) {
    double* un = slot + LY::UNI;
#pragma unroll
    for (int i = 0; i < 6; i++) { un[LY::U_L + i] = L[i]; un[LY::U_U + i] = Uu[i]; }
#pragma unroll
    for (int i = 0; i < 4; i++) un[LY::U_ID + i] = idg[i];
    un[LY::U_PERM] = __longlong_as_double((long long)code);
    un[LY::U_H] = f.h; un[LY::U_HKAP] = f.hkap; un[LY::U_VDEN] = f.vden; un[LY::U_E4C] = f.e4c;
    un[LY::U_CPSI] = bf.cPsi; un[LY::U_K2] = bf.k2; un[LY::U_GPHI] = bf.gPhi; un[LY::U_OCA] = bf.Oc_a; un[LY::U_OBA] = bf.Ob_a;
    un[LY::U_CSB2] = bf.csb2;
    un[LY::U_RSA] = ((ln.k * bf.eta > 240.0) && (-bf.taup * bf.H > 100.0 * bf.eta)) ? 1.0 : 0.0;     // perturbations.jl:216
  }
}

// Solve W U = r with the factorisation in `slot`: the lane's chain in rr[] (registers, overwritten by U), the five scalars
// in r5[] (same copy on every lane).  Same algebra and operation order as solve_reg().
template <class TR, bool CB = true>
__device__ __forceinline__ void solve_slot(const Lane& ln, const double* __restrict__ slot, double (&rr)[TR::MAXLEN], double (&r5)[5]) {
  typedef CtaLayout<TR> LY;
  constexpr int MAXLEN = TR::MAXLEN, NCH = TR::NCH;
  const double* col = slot + ln.base;
  const double* un = slot + LY::UNI;
  const int kind = ln.kind;
  const double hkl = col[LY::R_HK * NCH];
  double ibn = 0.0, rn = 0.0;
#pragma unroll
  for (int l = MAXLEN - 1; l >= 3; l--) {
    const double up = TR::top(kind, l) ? 0.0 : hkl * rl1_of<CB>(l);
    const double v = rr[l] - (up * ibn) * rn;
    rr[l] = v; rn = v; ibn = col[l * NCH];
  }
  const double r2 = rr[2] - (hkl * rl1_of<CB>(2) * ibn) * rn;
  const double ib0 = col[0], ib1 = col[NCH], ib2 = col[2 * NCH];
  const double r1 = rr[1] - (hkl * rl1_of<CB>(1) * ib2) * r2;
  const double r0 = rr[0] - (hkl * ib1) * r1;
  const double lo1 = -hkl * rl_of<CB>(1), lo2 = -hkl * rl_of<CB>(2);
  const double a0 = r0 * ib0, a1 = (r1 - lo1 * a0) * ib1, a2 = (r2 - lo2 * a1) * ib2;
  const int lT = ln.nq, lP = ln.nq + 1;
  const double sPsi = warp_sum(col[LY::R_WPSI * NCH] * a2);
  const double sPhi = warp_sum(col[LY::R_WPHI * NCH] * a0);
  const double sPi = shfl_d(a2, lT) + shfl_d(a2 + a0, lP);
  const double t1 = shfl_d(a1, lT);
  const double rPhi = r5[0], rdel = r5[1], rv = r5[2], rdb = r5[3], rvb = r5[4];
  const double h = un[LY::U_H], hk = un[LY::U_HKAP], vden = un[LY::U_VDEN], e4c = un[LY::U_E4C];
  const double vc = rv * vden, dc = rdel + hk * vc;
  double rhs[4];
  rhs[0] = -(rPhi + un[LY::U_CPSI] * sPsi);
  rhs[1] = -(un[LY::U_K2] * rPhi - un[LY::U_GPHI] * (un[LY::U_OCA] * dc + un[LY::U_OBA] * rdb + sPhi));
  rhs[2] = sPi;
  rhs[3] = -(hk * un[LY::U_CSB2] * rdb + e4c * t1 - rvb);
  // P rhs, forward and backward substitution
  const int code = (int)__double_as_longlong(un[LY::U_PERM]);
  double t[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int pi = (code >> (2 * i)) & 3;
    double v = rhs[0]; v = (pi == 1) ? rhs[1] : v; v = (pi == 2) ? rhs[2] : v; v = (pi == 3) ? rhs[3] : v;
    t[i] = v;
  }
  t[1] -= un[LY::U_L + 0] * t[0];
  t[2] -= un[LY::U_L + 1] * t[0]; t[3] -= un[LY::U_L + 3] * t[0];
  t[2] -= un[LY::U_L + 2] * t[1]; t[3] -= un[LY::U_L + 4] * t[1];
  t[3] -= un[LY::U_L + 5] * t[2];
  double y[4];
  y[3] = t[3] * un[LY::U_ID + 3];
  y[2] = (t[2] - un[LY::U_U + 5] * y[3]) * un[LY::U_ID + 2];
  y[1] = (t[1] - un[LY::U_U + 3] * y[2] - un[LY::U_U + 4] * y[3]) * un[LY::U_ID + 1];
  y[0] = (t[0] - un[LY::U_U + 0] * y[1] - un[LY::U_U + 1] * y[2] - un[LY::U_U + 2] * y[3]) * un[LY::U_ID + 0];
  r5[0] = rPhi + h * y[0];
  const double v = vc - hk * vden * y[1];
  r5[1] = rdel + hk * v - 3.0 * h * y[0];
  r5[2] = v;
  r5[3] = rdb - 3.0 * h * y[0] + hk * y[3];
  r5[4] = y[3];
  double U0 = a0, U1 = a1, U2 = a2;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    U0 += col[(LY::R_BETA + j) * NCH] * y[j]; U1 += col[(LY::R_BETA + 4 + j) * NCH] * y[j]; U2 += col[(LY::R_BETA + 8 + j) * NCH] * y[j];
  }
  rr[0] = U0; rr[1] = U1; rr[2] = U2;
  double Up = U2;
#pragma unroll
  for (int l = 3; l < MAXLEN; l++) {
    const double lo = TR::top(kind, l) ? -hkl : -hkl * rl_of<CB>(l);
    const double U = (rr[l] - lo * Up) * col[l * NCH];
    rr[l] = U; Up = U;
  }
}

}  // namespace bolt
