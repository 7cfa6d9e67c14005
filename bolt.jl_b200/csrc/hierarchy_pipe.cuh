// K1 "pipeline": ONE CTA PER k-MODE, six warps by ROLE (sm_100a, FP64) -- the register-resident truncations of source_grid.
//
// Same mathematics, controller and arithmetic building blocks as hierarchy_kernel_t (hierarchy_kernel.cuh; boltsolve
// src/perturbations.jl:25-33, hierarchy! :161-271, source_grid's sampling loop src/spectra.jl:13-18).  What changes is WHO does
// what WHEN.  Measured on the one-warp-per-mode kernel and its first CTA successor (profiles/r2_k1_roles.md): a KenCarp4 step
// costs ~25k cycles of which only ~3k are the inherently serial part (five stage back-solves + the smoothed error estimate).
// The rest is work that does not depend on the stage values or depends on them only through already finished stages:
//
//   warp 0     SOLVER      the serial spine only: rhs_s = P_s + a_{s,s-1} z_{s-1} (one FMA per row, z_{s-1} still in registers),
//                          back-solve with a factorisation that is already in shared memory, z_s, error norm, controller
//   warps 1-3  FACTORISERS W_s = I - gamma dt A(x_n + c_s dt) depends on (x_n, dt) only.  All five stage matrices of a step are
//                          factored together, one (stage, chain) pair per THREAD (5 x 18 = 90 of 96 threads busy instead of 18 of
//                          32 lanes five times over), the border sums through shared memory.  They work ONE STEP AHEAD: the
//                          controller keeps dt unchanged on ~93 % of the steps, so the solver posts the request for the step
//                          after the current one (x + dt, dt) speculatively; on a hit the next step starts with its factorisation
//                          ready, on a miss (rejected step, dt changed) the real request is posted and awaited.
//   warp 4     HELPER      P_{s+1} = u_n + sum_{j<s} a_{s+1,j} z_j as soon as z_{s-1} is published (flat over the state, off the
//                          spine), and the z_1..z_5 part of the error-estimate combination
//   warp 5     SAMPLER     dense output + source functions on the x grid for the PREVIOUS accepted step while the solver is
//                          already integrating the next one
//
// Hand-off: mbarriers in shared memory (arrive.release / try_wait.acquire), every one completing exactly once per step attempt
// (stage barriers), per request (factorisation ring) or per accepted step (sampler), so each party tracks phases by counting.
// A bounded spin turns a protocol error into a trap instead of a hung GPU.
#pragma once
#include "hierarchy_kernel.cuh"

namespace bolt {

__device__ __forceinline__ uint32_t pb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pb_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pb_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void pb_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(pb_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool pb_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(pb_smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// wait for the phase with the given parity to complete; a protocol error traps (kernel fails) instead of hanging the GPU
__device__ __forceinline__ void pb_wait(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
  for (int it = 0; it < (1 << 26); it++)
    if (pb_try(bar, parity)) return;
  __trap();
}
__device__ __forceinline__ void pb_named_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// LU of the 4x4 border system with partial pivoting (registers, select-based row swaps: no divergent branch).
// P M = L U;  L: l10 l20 l21 l30 l31 l32;  U: u01 u02 u03 u12 u13 u23;  idg: 1/u_ii;  code: perm[i] in bits 2i..2i+1.
__device__ __forceinline__ void pb_lu4(const double (&Min)[4][4], double (&L)[6], double (&Uu)[6], double (&idg)[4], int& code) {
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

constexpr int PIPE_THREADS = 192;     // solver | 3 factorisers | helper | sampler
constexpr int PIPE_NFACT = 96;

// Shared-memory layout of one CTA (offsets in doubles).
template <class TR>
struct PipeLayout {
  static constexpr int NCH = TR::NCH, MAXLEN = TR::MAXLEN;
  static constexpr int NA = MAXLEN * NCH + 8;        // one state array: interleaved [l][chain] + 5 scalars (as hierarchy_kernel_t)
  static constexpr int NFLAT = MAXLEN * NCH + 5;
  // state arrays: u / u_{n+1} (flip), z1|z6 of the previous/this step (flip), z2..z5, two partial-sum buffers
  static constexpr int A_UA = 0, A_UB = 1, A_ZA = 2, A_ZB = 3, A_Z1 = 4, A_Z2 = 5, A_Z3 = 6, A_Z4 = 7, A_P0 = 8, A_P1 = 9, NARR = 10;     // P1 doubles as the error-scale buffer
  // factor slot of one stage: per-chain rows (column = the chain's column): multipliers of the downward sweep m_l = up_l/d_{l+1},
  // inverse pivots 1/d_l, multipliers of the upward sweep g_l = lo_l/d_l, the beta vectors of rows 0..2, the lane weights ...
  static constexpr int R_M = 0, R_IB = MAXLEN, R_G = 2 * MAXLEN, R_BETA = 3 * MAXLEN, R_WPSI = 3 * MAXLEN + 12, R_WPHI = 3 * MAXLEN + 13, LROWS = 3 * MAXLEN + 14;
  // ... and a uniform block: the inverse of the 4x4 border matrix + the stage scalars the back-solve needs
  static constexpr int UNI = LROWS * NCH;
  static constexpr int U_MINV = 0, U_H = 16, U_HKAP = 17, U_VDEN = 18, U_E4C = 19, U_CPSI = 20, U_K2 = 21, U_GPHI = 22, U_OCA = 23, U_OBA = 24,
                       U_CSB2 = 25, U_RSA = 26, NUNI = 28;
  static constexpr int SLOT = ((UNI + NUNI + 1) / 2) * 2;
  static constexpr int NSLOT = 10;                   // 2 sets x 5 stages
  // factorisers' scratch per stage: 18 chains x 4 border-sum terms, the Theta chain's beta vectors + weights, the ThetaP term,
  // the stage scalars, the four table values
  static constexpr int C_SUM = 0, C_TV = 4 * (TR::NQ + 3), C_PV = C_TV + 14, C_UNI = C_TV + 16, C_BG = C_UNI + 12, CSTAGE = ((C_BG + 4 + 1) / 2) * 2;
  static constexpr int O_SLOTS = ((NARR * NA + 1) / 2) * 2;
  static constexpr int O_CON = O_SLOTS + NSLOT * SLOT;
  static constexpr int O_REQD = O_CON + 5 * CSTAGE;  // request ring: 4 x (x, dt, ik, spare)
  static constexpr int O_STEPD = O_REQD + 16;        // step descriptor for the helper: s1, flipU, flipZ, ik
  static constexpr int O_JOBD = O_STEPD + 4;         // sampler job: x, dt, s1, xn1, last, flipU, flipZ, ik
  static constexpr int O_FLAGS = O_JOBD + 8;         // [0] rsa seen by the sampler
  static constexpr int O_BARS = O_FLAGS + 2;
  // barriers
  static constexpr int B_REQ = 0, B_FULL = 4, B_STEP = 14, B_ZP = 15 /* +s, s = 1..4 */, B_PR = 20 /* +s, s = 2..5 */, B_ER = 26, B_SJOB = 27, B_SDONE = 28, B_SC = 29, NBARS = 30;      // B_ZP + 5 = 20: z6 and u_{n+1} published
  static constexpr int TOTAL = O_BARS + NBARS + 1;
};

// Back-solve W U = r with the factorisation in `slot`: the lane's chain in rr[] (registers, overwritten by U), the five scalars
// in r5[] (same copy on every lane).  Algebra of solve_reg(), arranged for the shortest dependency chain: one FMA per row in
// either sweep (the factorisers store the multipliers), the border system through its explicit inverse (4 independent FMA chains
// of depth 4 instead of a pivoted forward/backward substitution).
#ifdef K1P_PROF
#define PSOLVE_TS(i) if (ts) ts[i] = clock64()
#else
#define PSOLVE_TS(i)
#endif
template <class TR>
__device__ __forceinline__ void pipe_solve(const Lane& ln, const double* __restrict__ slot, double (&rr)[TR::MAXLEN], double (&r5)[5],
                                           long long* ts = nullptr) {
  typedef PipeLayout<TR> LY;
  constexpr int MAXLEN = TR::MAXLEN, NCH = TR::NCH;
  const double* col = slot + ln.base;
  const double* un = slot + LY::UNI;
  double w[MAXLEN];                         // w_l = v_l / d_l
  double v = rr[MAXLEN - 1];
  w[MAXLEN - 1] = __dmul_rn(v, col[(LY::R_IB + MAXLEN - 1) * NCH]);
#pragma unroll
  for (int l = MAXLEN - 2; l >= 0; l--) {
    v = __fma_rn(-col[(LY::R_M + l) * NCH], v, rr[l]);
    w[l] = __dmul_rn(v, col[(LY::R_IB + l) * NCH]);
  }
  PSOLVE_TS(0);
  const double g1 = col[(LY::R_G + 1) * NCH], g2 = col[(LY::R_G + 2) * NCH];
  const double a0 = w[0], a1 = __fma_rn(-g1, a0, w[1]), a2 = __fma_rn(-g2, a1, w[2]);
  const int lT = ln.nq, lP = ln.nq + 1;
  const double sPhi = warp_sum(col[LY::R_WPHI * NCH] * a0);
  const double sPsi = warp_sum(col[LY::R_WPSI * NCH] * a2);
  const double sPi = shfl_d(a2, lT) + shfl_d(a2 + a0, lP);
  const double t1 = shfl_d(a1, lT);
  PSOLVE_TS(1);
  const double rPhi = r5[0], rdel = r5[1], rv = r5[2], rdb = r5[3], rvb = r5[4];
  const double h = un[LY::U_H], hk = un[LY::U_HKAP], vden = un[LY::U_VDEN], e4c = un[LY::U_E4C];
  const double vc = rv * vden, dc = rdel + hk * vc;
  double rhs[4];
  rhs[0] = -(rPhi + un[LY::U_CPSI] * sPsi);
  rhs[1] = -(un[LY::U_K2] * rPhi - un[LY::U_GPHI] * (un[LY::U_OCA] * dc + un[LY::U_OBA] * rdb + sPhi));
  rhs[2] = sPi;
  rhs[3] = -(hk * un[LY::U_CSB2] * rdb + e4c * t1 - rvb);
  double y[4];
#pragma unroll
  for (int i = 0; i < 4; i++)
    y[i] = (un[LY::U_MINV + 4 * i] * rhs[0] + un[LY::U_MINV + 4 * i + 1] * rhs[1]) + (un[LY::U_MINV + 4 * i + 2] * rhs[2] + un[LY::U_MINV + 4 * i + 3] * rhs[3]);
  PSOLVE_TS(2);
  r5[0] = rPhi + h * y[0];
  const double vv = vc - hk * vden * y[1];
  r5[1] = rdel + hk * vv - 3.0 * h * y[0];
  r5[2] = vv;
  r5[3] = rdb - 3.0 * h * y[0] + hk * y[3];
  r5[4] = y[3];
  double U0 = (a0 + col[(LY::R_BETA + 0) * NCH] * y[0] + col[(LY::R_BETA + 1) * NCH] * y[1]) + (col[(LY::R_BETA + 2) * NCH] * y[2] + col[(LY::R_BETA + 3) * NCH] * y[3]);
  double U1 = (a1 + col[(LY::R_BETA + 4) * NCH] * y[0] + col[(LY::R_BETA + 5) * NCH] * y[1]) + (col[(LY::R_BETA + 6) * NCH] * y[2] + col[(LY::R_BETA + 7) * NCH] * y[3]);
  double U2 = (a2 + col[(LY::R_BETA + 8) * NCH] * y[0] + col[(LY::R_BETA + 9) * NCH] * y[1]) + (col[(LY::R_BETA + 10) * NCH] * y[2] + col[(LY::R_BETA + 11) * NCH] * y[3]);
  rr[0] = U0; rr[1] = U1; rr[2] = U2;
  PSOLVE_TS(3);
  double Up = U2;
#pragma unroll
  for (int l = 3; l < MAXLEN; l++) {
    const double U = __fma_rn(-col[(LY::R_G + l) * NCH], Up, w[l]);
    rr[l] = U; Up = U;
  }
  PSOLVE_TS(4);
}

// Chain part of the factorisation of W = I - h A(x_s) for ONE chain (thread): inverse pivots and beta vectors into the chain's
// slot column, its terms of the border sums into the stage scratch `con`.  Algebra of factor_reg().
template <class TR, bool CB = true>
__device__ __forceinline__ void pipe_factor_chain(const Lane& ln, const BgS& b, double h, double* __restrict__ col, double* __restrict__ con, bool rsa) {
  typedef PipeLayout<TR> LY;
  constexpr int MAXLEN = TR::MAXLEN, NCH = TR::NCH;
  const int kind = ln.kind;
  const bool photon = (kind == CH_T || kind == CH_P);
  const double hk = h * b.kappa * b.qe, hkap = h * b.kappa;
  const double dtau = photon ? -h * b.taup : 0.0;
  const double btr = 1.0 + h * (double)ln.len * b.iHeta + dtau;
  double ibn = 0.0, lo_next = 0.0;
#pragma unroll
  for (int l = MAXLEN - 1; l >= 3; l--) {
    const bool act = TR::act(kind, l), top = TR::top(kind, l);
    const double bd = top ? btr : 1.0 + dtau;
    const double up = top ? 0.0 : hk * rl1_of<CB>(l);
    const double lo = top ? -hk : -hk * rl_of<CB>(l);
    const double ml = up * ibn;
    const double rc = fast_rcp(bd - ml * lo_next);
    const double ibl = act ? rc : 0.0;
    col[(LY::R_M + l) * NCH] = ml; col[(LY::R_IB + l) * NCH] = ibl; col[(LY::R_G + l) * NCH] = lo * ibl;
    ibn = ibl; lo_next = act ? lo : 0.0;
  }
  const double up2 = hk * rl1_of<CB>(2), up1 = hk * rl1_of<CB>(1), up0 = hk;
  const double lo2 = -hk * rl_of<CB>(2), lo1 = -hk * rl_of<CB>(1);
  const double m2 = up2 * ibn;
  const double ib2 = fast_rcp((1.0 + dtau) - m2 * lo_next);
  const double m1 = up1 * ib2;
  const double ib1 = fast_rcp((1.0 + dtau) - m1 * lo2);
  const double m0 = up0 * ib1;
  const double ib0 = fast_rcp((1.0 + (kind == CH_P ? dtau : 0.0)) - m0 * lo1);
  col[(LY::R_M + 2) * NCH] = m2; col[(LY::R_M + 1) * NCH] = m1; col[LY::R_M * NCH] = m0;
  col[(LY::R_IB + 2) * NCH] = ib2; col[(LY::R_IB + 1) * NCH] = ib1; col[LY::R_IB * NCH] = ib0;
  col[(LY::R_G + 2) * NCH] = lo2 * ib2; col[(LY::R_G + 1) * NCH] = lo1 * ib1;
  double C0[4] = {0, 0, 0, 0}, C1[4] = {0, 0, 0, 0}, C2[4] = {0, 0, 0, 0};
  {
    const bool isM = kind == CH_M, isT = kind == CH_T, isP = kind == CH_P, isTN = isT || kind == CH_N;
    const double htp = h * b.taup;
    C0[0] = isM ? h * ln.df0 : (isTN ? -h : 0.0);
    C1[1] = isM ? -hkap * (1.0 / 3.0) * b.eq * ln.df0 : (isTN ? hkap * (1.0 / 3.0) : 0.0);
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
    col[(LY::R_BETA + j) * NCH] = be0[j]; col[(LY::R_BETA + 4 + j) * NCH] = be1[j]; col[(LY::R_BETA + 8 + j) * NCH] = be2[j];
  }
  col[LY::R_WPSI * NCH] = b.wPsi; col[LY::R_WPHI * NCH] = b.wPhi;
  // terms of the border sums: components (Phi', Psi) of the Psi / Phi' sums come from every chain ...
  double* cc = con + LY::C_SUM + 4 * ln.lane;
  cc[0] = b.wPsi * be2[0]; cc[1] = b.wPsi * be2[1]; cc[2] = b.wPhi * be0[0]; cc[3] = b.wPhi * be0[1];
  // ... components (Pi, v_b) are non-zero on the Theta chain only (and Pi on ThetaP)
  if (kind == CH_T) {
    double* tv = con + LY::C_TV;
#pragma unroll
    for (int j = 0; j < 4; j++) { tv[j] = be0[j]; tv[4 + j] = be1[j]; tv[8 + j] = be2[j]; }
    tv[12] = b.wPsi; tv[13] = b.wPhi;
    double* un = con + LY::C_UNI;
    un[0] = h; un[1] = hkap; un[2] = fast_rcp(1.0 + h); un[3] = b.cPsi; un[4] = b.k2; un[5] = b.gPhi; un[6] = b.Oc_a; un[7] = b.Ob_a;
    un[8] = b.csb2; un[9] = -3.0 * h * b.taup * b.R; un[10] = h * b.taup * b.R; un[11] = rsa ? 1.0 : 0.0;
  }
  if (kind == CH_P) con[LY::C_PV] = be2[2] + be0[2];
}

// Border system of one stage from the scratch sums S = (sum wPsi be2[0], sum wPsi be2[1], sum wPhi be0[0], sum wPhi be0[1]):
// assemble M (factor_reg), pivoted LU, stage scalars -> the slot's uniform block.
template <class TR>
__device__ __forceinline__ void pipe_border(const double (&S)[4], const double* __restrict__ con, double* __restrict__ unio) {
  typedef PipeLayout<TR> LY;
  const double* tv = con + LY::C_TV;
  const double* un = con + LY::C_UNI;
  const double h = un[0], hk = un[1], vden = un[2], cPsi = un[3], k2 = un[4], gPhi = un[5], Oc = un[6], Ob = un[7], csb2 = un[8], e4c = un[9], htR = un[10];
  double sPsi[4], sPhi[4], sPi[4], t1[4];
  sPsi[0] = S[0]; sPsi[1] = S[1]; sPhi[0] = S[2]; sPhi[1] = S[3];
  sPsi[2] = tv[12] * tv[8 + 2]; sPsi[3] = tv[12] * tv[8 + 3]; sPhi[2] = tv[13] * tv[2]; sPhi[3] = tv[13] * tv[3];
#pragma unroll
  for (int j = 0; j < 4; j++) { sPi[j] = tv[8 + j]; t1[j] = tv[4 + j]; }
  sPi[2] += con[LY::C_PV];
  const double dPhi_y[4] = {h, 0, 0, 0};
  const double dDel_y[4] = {-3.0 * h, -hk * hk * vden, 0, 0};
  const double dDb_y[4] = {-3.0 * h, 0, 0, hk};
  double M[4][4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    M[0][j] = dPhi_y[j] + cPsi * sPsi[j] + (j == 1 ? 1.0 : 0.0);
    M[1][j] = (j == 0 ? 1.0 : 0.0) - (j == 1 ? 1.0 : 0.0) + k2 * dPhi_y[j] - gPhi * (Oc * dDel_y[j] + Ob * dDb_y[j] + sPhi[j]);
    M[2][j] = (j == 2 ? 1.0 : 0.0) - sPi[j];
    M[3][j] = (j == 3 ? (1.0 + h - htR) : 0.0) + hk * ((j == 1 ? 1.0 : 0.0) + csb2 * dDb_y[j]) + e4c * t1[j];
  }
  // explicit inverse through the pivoted LU: column j of M^-1 solves M y = e_j
  double L[6], Uu[6], idg[4]; int code;
  pb_lu4(M, L, Uu, idg, code);
#pragma unroll
  for (int j = 0; j < 4; j++) {
    double t[4];
#pragma unroll
    for (int i = 0; i < 4; i++) t[i] = (((code >> (2 * i)) & 3) == j) ? 1.0 : 0.0;      // P e_j
    t[1] -= L[0] * t[0];
    t[2] -= L[1] * t[0]; t[3] -= L[3] * t[0];
    t[2] -= L[2] * t[1]; t[3] -= L[4] * t[1];
    t[3] -= L[5] * t[2];
    const double y3 = t[3] * idg[3];
    const double y2 = (t[2] - Uu[5] * y3) * idg[2];
    const double y1 = (t[1] - Uu[3] * y2 - Uu[4] * y3) * idg[1];
    const double y0 = (t[0] - Uu[0] * y1 - Uu[1] * y2 - Uu[2] * y3) * idg[0];
    unio[LY::U_MINV + j] = y0; unio[LY::U_MINV + 4 + j] = y1; unio[LY::U_MINV + 8 + j] = y2; unio[LY::U_MINV + 12 + j] = y3;
  }
  unio[LY::U_H] = h; unio[LY::U_HKAP] = hk; unio[LY::U_VDEN] = vden; unio[LY::U_E4C] = e4c;
  unio[LY::U_CPSI] = cPsi; unio[LY::U_K2] = k2; unio[LY::U_GPHI] = gPhi; unio[LY::U_OCA] = Oc; unio[LY::U_OBA] = Ob;
  unio[LY::U_CSB2] = csb2; unio[LY::U_RSA] = un[11];
}

// Development profile (-DK1P_PROF): cycle counters of the solver while it works on the FIRST work item (the largest k), dumped
// through the step-log buffer (BOLT_DEBUG_STEPS=<file>): row = category, cycles, count.
#ifdef K1P_PROF
#define PPROF_DECL long long prof[16] = {0}; long long pcnt[16] = {0}
#define PPROF_T(v) const long long v = clock64()
#define PPROF_ACC(i, t0) do { prof[i] += clock64() - (t0); pcnt[i]++; } while (0)
#define PPROF_DUMP(base, n_) do { if (p.dbg && (threadIdx.x & 31) == 0) for (int i_ = 0; i_ < (n_); i_++) { double* d_ = p.dbg + 4 * ((base) + i_); d_[0] = (base) + i_; d_[1] = (double)prof[i_] + 1e-9; d_[2] = (double)pcnt[i_]; d_[3] = 0; } } while (0)
#else
#define PPROF_DECL
#define PPROF_T(v)
#define PPROF_ACC(i, t0)
#define PPROF_DUMP(base, n_)
#endif

#ifndef K1P_MINBLOCKS
#define K1P_MINBLOCKS 2
#endif

template <class TR>
__global__ void __launch_bounds__(PIPE_THREADS, K1P_MINBLOCKS) hierarchy_pipe_kernel(SolveParams p) {
  typedef PipeLayout<TR> LY;
  static_assert(TR::MAXLEN > 0 && TR::NCH == TR::NQ + 4, "register-resident truncations with the compact layout");
  extern __shared__ double sm[];
  constexpr int NCH = TR::NCH, MAXLEN = TR::MAXLEN, na = LY::NA, NFLAT = LY::NFLAT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* const slots = sm + LY::O_SLOTS;
  double* const reqd = sm + LY::O_REQD;
  double* const stepd = sm + LY::O_STEPD;
  double* const jobd = sm + LY::O_JOBD;
  double* const flags = sm + LY::O_FLAGS;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(sm + LY::O_BARS);
  for (int i = threadIdx.x; i < LY::O_BARS; i += PIPE_THREADS) sm[i] = 0.0;       // state, slots (the idle lanes' zero column), scratch
  if (threadIdx.x == 0) {
#pragma unroll 1
    for (int i = 0; i < LY::NBARS; i++) pb_init(&bars[i], 1);
    // which resident CTA of this SM are we?  (p.counter[1 + smid], zeroed with the work queue)
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    flags[1] = (double)atomicAdd(p.counter + 1 + (smid & 255), 1);
  }
  __syncthreads();
  // Roles by warp.  Warps go to the four schedulers of the SM round-robin, so warps 2 and 3 are alone on theirs within a CTA:
  // the solver (the serial spine) takes one of them, alternating between the CTAs that share an SM, the light helper the other;
  // factorisers on warps 0, 1, 4, the sampler on warp 5.
  const bool odd = ((int)flags[1] & 1) != 0;
  const int w_solver = odd ? 3 : 2, w_helper = odd ? 2 : 3, w_sampler = 5;

  if (warp == 0 || warp == 1 || warp == 4) {
    // ------------------------------------------------ factorisers ------------------------------------------------
    const int fw = (warp == 4) ? 2 : warp;
    const int tf = fw * 32 + lane;                   // 0..95
    const int nchains = TR::NQ + 3;                  // 18
    const bool valid = tf < 5 * nchains;
    const int st = valid ? tf / nchains : 4, ch = valid ? tf % nchains : nchains;     // invalid threads: idle "chain" of stage 5
    double* const con_all = sm + LY::O_CON;
    int cur_ik = -1;
    Lane ln; ModeConst mc;
    const DevCosmo* cp = nullptr;
    ln.kind = CH_IDLE; ln.len = 0;
#pragma unroll 1
    for (uint32_t i = 0;; i++) {
      pb_wait(&bars[LY::B_REQ + (i & 3)], (i >> 2) & 1);
      const double* d = reqd + 4 * (i & 3);
      const double x = d[0], dt = d[1];
      const int ik = __double2int_rn(d[2]);
      if (ik < 0) break;
      const int set = i & 1;
      if (dt != 0.0) {
        if (ik != cur_ik) {
          cur_ik = ik; cp = p.cos_list[ik / p.nk_per];
          lane_setup<TR>(*cp, p, ln, ch);
          ln.k = p.k[ik];
          mode_const(*cp, ln, mc);
        }
        const DevCosmo& c = *cp;
        // phase 1: the four background tables at the five stage abscissae, one (stage, table) per thread
        if (tf < 20) {
          const int s5 = tf >> 2, wl = tf & 3;
          const int tab = 3 * wl + ((wl == 3) ? 2 : 0);
          con_all[s5 * LY::CSTAGE + LY::C_BG + wl] = spline_eval(c.tab[tab], c.n_x, c.x0, c.dx, x + KC_C[s5 + 1] * dt);
        }
        pb_named_sync(1, PIPE_NFACT);
        // phase 2: one (stage, chain) per thread
        {
          double* con = con_all + st * LY::CSTAGE;
          BgS b;
          b.H = con[LY::C_BG + 0]; b.eta = con[LY::C_BG + 1]; b.taup = con[LY::C_BG + 2]; b.csb2 = con[LY::C_BG + 3];
          bg_from_tables(ln, mc, x + KC_C[st + 1] * dt, b);
          const bool rsa = (ln.k * b.eta > 240.0) && (-b.taup * b.H > 100.0 * b.eta);       // perturbations.jl:216
          if (valid) pipe_factor_chain<TR>(ln, b, KC_GAMMA * dt, slots + (size_t)(set * 5 + st) * LY::SLOT + ch, con, rsa);
        }
        pb_named_sync(1, PIPE_NFACT);
        // phase 3 (first factoriser warp): border sums, one (stage, term) per lane; then one lane per stage assembles and factors
        if (fw == 0) {
          double v = 0.0;
          if (lane < 20) {
            const double* cs = con_all + (lane >> 2) * LY::CSTAGE + LY::C_SUM + (lane & 3);
#pragma unroll
            for (int q = 0; q < nchains; q++) v += cs[4 * q];
          }
          double S[4];
#pragma unroll
          for (int w = 0; w < 4; w++) S[w] = shfl_d(v, (lane & ~3) + w);
          if (lane < 20 && (lane & 3) == 0) {
            const int s5 = lane >> 2;
            pipe_border<TR>(S, con_all + s5 * LY::CSTAGE, slots + (size_t)(set * 5 + s5) * LY::SLOT + LY::UNI);
            pb_arrive(&bars[LY::B_FULL + set * 5 + s5]);
          }
        }
      } else if (fw == 0 && lane < 5) {
        pb_arrive(&bars[LY::B_FULL + set * 5 + lane]);       // placeholder request: nothing to factor, keep the phases in step
      }
    }
    return;
  }

  if (warp == w_helper) {
    // -------------------------------------------------- helper --------------------------------------------------
#pragma unroll 1
    for (uint32_t att = 0;; att++) {
      const uint32_t par = att & 1;
      pb_wait(&bars[LY::B_STEP], par);
      const double s1 = stepd[0];
      const bool flipU = stepd[1] != 0.0, flipZ = stepd[2] != 0.0;
      if (stepd[3] < 0.0) break;
      const double* U = sm + (size_t)(flipU ? LY::A_UB : LY::A_UA) * na;
      const double* Z0 = sm + (size_t)(flipZ ? LY::A_ZB : LY::A_ZA) * na;
      const double* Z1 = sm + (size_t)LY::A_Z1 * na; const double* Z2 = sm + (size_t)LY::A_Z2 * na;
      const double* Z3 = sm + (size_t)LY::A_Z3 * na; const double* Z4 = sm + (size_t)LY::A_Z4 * na;
      double* P0 = sm + (size_t)LY::A_P0 * na; double* P1 = sm + (size_t)LY::A_P1 * na;
      constexpr int NT = (NFLAT + 31) / 32;
      {
        const double a0 = KC_A[2][0] * s1;
#pragma unroll
        for (int t = 0; t < NT; t++) { const int i = lane + 32 * t; if (i < NFLAT) P0[i] = U[i] + a0 * Z0[i]; }
        __syncwarp(); if (lane == 0) pb_arrive(&bars[LY::B_PR + 2]);
      }
      pb_wait(&bars[LY::B_ZP + 1], par);
      {
        const double a0 = KC_A[3][0] * s1, a1 = KC_A[3][1];
#pragma unroll
        for (int t = 0; t < NT; t++) { const int i = lane + 32 * t; if (i < NFLAT) P1[i] = U[i] + a0 * Z0[i] + a1 * Z1[i]; }
        __syncwarp(); if (lane == 0) pb_arrive(&bars[LY::B_PR + 3]);
      }
      pb_wait(&bars[LY::B_ZP + 2], par);
      {
        const double a0 = KC_A[4][0] * s1, a1 = KC_A[4][1], a2 = KC_A[4][2];
#pragma unroll
        for (int t = 0; t < NT; t++) { const int i = lane + 32 * t; if (i < NFLAT) P0[i] = U[i] + a0 * Z0[i] + a1 * Z1[i] + a2 * Z2[i]; }
        __syncwarp(); if (lane == 0) pb_arrive(&bars[LY::B_PR + 4]);
      }
      pb_wait(&bars[LY::B_ZP + 3], par);
      {
        const double a0 = KC_A[5][0] * s1, a1 = KC_A[5][1], a2 = KC_A[5][2], a3 = KC_A[5][3];
#pragma unroll
        for (int t = 0; t < NT; t++) { const int i = lane + 32 * t; if (i < NFLAT) P1[i] = U[i] + a0 * Z0[i] + a1 * Z1[i] + a2 * Z2[i] + a3 * Z3[i]; }
        __syncwarp(); if (lane == 0) pb_arrive(&bars[LY::B_PR + 5]);
      }
      pb_wait(&bars[LY::B_ZP + 4], par);
      {
        // error-estimate combination without its z6 term (KC_E[1] = 0); the solver adds KC_E[5] z6 from registers
        const double e0 = KC_E[0] * s1;
#pragma unroll
        for (int t = 0; t < NT; t++) { const int i = lane + 32 * t; if (i < NFLAT) P0[i] = e0 * Z0[i] + KC_E[2] * Z2[i] + KC_E[3] * Z3[i] + KC_E[4] * Z4[i]; }
        __syncwarp(); if (lane == 0) pb_arrive(&bars[LY::B_ER]);
      }
      pb_wait(&bars[LY::B_ZP + 5], par);
      {
        // inverse error scales 1 / (abstol + reltol max(|u_n|, |u_{n+1}|)) for the solver's error norm (P5 in P1 is consumed)
        const double* UN = sm + (size_t)(flipU ? LY::A_UA : LY::A_UB) * na;
        const double reltol = p.reltol, abstol = p.abstol;
#pragma unroll
        for (int t = 0; t < NT; t++) { const int i = lane + 32 * t; if (i < NFLAT) P1[i] = fast_rcp(abstol + reltol * fmax(fabs(U[i]), fabs(UN[i]))); }
        __syncwarp(); if (lane == 0) pb_arrive(&bars[LY::B_SC]);
      }
    }
    return;
  }

  if (warp == w_sampler) {
    // -------------------------------------------------- sampler --------------------------------------------------
    Lane ln; ModeConst mc;
    int cur_ik = -1, ix = 0;
    const DevCosmo* cp = nullptr;
#pragma unroll 1
    for (uint32_t job = 0;; job++) {
      pb_wait(&bars[LY::B_SJOB], job & 1);
      const double x = jobd[0], dt = jobd[1], s1 = jobd[2], xn1 = jobd[3];
      const bool last = jobd[4] != 0.0, flipU = jobd[5] != 0.0, flipZ = jobd[6] != 0.0;
      const int ik = __double2int_rn(jobd[7]);
      if (ik < 0) break;
      if (ik != cur_ik) {
        cur_ik = ik; cp = p.cos_list[ik / p.nk_per]; ix = 0;
        lane_setup<TR>(*cp, p, ln);
        ln.k = p.k[ik];
        mode_const(*cp, ln, mc);
      }
      const DevCosmo& c = *cp;
      const double* U = sm + (size_t)(flipU ? LY::A_UB : LY::A_UA) * na;
      const double* UN = sm + (size_t)(flipU ? LY::A_UA : LY::A_UB) * na;
      const double* Z0 = sm + (size_t)(flipZ ? LY::A_ZB : LY::A_ZA) * na;
      const double* Z5 = sm + (size_t)(flipZ ? LY::A_ZA : LY::A_ZB) * na;
      bool rsa_flag = false;
      while (ix < c.n_x) {
        const double xs = c.x0 + c.dx * ix;
        if (!last && xs > xn1 + 1e-12) break;
        if (ix >= p.ix_first) {
          double th = (xs - x) / dt; if (th > 1.0) th = 1.0;
          Hermite hm = hermite_weights(th);
          sample_sources(c, ln, p, ik, ix, xs, hm, U, UN, Z0, s1, Z5, rsa_flag, &mc);
        }
        ix++;
      }
      __syncwarp();
      if (lane == 0) { if (rsa_flag) flags[0] = 1.0; pb_arrive(&bars[LY::B_SDONE]); }
    }
    return;
  }

  // ---------------------------------------------------- solver ----------------------------------------------------
  if (warp != w_solver) return;
  const int n = p.n;
  Lane ln;
  const bool fixed = (p.mode == BOLT_MODE_FIXED);
  const double reltol = p.reltol, abstol = p.abstol;
  uint32_t reqi = 0;          // factorisation requests posted so far
  uint32_t att = 0;           // step attempts so far (phase of the stage barriers)
  uint32_t jobs = 0, jobs_waited = 0;     // sampler jobs posted / known finished
  bool spec_ok = false; double spec_x = 0.0, spec_dt = 0.0; int spec_ik = -1;
  PPROF_DECL;
  auto post_req = [&](double x, double dt, int ik) {
    const uint32_t i = reqi++;
    if (lane == 0) { double* d = reqd + 4 * (i & 3); d[0] = x; d[1] = dt; d[2] = (double)ik; pb_arrive(&bars[LY::B_REQ + (i & 3)]); }
  };

  while (true) {
    int w = 0;
    if (lane == 0) w = atomicAdd(p.counter, 1);
    w = __shfl_sync(FULL, w, 0);
    if (w >= p.nk) break;
    const int ik = p.order[w];
    const DevCosmo& c = *p.cos_list[ik / p.nk_per];
    lane_setup<TR>(c, p, ln);
    ln.k = p.k[ik];
    const double x_begin = c.x0, x_end = 0.0;

    bool flipU = false, flipZ = false;
    double* const Z1 = sm + (size_t)LY::A_Z1 * na;
    double* const Z2 = sm + (size_t)LY::A_Z2 * na;
    double* const Z3 = sm + (size_t)LY::A_Z3 * na;
    double* const Z4 = sm + (size_t)LY::A_Z4 * na;
    double* const P0 = sm + (size_t)LY::A_P0 * na;
    double* const P1 = sm + (size_t)LY::A_P1 * na;
#define PSLOT_U  (sm + (size_t)(flipU ? LY::A_UB : LY::A_UA) * na)
#define PSLOT_UN (sm + (size_t)(flipU ? LY::A_UA : LY::A_UB) * na)
#define PSLOT_Z0 (sm + (size_t)(flipZ ? LY::A_ZB : LY::A_ZA) * na)
#define PSLOT_Z5 (sm + (size_t)(flipZ ? LY::A_ZA : LY::A_ZB) * na)
    double* U = PSLOT_U; double* UN = PSLOT_UN; double* Z0 = PSLOT_Z0; double* Z5 = PSLOT_Z5;
    // every other role is parked on a barrier (the sampler's last job of the previous mode was awaited): the state is ours
    for (int i = lane; i < LY::NARR * na; i += 32) sm[i] = 0.0;
    if (lane == 0) flags[0] = 0.0;
    __syncwarp();

    Bg b;
    eval_bg(c, ln, x_begin, b);
    initial_conditions(c, ln, b, U);
    rhs_full(c, ln, b, U, Z5);          // f(u0) in the z6 slot
    bool rsa_flag = (ln.k * b.eta > 240.0) && (-b.taup * b.H / b.eta > 100.0);

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
      // initial step, Hairer-Wanner as in OrdinaryDiffEq's ode_determine_initdt (same as hierarchy_kernel_t and the oracle)
      double* r = Z2;
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
      // the scratch values must not survive in the z slots' padded entries
      for (int i = lane; i < na; i += 32) { Z0[i] = 0.0; r[i] = 0.0; }
      __syncwarp();
    }
    flipZ = !flipZ; Z0 = PSLOT_Z0; Z5 = PSLOT_Z5;      // z1 slot <- f(u0): true z1 = s1 * Z0 with s1 = dt
    double s1 = dt;

    const double beta1 = 7.0 / 40.0, beta2 = 2.0 / 20.0, safety = 0.9, qmin = 0.2, qmax = 10.0;
    double lqold = -9.210340371976182;      // log(qold), qold = 1e-4
    const double inv_n = 1.0 / (double)n;
    const long long fixed_total = fixed ? llround((x_end - x_begin) / p.fixed_dt) : 0;
    long long fixed_left = fixed_total;
    const long long max_steps = p.max_steps > 0 ? p.max_steps : 1000000;
    const int lo_ = ln.base;

    while (true) {
      bool clamped = false;
      if (fixed) { if (fixed_left == 0) break; }
      else {
        if (x >= x_end) break;
        if (x + dt >= x_end) { const double dtn = x_end - x; s1 *= dtn / dt; dt = dtn; clamped = true; }
      }
      if (nsteps + nreject >= max_steps) { status = BOLT_K_MAXSTEPS; break; }

      // ---- factorisation of this step: the speculative request if it was right, else a fresh one; then speculate on the next
      PPROF_T(tstep0);
      uint32_t cur;
      if (spec_ok && spec_x == x && spec_dt == dt && spec_ik == ik) {
        cur = reqi - 1;
      } else {
        if (reqi > 0) {      // the wrong guess must have drained before its barriers are waited on again (phase hygiene)
          const uint32_t g = reqi - 1;
#pragma unroll 1
          for (int s5 = 0; s5 < 5; s5++) pb_wait(&bars[LY::B_FULL + (g & 1) * 5 + s5], (g >> 1) & 1);
        }
        cur = reqi; post_req(x, dt, ik);
      }
      {
        // the step after this one if it is accepted with the step size unchanged (the controller's dead zone: ~93 % of the steps)
        const bool last_now = fixed ? (fixed_left == 1) : clamped;
        double xn = fixed ? (x_begin + (double)(fixed_total - fixed_left + 1) * p.fixed_dt) : (x + dt);
        double dtn = dt;
        if (last_now || (!fixed && xn >= x_end)) { xn = x; dtn = 0.0; }          // nothing follows: placeholder request
        else if (!fixed && xn + dtn >= x_end) dtn = x_end - xn;
        post_req(xn, dtn, ik);
        spec_ok = dtn != 0.0; spec_x = xn; spec_dt = dtn; spec_ik = ik;
      }
      const uint32_t set = cur & 1, fpar = (cur >> 1) & 1, par = att & 1;
      att++;
      if (lane == 0) { stepd[0] = s1; stepd[1] = flipU ? 1.0 : 0.0; stepd[2] = flipZ ? 1.0 : 0.0; stepd[3] = (double)ik; pb_arrive(&bars[LY::B_STEP]); }

      PPROF_ACC(0, tstep0);
      bool accept = true; double E2 = 0.0, lE = 0.0;
      double rr[MAXLEN], r5[5], q5[5];
      {
        const double a0 = KC_A[1][0] * s1;
#pragma unroll
        for (int l = 0; l < MAXLEN; l++) rr[l] = U[lo_ + l * NCH] + a0 * Z0[lo_ + l * NCH];
#pragma unroll
        for (int j = 0; j < 5; j++) r5[j] = U[ln.iS + j] + a0 * Z0[ln.iS + j];
      }
#pragma unroll 1
      for (int s = 1; s <= 6; s++) {
        double* zs = (s == 1) ? Z1 : (s == 2) ? Z2 : (s == 3) ? Z3 : (s == 4) ? Z4 : Z5;
        const double* slot = slots + (size_t)(set * 5 + (s <= 5 ? s : 5) - 1) * LY::SLOT;
        PPROF_T(tst0);
        if (s >= 2) {
          // rhs_s = P_s + a_{s,s-1} z_{s-1};  "stage 6": err = E_part + (b - bhat)_6 z_6, smoothed by W^{-1} of the last stage
          const double* Pb = (s & 1) ? P1 : P0;
          const double al = (s <= 5) ? KC_A[s][s - 1] : KC_E[5];
          pb_wait(&bars[(s <= 5) ? LY::B_PR + s : LY::B_ER], par);
          PPROF_ACC(1, tst0);
#pragma unroll
          for (int l = 0; l < MAXLEN; l++) rr[l] = Pb[lo_ + l * NCH] + al * rr[l];
#pragma unroll
          for (int j = 0; j < 5; j++) r5[j] = Pb[ln.iS + j] + al * r5[j];
        }
        if (s <= 5) {
          // stage 5 writes the z6 and u_{n+1} slots, which were z1 and u_n of the previous step: the sampler must be done with them
          if (s == 5 && jobs_waited < jobs) { pb_wait(&bars[LY::B_SDONE], jobs_waited & 1); jobs_waited++; }
#pragma unroll
          for (int l = 0; l < MAXLEN; l++) zs[lo_ + l * NCH] = rr[l];     // park the right-hand side in the stage's own z slot
#pragma unroll
          for (int j = 0; j < 5; j++) q5[j] = r5[j];
          PPROF_T(twf0);
          pb_wait(&bars[LY::B_FULL + set * 5 + s - 1], fpar);
          PPROF_ACC(2, twf0);
          rsa_flag |= slot[LY::UNI + LY::U_RSA] != 0.0;
        } else if (fixed) break;
        PPROF_T(tso0);
#ifdef K1P_PROF
        long long tsv[5];
        pipe_solve<TR>(ln, slot, rr, r5, tsv);
        prof[9] += tsv[0] - tso0; prof[10] += tsv[1] - tsv[0]; prof[11] += tsv[2] - tsv[1]; prof[12] += tsv[3] - tsv[2]; prof[13] += tsv[4] - tsv[3];
        pcnt[9]++; pcnt[10]++; pcnt[11]++; pcnt[12]++; pcnt[13]++;
#else
        pipe_solve<TR>(ln, slot, rr, r5);
#endif
        PPROF_ACC(3, tso0);
        PPROF_T(tz0);
        if (s <= 5) {
          if (s == 5) {
            // u_{n+1} = U_6 (stiffly accurate)
#pragma unroll
            for (int l = 0; l < MAXLEN; l++) UN[lo_ + l * NCH] = rr[l];
            if (lane == 0) {
#pragma unroll
              for (int j = 0; j < 5; j++) UN[ln.iS + j] = r5[j];
            }
          }
#pragma unroll
          for (int l = 0; l < MAXLEN; l++) { const int idx = lo_ + l * NCH; const double z = (rr[l] - zs[idx]) * (1.0 / KC_GAMMA); rr[l] = z; zs[idx] = z; }
#pragma unroll
          for (int j = 0; j < 5; j++) r5[j] = (r5[j] - q5[j]) * (1.0 / KC_GAMMA);
          if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 5; j++) zs[ln.iS + j] = r5[j];
          }
          __syncwarp();
          if (lane == 0) pb_arrive(&bars[LY::B_ZP + s]);
        }
        PPROF_ACC(4, tz0);
        PPROF_ACC(5, tst0);
      }
      __syncwarp();
      PPROF_T(tn0);
      pb_wait(&bars[LY::B_SC], par);          // the helper's inverse error scales (every attempt, so the phases stay in step)
      if (!fixed) {
        double ssum = 0.0, ssum2 = 0.0;
#pragma unroll
        for (int l = 0; l < MAXLEN; l++) {
          const double q = rr[l] * P1[lo_ + l * NCH];
          if (l & 1) ssum2 += q * q; else ssum += q * q;
        }
        ssum += ssum2;
        if (ln.lane < 5) {
          double e = r5[0]; e = (ln.lane == 1) ? r5[1] : e; e = (ln.lane == 2) ? r5[2] : e; e = (ln.lane == 3) ? r5[3] : e; e = (ln.lane == 4) ? r5[4] : e;
          const double q = e * P1[ln.iS + ln.lane]; ssum += q * q;
        }
        // EEst = sqrt(E2).  The controller works on log EEst = log(E2)/2 directly: one log and one exp on the spine instead of
        // sqrt + 2 log + 2 exp + 3 divisions (~1250 cycles of dependent special-function code per step)
        E2 = warp_sum(ssum) * inv_n;
        if (!isfinite(E2)) { status = BOLT_K_NONFINITE; break; }
        lE = 0.5 * log(fmax(E2, 1e-12));                // log max(EEst, 1e-6): controller noise floor, DESIGN.md
        accept = E2 <= 1.0;
#ifndef K1P_PROF
        if (p.dbg && ik == 0 && ln.lane == 0 && nsteps + nreject < p.dbg_cap) {
          double* d = p.dbg + 4 * (nsteps + nreject); d[0] = x; d[1] = dt; d[2] = sqrt(E2); d[3] = accept ? 1.0 : 0.0;
        }
#endif
      }
      PPROF_ACC(6, tn0);
      PPROF_T(tc0);
      if (accept) {
        const bool last = fixed ? (fixed_left == 1) : clamped;
        const double xn1 = last ? x_end : (fixed ? (x_begin + (double)(fixed_total - fixed_left + 1) * p.fixed_dt) : (x + dt));
        // dense output and sources of this step: the sampler's job
        if (lane == 0) {
          jobd[0] = x; jobd[1] = dt; jobd[2] = s1; jobd[3] = xn1; jobd[4] = last ? 1.0 : 0.0; jobd[5] = flipU ? 1.0 : 0.0; jobd[6] = flipZ ? 1.0 : 0.0;
          jobd[7] = (double)ik;
          pb_arrive(&bars[LY::B_SJOB]);
        }
        jobs++;
        x = xn1; nsteps++;
        flipU = !flipU; flipZ = !flipZ; U = PSLOT_U; UN = PSLOT_UN; Z0 = PSLOT_Z0; Z5 = PSLOT_Z5;
        if (fixed) { fixed_left--; s1 = 1.0; }
        else {
          double q = exp(beta1 * lE - beta2 * lqold);            // EEst^beta1 / qold^beta2
          q = fmax(1.0 / qmax, fmin(1.0 / qmin, q / safety));
          lqold = fmax(lE, -9.210340371976182);                  // log max(EEst, 1e-4)
          if (q <= 1.2 && q >= 1.0) s1 = 1.0;                      // the controller's dead zone: step size unchanged
          else { const double dtn = dt / q; s1 = dtn / dt; dt = dtn; }
        }
      } else {
        nreject++;
        const double dtn = dt / fmin(1.0 / qmin, exp(beta1 * lE) / safety);
        s1 *= dtn / dt; dt = dtn;
        if (!(dt > 1e-14)) { status = BOLT_K_DT_UNDERFLOW; break; }
      }
      PPROF_ACC(7, tc0);
      PPROF_ACC(8, tstep0);
    }
#ifdef K1P_PROF
    if (w == 0) PPROF_DUMP(0, 14);
#endif
    // the last sample job must be finished before the state is read back / reused
    if (jobs_waited < jobs) { pb_wait(&bars[LY::B_SDONE], jobs_waited & 1); jobs_waited++; }
    rsa_flag |= flags[0] != 0.0;
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
  // no more modes: release every role
  __syncwarp();
  post_req(0.0, 0.0, -1);
  if (lane == 0) {
    stepd[3] = -1.0; pb_arrive(&bars[LY::B_STEP]);
    jobd[7] = -1.0; pb_arrive(&bars[LY::B_SJOB]);
  }
#undef PSLOT_U
#undef PSLOT_UN
#undef PSLOT_Z0
#undef PSLOT_Z5
}

}  // namespace bolt
