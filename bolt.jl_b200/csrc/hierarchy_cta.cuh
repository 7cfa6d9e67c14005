// K1, warp-specialised: ONE CTA PER k-MODE (sm_100a, FP64) -- the register-resident truncations of source_grid.
//
// Same mathematics, same controller and the same arithmetic building blocks as hierarchy_kernel_t (hierarchy_kernel.cuh;
// boltsolve src/perturbations.jl:25-33, hierarchy! :161-271, source_grid's sampling loop src/spectra.jl:13-18), but the work
// of a mode is split over the warps of a CTA by ROLE instead of being serialised in one warp:
//
//   warp 0      "solver"      the ESDIRK stage recursion: right-hand side assembly, the back-solve of every stage with a
//                             factorisation it finds ready in shared memory, error estimate, controller, dense output
//   warps 1, 2  "factorisers" the stage matrices W_s = I - gamma dt A(x_n + c_s dt) depend on (x_n, dt) only -- NOT on the
//                             solution -- so all five of a step are factored ahead of the solver: background splines at the
//                             stage abscissa, chain pivots, beta vectors, the 4x4 border system and its pivoted LU.
//                             Warp 1 takes stages 1,3,5, warp 2 stages 2,4.
//
// The serial depth of a step drops from 6 x (background + factor + solve) to 1 x factor + 6 x solve, and the per-role code is
// small enough to leave the 255-register regime of the one-warp kernel.  Hand-off through shared memory with mbarriers
// (producer arrive.release / consumer try_wait.acquire): `start` (solver -> factorisers: step descriptor x_n, dt, mode) and
// `full[s]` (factoriser -> solver: slot s is complete).  Every barrier completes exactly once per step attempt, so all
// parties track one phase bit.
#pragma once
#include "hierarchy_kernel.cuh"

namespace bolt {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}

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

// Development profile (-DK1C_PROF): cycle counters of the roles while they work on the FIRST work item (the largest k), dumped
// through the step-log buffer (BOLT_DEBUG_STEPS=<file>): row = category, cycles, count.
#ifdef K1C_PROF
#define PROF_DECL long long prof[16] = {0}; long long pcnt[16] = {0}
#define PROF_T(v) const long long v = clock64()
#define PROF_ACC(i, t0) do { prof[i] += clock64() - (t0); pcnt[i]++; } while (0)
#define PROF_DUMP(base, n_) do { if (p.dbg && (threadIdx.x & 31) == 0) for (int i_ = 0; i_ < (n_); i_++) { double* d_ = p.dbg + 4 * ((base) + i_); d_[0] = (base) + i_; d_[1] = (double)prof[i_] + 1e-9; d_[2] = (double)pcnt[i_]; d_[3] = 0; } } while (0)
#else
#define PROF_DECL
#define PROF_T(v)
#define PROF_ACC(i, t0)
#define PROF_DUMP(base, n_)
#endif

#ifndef K1C_MINBLOCKS
#define K1C_MINBLOCKS 5
#endif
constexpr int K1C_THREADS = 96;

template <class TR>
__global__ void __launch_bounds__(K1C_THREADS, K1C_MINBLOCKS) hierarchy_cta_kernel(SolveParams p) {
  typedef CtaLayout<TR> LY;
  static_assert(TR::MAXLEN > 0 && TR::NCH == TR::NQ + 4, "register-resident truncations with the compact layout");
  extern __shared__ double sm[];
  constexpr int NCH = TR::NCH, MAXLEN = TR::MAXLEN, na = LY::NA;
  const int warp = threadIdx.x >> 5;
  double* const slots = sm + LY::O_SLOTS;
  double* const desc = sm + LY::O_DESC;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(sm + LY::O_BARS);     // [0] start, [1..5] full[s]
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
#pragma unroll
    for (int s = 1; s <= 5; s++) mbar_init(&bars[s], 1);
  }
  __syncthreads();

  if (warp > 0) {
    // ------------------------------------------------ factorisers ------------------------------------------------
    const int fid = warp - 1;
    uint32_t ph = 0;
    int cur_ik = -1;
    Lane ln; ModeConst mc;
    const DevCosmo* cp = nullptr;
    PROF_DECL;
    while (true) {
      PROF_T(tw0);
      mbar_wait(&bars[0], ph);
      const double x = desc[0], dt = desc[1];
      const int ik = __double2int_rn(desc[2]);
#ifdef K1C_PROF
      if (cur_ik == p.order[0] && ik != cur_ik) PROF_DUMP(8 + 5 * fid, 5);
      if (ik == p.order[0] && cur_ik == ik) PROF_ACC(0, tw0);
#endif
      if (ik < 0) break;
      if (ik != cur_ik) {
        cur_ik = ik; cp = p.cos_list[ik / p.nk_per];
        lane_setup<TR>(*cp, p, ln);
        ln.k = p.k[ik];
        mode_const(*cp, ln, mc);
      }
#pragma unroll 1
      for (int s = 1 + fid; s <= 5; s += 2) {
#ifdef K1C_PROF
        long long ts[2]; PROF_T(tf0);
        factor_stage<TR>(*cp, ln, mc, x + KC_C[s] * dt, KC_GAMMA * dt, slots + (size_t)(s - 1) * LY::SLOT, ts);
        __syncwarp();
        if (ik == p.order[0]) { prof[1] += ts[0] - tf0; prof[2] += ts[1] - ts[0]; prof[3] += clock64() - ts[1]; pcnt[1]++; pcnt[2]++; pcnt[3]++; }
#else
        factor_stage<TR>(*cp, ln, mc, x + KC_C[s] * dt, KC_GAMMA * dt, slots + (size_t)(s - 1) * LY::SLOT);
        __syncwarp();
#endif
        if (ln.lane == 0) mbar_arrive(&bars[s]);
      }
      ph ^= 1;
    }
    return;
  }

  // ---------------------------------------------------- solver ----------------------------------------------------
  const int n = p.n;
  Lane ln;
  const bool fixed = (p.mode == BOLT_MODE_FIXED);
  const double reltol = p.reltol, abstol = p.abstol;
  uint32_t ph = 0;
  PROF_DECL;
  auto post = [&](double x, double dt, int ik) {     // publish the step descriptor and release the factorisers
    if ((threadIdx.x & 31) == 0) { desc[0] = x; desc[1] = dt; desc[2] = (double)ik; mbar_arrive(&bars[0]); }
  };

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

    // Array slots as in hierarchy_kernel_t: u/u_{n+1} and z1/z6 swap roles on every accepted step (no copies).
    bool flipU = false, flipZ = false;
    double* const Z2 = sm + (size_t)3 * na;
    double* const Z3 = sm + (size_t)4 * na;
    double* const Z4 = sm + (size_t)5 * na;
    double* const r = Z2;
#define CSLOT_U  (sm + (flipU ? (size_t)2 * na : (size_t)0))
#define CSLOT_Z1 (sm + (flipU ? (size_t)0 : (size_t)2 * na))
#define CSLOT_Z0 (sm + (flipZ ? (size_t)6 * na : (size_t)na))
#define CSLOT_Z5 (sm + (flipZ ? (size_t)na : (size_t)6 * na))
    double* U = CSLOT_U; double* Z0 = CSLOT_Z0; double* Z1 = CSLOT_Z1; double* Z5 = CSLOT_Z5;
    for (int i = ln.lane; i < LY::NARR * na; i += 32) sm[i] = 0.0;
    __syncwarp();

    Bg b;
    eval_bg(c, ln, x_begin, b);
    initial_conditions(c, ln, b, U);
    rhs_full(c, ln, b, U, Z5);          // f(u0) in the z6 slot
    bool rsa_flag = (ln.k * b.eta > 240.0) && (-b.taup * b.H / b.eta > 100.0);

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
      // initial step, Hairer-Wanner as in OrdinaryDiffEq's ode_determine_initdt (same as hierarchy_kernel_t and the oracle)
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
    flipZ = !flipZ; Z0 = CSLOT_Z0; Z5 = CSLOT_Z5;
    double s1 = dt;

    const double beta1 = 7.0 / 40.0, beta2 = 2.0 / 20.0, safety = 0.9, qmin = 0.2, qmax = 10.0;
    double qold = 1e-4;
    const long long fixed_total = fixed ? llround((x_end - x_begin) / p.fixed_dt) : 0;
    long long fixed_left = fixed_total;
    const long long max_steps = p.max_steps > 0 ? p.max_steps : 1000000;
    ModeConst mc; mode_const(c, ln, mc);
    const int lo_ = ln.base;
    constexpr int NFLAT = MAXLEN * NCH + 5;

    while (true) {
      bool clamped = false;
      if (fixed) { if (fixed_left == 0) break; }
      else {
        if (x >= x_end) break;
        if (x + dt >= x_end) { const double dtn = x_end - x; s1 *= dtn / dt; dt = dtn; clamped = true; }
      }
      if (nsteps + nreject >= max_steps) { status = BOLT_K_MAXSTEPS; break; }

      __syncwarp();
      PROF_T(tstep0);
      post(x, dt, ik);          // the factorisers start on the five stage matrices of this attempt
      bool accept = true; double EEst = 0.0, q11 = 0.0;
      double rr[MAXLEN], r5[5];
      for (int s = 1; s <= 6; s++) {
        PROF_T(ta0);
        double* zout = sm + (size_t)((s == 1) ? (flipU ? 0 : 2) : (s >= 5) ? (flipZ ? 1 : 6) : s + 1) * na;
        const double* slot = slots + (size_t)((s <= 5 ? s : 5) - 1) * LY::SLOT;
        if (s <= 5) {
          // right-hand side u_n + sum_j a_sj z_j, flat over the state, parked in the stage's own z slot
          const double a0 = KC_A[s][0] * s1, a1 = KC_A[s][1], a2 = KC_A[s][2], a3 = KC_A[s][3], a4 = KC_A[s][4];
#pragma unroll
          for (int t = 0; t < (NFLAT + 31) / 32; t++) {
            const int i = ln.lane + 32 * t;
            if (i < NFLAT) zout[i] = U[i] + a0 * Z0[i] + a1 * Z1[i] + a2 * Z2[i] + a3 * Z3[i] + a4 * Z4[i];
          }
          __syncwarp();
#pragma unroll
          for (int l = 0; l < MAXLEN; l++) rr[l] = zout[lo_ + l * NCH];
#pragma unroll
          for (int j = 0; j < 5; j++) r5[j] = zout[ln.iS + j];
          PROF_ACC(0, ta0);
          PROF_T(tw0);
          mbar_wait(&bars[s], ph);                    // W_s is factored
          PROF_ACC(1, tw0);
          rsa_flag |= slot[LY::UNI + LY::U_RSA] != 0.0;
        } else {
          // "stage 7": err = sum (b - bhat)_j z_j, smoothed by W^{-1} of the last stage; u_{n+1} = u_n + sum b_j z_j -> z2 slot
          const double e0 = KC_E[0] * s1, b0 = KC_A[5][0] * s1;
#pragma unroll
          for (int l = 0; l < MAXLEN; l++) {
            const int idx = lo_ + l * NCH;
            const double z0 = Z0[idx], z2 = Z2[idx], z3 = Z3[idx], z4 = Z4[idx], z5 = Z5[idx];
            rr[l] = e0 * z0 + KC_E[2] * z2 + KC_E[3] * z3 + KC_E[4] * z4 + KC_E[5] * z5;
            Z1[idx] = U[idx] + b0 * z0 + KC_A[5][2] * z2 + KC_A[5][3] * z3 + KC_A[5][4] * z4 + KC_GAMMA * z5;
          }
#pragma unroll
          for (int j = 0; j < 5; j++) {
            const int idx = ln.iS + j;
            const double z0 = Z0[idx], z2 = Z2[idx], z3 = Z3[idx], z4 = Z4[idx], z5 = Z5[idx];
            r5[j] = e0 * z0 + KC_E[2] * z2 + KC_E[3] * z3 + KC_E[4] * z4 + KC_E[5] * z5;
            Z1[idx] = U[idx] + b0 * z0 + KC_A[5][2] * z2 + KC_A[5][3] * z3 + KC_A[5][4] * z4 + KC_GAMMA * z5;
          }
          if (fixed) break;
          PROF_ACC(4, ta0);
        }
        PROF_T(ts0);
        solve_slot<TR>(ln, slot, rr, r5);
        PROF_ACC(2, ts0);
        PROF_T(tz0);
        if (s <= 5) {
#pragma unroll
          for (int l = 0; l < MAXLEN; l++) { const int idx = lo_ + l * NCH; zout[idx] = (rr[l] - zout[idx]) * (1.0 / KC_GAMMA); }
          double zz[5];
#pragma unroll
          for (int j = 0; j < 5; j++) zz[j] = (r5[j] - zout[ln.iS + j]) * (1.0 / KC_GAMMA);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 5; j++) zout[ln.iS + j] = zz[j];
          __syncwarp();
          PROF_ACC(3, tz0);
        }
      }
      ph ^= 1;
      __syncwarp();
      PROF_T(tn0);
      if (!fixed) {
        double ssum = 0.0;
#pragma unroll
        for (int l = 0; l < MAXLEN; l++) {
          const int idx = lo_ + l * NCH;
          const double sc = abstol + reltol * fmax(fabs(U[idx]), fabs(Z1[idx]));
          const double q = rr[l] * fast_rcp(sc); ssum += q * q;
        }
        if (ln.lane < 5) {
          double e = r5[0]; e = (ln.lane == 1) ? r5[1] : e; e = (ln.lane == 2) ? r5[2] : e; e = (ln.lane == 3) ? r5[3] : e; e = (ln.lane == 4) ? r5[4] : e;
          const int idx = ln.iS + ln.lane;
          const double sc = abstol + reltol * fmax(fabs(U[idx]), fabs(Z1[idx]));
          const double q = e * fast_rcp(sc); ssum += q * q;
        }
        EEst = sqrt(warp_sum(ssum) / n);
        if (!isfinite(EEst)) { status = BOLT_K_NONFINITE; break; }
        q11 = exp(beta1 * log(fmax(EEst, 1e-6)));       // controller noise floor: DESIGN.md
        accept = EEst <= 1.0;
#ifndef K1C_PROF
        if (p.dbg && ik == 0 && ln.lane == 0 && nsteps + nreject < p.dbg_cap) {
          double* d = p.dbg + 4 * (nsteps + nreject); d[0] = x; d[1] = dt; d[2] = EEst; d[3] = accept ? 1.0 : 0.0;
        }
#endif
      }
      PROF_ACC(5, tn0);
      PROF_T(tsm0);
      if (accept) {
        const bool last = fixed ? (fixed_left == 1) : clamped;
        const double xn1 = last ? x_end : (fixed ? (x_begin + (double)(fixed_total - fixed_left + 1) * p.fixed_dt) : (x + dt));
        while (ix < c.n_x) {
          const double xs = c.x0 + c.dx * ix;
          if (!last && xs > xn1 + 1e-12) break;
          if (ix >= p.ix_first) {
            double th = (xs - x) / dt; if (th > 1.0) th = 1.0;
            Hermite hm = hermite_weights(th);
            sample_sources(c, ln, p, ik, ix, xs, hm, U, Z1, Z0, s1, Z5, rsa_flag, &mc);
          }
          ix++;
        }
        x = xn1; nsteps++;
        flipU = !flipU; flipZ = !flipZ; U = CSLOT_U; Z1 = CSLOT_Z1; Z0 = CSLOT_Z0; Z5 = CSLOT_Z5;
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
      PROF_ACC(6, tsm0);
      PROF_ACC(7, tstep0);
    }
#ifdef K1C_PROF
    if (w == 0) PROF_DUMP(0, 8);
#endif
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
  __syncwarp();
  post(0.0, 0.0, -1);       // no more modes: the factorisers leave
#undef CSLOT_U
#undef CSLOT_Z1
#undef CSLOT_Z0
#undef CSLOT_Z5
}

}  // namespace bolt
