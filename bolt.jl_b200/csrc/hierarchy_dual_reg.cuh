// K1 with forward-mode partials on the register-resident stage solver (source_grid's truncations).
//
// Same mathematics as hierarchy_dual.cuh -- W U = r, W S_j = r_j + h G_j with ONE factorisation per stage -- but the value
// solve and all NP sensitivity solves go through factor_reg()/solve_reg() of hierarchy_kernel.cuh: the chain of the system
// being solved lives in a register array, sweeps are fully unrolled.  The dual state (7 arrays x (1+NP) components) stays in
// shared memory in the compact interleaved layout [l][chain], component-major: NCH = nq+4 columns, the last one a dummy
// column shared by the idle lanes nq+3..31 that only ever holds zeros, so that no load/store needs an idle-lane guard
// (guarded version measured 15 % slower); 67 KB per warp at NP = 4 (three warps per SM), 92 KB at NP = 6 (two).
#pragma once
#include "hierarchy_dual.cuh"

namespace bolt {

// row l of G = (dA/dp) U for the lane's chain, U in registers (compile-time l)
template <class TR, int NP>
__device__ __forceinline__ Dual<NP> g_row_reg(const Lane& ln, const BgD<NP>& b, const MetricG<NP>& m, int l,
                                              const double (&u)[TR::MAXLEN]) {
  typedef Dual<NP> T;
  const int kind = ln.kind;
  const bool photon = (kind == CH_T || kind == CH_P);
  const T kq = b.kappa * b.qe;
  const double um = (l > 0) ? u[l > 0 ? l - 1 : 0] : 0.0, up = (l + 1 < TR::MAXLEN) ? u[l + 1 < TR::MAXLEN ? l + 1 : l] : 0.0, uc = u[l];
  // truncation row (perturbations.jl:212, 243, 263-264)
  T damp = (double)ln.len / (b.H * b.eta);
  if (photon) damp = damp - b.taup;
  const T rtop = kq * um - damp * uc;
  // interior rows
  T r;
  if (l == 0) {
    r = -(kq * up);
    if (kind == CH_M) r = r + m.dPhi * ln.df0;
    else if (kind == CH_P) r = r + b.taup * (uc - m.Pi * 0.5);
    else r = r - m.dPhi;
  } else {
    const double rl = RLc(l);
    r = kq * (rl * um - (1.0 - rl) * up);
    if (l == 1) {
      if (kind == CH_M) r = r - b.kappa * (1.0 / 3.0) * b.eq * m.Psi * ln.df0;
      else if (kind != CH_P) r = r + b.kappa * (1.0 / 3.0) * m.Psi;
      if (kind == CH_T) r = r + b.taup * (m.v_b * (1.0 / 3.0));
    }
    if (photon) r = r + b.taup * (uc - (l == 2 ? m.Pi * 0.1 : 0.0));
  }
  const bool top = TR::top(kind, l), act = TR::act(kind, l);
  T out;
  out.v = act ? (top ? rtop.v : r.v) : 0.0;
#pragma unroll
  for (int j = 0; j < NP; j++) out.d[j] = act ? (top ? rtop.d[j] : r.d[j]) : 0.0;
  return out;
}

template <class TR, int NP> __host__ __device__ constexpr size_t k1_dualreg_smem_doubles() {
  return (size_t)7 * (1 + NP) * (TR::MAXLEN * TR::NCH + 8) + k1_extra_doubles<TR>() + (TR::MAXLEN * TR::NCH + 8);   // + factor scratch + rhs scratch
}

template <class TR, int NP>
__global__ void __launch_bounds__(32) hierarchy_dual_reg_kernel(SolveParams p) {
  extern __shared__ double sm[];
  typedef Dual<NP> T;
  constexpr int ND = 1 + NP, NCH = TR::NCH, MAXLEN = TR::MAXLEN;
  constexpr int na = MAXLEN * NCH + 8;            // doubles per component array
  constexpr size_t astr = (size_t)ND * na;        // stride between state arrays
  const int n = p.n;
  Lane ln;
  const bool fixed = (p.mode == BOLT_MODE_FIXED);
  const double reltol = p.reltol, abstol = p.abstol;

  while (true) {
    int w = 0;
    if (threadIdx.x == 0) w = atomicAdd(p.counter, 1);
    w = __shfl_sync(FULL, w, 0);
    if (w >= p.nk) break;
    const int ik = p.order[w];
    const DevCosmo& c = *p.cos_list[ik / p.nk_per];
    lane_setup<TR>(c, p, ln);
    ln.k = p.k[ik];
    const bool live = (TR::NCH == TR::NQ + 4) || ln.kind != CH_IDLE;   // with the dummy column every lane runs unguarded
    const int lo_ = ln.base;                       // column of the lane (idle lanes: the shared all-zero column)
    const double x_begin = c.x0, x_end = 0.0;

    bool flipU = false, flipZ = false;
    const DArr<NP> Z2{sm + 3 * astr, na}, Z3{sm + 4 * astr, na}, Z4{sm + 5 * astr, na};
#define RSLOT_U  DArr<NP>{sm + (flipU ? 2 * astr : 0), na}
#define RSLOT_Z1 DArr<NP>{sm + (flipU ? 0 : 2 * astr), na}
#define RSLOT_Z0 DArr<NP>{sm + (flipZ ? 6 * astr : astr), na}
#define RSLOT_Z5 DArr<NP>{sm + (flipZ ? astr : 6 * astr), na}
    DArr<NP> U = RSLOT_U, Z0 = RSLOT_Z0, Z1 = RSLOT_Z1, Z5 = RSLOT_Z5;
    for (int i = ln.lane; i < 7 * (int)astr; i += 32) sm[i] = 0.0;      // padded rows must read as zero
    __syncwarp();

    BgD<NP> bd;
    eval_bg_d<NP>(c, ln, x_begin, bd);
    initial_conditions_d<NP>(c, ln, bd, U);
    rhs_full_d<NP>(c, ln, bd, U, Z5, false);       // f(u0) with its partials
    bool rsa_flag = (ln.k * bd.eta.v > 240.0) && (-bd.taup.v * bd.H.v / bd.eta.v > 100.0);

    int ix = 0;
    int status = BOLT_K_OK;
    long long nsteps = 0, nreject = 0;
    double x = x_begin, dt;
    double* const scr = Z2.p;                      // scratch for the initial-step heuristic (value component of the z3 slot)
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
      const double d0 = sqrt(sumsq_scaled(U.p, U.p, U.p) / n), d1 = sqrt(sumsq_scaled(Z5.p, U.p, U.p) / n);
      double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
      dt0 = fmin(dt0, x_end - x_begin);
#pragma unroll 1
      for (int l = 0; l < ln.len; l++) { const int idx = ln.base + l * ln.stride; scr[idx] = U.p[idx] + dt0 * Z5.p[idx]; }
      if (ln.lane < 5) { const int idx = ln.iS + ln.lane; scr[idx] = U.p[idx] + dt0 * Z5.p[idx]; }
      __syncwarp();
      Bg b1; eval_bg(c, ln, x_begin + dt0, b1);
      rhs_full(c, ln, b1, scr, Z0.p);
#pragma unroll 1
      for (int l = 0; l < ln.len; l++) { const int idx = ln.base + l * ln.stride; Z0.p[idx] -= Z5.p[idx]; }
      if (ln.lane < 5) { const int idx = ln.iS + ln.lane; Z0.p[idx] -= Z5.p[idx]; }
      __syncwarp();
      const double d2 = sqrt(sumsq_scaled(Z0.p, U.p, U.p) / n) / dt0;
      const double dm = fmax(d1, d2);
      const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(dm)) / 5.0);
      dt = fmin(100.0 * dt0, dt1);
#pragma unroll 1
      for (int l = 0; l < ln.len; l++) scr[ln.base + l * ln.stride] = 0.0;
      if (ln.lane < 5) scr[ln.iS + ln.lane] = 0.0;
      __syncwarp();
    }
    flipZ = !flipZ; Z0 = RSLOT_Z0; Z5 = RSLOT_Z5;
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

      RegFactor<TR> f;
      f.fs = sm + (size_t)7 * (1 + NP) * (TR::MAXLEN * TR::NCH + 8) + ln.lane;
      double* const scr2 = sm + (size_t)7 * (1 + NP) * (TR::MAXLEN * TR::NCH + 8) + k1_extra_doubles<TR>();   // right-hand side of a partial system
      constexpr int NFLAT = MAXLEN * NCH + 5;     // the elementwise passes run flat over the state (7 lane-strided elements per lane)
      BgS bf;
      double rr[MAXLEN], r5[5];
      bool accept = true; double EEst = 0.0, q11 = 0.0;
      const double h = KC_GAMMA * dt;
      for (int s = 1; s < 6; s++) {
        const double a0 = KC_A[s][0] * s1, a1 = KC_A[s][1], a2 = KC_A[s][2], a3 = KC_A[s][3], a4 = KC_A[s][4];
        const DArr<NP> zout = (s == 1) ? Z1 : (s == 2) ? Z2 : (s == 3) ? Z3 : (s == 4) ? Z4 : Z5;
        // component `comp` of rhs_s = u_n + sum_j a_sj z_j  (coefficients of stages >= s are zero: branch-free -- skipping the
        // zero terms with uniform predicates was measured 20 % slower)
        auto rhs_at = [&](int comp, int idx) {
          const size_t o = (size_t)comp * na + idx;
          return U.p[o] + a0 * Z0.p[o] + a1 * Z1.p[o] + a2 * Z2.p[o] + a3 * Z3.p[o] + a4 * Z4.p[o];
        };
        const double xs = x + KC_C[s] * dt;
        const double tv = bg_stage_prefetch<NP>(c, ln, xs);      // table loads in flight while the value system is solved
        // ---- value: W U = r in registers ----
        // every right-hand side is formed ONCE and parked in the stage's own (not yet written) z slot, so that
        // z = (solution - rhs)/gamma needs neither a register copy nor a second pass over the six state arrays
#pragma unroll
        for (int t = 0; t < (NFLAT + 31) / 32; t++) { const int i = ln.lane + 32 * t; if (i < NFLAT) zout.p[i] = rhs_at(0, i); }
        eval_bg_fast(c, ln, mc, xs, bf);
        __syncwarp();
#pragma unroll
        for (int l = 0; l < MAXLEN; l++) rr[l] = zout.p[lo_ + l * NCH];
#pragma unroll
        for (int j = 0; j < 5; j++) r5[j] = zout.p[ln.iS + j];
        rsa_flag |= (ln.k * bf.eta > 240.0) && (-bf.taup * bf.H > 100.0 * bf.eta);
        factor_reg<TR, false>(ln, bf, h, f);
        solve_reg<TR, false>(ln, bf, f, rr, r5);          // rr, r5 = stage value U
        // ---- G = (dA/dp) U in dual arithmetic on the plain stage value; h G_j -> zout_j ----
        eval_bg_d_stage<NP>(c, ln, xs, tv, bd);
        {
          MetricG<NP> m;
          m.Phi = r5[0]; m.delta = r5[1]; m.v = r5[2]; m.delta_b = r5[3]; m.v_b = r5[4];
          const T sPsi = warp_sum_T(bd.wPsi * rr[2]), sPhi = warp_sum_T(bd.wPhi * rr[0]);
          double pi = 0.0;
          if (ln.kind == CH_T) pi = rr[2]; else if (ln.kind == CH_P) pi = rr[2] + rr[0];
          m.Pi = warp_sum(pi);
          m.Psi = -(bd.cPsi * sPsi) - m.Phi;
          m.dPhi = m.Psi - bd.k2 * m.Phi + bd.gPhi * (cs_d<NP>(c, BOLT_S_Omega_c) * (m.delta / bd.a) + cs_d<NP>(c, BOLT_S_Omega_b) * (m.delta_b / bd.a) + sPhi);
          const double T1 = shfl_d(rr[1], ln.nq);
#pragma unroll
          for (int l = 0; l < MAXLEN; l++) {
            const T g = g_row_reg<TR, NP>(ln, bd, m, l, rr);
            if (live) {
              const int idx = lo_ + l * NCH;
#pragma unroll
              for (int j = 0; j < NP; j++) zout.p[(size_t)(1 + j) * na + idx] = h * g.d[j];
            }
          }
          const T g0 = m.dPhi, g1 = bd.kappa * m.v - 3.0 * m.dPhi, g2 = -(bd.kappa * m.Psi) - m.v, g3 = bd.kappa * m.v_b - 3.0 * m.dPhi;
          const T g4 = -(bd.kappa * (m.Psi + bd.csb2 * m.delta_b)) + bd.taup * bd.R * (3.0 * T1 + m.v_b) - m.v_b;
#pragma unroll
          for (int j = 0; j < NP; j++) {      // every lane stores the same scalars (see the value kernel)
            double* zp = zout.p + (size_t)(1 + j) * na + ln.iS;
            zp[0] = h * g0.d[j]; zp[1] = h * g1.d[j]; zp[2] = h * g2.d[j]; zp[3] = h * g3.d[j]; zp[4] = h * g4.d[j];
          }
        }
        // value stage increment (U is still in rr / r5)
#pragma unroll
        for (int l = 0; l < MAXLEN; l++) if (live) { const int idx = lo_ + l * NCH; zout.p[idx] = (rr[l] - zout.p[idx]) * (1.0 / KC_GAMMA); }
        {
          double zz[5];
#pragma unroll
          for (int j = 0; j < 5; j++) zz[j] = (r5[j] - zout.p[ln.iS + j]) * (1.0 / KC_GAMMA);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 5; j++) zout.p[ln.iS + j] = zz[j];
        }
        __syncwarp();       // h G_j (written by the chain owners) is read flat below
        // ---- partials: W S_j = r_j + h G_j with the factorisation in hand ----
#pragma unroll 1
        for (int j = 1; j <= NP; j++) {
          double* zj = zout.p + (size_t)j * na;
          // system right-hand side r_j + h G_j (h G_j waits in zj); r_j is parked in zj for the stage increment
#pragma unroll
          for (int t = 0; t < (NFLAT + 31) / 32; t++) {
            const int i = ln.lane + 32 * t;
            if (i < NFLAT) { const double rj = rhs_at(j, i); scr2[i] = rj + zj[i]; zj[i] = rj; }
          }
          __syncwarp();
#pragma unroll
          for (int l = 0; l < MAXLEN; l++) rr[l] = scr2[lo_ + l * NCH];
#pragma unroll
          for (int q = 0; q < 5; q++) r5[q] = scr2[ln.iS + q];
          double rj5[5];
          solve_reg<TR, false>(ln, bf, f, rr, r5);
#pragma unroll
          for (int l = 0; l < MAXLEN; l++) if (live) { const int idx = lo_ + l * NCH; zj[idx] = (rr[l] - zj[idx]) * (1.0 / KC_GAMMA); }
#pragma unroll
          for (int q = 0; q < 5; q++) rj5[q] = (r5[q] - zj[ln.iS + q]) * (1.0 / KC_GAMMA);
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 5; q++) zj[ln.iS + q] = rj5[q];
        }
      }
      // u_{n+1} for every component (z2 slot)
      {
        const double b0 = KC_A[5][0] * s1;
        auto unew = [&](int idx) {
#pragma unroll
          for (int j = 0; j < ND; j++) {
            const size_t o = (size_t)j * na + idx;
            Z1.p[o] = U.p[o] + b0 * Z0.p[o] + KC_A[5][2] * Z2.p[o] + KC_A[5][3] * Z3.p[o] + KC_A[5][4] * Z4.p[o] + KC_GAMMA * Z5.p[o];
          }
        };
        __syncwarp();       // the last stage's increments were written by the chain owners
#pragma unroll
        for (int t = 0; t < (NFLAT + 31) / 32; t++) { const int i = ln.lane + 32 * t; if (i < NFLAT) unew(i); }
        __syncwarp();
      }
      if (!fixed) {
        // error norm over value and partials (DiffEqBase semantics, see hierarchy_dual.cuh), each component's estimate
        // smoothed by W^{-1} of the last stage
        const double e0 = KC_E[0] * s1;
        auto err_at = [&](int comp, int idx) {
          const size_t o = (size_t)comp * na + idx;
          return e0 * Z0.p[o] + KC_E[2] * Z2.p[o] + KC_E[3] * Z3.p[o] + KC_E[4] * Z4.p[o] + KC_E[5] * Z5.p[o];
        };
        auto inv_scale = [&](int idx) {
          double n0 = 0.0, n1 = 0.0;
#pragma unroll
          for (int j = 0; j < ND; j++) { const double a = U.p[(size_t)j * na + idx], b2 = Z1.p[(size_t)j * na + idx]; n0 += a * a; n1 += b2 * b2; }
          return fast_rcp(abstol + reltol * sqrt(fmax(n0, n1)));
        };
        double isc[MAXLEN], isc5[5];
#pragma unroll
        for (int l = 0; l < MAXLEN; l++) isc[l] = live ? inv_scale(lo_ + l * NCH) : 0.0;
#pragma unroll
        for (int j = 0; j < 5; j++) isc5[j] = inv_scale(ln.iS + j);
        double ssum = 0.0;
#pragma unroll 1
        for (int j = 0; j < ND; j++) {
#pragma unroll
          for (int l = 0; l < MAXLEN; l++) rr[l] = live ? err_at(j, lo_ + l * NCH) : 0.0;
#pragma unroll
          for (int q = 0; q < 5; q++) r5[q] = err_at(j, ln.iS + q);
          solve_reg<TR, false>(ln, bf, f, rr, r5);
#pragma unroll
          for (int l = 0; l < MAXLEN; l++) { const double q = rr[l] * isc[l]; ssum += q * q; }
          if (ln.lane == 0) {
#pragma unroll
            for (int q = 0; q < 5; q++) { const double e = r5[q] * isc5[q]; ssum += e * e; }
          }
        }
        EEst = sqrt(warp_sum(ssum) / ((double)n * p.out_nd));   // totallength of the CALLER's dual state: partials not carried (A, n) are zeros that still count
        if (!isfinite(EEst)) { status = BOLT_K_NONFINITE; break; }
        q11 = exp(beta1 * log(fmax(EEst, 1e-6)));
        accept = EEst <= 1.0;
      }
      __syncwarp();
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
        flipU = !flipU; flipZ = !flipZ; U = RSLOT_U; Z1 = RSLOT_Z1; Z0 = RSLOT_Z0; Z5 = RSLOT_Z5;
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
        for (int j = 0; j < ND; j++) out[(size_t)(ln.rbase + l * ln.rstride) * p.out_nd + (j ? p.comp_map[j - 1] : 0)] = U.p[(size_t)j * na + ln.base + l * ln.stride];
      if (ln.lane < 5) for (int j = 0; j < ND; j++) out[(size_t)(ln.riS + ln.lane) * p.out_nd + (j ? p.comp_map[j - 1] : 0)] = U.p[(size_t)j * na + ln.iS + ln.lane];
    }
    if (ln.lane == 0) {
      if (p.status) p.status[ik] = status;
      if (p.nsteps) p.nsteps[ik] = nsteps;
      if (p.nreject) p.nreject[ik] = nreject;
    }
    __syncwarp();
  }
}

}  // namespace bolt
