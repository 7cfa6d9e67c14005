// K1 with forward-mode partials, ONE CTA PER k-MODE: the value system and each of the NP sensitivity systems on a warp of its
// own (sm_100a, FP64) -- source_grid's truncations.
//
// Mathematics of hierarchy_dual.cuh / hierarchy_dual_reg.cuh (the reference integrates a Dual-valued state,
// examples/plot_deriv_cl.jl:28-33): every implicit stage is  W U = r,  W S_j = r_j + h G_j,  G_j = (dA/dp_j)(x_s) U  with the SAME
// matrix W = I - h A(x_s).  The one-warp kernels solve the 1 + NP systems of a mode one after the other (88 us per step at three
// warps per SM, Dual<NP> register spills of 2-5 KB per thread).  Here the systems of a mode run SIDE BY SIDE:
//
//   warp 0      value:  r = u_n + sum a_sj z_j (flat), background, factorisation of W into a shared-memory slot, back-solve,
//               publishes the stage value U_s; runs ONE STAGE AHEAD of the others (two slots / two U_s buffers)
//   warp j      sensitivity j (j = 1..NP):  r_j (flat, over its own component arrays) and the Dual<1> background of ITS partial
//               while warp 0 factors; after the stage barrier G_j = (dA/dp_j) U_s in Dual<1> arithmetic on a single-partial view of
//               the cosmology, back-solve of r_j + h G_j with the slot, z_j
//
// One named barrier per stage, three at the end of a step (u_{n+1} of every component -> error contributions -> decision).  The
// error norm runs over value and partials with the combined per-element scale (DiffEqBase semantics, hierarchy_dual.cuh).  The
// state is the component-major layout of hierarchy_dual_reg.cuh, so initial conditions, f(u_0), dense output and both source
// functions in dual arithmetic are the very functions of hierarchy_dual.cuh (run by warp 0 on the full Dual<NP> view).
#pragma once
#include "hierarchy_dual_reg.cuh"
#include "stage_slot.cuh"

namespace bolt {

__device__ __forceinline__ void dcta_sync(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

template <class TR, int NP> __host__ __device__ constexpr size_t k1_dual_cta_smem_doubles() {
  typedef CtaLayout<TR> LY;
  // 7 state arrays x (1+NP) components | 2 factor slots | 2 stage-value buffers | one right-hand-side scratch per warp | control
  return (size_t)7 * (1 + NP) * LY::NA + 2 * LY::SLOT + 2 * LY::NA + (size_t)(1 + NP) * LY::NA + 32;
}

template <class TR, int NP>
__global__ void __launch_bounds__(32 * (1 + NP), (NP <= 4) ? 2 : 1) hierarchy_dual_cta_kernel(SolveParams p) {
  extern __shared__ double sm[];
  typedef CtaLayout<TR> LY;
  constexpr int ND = 1 + NP, NCH = TR::NCH, MAXLEN = TR::MAXLEN, NTH = 32 * ND;
  constexpr int na = LY::NA;
  constexpr size_t astr = (size_t)ND * na;
  constexpr int NFLAT = MAXLEN * NCH + 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* const slots = sm + 7 * astr;
  double* const US = slots + 2 * LY::SLOT;
  double* const scr = US + 2 * na + (size_t)warp * na;        // this warp's right-hand-side scratch
  double* const ctrl = US + 2 * na + (size_t)ND * na;         // [0] next work item, [1] accept, [2] dt, [3] s1, [4] x, [5] status, [6] stop, [8..] error parts
  const int n = p.n;
  Lane ln;
  const bool fixed = (p.mode == BOLT_MODE_FIXED);
  const double reltol = p.reltol, abstol = p.abstol;
  for (int i = threadIdx.x; i < 2 * LY::SLOT; i += NTH) slots[i] = 0.0;      // the idle lanes' all-zero column

  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) ctrl[0] = (double)atomicAdd(p.counter, 1);
    __syncthreads();
    const int w = (int)ctrl[0];
    if (w >= p.nk) break;
    const int ik = p.order[w];
    const DevCosmo& c = *p.cos_list[ik / p.nk_per];                           // the Dual<NP> view (warp 0: IC, sampling, output)
    const DevCosmo& cj = *p.view_list[warp > 0 ? warp - 1 : 0];               // this warp's single-partial view
    lane_setup<TR>(c, p, ln);
    ln.k = p.k[ik];
    const int lo_ = ln.base;
    const double x_begin = c.x0, x_end = 0.0;

    bool flipU = false, flipZ = false;
    const DArr<NP> Z2{sm + 3 * astr, na}, Z3{sm + 4 * astr, na}, Z4{sm + 5 * astr, na};
#define CSL_U  DArr<NP>{sm + (flipU ? 2 * astr : 0), na}
#define CSL_Z1 DArr<NP>{sm + (flipU ? 0 : 2 * astr), na}
#define CSL_Z0 DArr<NP>{sm + (flipZ ? 6 * astr : astr), na}
#define CSL_Z5 DArr<NP>{sm + (flipZ ? astr : 6 * astr), na}
    DArr<NP> U = CSL_U, Z0 = CSL_Z0, Z1 = CSL_Z1, Z5 = CSL_Z5;
    for (int i = threadIdx.x; i < 7 * (int)astr; i += NTH) sm[i] = 0.0;       // padded rows must read as zero
    __syncthreads();

    double dt = 0.0;
    bool rsa_flag = false;
    if (warp == 0) {
      // initial conditions, f(u_0) and the initial step on the full dual view, exactly as the one-warp kernel
      BgD<NP> bd;
      eval_bg_d<NP>(c, ln, x_begin, bd);
      initial_conditions_d<NP>(c, ln, bd, U);
      rhs_full_d<NP>(c, ln, bd, U, Z5, false);
      rsa_flag = (ln.k * bd.eta.v > 240.0) && (-bd.taup.v * bd.H.v / bd.eta.v > 100.0);
      double* const sc0 = Z2.p;
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
      if (fixed) dt = p.fixed_dt;
      else {
        const double d0 = sqrt(sumsq_scaled(U.p, U.p, U.p) / n), d1 = sqrt(sumsq_scaled(Z5.p, U.p, U.p) / n);
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        dt0 = fmin(dt0, x_end - x_begin);
#pragma unroll 1
        for (int l = 0; l < ln.len; l++) { const int idx = ln.base + l * ln.stride; sc0[idx] = U.p[idx] + dt0 * Z5.p[idx]; }
        if (ln.lane < 5) { const int idx = ln.iS + ln.lane; sc0[idx] = U.p[idx] + dt0 * Z5.p[idx]; }
        __syncwarp();
        Bg b1; eval_bg(c, ln, x_begin + dt0, b1);
        rhs_full(c, ln, b1, sc0, Z0.p);
#pragma unroll 1
        for (int l = 0; l < ln.len; l++) { const int idx = ln.base + l * ln.stride; Z0.p[idx] -= Z5.p[idx]; }
        if (ln.lane < 5) { const int idx = ln.iS + ln.lane; Z0.p[idx] -= Z5.p[idx]; }
        __syncwarp();
        const double d2 = sqrt(sumsq_scaled(Z0.p, U.p, U.p) / n) / dt0;
        const double dm = fmax(d1, d2);
        const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(dm)) / 5.0);
        dt = fmin(100.0 * dt0, dt1);
        for (int i = lane; i < na; i += 32) { sc0[i] = 0.0; Z0.p[i] = 0.0; }
        __syncwarp();
      }
      if (lane == 0) ctrl[2] = dt;
    }
    __syncthreads();
    dt = ctrl[2];
    flipZ = !flipZ; Z0 = CSL_Z0; Z5 = CSL_Z5;
    double s1 = dt, x = x_begin;

    int ix = 0;
    int status = BOLT_K_OK;
    long long nsteps = 0, nreject = 0;
    const double beta1 = 7.0 / 40.0, beta2 = 2.0 / 20.0, safety = 0.9, qmin = 0.2, qmax = 10.0;
    double qold = 1e-4;
    const long long fixed_total = fixed ? llround((x_end - x_begin) / p.fixed_dt) : 0;
    long long fixed_left = fixed_total;
    const long long max_steps = p.max_steps > 0 ? p.max_steps : 1000000;
    ModeConst mc; mode_const(c, ln, mc);

    while (true) {
      // every warp evaluates the same loop conditions on identical copies of (x, dt, counters)
      bool clamped = false;
      if (fixed) { if (fixed_left == 0) break; }
      else {
        if (x >= x_end) break;
        if (x + dt >= x_end) { const double dtn = x_end - x; s1 *= dtn / dt; dt = dtn; clamped = true; }
      }
      if (nsteps + nreject >= max_steps) { status = BOLT_K_MAXSTEPS; break; }
      const double h = KC_GAMMA * dt;
      double rr[MAXLEN], r5[5];
      const size_t coff = (size_t)warp * na;                 // this warp's component inside a state array

#pragma unroll 1
      for (int s = 1; s < 6; s++) {
        const double a0 = KC_A[s][0] * s1, a1 = KC_A[s][1], a2 = KC_A[s][2], a3 = KC_A[s][3], a4 = KC_A[s][4];
        const DArr<NP> zout = (s == 1) ? Z1 : (s == 2) ? Z2 : (s == 3) ? Z3 : (s == 4) ? Z4 : Z5;
        double* const zc = zout.p + coff;
        const double* Uc = U.p + coff; const double* Z0c = Z0.p + coff; const double* Z1c = Z1.p + coff;
        const double* Z2c = Z2.p + coff; const double* Z3c = Z3.p + coff; const double* Z4c = Z4.p + coff;
        const double xs = x + KC_C[s] * dt;
        double* const slot = slots + (size_t)(s & 1) * LY::SLOT;
        double* const us = US + (size_t)(s & 1) * na;
        // r (this warp's component) = u_n + sum_j a_sj z_j, flat over the state; parked in the stage's own z slot
#pragma unroll
        for (int t = 0; t < (NFLAT + 31) / 32; t++) {
          const int i = lane + 32 * t;
          if (i < NFLAT) zc[i] = Uc[i] + a0 * Z0c[i] + a1 * Z1c[i] + a2 * Z2c[i] + a3 * Z3c[i] + a4 * Z4c[i];
        }
        if (warp == 0) {
          factor_stage<TR>(c, ln, mc, xs, h, slot);           // background + chain pivots + betas + 4x4 LU -> slot
          __syncwarp();
          rsa_flag |= slot[LY::UNI + LY::U_RSA] != 0.0;
#pragma unroll
          for (int l = 0; l < MAXLEN; l++) rr[l] = zc[lo_ + l * NCH];
#pragma unroll
          for (int j = 0; j < 5; j++) r5[j] = zc[ln.iS + j];
          solve_slot<TR>(ln, slot, rr, r5);                   // rr, r5 = the stage value U_s
#pragma unroll
          for (int l = 0; l < MAXLEN; l++) us[lo_ + l * NCH] = rr[l];
          if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 5; j++) us[ln.iS + j] = r5[j];
          }
          dcta_sync(NTH);                                     // slot and U_s are published; the sensitivity warps start on them
        } else {
          // the Dual<1> background of this warp's partial, while warp 0 factors
          const double tv = bg_stage_prefetch<1>(cj, ln, xs);
          BgD<1> bd;
          eval_bg_d_stage<1>(cj, ln, xs, tv, bd);
          dcta_sync(NTH);
          // G_j = (dA/dp_j) U_s on the plain stage value; right-hand side r_j + h G_j
          double uu[MAXLEN], u5[5];
#pragma unroll
          for (int l = 0; l < MAXLEN; l++) uu[l] = us[lo_ + l * NCH];
#pragma unroll
          for (int j = 0; j < 5; j++) u5[j] = us[ln.iS + j];
          MetricG<1> m;
          m.Phi = u5[0]; m.delta = u5[1]; m.v = u5[2]; m.delta_b = u5[3]; m.v_b = u5[4];
          const Dual<1> sPsi = warp_sum_T(bd.wPsi * uu[2]), sPhi = warp_sum_T(bd.wPhi * uu[0]);
          double pi = 0.0;
          if (ln.kind == CH_T) pi = uu[2]; else if (ln.kind == CH_P) pi = uu[2] + uu[0];
          m.Pi = warp_sum(pi);
          m.Psi = -(bd.cPsi * sPsi) - m.Phi;
          m.dPhi = m.Psi - bd.k2 * m.Phi + bd.gPhi * (cs_d<1>(cj, BOLT_S_Omega_c) * (m.delta / bd.a) + cs_d<1>(cj, BOLT_S_Omega_b) * (m.delta_b / bd.a) + sPhi);
          const double T1 = shfl_d(uu[1], ln.nq);
#pragma unroll
          for (int l = 0; l < MAXLEN; l++) {
            const Dual<1> g = g_row_reg<TR, 1>(ln, bd, m, l, uu);
            rr[l] = zc[lo_ + l * NCH] + h * g.d[0];
          }
          const Dual<1> g0 = m.dPhi, g1 = bd.kappa * m.v - 3.0 * m.dPhi, g2 = -(bd.kappa * m.Psi) - m.v, g3 = bd.kappa * m.v_b - 3.0 * m.dPhi;
          const Dual<1> g4 = -(bd.kappa * (m.Psi + bd.csb2 * m.delta_b)) + bd.taup * bd.R * (3.0 * T1 + m.v_b) - m.v_b;
          r5[0] = zc[ln.iS] + h * g0.d[0]; r5[1] = zc[ln.iS + 1] + h * g1.d[0]; r5[2] = zc[ln.iS + 2] + h * g2.d[0];
          r5[3] = zc[ln.iS + 3] + h * g3.d[0]; r5[4] = zc[ln.iS + 4] + h * g4.d[0];
          solve_slot<TR>(ln, slot, rr, r5);                   // rr, r5 = S_j of this stage
        }
        // stage increment z = (solution - r)/gamma (r waits in the z slot)
#pragma unroll
        for (int l = 0; l < MAXLEN; l++) { const int idx = lo_ + l * NCH; zc[idx] = (rr[l] - zc[idx]) * (1.0 / KC_GAMMA); }
        {
          double zz[5];
#pragma unroll
          for (int j = 0; j < 5; j++) zz[j] = (r5[j] - zc[ln.iS + j]) * (1.0 / KC_GAMMA);
          __syncwarp();
          if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 5; j++) zc[ln.iS + j] = zz[j];
          }
        }
        __syncwarp();
      }
      // ---- end of step: u_{n+1} per component, error contributions, decision ----
      {
        const double b0 = KC_A[5][0] * s1;
        const double* Uc = U.p + coff; const double* Z0c = Z0.p + coff; const double* Z2c = Z2.p + coff;
        const double* Z3c = Z3.p + coff; const double* Z4c = Z4.p + coff; const double* Z5c = Z5.p + coff;
        double* Z1c = Z1.p + coff;
#pragma unroll
        for (int t = 0; t < (NFLAT + 31) / 32; t++) {
          const int i = lane + 32 * t;
          if (i < NFLAT) Z1c[i] = Uc[i] + b0 * Z0c[i] + KC_A[5][2] * Z2c[i] + KC_A[5][3] * Z3c[i] + KC_A[5][4] * Z4c[i] + KC_GAMMA * Z5c[i];
        }
      }
      dcta_sync(NTH);                  // every component of u_{n+1} is in place (the combined error scale needs all of them)
      if (!fixed) {
        const double e0 = KC_E[0] * s1;
        const double* Z0c = Z0.p + coff; const double* Z2c = Z2.p + coff; const double* Z3c = Z3.p + coff;
        const double* Z4c = Z4.p + coff; const double* Z5c = Z5.p + coff;
        auto err_at = [&](int idx) { return e0 * Z0c[idx] + KC_E[2] * Z2c[idx] + KC_E[3] * Z3c[idx] + KC_E[4] * Z4c[idx] + KC_E[5] * Z5c[idx]; };
        auto inv_scale = [&](int idx) {
          double n0 = 0.0, n1 = 0.0;
#pragma unroll
          for (int j = 0; j < ND; j++) { const double a = U.p[(size_t)j * na + idx], b2 = Z1.p[(size_t)j * na + idx]; n0 += a * a; n1 += b2 * b2; }
          return fast_rcp(abstol + reltol * sqrt(fmax(n0, n1)));
        };
#pragma unroll
        for (int l = 0; l < MAXLEN; l++) rr[l] = err_at(lo_ + l * NCH);
#pragma unroll
        for (int q = 0; q < 5; q++) r5[q] = err_at(ln.iS + q);
        solve_slot<TR>(ln, slots + (size_t)(5 & 1) * LY::SLOT, rr, r5);      // smoothed by W^{-1} of the last stage
        double ssum = 0.0;
#pragma unroll
        for (int l = 0; l < MAXLEN; l++) { const double q = rr[l] * inv_scale(lo_ + l * NCH); ssum += q * q; }
        if (lane == 0) {
#pragma unroll
          for (int q = 0; q < 5; q++) { const double e = r5[q] * inv_scale(ln.iS + q); ssum += e * e; }
        }
        ssum = warp_sum(ssum);
        if (lane == 0) ctrl[8 + warp] = ssum;
      }
      dcta_sync(NTH);
      if (warp == 0) {
        bool accept = true; double EEst = 0.0, q11 = 0.0;
        int st = BOLT_K_OK;
        if (!fixed) {
          double tot = 0.0;
#pragma unroll
          for (int j = 0; j < ND; j++) tot += ctrl[8 + j];
          EEst = sqrt(tot / ((double)n * p.out_nd));      // totallength of the CALLER's dual state (uncarried partials are zeros that count)
          if (!isfinite(EEst)) st = BOLT_K_NONFINITE;
          q11 = exp(beta1 * log(fmax(EEst, 1e-6)));
          accept = EEst <= 1.0;
          if (p.dbg && ik == 0 && lane == 0 && nsteps + nreject < p.dbg_cap) {
            double* d = p.dbg + 4 * (nsteps + nreject); d[0] = x; d[1] = dt; d[2] = EEst; d[3] = accept ? 1.0 : 0.0;
          }
        }
        double xn = x, dtn = dt, s1n = s1;
        if (st == BOLT_K_OK) {
          if (accept) {
            const bool last = fixed ? (fixed_left == 1) : clamped;
            const double xn1 = last ? x_end : (fixed ? (x_begin + (double)(fixed_total - fixed_left + 1) * p.fixed_dt) : (x + dt));
            xn = xn1;
            if (fixed) s1n = 1.0;
            else {
              double q = q11 * exp(-beta2 * log(qold));
              q = fmax(1.0 / qmax, fmin(1.0 / qmin, q / safety));
              if (q <= 1.2 && q >= 1.0) q = 1.0;
              qold = fmax(EEst, 1e-4);
              dtn = dt / q; s1n = dtn / dt;
            }
          } else {
            dtn = dt / fmin(1.0 / qmin, q11 / safety);
            s1n = s1 * (dtn / dt);
            if (!(dtn > 1e-14)) st = BOLT_K_DT_UNDERFLOW;
          }
        }
        __syncwarp();
        if (lane == 0) { ctrl[1] = accept ? 1.0 : 0.0; ctrl[2] = dtn; ctrl[3] = s1n; ctrl[4] = xn; ctrl[5] = (double)st; }
      }
      dcta_sync(NTH);
      {
        const bool accept = ctrl[1] != 0.0;
        const int st = (int)ctrl[5];
        if (st != BOLT_K_OK) { status = st; break; }
        if (accept) {
          // dense output on the grid rows this step crossed: EVERY warp samples its own component (warp 0 the value, warp j the
          // Dual<1> sources of its partial on its single-partial view) instead of warp 0 doing all 1 + NP while the others wait
          const bool last = fixed ? (fixed_left == 1) : clamped;
          const double xn1 = ctrl[4];
          bool sampled = false;
          while (ix < c.n_x) {
            const double xq = c.x0 + c.dx * ix;
            if (!last && xq > xn1 + 1e-12) break;
            if (ix >= p.ix_first) {
              double th = (xq - x) / dt; if (th > 1.0) th = 1.0;
              Hermite hm = hermite_weights(th);
              if (warp == 0) sample_sources(c, ln, p, ik, ix, xq, hm, U.p, Z1.p, Z0.p, s1, Z5.p, rsa_flag, &mc);
              else {
                const int cs = warp * na;      // Dual<1> view of (value, this warp's partial) of every state array
                sample_sources_d<1>(cj, ln, p, ik, ix, xq, hm, DArr<1>{U.p, cs}, DArr<1>{Z1.p, cs}, DArr<1>{Z0.p, cs}, s1, DArr<1>{Z5.p, cs},
                                    rsa_flag, false, &p.comp_map[warp - 1]);
              }
              sampled = true;
            }
            ix++;
          }
          if (sampled) dcta_sync(NTH);        // nobody overwrites u_n / z_1 (next step's stage 1) before every warp has read them
          x = xn1; nsteps++;
          flipU = !flipU; flipZ = !flipZ; U = CSL_U; Z1 = CSL_Z1; Z0 = CSL_Z0; Z5 = CSL_Z5;
          if (fixed) fixed_left--;
        } else nreject++;
        dt = ctrl[2]; s1 = ctrl[3];
      }
    }
    __syncthreads();
    if (warp == 0) {
      if (rsa_flag && status == BOLT_K_OK) status = BOLT_K_RSA_TRIGGERED;
      if (p.u_final) {   // [nk][n][nd]
        double* out = p.u_final + (size_t)ik * n * p.out_nd;
#pragma unroll 1
        for (int l = 0; l < ln.len; l++)
          for (int j = 0; j < ND; j++) out[(size_t)(ln.rbase + l * ln.rstride) * p.out_nd + (j ? p.comp_map[j - 1] : 0)] = U.p[(size_t)j * na + ln.base + l * ln.stride];
        if (ln.lane < 5) for (int j = 0; j < ND; j++) out[(size_t)(ln.riS + ln.lane) * p.out_nd + (j ? p.comp_map[j - 1] : 0)] = U.p[(size_t)j * na + ln.iS + ln.lane];
      }
      if (lane == 0) {
        if (p.status) p.status[ik] = status;
        if (p.nsteps) p.nsteps[ik] = nsteps;
        if (p.nreject) p.nreject[ik] = nreject;
      }
    }
#undef CSL_U
#undef CSL_Z1
#undef CSL_Z0
#undef CSL_Z5
  }
}

}  // namespace bolt
