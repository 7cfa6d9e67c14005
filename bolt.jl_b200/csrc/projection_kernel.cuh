// K2 -- spherical-Bessel line-of-sight projection and k-integration to C_l (sm_100a, FP64).
//
// Replaces, for all requested multipoles at once:
//   bessel_interpolator       src/spectra.jl:49-58   (5001-point j_l table + cubic B-spline prefilter)
//   Tl / _Tl_integrand        src/spectra.jl:70-82   (left-Riemann LOS sum over x_grid[x_i .. N-1])
//   cltt / clte / clee        src/spectra.jl:84-160  (midpoint rule on quadratic_k(kmin,kmax,n))
//   the bilinear source interpolant with Line() extrapolation in k   src/spectra.jl:21,40
//
// Pipeline (all on one stream):
//   bessel_table_kernel   j_l(i*dg) for every requested l, one thread per abscissa (Miller recurrence)
//   bessel_prefilter_kernel   B-spline coefficients per l (Thomas sweep with constant tridiagonal)
//   dense_source_kernel   S(x_i, kbar_j)*dx_i on the dense midpoint grid, laid out [x][k] (k fastest)
//   project_kernel        CTA = (group of NL multipoles) x (slice of kbar); the NL coefficient tables are
//                         staged in shared memory, the B-spline weights of each (kbar, x) are computed
//                         once and applied to all NL tables and to both sources; TT, TE, EE from one pass.
//   cl_finalize_kernel    deterministic sum over k slices, times 4 pi.
#pragma once
#include "common.cuh"

namespace bolt {

constexpr int BESSEL_NB = 5001;          // spectra.jl:53  (bessel_argmin:dg:bessel_argmax with dg = xmax/5000)
constexpr int BESSEL_NC = BESSEL_NB + 3; // ROW STRIDE of the coefficient tables: n+2 coefficients (Interpolations.jl padding) + one unused
                                         // double, so that a row is 40032 bytes = a multiple of 16: the unit of a bulk (TMA) copy

// Stage NL consecutive coefficient tables (rows e0.. of Cf, clamped at the last row) into shared memory with bulk asynchronous copies
// (cp.async.bulk, the 1-D TMA path): one elected thread arms an mbarrier with the byte count and issues one 40 KB copy per table;
// the copy engine moves the data while the CTA loads its other inputs; every thread then waits on the barrier.
__device__ __forceinline__ void stage_tables_bulk(double* tabs, const double* __restrict__ Cf, int e0, int nell, int NL,
                                                  unsigned long long* bar) {
  const unsigned bar_s = (unsigned)__cvta_generic_to_shared(bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned bytes = (unsigned)(BESSEL_NC * sizeof(double));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes * (unsigned)NL) : "memory");
    for (int l = 0; l < NL; l++) {
      const int e = min(e0 + l, nell - 1);
      const unsigned dst = (unsigned)__cvta_generic_to_shared(tabs + (size_t)l * BESSEL_NC);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "l"(Cf + (size_t)e * BESSEL_NC), "r"(bytes), "r"(bar_s) : "memory");
    }
  }
}
__device__ __forceinline__ void stage_tables_wait(unsigned long long* bar) {
  const unsigned bar_s = (unsigned)__cvta_generic_to_shared(bar);
  unsigned ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar_s) : "memory");
  }
}

// j_l(x) at x = i*dg for l in ells[] (ascending).  Miller's downward recurrence, normalised with the larger of
// j_0, j_1; two passes so that rescaling never loses already-emitted values.  (SpecialFunctions.sphericalbesselj
// is not vendored; any >= 1e-14 accurate j_l is equivalent)
__global__ void bessel_table_kernel(const int* __restrict__ ells, int nell, double dg, double* __restrict__ J) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BESSEL_NB) return;
  const double x = dg * (double)i;
  if (i == 0) {
    for (int e = 0; e < nell; e++) J[(size_t)e * BESSEL_NB] = (ells[e] == 0) ? 1.0 : 0.0;
    return;
  }
  const int lmax = ells[nell - 1];
  const double big = fmax((double)lmax, x);
  const int nstart = (int)(big + 30.0 + 10.0 * sqrt(big));
  const double ix = 1.0 / x;
  double jp = 0.0, jc = 1e-280;
  int nres = 0;
  for (int n = nstart; n >= 1; n--) {
    const double jn = (double)(2 * n + 1) * ix * jc - jp;
    jp = jc; jc = jn;
    if (fabs(jc) > 1e250) { jc *= 1e-250; jp *= 1e-250; nres++; }
  }
  double s, c; sincos(x, &s, &c);
  const double j0 = s * ix, j1 = s * ix * ix - c * ix;
  const double norm = (fabs(j0) >= fabs(j1)) ? (j0 / jc) : (j1 / jp);
  jp = 0.0; jc = 1e-280;
  int left = nres, e = nell - 1;
  for (int n = nstart; n >= 1; n--) {
    const double jn = (double)(2 * n + 1) * ix * jc - jp;
    jp = jc; jc = jn;
    if (fabs(jc) > 1e250) { jc *= 1e-250; jp *= 1e-250; left--; }
    while (e >= 0 && ells[e] == n - 1) {
      const double sc = (left == 0) ? 1.0 : ((left == 1) ? 1e-250 : 0.0);
      J[(size_t)e * BESSEL_NB + i] = (jc * sc) * norm;
      e--;
    }
  }
}

// Interpolations.jl prefilter for BSpline(Cubic(Line(OnGrid()))) (src/util.jl:11): c[1] = y[0], c[n] = y[n-1],
// interior rows (1/6, 2/3, 1/6), c[0] = 2c[1]-c[2], c[n+1] = 2c[n]-c[n-1].  One thread per multipole.
// cp[] / iden[] are the (l-independent) Thomas multipliers, precomputed on the host.
__global__ void bessel_prefilter_kernel(const double* __restrict__ J, int nell, const double* __restrict__ cp,
                                        const double* __restrict__ iden, double* __restrict__ Cf) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nell) return;
  const double* y = J + (size_t)e * BESSEL_NB;
  double* c = Cf + (size_t)e * BESSEL_NC;
  const int n = BESSEL_NB, m = n - 2;
  const double a = 1.0 / 6.0;
  const double y0 = y[0], yl = y[n - 1];
  // forward sweep, dp stored in place of the coefficients c[2..n-1]
  double dp = 0.0;
  for (int i = 0; i < m; i++) {
    double r = y[i + 1];
    if (i == 0) r -= a * y0;
    if (i == m - 1) r -= a * yl;
    dp = (r - a * dp) * iden[i];
    c[i + 2] = dp;
  }
  double cn = c[m + 1];
  for (int i = m - 2; i >= 0; i--) { cn = c[i + 2] - cp[i] * cn; c[i + 2] = cn; }
  c[1] = y0; c[n] = yl;
  c[0] = 2.0 * y0 - c[2];
  c[n + 1] = 2.0 * yl - c[n - 1];
}

// Dense midpoint grid of cltt (spectra.jl:88-93): kbar_j = (kd[j]+kd[j+1])/2, weight_j = A (kbar/0.05)^(n-1) dk/kbar,
// kd = quadratic_k(kmin,kmax,n_kd) (spectra.jl:60-63); plus the bracket in the coarse k grid (Line() extrapolation).
__global__ void dense_k_kernel(const double* __restrict__ kc, int nk, double kd_min, double kd_max, int n_kd, double A, double ns,
                               double dg, double* __restrict__ kscaled, double* __restrict__ wk, int* __restrict__ jlo,
                               double* __restrict__ wlerp) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_kd - 1) return;
  const double r0 = (double)(j + 1) / n_kd, r1 = (double)(j + 2) / n_kd;
  const double k0 = kd_min + (kd_max - kd_min) * (r0 * r0), k1 = kd_min + (kd_max - kd_min) * (r1 * r1);
  const double k = (k0 + k1) / 2.0, dk = k1 - k0;
  const double Pprim = A * pow(k / 0.05, ns - 1.0);
  kscaled[j] = k / dg;
  wk[j] = Pprim * dk / k;
  int lo = 0, hi = nk - 2;   // largest lo in [0, nk-2] with kc[lo] <= k (0 if none)
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (kc[mid] <= k) lo = mid; else hi = mid - 1; }
  jlo[j] = lo;
  wlerp[j] = (k - kc[lo]) / (kc[lo + 1] - kc[lo]);
}

// SD[i][j] = ((1-w) S[jlo][ix_start+i] + w S[jlo+1][ix_start+i]) * dx_i, layout [nrows][ld] with j fastest.
__global__ void dense_source_kernel(const double* __restrict__ S, int n_x, int ix_start, int nrows, const int* __restrict__ jlo,
                                    const double* __restrict__ wlerp, int nkd1, int ld, double x0, double dx, double* __restrict__ SD) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= nkd1 || i >= nrows) return;
  const int ix = ix_start + i;
  const double dxi = (x0 + dx * (ix + 1)) - (x0 + dx * ix);
  const double w = wlerp[j]; const int lo = jlo[j];
  const double s = (1.0 - w) * S[(size_t)lo * n_x + ix] + w * S[(size_t)(lo + 1) * n_x + ix];
  SD[(size_t)i * ld + j] = s * dxi;
}

// chi_i = eta0 - eta(x_i) for the LOS rows (spectra.jl:81)
__global__ void chi_kernel(const DevCosmo* cos, int ix_start, int nrows, double* __restrict__ chi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  const DevCosmo& c = *cos;
  const double x = c.x0 + c.dx * (ix_start + i);
  chi[i] = c.s[BOLT_S_eta0] - spline_eval(c.tab[BOLT_T_eta], c.n_x, c.x0, c.dx, x);
}

struct ProjectParams {
  const double* Cf;        // [nell][BESSEL_NC]
  const int* ells;         // [nell]
  int nell;
  const double* chi;       // [nrows]
  int nrows;
  const double* kscaled;   // [nkd1]
  const double* wk;        // [nkd1]
  int nkd1, ld;
  const double* SD_T;      // [nrows][ld] or null
  const double* SD_P;      // [nrows][ld] or null
  int nsplit;
  double* partial;         // [nell][nsplit][3]
};

#ifndef K2_WINDOW_SHIFTS
#define K2_WINDOW_SHIFTS 0
#endif
#ifndef K2_WINDOW
#define K2_WINDOW 1
#endif
#ifndef K2_ROW_UNROLL
#define K2_ROW_UNROLL 2
#endif
constexpr int k2_row_unroll = K2_ROW_UNROLL;
template <int NL, int NT>
__global__ void __launch_bounds__(NT, (NL <= 2 ? 2 : 1)) project_kernel(ProjectParams p) {
  extern __shared__ __align__(128) double smem[];
  double* tabs = smem;                                   // [NL][BESSEL_NC]
  double* chi = smem + (size_t)NL * BESSEL_NC;           // [nrows]
  __shared__ double red[3 * NL][NT / 32];
  __shared__ __align__(8) unsigned long long stage_bar;
  const int g = blockIdx.x, split = blockIdx.y;
  const int e0 = g * NL;
  stage_tables_bulk(tabs, p.Cf, e0, p.nell, NL, &stage_bar);      // 160 KB by the copy engine ...
  for (int i = threadIdx.x; i < p.nrows; i += NT) chi[i] = p.chi[i];   // ... while the threads fetch chi
  stage_tables_wait(&stage_bar);
  __syncthreads();

  const int per = (p.nkd1 + p.nsplit - 1) / p.nsplit;
  const int jbeg = split * per, jend = min(p.nkd1, jbeg + per);
  // per-warp sums in shared memory (red[.][warp]); each k-iteration adds its warp-reduced contribution (see project_kernel_dual)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane; i < 3 * NL; i += 32) red[i][warp] = 0.0;
  __syncwarp();
  const bool hasT = p.SD_T != nullptr, hasP = p.SD_P != nullptr;
  for (int j0 = jbeg; j0 < jend; j0 += NT) {         // warp-uniform trip count: every lane takes part in the reductions
    const int j = min(j0 + (int)threadIdx.x, jend - 1);
    const bool livej = j0 + (int)threadIdx.x < jend;
    const double ks = p.kscaled[j];
    double th[NL], ep[NL];
#pragma unroll
    for (int l = 0; l < NL; l++) { th[l] = 0; ep[l] = 0; }
    const double* sT = p.SD_T + j; const double* sP = p.SD_P + j;
    // Coefficient window: chi_i decreases with i, so the table cell ii never increases, and for ~40 % of the (k, x) pairs it does
    // not move at all between consecutive rows (it moves by ks * dchi / dg).  The four coefficients of each multipole are kept in
    // registers and re-read from shared memory only when the cell changes -- same operands, same operation order: bit-identical
    // to fetching all four every time.  Measured (2000-mode C3 grid, row loop unrolled by 2 in every case): no window 14.8 ms;
    // re-read on any move (this) 13.0 ms; additionally SHIFTING the window by one or two cells (K2_WINDOW_SHIFTS=1: fewer shared
    // loads, but 24-34 register moves per shift and a warp executes the union of its lanes' cases) 13.9 ms.  Unrolling the row
    // loop by 2 lets the next row's argument / weight arithmetic overlap the current row's loads (13.7 -> 13.0 ms; by 4 the
    // kernel hits the 128-register cap of a 512-thread CTA: 14.1 ms).
    double cw[NL][4];
    int ii_prev = -1000;
#pragma unroll (k2_row_unroll)
    for (int i = 0; i < p.nrows; i++) {
      const double t = ks * chi[i];
      int ii = (int)t;                       // t >= 0: truncation == floor
      ii = min(ii, BESSEL_NB - 2);
      const double d = t - (double)ii, e = 1.0 - d;
      const double d2 = d * d, e2 = e * e;
      const double w0 = e2 * e * (1.0 / 6.0);
      const double w1 = 2.0 / 3.0 - d2 + d2 * d * 0.5;
      const double w2 = 2.0 / 3.0 - e2 + e2 * e * 0.5;
      const double w3 = d2 * d * (1.0 / 6.0);
      const double vT = hasT ? __ldg(sT + (size_t)i * p.ld) : 0.0;
      const double vP = hasP ? __ldg(sP + (size_t)i * p.ld) : 0.0;
      const int delta = ii_prev - ii;
      if (!K2_WINDOW || delta != 0) {
        const double* c = tabs + ii;
#if K2_WINDOW_SHIFTS
        if (delta == 1) {
#pragma unroll
          for (int l = 0; l < NL; l++) { cw[l][3] = cw[l][2]; cw[l][2] = cw[l][1]; cw[l][1] = cw[l][0]; cw[l][0] = c[l * BESSEL_NC]; }
        } else if (delta == 2) {
#pragma unroll
          for (int l = 0; l < NL; l++) { cw[l][3] = cw[l][1]; cw[l][2] = cw[l][0]; cw[l][1] = c[l * BESSEL_NC + 1]; cw[l][0] = c[l * BESSEL_NC]; }
        } else
#endif
        {
#pragma unroll
          for (int l = 0; l < NL; l++) { cw[l][0] = c[l * BESSEL_NC]; cw[l][1] = c[l * BESSEL_NC + 1]; cw[l][2] = c[l * BESSEL_NC + 2]; cw[l][3] = c[l * BESSEL_NC + 3]; }
        }
        ii_prev = ii;
      }
#pragma unroll
      for (int l = 0; l < NL; l++) {
        const double bes = cw[l][0] * w0 + cw[l][1] * w1 + cw[l][2] * w2 + cw[l][3] * w3;
        th[l] += bes * vT;
        ep[l] += bes * vP;
      }
    }
    const double w = livej ? p.wk[j] : 0.0;
#pragma unroll
    for (int l = 0; l < NL; l++) {
      const double a = warp_sum(th[l] * th[l] * w), b = warp_sum(th[l] * ep[l] * w), c = warp_sum(ep[l] * ep[l] * w);
      if (lane == 0) { red[3 * l][warp] += a; red[3 * l + 1][warp] += b; red[3 * l + 2][warp] += c; }
    }
  }
  // block reduction (fixed order: deterministic)
  __syncthreads();
  if (threadIdx.x < 3 * NL) {
    double s = 0.0;
    for (int w = 0; w < NT / 32; w++) s += red[threadIdx.x][w];
    const int l = threadIdx.x / 3, comp = threadIdx.x - 3 * l;
    const int e = e0 + l;
    if (e < p.nell) p.partial[((size_t)e * p.nsplit + split) * 3 + comp] = s;
  }
}

// C_l = 4 pi sum_k (...)  (spectra.jl:95,113,129) with the spin factor of the E mode (spectra.jl:101,118)
__global__ void cl_finalize_kernel(const double* __restrict__ partial, const int* __restrict__ ells, int nell, int nsplit,
                                   double* __restrict__ tt, double* __restrict__ te, double* __restrict__ ee) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nell) return;
  double a = 0, b = 0, c = 0;
  for (int s = 0; s < nsplit; s++) {
    a += partial[((size_t)e * nsplit + s) * 3 + 0];
    b += partial[((size_t)e * nsplit + s) * 3 + 1];
    c += partial[((size_t)e * nsplit + s) * 3 + 2];
  }
  const double l = (double)ells[e];
  const double lfac = sqrt((l + 2.0) * (l + 1.0) * l * (l - 1.0));
  if (tt) tt[e] = 4.0 * M_PI * a;
  if (te) te[e] = 4.0 * M_PI * (b * lfac);
  if (ee) ee[e] = 4.0 * M_PI * (c * lfac * lfac);
}


// ---------------------------------------------------------------------------------------------------
// K2 with forward-mode partials.  What carries partials (src/spectra.jl:46-47,52,61,66 strip them from the k grids and
// from the Bessel table range): the sources S(x,k), chi_i = eta0 - eta(x_i) -- so the Bessel ARGUMENT k*chi_i is dual and
// the spline's derivative is needed -- and the primordial weight A (k/0.05)^(n-1).
//   dTheta_l/dp = sum_i [ j~'(t_i) * ks * dchi_i/dp * S_i + j~(t_i) * dS_i/dp ] dx_i ,   t_i = ks*chi_i  (index units)
//   dC_l/dp     = 4 pi sum_k [ 2 Theta dTheta w + Theta^2 dw/dp ]
// Component-major device layouts: SD[comp][row][k], chi[comp][row], wk[comp][k]; partial[l][split][3][nd].
// ---------------------------------------------------------------------------------------------------
__global__ void dense_source_kernel_nd(const double* __restrict__ S, int n_x, int nd, int comp, int ix_start, int nrows,
                                       const int* __restrict__ jlo, const double* __restrict__ wlerp, int nkd1, int ld, double x0,
                                       double dx, double* __restrict__ SD) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= nkd1 || i >= nrows) return;
  const int ix = ix_start + i;
  const double dxi = (x0 + dx * (ix + 1)) - (x0 + dx * ix);
  const double w = wlerp[j]; const int lo = jlo[j];
  const double s = (1.0 - w) * S[((size_t)lo * n_x + ix) * nd + comp] + w * S[((size_t)(lo + 1) * n_x + ix) * nd + comp];
  SD[(size_t)i * ld + j] = s * dxi;
}

__global__ void chi_kernel_nd(const DevCosmo* cos, int ix_start, int nrows, double* __restrict__ chi /* [nd][nrows] */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  const DevCosmo& c = *cos;
  const double x = c.x0 + c.dx * (ix_start + i);
  chi[i] = c.s[BOLT_S_eta0] - spline_eval(c.tab[BOLT_T_eta], c.n_x, c.x0, c.dx, x);
  for (int j = 0; j < c.np; j++)
    chi[(size_t)(1 + j) * nrows + i] = c.ds[BOLT_S_eta0][j] - spline_eval(c.dtab[BOLT_T_eta] + (size_t)j * (c.n_x + 2), c.n_x, c.x0, c.dx, x);
}

// partials of the k weights: d/dp [A (k/0.05)^(n-1) dk/k] = w (dA/A + ln(k/0.05) dn)
__global__ void dense_k_partials_kernel(const DevCosmo* cos, const double* __restrict__ kscaled, double dg, const double* __restrict__ wk,
                                        int nkd1, double* __restrict__ dwk /* [np][nkd1] */) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nkd1) return;
  const DevCosmo& c = *cos;
  const double k = kscaled[j] * dg, lk = log(k / 0.05), w = wk[j];
  for (int q = 0; q < c.np; q++) dwk[(size_t)q * nkd1 + j] = w * (c.ds[BOLT_S_A][q] / c.s[BOLT_S_A] + lk * c.ds[BOLT_S_n][q]);
}

struct ProjectParamsD {
  ProjectParams v;
  const double* chi_d;     // [np][nrows]
  const double* dwk;       // [np][nkd1]
  const double* SD_T_d;    // [np][nrows][ld] or null
  const double* SD_P_d;    // [np][nrows][ld] or null
  double* partial_d;       // [nell][nsplit][3][np]
};

template <int NL, int NT, int NP>
__global__ void __launch_bounds__(NT) project_kernel_dual(ProjectParamsD pd) {
  const ProjectParams& p = pd.v;
  extern __shared__ __align__(128) double smem[];
  double* tabs = smem;                                   // [NL][BESSEL_NC]
  double* chi = smem + (size_t)NL * BESSEL_NC;           // [1+NP][nrows]
  __shared__ double red[3 * NL * (1 + NP)][NT / 32];
  __shared__ __align__(8) unsigned long long stage_bar;
  const int g = blockIdx.x, split = blockIdx.y;
  const int e0 = g * NL;
  stage_tables_bulk(tabs, p.Cf, e0, p.nell, NL, &stage_bar);
  for (int i = threadIdx.x; i < p.nrows; i += NT) {
    chi[i] = p.chi[i];
#pragma unroll
    for (int q = 0; q < NP; q++) chi[(size_t)(1 + q) * p.nrows + i] = pd.chi_d[(size_t)q * p.nrows + i];
  }
  stage_tables_wait(&stage_bar);
  __syncthreads();
  const int per = (p.nkd1 + p.nsplit - 1) / p.nsplit;
  const int jbeg = split * per, jend = min(p.nkd1, jbeg + per);
  // The 3 x NL x (1+NP) accumulators of a thread (42 doubles at NL = 2, NP = 6) would be live through the whole row loop; they are
  // needed once per dense wavenumber only, so each warp keeps ITS sums in shared memory (red[.][warp]) and adds the warp-reduced
  // contribution of every k-iteration there (42 reductions per ~120k row-loop instructions): 239 -> fewer registers, 8 -> 12 warps.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane; i < 3 * NL * (1 + NP); i += 32) red[i][warp] = 0.0;
  __syncwarp();
  const bool hasT = p.SD_T != nullptr, hasP = p.SD_P != nullptr;
  const size_t cstr = (size_t)p.nrows * p.ld;
  for (int j0 = jbeg; j0 < jend; j0 += NT) {       // warp-uniform trip count: every lane takes part in the reductions
    const int j = min(j0 + (int)threadIdx.x, jend - 1);
    const bool livej = j0 + (int)threadIdx.x < jend;
    const double ks = p.kscaled[j];
    double th[NL][1 + NP], ep[NL][1 + NP];
#pragma unroll
    for (int l = 0; l < NL; l++)
#pragma unroll
      for (int q = 0; q <= NP; q++) { th[l][q] = 0.0; ep[l][q] = 0.0; }
    for (int i = 0; i < p.nrows; i++) {
      const double t = ks * chi[i];
      int ii = (int)t;
      ii = min(ii, BESSEL_NB - 2);
      const double d = t - (double)ii, e = 1.0 - d;
      const double d2 = d * d, e2 = e * e;
      const double w0 = e2 * e * (1.0 / 6.0), w1 = 2.0 / 3.0 - d2 + d2 * d * 0.5, w2 = 2.0 / 3.0 - e2 + e2 * e * 0.5, w3 = d2 * d * (1.0 / 6.0);
      const double g0 = -0.5 * e2, g1 = -2.0 * d + 1.5 * d2, g2 = 2.0 * e - 1.5 * e2, g3 = 0.5 * d2;     // d(weights)/dt
      const size_t off = (size_t)i * p.ld + j;
      const double vT = hasT ? __ldg(p.SD_T + off) : 0.0, vP = hasP ? __ldg(p.SD_P + off) : 0.0;
      double tq[NP], dT[NP], dP[NP];
#pragma unroll
      for (int q = 0; q < NP; q++) {
        tq[q] = ks * chi[(size_t)(1 + q) * p.nrows + i];
        dT[q] = hasT ? __ldg(pd.SD_T_d + (size_t)q * cstr + off) : 0.0;
        dP[q] = hasP ? __ldg(pd.SD_P_d + (size_t)q * cstr + off) : 0.0;
      }
#pragma unroll
      for (int l = 0; l < NL; l++) {
        const double* c = tabs + l * BESSEL_NC + ii;
        const double c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3];
        const double bes = c0 * w0 + c1 * w1 + c2 * w2 + c3 * w3;
        const double dbes = c0 * g0 + c1 * g1 + c2 * g2 + c3 * g3;
        th[l][0] += bes * vT; ep[l][0] += bes * vP;
#pragma unroll
        for (int q = 0; q < NP; q++) {
          th[l][1 + q] += dbes * tq[q] * vT + bes * dT[q];
          ep[l][1 + q] += dbes * tq[q] * vP + bes * dP[q];
        }
      }
    }
    const double w = livej ? p.wk[j] : 0.0;
    auto add = [&](int a, int l, int q, double v) {
      const double sv = warp_sum(v);
      if (lane == 0) red[(a * NL + l) * (1 + NP) + q][warp] += sv;
    };
#pragma unroll
    for (int l = 0; l < NL; l++) {
      const double T0 = th[l][0], E0 = ep[l][0];
      add(0, l, 0, T0 * T0 * w); add(1, l, 0, T0 * E0 * w); add(2, l, 0, E0 * E0 * w);
#pragma unroll
      for (int q = 0; q < NP; q++) {
        const double dw = livej ? pd.dwk[(size_t)q * p.nkd1 + j] : 0.0;
        add(0, l, 1 + q, 2.0 * T0 * th[l][1 + q] * w + T0 * T0 * dw);
        add(1, l, 1 + q, (th[l][1 + q] * E0 + T0 * ep[l][1 + q]) * w + T0 * E0 * dw);
        add(2, l, 1 + q, 2.0 * E0 * ep[l][1 + q] * w + E0 * E0 * dw);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 3 * NL * (1 + NP)) {
    double s = 0.0;
    for (int w = 0; w < NT / 32; w++) s += red[threadIdx.x][w];
    const int q = threadIdx.x % (1 + NP), al = threadIdx.x / (1 + NP), l = al % NL, a = al / NL;
    const int e = e0 + l;
    if (e < p.nell) {
      if (q == 0) p.partial[((size_t)e * p.nsplit + split) * 3 + a] = s;
      else pd.partial_d[(((size_t)e * p.nsplit + split) * 3 + a) * NP + (q - 1)] = s;
    }
  }
}

// C_l and its partials: out arrays [nell][nd]
__global__ void cl_finalize_kernel_nd(const double* __restrict__ partial, const double* __restrict__ partial_d, const int* __restrict__ ells,
                                      int nell, int nsplit, int np, double* __restrict__ tt, double* __restrict__ te, double* __restrict__ ee) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nell) return;
  const int nd = 1 + np;
  const double l = (double)ells[e];
  const double lfac = sqrt((l + 2.0) * (l + 1.0) * l * (l - 1.0));
  for (int q = 0; q < nd; q++) {
    double a = 0, b = 0, c = 0;
    for (int s = 0; s < nsplit; s++) {
      const size_t base = ((size_t)e * nsplit + s) * 3;
      if (q == 0) { a += partial[base]; b += partial[base + 1]; c += partial[base + 2]; }
      else { a += partial_d[base * np + (q - 1)]; b += partial_d[(base + 1) * np + (q - 1)]; c += partial_d[(base + 2) * np + (q - 1)]; }
    }
    if (tt) tt[(size_t)e * nd + q] = 4.0 * M_PI * a;
    if (te) te[(size_t)e * nd + q] = 4.0 * M_PI * (b * lfac);
    if (ee) ee[(size_t)e * nd + q] = 4.0 * M_PI * (c * lfac * lfac);
  }
}

// plin from the state at x = 0 (spectra.jl:163-198), one thread per k
__global__ void plin_kernel(const DevCosmo* cos, const double* __restrict__ kk, int nk, const double* __restrict__ u_final,
                            int L, int Lnu, int Lm, double* __restrict__ pk) {
  const int ik = blockIdx.x * blockDim.x + threadIdx.x;
  if (ik >= nk) return;
  const DevCosmo& c = *cos;
  const int nq = c.nq;
  const int iM = 2 * (L + 1) + (Lnu + 1), iS = iM + (Lm + 1) * nq, n = iS + 5;
  const double* res = u_final + (size_t)ik * n;
  const double k = kk[ik], x = 0.0, a = 1.0;
  const double rho0M = spline_eval(c.tab[BOLT_T_rho0M], c.n_x, c.x0, c.dx, x);
  const double Hx = spline_eval(c.tab[BOLT_T_H], c.n_x, c.x0, c.dx, x);
  const double m = c.s[BOLT_S_Sum_m_nu];
  double rho = 0.0, th = 0.0;
  for (int i = 0; i < nq; i++) {
    const double q = c.q[i], eps = sqrt(q * q + (a * m) * (a * m));
    rho += c.wq[i] * eps * res[iM + i];                 // rho_sigma (:170-172)
    th += c.wq[i] * q * res[iM + nq + i];               // theta     (:174-175)
  }
  const double Mrho = rho / rho0M;
  const double Mtheta = k * th / rho0M;
  const double dcN = res[iS + 1], dbN = res[iS + 3], vcN = res[iS + 2], vbN = res[iS + 4];
  const double vmnuN = -Mtheta / k;
  const double hh = c.s[BOLT_S_h], Om_r = c.s[BOLT_S_Omega_r], N_nu = c.s[BOLT_S_N_nu];
  const double Tg = pow(15.0 / (M_PI * M_PI) * c.s[BOLT_S_rho_crit] * Om_r, 0.25);
  const double zeta = 1.2020569;
  const double nufac = (90.0 * zeta / (11.0 * pow(M_PI, 4.0))) * (Om_r * hh * hh / Tg) * pow(N_nu / 3.0, 0.75);
  const double Om_nu = m * nufac / (hh * hh);
  const double Om_c = c.s[BOLT_S_Omega_c], Om_b = c.s[BOLT_S_Omega_b];
  const double Om_m = Om_c + Om_b + Om_nu;
  const double dc = dcN - 3.0 * Hx * vcN / k, db = dbN - 3.0 * Hx * vbN / k;
  const double dmnu = Mrho - 3.0 * Hx * vmnuN / k;
  const double dm = (Om_c * dc + Om_b * db + Om_nu * dmnu) / Om_m;
  const double Pprim = c.s[BOLT_S_A] * pow(k / 0.05, c.s[BOLT_S_n] - 1.0);
  pk[ik] = (2.0 * M_PI * M_PI / (k * k * k)) * dm * dm * Pprim;
}

}  // namespace bolt
