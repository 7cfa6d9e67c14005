// Batched input-table generator on the device: Background + RECFAST + reionization + optical depth for a
// BATCH of cosmologies, producing exactly what bolt_cosmo_upload consumes (12 cubic-B-spline coefficient tables + 13 scalars per
// cosmology) -- what the reference computes on the host, one cosmology at a time, in
//   src/background.jl:5-128 (rho_P_0, H_a, eta, Background), src/ionization/recfast.jl:22-536 (RECFAST, recfastsolve, Xe/Tmat
//   accessors, tanh reionization), src/ionization/ionization.jl:107-137 (tau, tau', g), src/ionization/recfast.jl:674-726
//   (IonizationHistory), src/util.jl:11-13 (spline, spline_d, spline_dd).
// With K1/K2 at ~57 ms per spectrum set the ~2 s per cosmology of a host generator dominates an emulator / MCMC batch
// (BASELINE config 5); here a thread integrates the recombination ODEs of one cosmology, so a batch costs what one costs.
//
// Numerics.  The reference integrates RECFAST with Tsit5 + dense output and finds the two switch redshifts with a Falsi root
// finder; here: Dormand-Prince 5(4) with steps that END on the x-grid redshifts (no dense output), bisection for the roots, the
// post-reionization matter temperature integrated together with the un-reionized solution it reads.  Same equations, same
// switches; agreement with the harness generator (hostgen/, pinned by the reference's Fortran RECFAST fixture) is asserted in
// tests/test_gpu_hostgen.py.  Value-only (no partials).  Standalone translation unit: no context needed.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/bolt_cuda.h"

namespace {

constexpr int HG_MAXQ = 32;
thread_local std::string g_hg_err;

struct RecConsts {      // RECFAST's constant block for one cosmology (recfast.jl:22-121)
  double C, k_B, h_P, m_H, not4, Lambda, Lambda_He, A2P_s, A2P_t, L_He_2p, L_He_2Pt, L_He_2St, L_He2St_ion, sigma_He_2Ps, sigma_He_2Pt;
  double AGauss1, AGauss2, zGauss1, zGauss2, wGauss1, wGauss2, a_PPB, b_PPB, c_PPB, d_PPB, a_VF, b_VF, T_0, T_1, a_trip, b_trip;
  double CDB, CDB_He, CB1, CB1_He1, CB1_He2, CR, CK, CK_He, CL, CL_He, CT, Bfact, H_frac, fu, b_He;
  double Yp, HO, Tnow, mu_H, mu_T, fHe, Nnow;
  int Hswitch, Heswitch;
};

struct HgCosmo {        // per cosmology
  double h, Om_r, Om_b, Om_c, A, ns, Yp, N_nu, m_nu;       // CosmoParams (src/Bolt.jl:56-66)
  double H0, rho_crit, Om_L, T_nu, eta0;
  RecConsts r;
};

struct HgGrid { double x0, dx; int n_x, nq; double pts[HG_MAXQ], wts[HG_MAXQ]; };

// unit constants of the reference's "Mpc units" (src/Bolt.jl:49-51, ionization.jl:34-39, recfast.jl:4-5; CODATA 2018)
constexpr double C_SI = 299792458.0, HBAR_SI = 6.62607015e-34 / (2.0 * 3.14159265358979323846), KB_SI = 1.380649e-23, EV_SI = 1.602176634e-19;
constexpr double G_SI = 6.67430e-11, MPC_SI = 1.0e6 * (149597870700.0 * 648000.0 / 3.14159265358979323846);
constexpr double km_s_Mpc_100 = 100.0e3 / C_SI;
constexpr double G_natural = G_SI * HBAR_SI / (C_SI * C_SI * C_SI) / (MPC_SI * MPC_SI);
constexpr double m_H_nat = 1.67262192369e-27 * C_SI / HBAR_SI * MPC_SI;
constexpr double sigma_T_nat = 6.6524587321e-29 / (MPC_SI * MPC_SI);
constexpr double H0_unit = MPC_SI / C_SI;                       // one natural time unit in seconds
constexpr double Kelvin_unit = HBAR_SI * C_SI / (KB_SI * MPC_SI);
constexpr double ZETA3 = 1.2020569;
constexpr double PI = 3.14159265358979323846;

// ---------------------------------------------------------------- background (background.jl:21-77) ----------------------
__host__ __device__ inline double hg_to_ui(double lq, double lqmi, double lqma) { return -1.0 + 2.0 / (lqma - lqmi) * (lq - lqmi); }
__host__ __device__ inline double hg_from_ui(double x, double lqmi, double lqma) { return lqmi + (lqma - lqmi) / 2.0 * (x + 1.0); }
__host__ __device__ inline double hg_dxdq(double q, double lqmi, double lqma) { return (1.0 + hg_to_ui(1.0 + lqmi, lqmi, lqma)) / (q * log(10.0)); }
__host__ __device__ inline double hg_xq2q(double x, double lqmi, double lqma) { return pow(10.0, hg_from_ui(x, lqmi, lqma)); }

__host__ __device__ inline double hg_rho_nu(const HgCosmo& c, const HgGrid& g, double a) {       // rhoP_0, background.jl:33-48 (density only)
  const double lqmi = log10(c.T_nu / 30.0), lqma = log10(c.T_nu * 30.0);
  double s = 0.0;
  for (int i = 0; i < g.nq; i++) {
    const double q = hg_xq2q(g.pts[i], lqmi, lqma);
    const double eps = sqrt(q * q + (a * c.m_nu) * (a * c.m_nu));
    const double f0 = 2.0 / (8.0 * PI * PI * PI) / (exp(q / c.T_nu) + 1.0);
    s += q * q * eps * (f0 / hg_dxdq(q, lqmi, lqma) * g.wts[i]);
  }
  return 4.0 * PI / (a * a * a * a) * s;
}
__host__ __device__ inline double hg_calH(const HgCosmo& c, const HgGrid& g, double a) {         // a * H_a, background.jl:58-66
  const double rad = c.Om_r * (1.0 + (2.0 / 3.0) * (7.0 * c.N_nu / 8.0) * pow(4.0 / 11.0, 4.0 / 3.0));
  return a * c.H0 * sqrt((c.Om_c + c.Om_b) / (a * a * a) + hg_rho_nu(c, g, a) / c.rho_crit + rad / (a * a * a * a) + c.Om_L);
}
__host__ __device__ inline double hg_eta(const HgCosmo& c, const HgGrid& g, double x) {          // background.jl:73-77
  const double logamin = -13.75, logamax = log10(exp(x));
  double s = 0.0;
  for (int i = 0; i < g.nq; i++) {
    const double ap = hg_xq2q(g.pts[i], logamin, logamax);
    s += 1.0 / (ap * hg_calH(c, g, ap)) / hg_dxdq(ap, logamin, logamax) * g.wts[i];
  }
  return s;
}

// samples on the x grid into the series calH, eta, rho0M of Y[cos][BOLT_NTABLES][n_x]
__global__ void hg_background_kernel(const HgCosmo* __restrict__ cos, HgGrid g, double* __restrict__ Y) {
  const int ic = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.n_x) return;
  const HgCosmo& c = cos[ic];
  const double x = g.x0 + g.dx * i, a = exp(x);
  double* y = Y + (size_t)ic * BOLT_NTABLES * g.n_x;
  y[(size_t)BOLT_T_H * g.n_x + i] = hg_calH(c, g, a);
  y[(size_t)BOLT_T_eta * g.n_x + i] = hg_eta(c, g, x);
  y[(size_t)BOLT_T_rho0M * g.n_x + i] = hg_rho_nu(c, g, a);
}

// ---------------------------------------------------------------- splines (util.jl:11-13) -------------------------------
// Interpolations.jl prefilter of BSpline(Cubic(Line(OnGrid()))): one thread per (cosmology, table) series.  cp / iden: the
// Thomas multipliers of the (1/6, 2/3, 1/6) interior system (they depend on n only).
__global__ void hg_prefilter_kernel(const double* __restrict__ Y, int nseries, int n, const double* __restrict__ cp, const double* __restrict__ iden,
                                    double* __restrict__ Cf, int ystride, int cstride) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseries) return;
  const double* y = Y + (size_t)s * ystride;
  double* c = Cf + (size_t)s * cstride;
  const int m = n - 2;
  const double a = 1.0 / 6.0, y0 = y[0], yl = y[n - 1];
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
// gradient / hessian of a spline AT its knots (what spline_d / spline_dd resample): out[i] from coefficients c[0..n+1]
__global__ void hg_knot_derivs_kernel(const double* __restrict__ Cf, int nseries, int n, int cstride, double dx, double* __restrict__ G,
                                      double* __restrict__ H, int ostride) {
  const int s = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseries || i >= n) return;
  const double* c = Cf + (size_t)s * cstride;
  if (G) G[(size_t)s * ostride + i] = (c[i + 2] - c[i]) / (2.0 * dx);
  if (H) H[(size_t)s * ostride + i] = (c[i] - 2.0 * c[i + 1] + c[i + 2]) / (dx * dx);
}
__device__ inline double hg_spline(const double* __restrict__ c, int n, double x0, double dx, double x) {
  double t = (x - x0) / dx;
  int i = (int)floor(t);
  i = max(0, min(i, n - 2));
  const double d = t - (double)i, e = 1.0 - d;
  return c[i] * (e * e * e / 6.0) + c[i + 1] * (2.0 / 3.0 - d * d + d * d * d / 2.0) + c[i + 2] * (2.0 / 3.0 - e * e + e * e * e / 2.0) + c[i + 3] * (d * d * d / 6.0);
}

// ---------------------------------------------------------------- RECFAST (recfast.jl:156-536) --------------------------
struct HSpl { const double* cH; const double* cHp; int n; double x0, dx; };
__device__ inline void hg_Hz(const HSpl& hs, double z, double& Hz, double& dHdz) {           // recfast.jl:313-323
  const double a = 1.0 / (1.0 + z), xa = log(a);
  const double Hc = hg_spline(hs.cH, hs.n, hs.x0, hs.dx, xa);
  Hz = Hc / a / H0_unit;
  dHdz = (-hg_spline(hs.cHp, hs.n, hs.x0, hs.dx, xa) + Hc) / H0_unit;
}
// ion_recfast (recfast.jl:156-310): y = (x_H, x_He, Tmat) -> f = dy/dz.  false: trial state outside the domain (step is rejected).
__device__ bool hg_ion_recfast(const RecConsts& r, const HSpl& hs, double z, const double (&y)[3], double (&f)[3]) {
  const double x_H = y[0], x_He = y[1], Tmat = y[2];
  if (!(Tmat > 0.0) || !(x_H + r.fHe * x_He > 0.0)) return false;
  const double x = x_H + r.fHe * x_He;
  const double zp1 = 1.0 + z;
  const double n = r.Nnow * zp1 * zp1 * zp1, n_He = r.fHe * n;
  const double Trad = r.Tnow * zp1;
  double Hz, dHdz; hg_Hz(hs, z, Hz, dHdz);
  const double Rdown = 1e-19 * r.a_PPB * pow(Tmat / 1e4, r.b_PPB) / (1.0 + r.c_PPB * pow(Tmat / 1e4, r.d_PPB));
  const double CRT15 = (r.CR * Tmat) * sqrt(r.CR * Tmat);
  const double Rup = Rdown * CRT15 * exp(-r.CDB / Tmat);
  const double sq_0 = sqrt(Tmat / r.T_0), sq_1 = sqrt(Tmat / r.T_1);
  double Rdown_He = r.a_VF / (sq_0 * pow(1.0 + sq_0, 1.0 - r.b_VF));
  Rdown_He = Rdown_He / pow(1.0 + sq_1, 1.0 + r.b_VF);
  const double Rup_He = 4.0 * Rdown_He * CRT15 * exp(-r.CDB_He / Tmat);
  const double He_Boltz = (r.Bfact / Tmat > 680.0) ? exp(680.0) : exp(r.Bfact / Tmat);
  double Kc = r.CK / Hz;
  if (r.Hswitch != 0) {
    const double l = log(zp1), g1 = (l - r.zGauss1) / r.wGauss1, g2 = (l - r.zGauss2) / r.wGauss2;
    Kc *= 1.0 + r.AGauss1 * exp(-g1 * g1) + r.AGauss2 * exp(-g2 * g2);
  }
  double Rdown_trip = r.a_trip / (sq_0 * pow(1.0 + sq_0, 1.0 - r.b_trip));
  Rdown_trip = Rdown_trip / pow(1.0 + sq_1, 1.0 + r.b_trip);
  double Rup_trip = Rdown_trip * exp(-r.h_P * r.C * r.L_He2St_ion / (r.k_B * Tmat));
  Rup_trip = Rup_trip * CRT15 * (4.0 / 3.0);
  const int Heflag = ((x_He < 5.e-9) || (x_He > 0.980)) ? 0 : r.Heswitch;
  double CfHe_t = 0.0, K_He;
  if (Heflag == 0) K_He = r.CK_He / Hz;
  else {
    const double tauHe_s = r.A2P_s * r.CK_He * 3.0 * n_He * (1.0 - x_He) / Hz;
    const double pHe_s = (1.0 - exp(-tauHe_s)) / tauHe_s;
    K_He = 1.0 / (r.A2P_s * pHe_s * 3.0 * n_He * (1.0 - x_He));
    if (((Heflag == 2) || (Heflag >= 5)) && (x_H < 0.9999999)) {
      double Doppler = 2.0 * r.k_B * Tmat / (r.m_H * r.not4 * r.C * r.C);
      Doppler = r.C * r.L_He_2p * sqrt(Doppler);
      const double gamma_2Ps = 3.0 * r.A2P_s * r.fHe * (1.0 - x_He) * r.C * r.C / (sqrt(PI) * r.sigma_He_2Ps * 8.0 * PI * Doppler * (1.0 - x_H)) /
                               ((r.C * r.L_He_2p) * (r.C * r.L_He_2p));
      const double AHcon = r.A2P_s / (1.0 + 0.36 * pow(gamma_2Ps, r.b_He));
      K_He = 1.0 / ((r.A2P_s * pHe_s + AHcon) * 3.0 * n_He * (1.0 - x_He));
    }
    if (Heflag >= 3) {
      double tauHe_t = r.A2P_t * n_He * (1.0 - x_He) * 3.0;
      tauHe_t = tauHe_t / (8.0 * PI * Hz * r.L_He_2Pt * r.L_He_2Pt * r.L_He_2Pt);
      const double pHe_t = (1.0 - exp(-tauHe_t)) / tauHe_t;
      const double CL_PSt = r.h_P * r.C * (r.L_He_2Pt - r.L_He_2St) / r.k_B;
      if ((Heflag == 3) || (Heflag == 5) || (x_H > 0.99999)) {
        CfHe_t = r.A2P_t * pHe_t * exp(-CL_PSt / Tmat);
        CfHe_t = CfHe_t / (Rup_trip + CfHe_t);
      } else {
        double Doppler = 2.0 * r.k_B * Tmat / (r.m_H * r.not4 * r.C * r.C);
        Doppler = r.C * r.L_He_2Pt * sqrt(Doppler);
        const double gamma_2Pt = (3.0 * r.A2P_t * r.fHe * (1.0 - x_He) * r.C * r.C / (sqrt(PI) * r.sigma_He_2Pt * 8.0 * PI * Doppler * (1.0 - x_H)) /
                                  ((r.C * r.L_He_2Pt) * (r.C * r.L_He_2Pt)));
        const double AHcon = r.A2P_t / (1.0 + 0.66 * pow(gamma_2Pt, 0.9)) / 3.0;
        CfHe_t = (r.A2P_t * pHe_t + AHcon) * exp(-CL_PSt / Tmat);
        CfHe_t = CfHe_t / (Rup_trip + CfHe_t);
      }
    }
  }
  const double timeTh = (1.0 / (r.CT * Trad * Trad * Trad * Trad)) * (1.0 + x + r.fHe) / x;
  const double timeH = 2.0 / (3.0 * r.HO * zp1 * sqrt(zp1));
  double f1, f2, f3;
  if (x_H > 0.99) f1 = 0.0;
  else if (x_H > 0.985) f1 = (x * x_H * n * Rdown - Rup * (1.0 - x_H) * exp(-r.CL / Tmat)) / (Hz * zp1);
  else
    f1 = ((x * x_H * n * Rdown - Rup * (1.0 - x_H) * exp(-r.CL / Tmat)) * (1.0 + Kc * r.Lambda * n * (1.0 - x_H))) /
         (Hz * zp1 * (1.0 / r.fu + Kc * r.Lambda * n * (1.0 - x_H) / r.fu + Kc * Rup * n * (1.0 - x_H)));
  if (x_He < 1e-15) f2 = 0.0;
  else {
    f2 = ((x * x_He * n * Rdown_He - Rup_He * (1.0 - x_He) * exp(-r.CL_He / Tmat)) * (1.0 + K_He * r.Lambda_He * n_He * (1.0 - x_He) * He_Boltz)) /
         (Hz * zp1 * (1.0 + K_He * (r.Lambda_He + Rup_He) * n_He * (1.0 - x_He) * He_Boltz));
    if (Heflag >= 3)
      f2 += (x * x_He * n * Rdown_trip - (1.0 - x_He) * 3.0 * Rup_trip * exp(-r.h_P * r.C * r.L_He_2St / (r.k_B * Tmat))) * CfHe_t / (Hz * zp1);
  }
  if (timeTh < r.H_frac * timeH) {
    const double epsilon = Hz * (1.0 + x + r.fHe) / (r.CT * Trad * Trad * Trad * x);
    f3 = r.Tnow + epsilon * ((1.0 + r.fHe) / (1.0 + r.fHe + x)) * ((f1 + r.fHe * f2) / x) - epsilon * dHdz / Hz + 3.0 * epsilon / zp1;
  } else
    f3 = r.CT * (Trad * Trad * Trad * Trad) * x / (1.0 + x + r.fHe) * (Tmat - Trad) / (Hz * zp1) + 2.0 * Tmat / zp1;
  f[0] = f1; f[1] = f2; f[2] = f3;
  return true;
}
__device__ inline double hg_saha_rhs(const RecConsts& r, double z, double CB) { return exp(1.5 * log(r.CR * r.Tnow / (1.0 + z)) - CB / (r.Tnow * (1.0 + z))) / r.Nnow; }
__device__ inline double hg_xH_saha(const RecConsts& r, double z) { const double s = hg_saha_rhs(r, z, r.CB1); return 0.5 * (sqrt(s * s + 4.0 * s) - s); }         // recfast.jl:380-384
__device__ inline double hg_xHe_saha(const RecConsts& r, double z) {
  const double s = 4.0 * hg_saha_rhs(r, z, r.CB1_He1);
  return 0.5 * (sqrt((s - 1.0) * (s - 1.0) + 4.0 * (1.0 + r.fHe) * s) - (s - 1.0));
}
// reionization_Xe (recfast.jl:478-488; z_re hard-coded there)
__device__ inline double hg_reio_Xe(const RecConsts& r, double z, double x_orig) {
  const double X_fin = 1.0 + r.Yp / (r.not4 * (1.0 - r.Yp));
  const double zre = 7.6711, al = 1.5, dH = 0.5, zHe = 3.5, dHe = 0.5, fHe = X_fin - 1.0;
  const double x_reio_H = (X_fin - x_orig) / 2.0 * (1.0 + tanh((pow(1.0 + zre, al) - pow(1.0 + z, al)) / (al * pow(1.0 + zre, al - 1.0)) / dH)) + x_orig;
  return x_reio_H + fHe / 2.0 * (1.0 + tanh((zHe - z) / dHe));
}

// The generalised right-hand side of the device integration.  phase 0: He evolution with Saha hydrogen (ion_recfast_H_Saha,
// recfast.jl:360-366: state x_He, Tmat); phase 1: full (x_H, x_He, Tmat); phase 2: phase 1 plus the post-reionization matter
// temperature (recfast.jl:506-536), which reads the un-reionized X_e of the same state.
__device__ bool hg_rhs(const RecConsts& r, const HSpl& hs, int phase, double z, const double (&u)[4], double (&du)[4]) {
  double y[3], f[3];
  if (phase == 0) { y[0] = hg_xH_saha(r, z); y[1] = u[0]; y[2] = u[1]; }
  else { y[0] = u[0]; y[1] = u[1]; y[2] = u[2]; }
  if (!hg_ion_recfast(r, hs, z, y, f)) return false;
  if (phase == 0) { du[0] = f[1]; du[1] = f[2]; du[2] = 0.0; du[3] = 0.0; return true; }
  du[0] = f[0]; du[1] = f[1]; du[2] = f[2]; du[3] = 0.0;
  if (phase == 2) {
    const double x_reio = hg_reio_Xe(r, z, u[0] + r.fHe * u[1]);
    double Hz, dH; hg_Hz(hs, z, Hz, dH);
    const double Trad = r.Tnow * (1.0 + z);
    du[3] = r.CT * Trad * Trad * Trad * Trad * x_reio / (1.0 + x_reio + r.fHe) * (u[3] - Trad) / (Hz * (1.0 + z)) + 2.0 * u[3] / (1.0 + z);
  }
  return true;
}

// Dormand-Prince 5(4) from z to z_end (z_end < z), adaptive, ending exactly on z_end.  nv active variables.  hs: step memory.
__device__ bool hg_integrate(const RecConsts& r, const HSpl& hs, int phase, int nv, double& z, double z_end, double (&u)[4], double& hstep,
                             double rtol, double atol) {
  const double a21 = 1.0 / 5, a31 = 3.0 / 40, a32 = 9.0 / 40, a41 = 44.0 / 45, a42 = -56.0 / 15, a43 = 32.0 / 9;
  const double a51 = 19372.0 / 6561, a52 = -25360.0 / 2187, a53 = 64448.0 / 6561, a54 = -212.0 / 729;
  const double a61 = 9017.0 / 3168, a62 = -355.0 / 33, a63 = 46732.0 / 5247, a64 = 49.0 / 176, a65 = -5103.0 / 18656;
  const double b1 = 35.0 / 384, b3 = 500.0 / 1113, b4 = 125.0 / 192, b5 = -2187.0 / 6784, b6 = 11.0 / 84;
  const double e1 = b1 - 5179.0 / 57600, e3 = b3 - 7571.0 / 16695, e4 = b4 - 393.0 / 640, e5 = b5 + 92097.0 / 339200, e6 = b6 - 187.0 / 2100, e7 = -1.0 / 40;
  double k1[4], k2[4], k3[4], k4[4], k5[4], k6[4], k7[4], t[4], un[4];
  if (!hg_rhs(r, hs, phase, z, u, k1)) return false;
  int guard = 0;
  while (z > z_end && guard++ < 200000) {
    double h = -fmin(hstep, z - z_end);                 // negative: z decreases
    const bool to_end = (z + h <= z_end * (1.0 + 1e-15) + 1e-300) || (hstep >= z - z_end);
    if (to_end) h = z_end - z;
    bool ok = true;
    for (int i = 0; i < 4; i++) t[i] = u[i] + h * a21 * k1[i];
    ok = ok && hg_rhs(r, hs, phase, z + h / 5, t, k2);
    if (ok) { for (int i = 0; i < 4; i++) t[i] = u[i] + h * (a31 * k1[i] + a32 * k2[i]); ok = hg_rhs(r, hs, phase, z + 3 * h / 10, t, k3); }
    if (ok) { for (int i = 0; i < 4; i++) t[i] = u[i] + h * (a41 * k1[i] + a42 * k2[i] + a43 * k3[i]); ok = hg_rhs(r, hs, phase, z + 4 * h / 5, t, k4); }
    if (ok) { for (int i = 0; i < 4; i++) t[i] = u[i] + h * (a51 * k1[i] + a52 * k2[i] + a53 * k3[i] + a54 * k4[i]); ok = hg_rhs(r, hs, phase, z + 8 * h / 9, t, k5); }
    if (ok) { for (int i = 0; i < 4; i++) t[i] = u[i] + h * (a61 * k1[i] + a62 * k2[i] + a63 * k3[i] + a64 * k4[i] + a65 * k5[i]); ok = hg_rhs(r, hs, phase, z + h, t, k6); }
    if (ok) { for (int i = 0; i < 4; i++) un[i] = u[i] + h * (b1 * k1[i] + b3 * k3[i] + b4 * k4[i] + b5 * k5[i] + b6 * k6[i]); ok = hg_rhs(r, hs, phase, z + h, un, k7); }
    double err = 1e300;
    if (ok) {
      err = 0.0;
      for (int i = 0; i < nv; i++) {
        const double e = h * (e1 * k1[i] + e3 * k3[i] + e4 * k4[i] + e5 * k5[i] + e6 * k6[i] + e7 * k7[i]);
        const double sc = atol + rtol * fmax(fabs(u[i]), fabs(un[i]));
        err = fmax(err, fabs(e) / sc);
      }
      if (!(err == err)) { err = 1e300; ok = false; }
    }
    if (ok && err <= 1.0) {
      z = to_end ? z_end : z + h;
      for (int i = 0; i < 4; i++) { u[i] = un[i]; k1[i] = k7[i]; }           // FSAL
      const double fac = fmin(5.0, fmax(0.2, 0.9 * pow(fmax(err, 1e-10), -0.2)));
      if (!to_end || fac < 1.0) hstep = fabs(h) * fac; else hstep = fmax(hstep, fabs(h) * fac);
    } else {
      hstep = fabs(h) * (ok ? fmax(0.2, 0.9 * pow(err, -0.2)) : 0.25);
      if (!(hstep > 1e-12 * fmax(1.0, z))) return false;
    }
  }
  return z <= z_end;
}

// One thread per cosmology: X_e(x_i) and T_mat(x_i) on the whole x grid (recfastsolve + Xe_RECFAST/Tmat_RECFAST + the tanh
// reionization history + IonizationHistory's sampling loop, recfast.jl:392-475, 506-536, 680-690).
__global__ void hg_recfast_kernel(const HgCosmo* __restrict__ cos, int ncos, HgGrid g, const double* __restrict__ CfH, const double* __restrict__ CfHp,
                                  int cstride, double* __restrict__ S, int* __restrict__ status) {
  const int ic = blockIdx.x * blockDim.x + threadIdx.x;
  if (ic >= ncos) return;
  const RecConsts& r = cos[ic].r;
  HSpl hs{CfH + (size_t)ic * cstride, CfHp + (size_t)ic * cstride, g.n_x, g.x0, g.dx};
  double* Xe = S + (size_t)ic * 3 * g.n_x; double* Tm = Xe + g.n_x;
  const double zinitial = 10000.0, zfinal = 0.0, zre_ini = 50.0;
  // switch redshifts (recfast.jl:350-357, 386, 451-456): bisection to machine precision
  auto root = [&](int which, double lo, double hi) {
    auto fn = [&](double z) { return which == 0 ? (hg_xHe_saha(r, z) - 1.0) / r.fHe - 0.99 : hg_xH_saha(r, z) - 0.985; };
    const double flo = fn(lo);
    for (int it = 0; it < 200; it++) {
      const double mid = 0.5 * (lo + hi);
      if (mid == lo || mid == hi) break;
      if ((fn(mid) > 0.0) == (flo > 0.0)) lo = mid; else hi = mid;
    }
    return 0.5 * (lo + hi);
  };
  const double z_He = root(0, zfinal, fmin(zinitial, 3500.0));
  const double z_HHe = root(1, zfinal, z_He);
  const double xinitial = log(1.0 / (1.0 + zinitial));
  const double Xe_initial = 1.0 + 2.0 * r.fHe;                      // Xe_RECFAST(zinitial > 8000)
  const double rtol = 1e-9, atol = 1e-13;
  double u[4] = {0, 0, 0, 0}, z = 0.0, hstep = 1.0;
  int phase = -1, ok = 1;
  for (int i = 0; i < g.n_x; i++) {
    const double x = g.x0 + g.dx * i;
    if (x < xinitial) { Xe[i] = Xe_initial; Tm[i] = r.Tnow * (1.0 + (1.0 / exp(x) - 1.0)); continue; }
    const double zi = 1.0 / exp(x) - 1.0;
    double xe, tm;
    if (zi > z_He) {            // analytic branches of Xe_RECFAST (recfast.jl:392-403)
      if (zi > 8000.0) xe = 1.0 + 2.0 * r.fHe;
      else if (zi > 5000.0) { const double s = hg_saha_rhs(r, zi, r.CB1_He2); xe = 0.5 * (sqrt((s - 1.0 - r.fHe) * (s - 1.0 - r.fHe) + 4.0 * (1.0 + 2.0 * r.fHe) * s) - (s - 1.0 - r.fHe)); }
      else if (zi > 3500.0) xe = 1.0 + r.fHe;
      else xe = hg_xHe_saha(r, zi);
      tm = r.Tnow * (1.0 + zi);
    } else {
      if (phase < 0) {           // init_He_evolution (recfast.jl:368-376)
        phase = 0; z = z_He; u[0] = (hg_xHe_saha(r, z_He) - 1.0) / r.fHe; u[1] = r.Tnow * (1.0 + z_He); hstep = 1.0;
      }
      if (phase == 0 && zi <= z_HHe) {      // hand over to the full system at z_H_He_evo_start (recfast.jl:463-471)
        if (z > z_HHe) ok &= hg_integrate(r, hs, 0, 2, z, z_HHe, u, hstep, rtol, atol) ? 1 : 0;
        const double xhe = u[0], t0 = u[1];
        u[0] = hg_xH_saha(r, z_HHe); u[1] = xhe; u[2] = t0; phase = 1; hstep = fmin(hstep, 1.0);
      }
      if (phase == 1 && zi <= zre_ini) {    // the reionization temperature starts from Tmat_RECFAST(zre_ini) (recfast.jl:530)
        if (z > zre_ini) ok &= hg_integrate(r, hs, 1, 3, z, zre_ini, u, hstep, rtol, atol) ? 1 : 0;
        u[3] = u[2]; phase = 2;
      }
      if (z > zi) ok &= hg_integrate(r, hs, phase, phase == 0 ? 2 : (phase == 1 ? 3 : 4), z, fmax(zi, 0.0), u, hstep, rtol, atol) ? 1 : 0;
      if (phase == 0) { xe = hg_xH_saha(r, zi) + r.fHe * u[0]; tm = u[1]; }
      else if (phase == 1) { xe = u[0] + r.fHe * u[1]; tm = u[2]; }
      else { xe = hg_reio_Xe(r, zi, u[0] + r.fHe * u[1]); tm = u[3]; }
    }
    Xe[i] = xe; Tm[i] = tm;
  }
  status[ic] = ok ? 0 : 1;
}

// tau', tau (reverse cumulative trapezoid), g  (ionization.jl:107-137); one thread per cosmology for the scan
__global__ void hg_tau_kernel(const HgCosmo* __restrict__ cos, int ncos, HgGrid g, double* __restrict__ Y, const double* __restrict__ XE) {
  const int ic = blockIdx.x * blockDim.x + threadIdx.x;
  if (ic >= ncos) return;
  const HgCosmo& c = cos[ic];
  double* y = Y + (size_t)ic * BOLT_NTABLES * g.n_x;
  const double* calH = y + (size_t)BOLT_T_H * g.n_x;
  const double* Xe = XE + (size_t)ic * 3 * g.n_x;
  double* tau = y + (size_t)BOLT_T_tau * g.n_x; double* gg = y + (size_t)BOLT_T_g * g.n_x;
  auto taup = [&](int i) {
    const double x = g.x0 + g.dx * i, a = exp(x);
    const double n_H = c.Om_b * c.rho_crit / (m_H_nat * a * a * a) * (1.0 - c.Yp);
    return -Xe[i] * n_H * a * sigma_T_nat / calH[i];
  };
  double cum = 0.0, prev = taup(g.n_x - 1);
  tau[g.n_x - 1] = 0.0; gg[g.n_x - 1] = -prev;
  for (int i = g.n_x - 2; i >= 0; i--) {
    const double cur = taup(i);
    const double xr0 = g.x0 + g.dx * (i + 1), xr1 = g.x0 + g.dx * i;
    cum += (xr1 - xr0) * (cur + prev) / 2.0;
    tau[i] = cum; gg[i] = -cur * exp(-cum);
    prev = cur;
  }
}
// csb2 samples (recfast.jl:706-712): csb2_pre * (Tmat - dTmat/3), csb2_pre = C^-2 k_B/m_H (1/mu_T + (1 - Yp) Xe)
// S[cos][3][n_x] = Xe, Tmat, dTmat/dx at the knots
__global__ void hg_csb2_kernel(const HgCosmo* __restrict__ cos, HgGrid g, const double* __restrict__ S, double* __restrict__ Y) {
  const int ic = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.n_x) return;
  const RecConsts& r = cos[ic].r;
  const double* sc = S + (size_t)ic * 3 * g.n_x;
  const double pre = 1.0 / (r.C * r.C) * r.k_B / r.m_H * (1.0 / r.mu_T + (1.0 - r.Yp) * sc[i]);
  Y[((size_t)ic * BOLT_NTABLES + BOLT_T_csb2) * g.n_x + i] = pre * (sc[g.n_x + i] - sc[2 * (size_t)g.n_x + i] / 3.0);
}

void fill_consts(HgCosmo& c) {
  c.H0 = c.h * km_s_Mpc_100;                                                   // background.jl:5
  c.rho_crit = (3.0 / (8.0 * PI)) * c.H0 * c.H0 / G_natural;                   // :6
  const double Tg = pow(15.0 / (PI * PI) * c.rho_crit * c.Om_r, 0.25);
  c.T_nu = pow(c.N_nu / 3.0, 0.25) * pow(4.0 / 11.0, 1.0 / 3.0) * Tg;          // :22
  const double nufac = (90.0 * ZETA3 / (11.0 * PI * PI * PI * PI)) * (c.Om_r * c.h * c.h / Tg) * pow(c.N_nu / 3.0, 0.75);
  const double Om_nu = c.m_nu * nufac / (c.h * c.h);
  c.Om_L = 1.0 - (c.Om_r * (1.0 + (2.0 / 3.0) * (7.0 * c.N_nu / 8.0) * pow(4.0 / 11.0, 4.0 / 3.0)) + c.Om_b + c.Om_c + Om_nu);   // :7-18
  RecConsts& s = c.r;                                                          // recfast.jl:22-121
  s.C = 2.99792458e8; s.k_B = 1.380658e-23; s.h_P = 6.6260755e-34;
  const double m_e = 9.1093897e-31; s.m_H = 1.673575e-27; s.not4 = 3.9715e0;
  const double sigma = 6.6524616e-29, a_rad = 7.565914e-16, G = 6.6742e-11;
  s.Lambda = 8.2245809e0; s.Lambda_He = 51.3e0;
  const double L_H_ion = 1.096787737e7, L_H_alpha = 8.225916453e6, L_He1_ion = 1.98310772e7, L_He2_ion = 4.389088863e7, L_He_2s = 1.66277434e7;
  s.L_He_2p = 1.71134891e7; s.A2P_s = 1.798287e9; s.A2P_t = 177.58e0;
  s.L_He_2Pt = 1.690871466e7; s.L_He_2St = 1.5985597526e7; s.L_He2St_ion = 3.8454693845e6;
  s.sigma_He_2Ps = 1.436289e-22; s.sigma_He_2Pt = 1.484872e-22;
  s.AGauss1 = -0.14e0; s.AGauss2 = 0.079e0; s.zGauss1 = 7.28e0; s.zGauss2 = 6.73e0; s.wGauss1 = 0.18e0; s.wGauss2 = 0.33e0;
  s.a_PPB = 4.309; s.b_PPB = -0.6166; s.c_PPB = 0.6703; s.d_PPB = 0.5300;
  s.a_VF = pow(10.0, -16.744); s.b_VF = 0.711; s.T_0 = pow(10.0, 0.477121); s.T_1 = pow(10.0, 5.114);
  s.a_trip = pow(10.0, -16.306); s.b_trip = 0.761;
  const double Lalpha = 1.0 / L_H_alpha, Lalpha_He = 1.0 / s.L_He_2p;
  s.CDB = s.h_P * s.C * (L_H_ion - L_H_alpha) / s.k_B; s.CDB_He = s.h_P * s.C * (L_He1_ion - L_He_2s) / s.k_B;
  s.CB1 = s.h_P * s.C * L_H_ion / s.k_B; s.CB1_He1 = s.h_P * s.C * L_He1_ion / s.k_B; s.CB1_He2 = s.h_P * s.C * L_He2_ion / s.k_B;
  s.CR = 2.0 * PI * (m_e / s.h_P) * (s.k_B / s.h_P);
  s.CK = Lalpha * Lalpha * Lalpha / (8.0 * PI); s.CK_He = Lalpha_He * Lalpha_He * Lalpha_He / (8.0 * PI);
  s.CL = s.C * s.h_P / (s.k_B * Lalpha); s.CL_He = s.C * s.h_P / (s.k_B / L_He_2s);
  s.CT = (8.0 / 3.0) * (sigma / (m_e * s.C)) * a_rad;
  s.Bfact = s.h_P * s.C * (s.L_He_2p - L_He_2s) / s.k_B;
  s.H_frac = 1e-3; s.Hswitch = 1; s.Heswitch = 6; s.fu = 1.125; s.b_He = 0.86;
  s.Yp = c.Yp;
  s.HO = c.H0 / H0_unit;
  s.Tnow = pow(15.0 / (PI * PI) * c.rho_crit * c.Om_r, 0.25) * Kelvin_unit;
  s.mu_H = 1.0 / (1.0 - c.Yp); s.mu_T = s.not4 / (s.not4 - (s.not4 - 1.0) * c.Yp);
  s.fHe = c.Yp / (s.not4 * (1.0 - c.Yp));
  s.Nnow = 3.0 * s.HO * s.HO * c.Om_b / (8.0 * PI * G * s.mu_H * s.m_H);
}

#define HG_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_hg_err = std::string(#call) + ": " + cudaGetErrorString(e_); rc = BOLT_ERR_CUDA; goto done; } } while (0)

}  // namespace

extern "C" {

const char* bolt_hostgen_last_error(void) { return g_hg_err.c_str(); }

int bolt_hostgen_batch(int device_ordinal, const double* params, int ncos, double x0, double dx, int n_x, const double* quad_pts,
                       const double* quad_wts, int nq, double* tables_out, double* scalars_out, int32_t* status_out) {
  if (!params || ncos < 1 || n_x < 8 || !quad_pts || !quad_wts || nq < 1 || nq > HG_MAXQ || !tables_out || !scalars_out) {
    g_hg_err = "bolt_hostgen_batch: bad arguments"; return BOLT_ERR_ARG;
  }
  int rc = BOLT_OK;
  const int nc = n_x + 2, m = n_x - 2;
  std::vector<HgCosmo> hc(ncos);
  HgGrid g; g.x0 = x0; g.dx = dx; g.n_x = n_x; g.nq = nq;
  for (int i = 0; i < nq; i++) { g.pts[i] = quad_pts[i]; g.wts[i] = quad_wts[i]; }
  for (int i = 0; i < ncos; i++) {
    const double* p = params + (size_t)i * 9;
    HgCosmo& c = hc[i];
    c.h = p[0]; c.Om_r = p[1]; c.Om_b = p[2]; c.Om_c = p[3]; c.A = p[4]; c.ns = p[5]; c.Yp = p[6]; c.N_nu = p[7]; c.m_nu = p[8];
    fill_consts(c);
    c.eta0 = hg_eta(c, g, 0.0);                                                  // background.jl:86
  }
  std::vector<double> cp(m), iden(m);
  { const double a = 1.0 / 6.0, b = 2.0 / 3.0;
    for (int i = 0; i < m; i++) { const double den = (i == 0) ? b : b - a * cp[i - 1]; cp[i] = a / den; iden[i] = 1.0 / den; } }
  HgCosmo* d_cos = nullptr; double *d_cp = nullptr, *d_iden = nullptr, *d_Y = nullptr, *d_C = nullptr, *d_S = nullptr, *d_CT = nullptr; int* d_st = nullptr;
  // Y[cos][12][n_x]: samples of every table series; C[cos][12][n_x+2]: their spline coefficients (the output);
  // S[cos][3][n_x]: Xe, Tmat, dTmat/dx; CT[cos][n_x+2]: coefficients of the Tmat spline
  const size_t nY = (size_t)ncos * BOLT_NTABLES * n_x, nC = (size_t)ncos * BOLT_NTABLES * nc;
  std::vector<int> st(ncos, 0);
  {
    HG_OK(cudaSetDevice(device_ordinal));
    HG_OK(cudaMalloc(&d_cos, ncos * sizeof(HgCosmo))); HG_OK(cudaMemcpy(d_cos, hc.data(), ncos * sizeof(HgCosmo), cudaMemcpyHostToDevice));
    HG_OK(cudaMalloc(&d_cp, m * 8)); HG_OK(cudaMalloc(&d_iden, m * 8));
    HG_OK(cudaMemcpy(d_cp, cp.data(), m * 8, cudaMemcpyHostToDevice)); HG_OK(cudaMemcpy(d_iden, iden.data(), m * 8, cudaMemcpyHostToDevice));
    HG_OK(cudaMalloc(&d_Y, nY * 8)); HG_OK(cudaMalloc(&d_C, nC * 8));
    HG_OK(cudaMalloc(&d_S, (size_t)ncos * 3 * n_x * 8)); HG_OK(cudaMalloc(&d_CT, (size_t)ncos * nc * 8));
    HG_OK(cudaMalloc(&d_st, ncos * sizeof(int)));
    const int ystr = BOLT_NTABLES * n_x, cstr = BOLT_NTABLES * nc;
    auto Ys = [&](int t) { return d_Y + (size_t)t * n_x; };      // series t of cosmology 0 (cosmology stride ystr)
    auto Cs = [&](int t) { return d_C + (size_t)t * nc; };
    const dim3 gp((n_x + 127) / 128, ncos);
    auto prefilter = [&](int t) { hg_prefilter_kernel<<<(ncos + 63) / 64, 64>>>(Ys(t), ncos, n_x, d_cp, d_iden, Cs(t), ystr, cstr); };
    auto derivs = [&](int t, int tg, int th) {      // knot gradient / hessian of series t's spline -> the SAMPLES of series tg / th
      hg_knot_derivs_kernel<<<gp, 128>>>(Cs(t), ncos, n_x, cstr, dx, Ys(tg), Ys(th), ystr);
    };
    // background: calH, eta, rho0M and calH', calH'' (background.jl:95-101)
    hg_background_kernel<<<gp, 128>>>(d_cos, g, d_Y);
    prefilter(BOLT_T_H); prefilter(BOLT_T_eta); prefilter(BOLT_T_rho0M);
    derivs(BOLT_T_H, BOLT_T_Hp, BOLT_T_Hpp);
    prefilter(BOLT_T_Hp); prefilter(BOLT_T_Hpp);
    // recombination + reionization: Xe, Tmat on the grid (one thread per cosmology)
    hg_recfast_kernel<<<(ncos + 31) / 32, 32>>>(d_cos, ncos, g, Cs(BOLT_T_H), Cs(BOLT_T_Hp), cstr, d_S, d_st);
    // tau, g and their derivative tables (ionization.jl:107-137, recfast.jl:696-726)
    hg_tau_kernel<<<(ncos + 31) / 32, 32>>>(d_cos, ncos, g, d_Y, d_S);
    prefilter(BOLT_T_tau); prefilter(BOLT_T_g);
    derivs(BOLT_T_tau, BOLT_T_taup, BOLT_T_taupp); derivs(BOLT_T_g, BOLT_T_gp, BOLT_T_gpp);
    prefilter(BOLT_T_taup); prefilter(BOLT_T_taupp); prefilter(BOLT_T_gp); prefilter(BOLT_T_gpp);
    // baryon sound speed from the Tmat spline and its knot gradient (recfast.jl:700-712)
    hg_prefilter_kernel<<<(ncos + 63) / 64, 64>>>(d_S + n_x, ncos, n_x, d_cp, d_iden, d_CT, 3 * n_x, nc);
    hg_knot_derivs_kernel<<<gp, 128>>>(d_CT, ncos, n_x, nc, dx, d_S + 2 * (size_t)n_x, nullptr, 3 * n_x);
    hg_csb2_kernel<<<gp, 128>>>(d_cos, g, d_S, d_Y);
    prefilter(BOLT_T_csb2);
    HG_OK(cudaGetLastError());
    HG_OK(cudaDeviceSynchronize());
    HG_OK(cudaMemcpy(tables_out, d_C, nC * 8, cudaMemcpyDeviceToHost));
    HG_OK(cudaMemcpy(st.data(), d_st, ncos * sizeof(int), cudaMemcpyDeviceToHost));
  }
  for (int i = 0; i < ncos; i++) {
    const HgCosmo& c = hc[i];
    double* s = scalars_out + (size_t)i * BOLT_NSCALARS;
    s[BOLT_S_h] = c.h; s[BOLT_S_Omega_r] = c.Om_r; s[BOLT_S_Omega_b] = c.Om_b; s[BOLT_S_Omega_c] = c.Om_c; s[BOLT_S_A] = c.A; s[BOLT_S_n] = c.ns;
    s[BOLT_S_Y_p] = c.Yp; s[BOLT_S_N_nu] = c.N_nu; s[BOLT_S_Sum_m_nu] = c.m_nu;
    s[BOLT_S_H0] = c.H0; s[BOLT_S_eta0] = c.eta0; s[BOLT_S_rho_crit] = c.rho_crit; s[BOLT_S_Omega_L] = c.Om_L;
    if (status_out) status_out[i] = st[i];
  }
done:
  cudaFree(d_cos); cudaFree(d_cp); cudaFree(d_iden); cudaFree(d_Y); cudaFree(d_C); cudaFree(d_S); cudaFree(d_CT); cudaFree(d_st);
  return rc;
}

}  // extern "C"
