// Shared device-side definitions of libbolt_cuda (sm_100a).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/bolt_cuda.h"
#include "dual.cuh"

namespace bolt {

constexpr int MAX_NQ = 29;        // chains = nq + 3 must fit in one warp
constexpr int MAX_L = 127;        // ratio table size
constexpr unsigned FULL = 0xffffffffu;

// One cosmology, resident in HBM (uploaded once; 12 x (n_x+2) spline coefficients + momentum grid).
struct DevCosmo {
  int n_x, nq, nd;
  double x0, dx, inv_dx;
  double s[BOLT_NSCALARS];
  const double* tab[BOLT_NTABLES];   // value tables, each n_x+2 doubles
  double Omega_nu;                   // 7(2/3)N_ν/8 (4/11)^{4/3} Ω_r      perturbations.jl:171
  double q[MAX_NQ];                  // q_i                               perturbations.jl:166
  double wq[MAX_NQ];                 // 4π q_i² f0(q_i)/dxdq(q_i) w_i     perturbations.jl:138-142
  double df0[MAX_NQ];                // dlnf0dlnq(q_i)                    background.jl:27-30
  double eta_end;                    // η(x_grid[end])                    perturbations.jl:401
  // forward-mode partials (nd = 1 + np > 1 only): same quantities, d/dp_j
  int np;
  const double* dtab[BOLT_NTABLES];  // partial tables, component-major [np][n_x+2]
  double ds[BOLT_NSCALARS][MAX_NP];
  double dq[MAX_NQ][MAX_NP], dwq[MAX_NQ][MAX_NP];
  double dOmega_nu[MAX_NP], deta_end[MAX_NP];
};

// Interpolations.jl BSpline(Cubic(Line(OnGrid()))) on the uniform x grid (src/util.jl:11).
__device__ __forceinline__ double spline_eval(const double* __restrict__ c, int n_x, double x0, double dx, double x) {
  double t = (x - x0) / dx;
  int i = (int)floor(t);
  i = max(0, min(i, n_x - 2));
  double d = t - (double)i, e = 1.0 - d;
  double w0 = e * e * e * (1.0 / 6.0);
  double w1 = 2.0 / 3.0 - d * d + d * d * d * 0.5;
  double w2 = 2.0 / 3.0 - e * e + e * e * e * 0.5;
  double w3 = d * d * d * (1.0 / 6.0);
  return c[i] * w0 + c[i + 1] * w1 + c[i + 2] * w2 + c[i + 3] * w3;
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

}  // namespace bolt
