// Bessel-moment tables and the Filon line-of-sight rule on the device.
//
// What the reference has (src/bessel/): moments  I_m(x) = int_0^x t^m j_nu(t) dt  (moments.jl:56-112), a 2,000,000-node cubic
// B-spline table of (I_0 .. I_{order-1}) with a Maclaurin branch below and a Lommel-asymptotic branch above the table
// (interpolator.jl:26-110), and the Filon rule that integrates a quadratic source piece against j_nu(kx) with differences of
// the moments (integrator.jl:7-38).  The table is built on the host, one node at a time, summing a 1F2 with Weniger's sequence
// transformation in Double64 for x < 50 (weniger.jl:50-235).
//
// Device design (NOT a translation of that):
//  * small arguments (x < weniger_cut): the moments are INTEGRATED, not summed.  A fine prefix table P[i] = I_m(X0 + i/64) is
//    built once per table by 8-point Gauss-Legendre on every 1/64-wide interval (the integrand t^m j_nu(t) is entire; the
//    quadrature error is ~1e-25) and a double-double prefix sum; I_m(x) = P[floor] + one more Gauss-Legendre piece.  Below
//    X0 = 4 the Maclaurin series of the 1F2 is used directly (all terms < 1.2 there: no cancellation).  Every node is
//    independent -> one thread per node; no extended-precision recurrences.
//  * large arguments: the Lommel asymptotic form (moments.jl:61-70) with j_nu, j_{nu-1} by upward recurrence from sincos.
//  * B-spline prefilter of the 2e6 x order node values: the inverse of the (1,4,1)/6 operator decays like 0.268^d, so interior
//    coefficients are an 81-tap convolution (one thread per node, 1e-23 truncation) and the two ends are closed exactly
//    (Interpolations.jl's Line(OnGrid()) condition gives c_0 = y_0) by a 40-unknown Thomas solve each.
//  * the table (64 MB) stays resident in the 126 MB L2 while Filon kernels gather from it; coefficients are stored as 32-byte
//    node records so one evaluation is four 256-bit loads (LDG.E.256) from 128 consecutive bytes.
// Standalone translation unit: no context needed (like hostgen_batch.cu).
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/bolt_cuda.h"

struct bolt_moment_table {
  int device, nu, order, N;
  double xmin, xmax, cut;
  double pref[4], c2;
  double* d_coef;          // [N+2][4] cubic B-spline coefficients (one padding node each side; 32-byte node records)
  cudaStream_t stream;
};

namespace {

thread_local std::string g_bm_err;
constexpr int MAXO = 4;
constexpr double X0_SMALL = 4.0;       // Maclaurin below, prefix + quadrature above
constexpr double H_FINE = 1.0 / 64.0;
constexpr int CSTRIDE = 4;             // doubles per node record of the coefficient table
constexpr int PF_W = 40;               // prefilter half window: 0.268^40 = 1.3e-23
constexpr double PI = 3.14159265358979323846;

#define BM_CHECK(call)                                                                                             \
  do {                                                                                                             \
    cudaError_t e_ = (call);                                                                                       \
    if (e_ != cudaSuccess) {                                                                                       \
      g_bm_err = std::string(#call) + ": " + cudaGetErrorString(e_);                                               \
      return BOLT_ERR_CUDA;                                                                                        \
    }                                                                                                              \
  } while (0)

struct MomentSpec {        // what every evaluator needs
  int nu, order;
  double power[MAXO];      // m_0 .. m_{order-1} (0,1,2,.. for tables; real-valued for the J_nu building blocks)
  double pref[MAXO];       // asymptotic constants  int_0^inf t^m j_nu(t) dt
  double c2;               // (nu + 1/2) log 2 + lgamma(nu + 3/2)
};

struct Prefix {            // fine prefix table of the small-argument evaluator
  const double* hi;        // [n+1][order]
  const double* lo;
  int n;                   // intervals
};

struct TableView {
  const double* coef;
  int N;
  double xmin, xmax, inv_h;
};

// ---------------------------------------------------------------------------------------------------------------------
// pointwise pieces
// ---------------------------------------------------------------------------------------------------------------------
__device__ inline void sph_jn_pair(double t, int nu, double& jn, double& jnm1) {      // t >~ nu: upward recurrence is stable
  double s, c;
  sincos(t, &s, &c);
  const double ti = 1.0 / t;
  double a = s * ti, b = (s * ti - c) * ti;
  for (int n = 1; n < nu; n++) {
    const double nx = (2 * n + 1) * ti * b - a;
    a = b;
    b = nx;
  }
  jn = b;
  jnm1 = a;
}

__device__ inline double pow_m(double x, double p) {      // x^p, exact products for the small integers the tables use
  if (p == 0.0) return 1.0;
  if (p == 1.0) return x;
  if (p == 2.0) return x * x;
  if (p == 3.0) return x * x * x;
  return pow(x, p);
}

// Lommel series S(t; a, mu) = 1 + sum_j prod_{i<=j} (mu^2 - (a - 1/2 - 2i)^2) / t^2   (moments.jl:5-21 without the t^a t^-1/2)
__device__ inline double lommel_series(double t, double a, double mu) {
  double s = 1.0, sk = 1.0;
  const double ti2 = 1.0 / (t * t), mu2 = mu * mu, am1 = a - 0.5;
  for (int j = 0; j <= 20; j++) {
    const double d = am1 - 2.0 * j;
    sk *= (mu2 - d * d) * ti2;
    s += sk;
    if (fabs(sk) < 1e-16 * fabs(s)) break;
  }
  return s;
}

// I_m(x) for large x (moments.jl:61-70 with J_{nu +- 1/2} = sqrt(2x/pi) j): pref + (m+nu-1) j_nu x^{m-1} S1 - j_{nu-1} x^m S2
__device__ inline void moments_asymp(const MomentSpec& sp, double x, double* out) {
  double jn, jm;
  sph_jn_pair(x, sp.nu, jn, jm);
  for (int o = 0; o < sp.order; o++) {
    const double m = sp.power[o];
    const double xm = pow_m(x, m);
    const double S1 = lommel_series(x, m - 2.0, sp.nu - 0.5), S2 = lommel_series(x, m - 1.0, sp.nu + 0.5);
    out[o] = sp.pref[o] + ((m + sp.nu - 1.0) * jn * (xm / x) * S1 - jm * xm * S2);
  }
}

// I_m(x) from the Maclaurin series of 1F2((1+m+nu)/2; (3+m+nu)/2, nu+3/2; -x^2/4)   (moments.jl:79-83, weniger.jl:238-249)
__device__ inline void moments_maclaurin(const MomentSpec& sp, double x, double* out) {
  if (!(x > 0.0)) {
    for (int o = 0; o < sp.order; o++) out[o] = 0.0;
    return;
  }
  const double z = -0.25 * x * x, lx = log(x), b2 = sp.nu + 1.5;
  for (int o = 0; o < sp.order; o++) {
    const double m = sp.power[o], a = 0.5 * (1.0 + m + sp.nu), b1 = a + 1.0;
    double S0 = 1.0, S1 = 1.0 + a * z / (b1 * b2);
    for (int j = 1; j < 200; j++) {
      if (j > 1 && !(fabs(S0 - S1) > 10.0 * 2.220446049250313e-16 * fmax(fabs(S0), fabs(S1)))) break;
      const double r = (a + j) / ((j + 1.0) * (b1 + j) * (b2 + j));
      const double nx = S1 + (S1 - S0) * r * z;
      S0 = S1;
      S1 = nx;
    }
    out[o] = (1.2533141373155003 / (m + sp.nu + 1.0)) * exp((m + sp.nu + 1.0) * lx - sp.c2) * S1;      // sqrt(pi/2)
  }
}

// int_a^b t^m j_nu(t) dt, 8-point Gauss-Legendre, b - a <= 1/64 (a >= 4)
__device__ inline void gl8_piece(const MomentSpec& sp, double a, double b, double* inc) {
  const double GX[4] = {0.1834346424956498, 0.5255324099163290, 0.7966664774136267, 0.9602898564975363};
  const double GW[4] = {0.3626837833783620, 0.3137066458778873, 0.2223810344533745, 0.1012285362903763};
  const double c = 0.5 * (a + b), hw = 0.5 * (b - a);
  for (int o = 0; o < sp.order; o++) inc[o] = 0.0;
#pragma unroll
  for (int g = 0; g < 8; g++) {
    const double t = (g & 1) ? c + hw * GX[g >> 1] : c - hw * GX[g >> 1];
    double jn, jm;
    sph_jn_pair(t, sp.nu, jn, jm);
    const double w = GW[g >> 1] * jn;
    for (int o = 0; o < sp.order; o++) inc[o] = fma(w, pow_m(t, sp.power[o]), inc[o]);
  }
  for (int o = 0; o < sp.order; o++) inc[o] *= hw;
}

__device__ inline void moments_small(const MomentSpec& sp, const Prefix& P, double x, double* out) {
  if (x < X0_SMALL) {
    moments_maclaurin(sp, x, out);
    return;
  }
  int i = (int)((x - X0_SMALL) * (1.0 / H_FINE));
  if (i > P.n) i = P.n;
  double inc[MAXO];
  gl8_piece(sp, X0_SMALL + i * H_FINE, x, inc);
  for (int o = 0; o < sp.order; o++) out[o] = P.hi[(size_t)i * sp.order + o] + (P.lo[(size_t)i * sp.order + o] + inc[o]);
}

__device__ inline void ld256(const double* p, double* r) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r[0]), "=d"(r[1]), "=d"(r[2]), "=d"(r[3]) : "l"(p));
}

__device__ inline void table_interp(const TableView& T, int order, double x, double* out) {
  const double t = (x - T.xmin) * T.inv_h;
  int i = (int)floor(t);
  i = max(0, min(i, T.N - 2));
  const double u = t - i, u2 = u * u, u3 = u2 * u, v = 1.0 - u;
  const double w0 = v * v * v * (1.0 / 6.0), w1 = (3.0 * u3 - 6.0 * u2 + 4.0) * (1.0 / 6.0);
  const double w2 = (-3.0 * u3 + 3.0 * u2 + 3.0 * u + 1.0) * (1.0 / 6.0), w3 = u3 * (1.0 / 6.0);
  // node records are 4 doubles (32 B, aligned) whatever the order: one 256-bit load per node.  A divergent gather costs one L1
  // wavefront per lane per load instruction, so 4 wide loads instead of 4 x order narrow ones is what this kernel is bound by.
  const double* c = T.coef + (size_t)i * CSTRIDE;
  double r0[4], r1[4], r2[4], r3[4];
  ld256(c, r0);
  ld256(c + CSTRIDE, r1);
  ld256(c + 2 * CSTRIDE, r2);
  ld256(c + 3 * CSTRIDE, r3);
#pragma unroll
  for (int o = 0; o < MAXO; o++)
    if (o < order) out[o] = w0 * r0[o] + w1 * r1[o] + w2 * r2[o] + w3 * r3[o];
}

// the three-branch call of the reference's MomentTable (interpolator.jl:26-34)
__device__ inline void table_call(const MomentSpec& sp, const TableView& T, double x, double* out) {
  if (x >= T.xmin && x <= T.xmax) table_interp(T, sp.order, x, out);
  else if (x > T.xmax) moments_asymp(sp, x, out);
  else moments_maclaurin(sp, x, out);
}

// ---------------------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------------------
__device__ inline void dd_add(double& hi, double& lo, double x) {      // (hi, lo) += x, error-free two-sum
  const double s = hi + x, bb = s - hi, e = (hi - (s - bb)) + (x - bb);
  hi = s;
  lo += e;
}

// one block: thread t integrates a contiguous run of fine intervals, then the run totals are chained (double-double throughout)
__global__ void prefix_kernel(MomentSpec sp, int n, double* __restrict__ hi, double* __restrict__ lo) {
  __shared__ double tot_hi[256][MAXO], tot_lo[256][MAXO];
  const int t = threadIdx.x, per = (n + 255) / 256, i0 = min(t * per, n), i1 = min(i0 + per, n);
  double ah[MAXO] = {0, 0, 0, 0}, al[MAXO] = {0, 0, 0, 0};
  for (int i = i0; i < i1; i++) {
    double inc[MAXO];
    gl8_piece(sp, X0_SMALL + i * H_FINE, X0_SMALL + (i + 1) * H_FINE, inc);
    for (int o = 0; o < sp.order; o++) {
      dd_add(ah[o], al[o], inc[o]);
      hi[(size_t)(i + 1) * sp.order + o] = ah[o];      // local prefix, offset added below
      lo[(size_t)(i + 1) * sp.order + o] = al[o];
    }
  }
  for (int o = 0; o < MAXO; o++) { tot_hi[t][o] = ah[o]; tot_lo[t][o] = al[o]; }
  __syncthreads();
  if (t == 0) {
    double base[MAXO];
    moments_maclaurin(sp, X0_SMALL, base);
    double bh[MAXO], bl[MAXO];
    for (int o = 0; o < sp.order; o++) {
      bh[o] = base[o];
      bl[o] = 0.0;
      hi[o] = bh[o];
      lo[o] = 0.0;
    }
    for (int q = 0; q < 256; q++)
      for (int o = 0; o < sp.order; o++) {
        const double th = tot_hi[q][o], tl = tot_lo[q][o];
        tot_hi[q][o] = bh[o];      // exclusive offset of run q
        tot_lo[q][o] = bl[o];
        dd_add(bh[o], bl[o], th);
        bl[o] += tl;
      }
  }
  __syncthreads();
  for (int i = i0; i < i1; i++)
    for (int o = 0; o < sp.order; o++) {
      const size_t id = (size_t)(i + 1) * sp.order + o;
      double h = tot_hi[t][o], l = tot_lo[t][o];
      dd_add(h, l, hi[id]);
      l += lo[id];
      const double s = h + l;      // renormalise
      hi[id] = s;
      lo[id] = l - (s - h);
    }
}

enum { METHOD_SMALL = 0, METHOD_ASYMP = 1, METHOD_MACLAURIN = 2 };

__global__ void direct_kernel(MomentSpec sp, Prefix P, int method, const double* __restrict__ x, int n, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r[MAXO];
  if (method == METHOD_SMALL) moments_small(sp, P, x[i], r);
  else if (method == METHOD_ASYMP) moments_asymp(sp, x[i], r);
  else moments_maclaurin(sp, x[i], r);
  for (int o = 0; o < sp.order; o++) out[(size_t)i * sp.order + o] = r[o];
}

// node values of the table: small-argument evaluator below the cut, asymptotic form above (interpolator.jl:96-108)
__global__ void fill_kernel(MomentSpec sp, Prefix P, double xmin, double xmax, int N, double cut, double* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double f = (double)i / (double)(N - 1), x = (1.0 - f) * xmin + f * xmax;
  double r[MAXO];
  if (x < cut) moments_small(sp, P, x, r);
  else moments_asymp(sp, x, r);
  for (int o = 0; o < sp.order; o++) y[(size_t)i * sp.order + o] = r[o];
}

// interior B-spline coefficients: c_i = sqrt(3) sum_d z^|d| y_{i+d}, z = sqrt(3) - 2, for PF_W <= i <= N-1-PF_W
__global__ void prefilter_interior_kernel(const double* __restrict__ y, int N, int order, double* __restrict__ coef) {
  extern __shared__ double tile[];      // [blockDim + 2 PF_W][order]
  const int base = PF_W + blockIdx.x * blockDim.x;
  const int span = blockDim.x + 2 * PF_W;
  for (int j = threadIdx.x; j < span * order; j += blockDim.x) {
    const long g = (long)(base - PF_W) * order + j;
    tile[j] = (g < (long)N * order) ? y[g] : 0.0;
  }
  __syncthreads();
  const int i = base + threadIdx.x;
  if (i > N - 1 - PF_W) return;
  const double z = -0.2679491924311227, r3 = 1.7320508075688772;
  for (int o = 0; o < order; o++) {
    // Horner from the outside in: acc = y_{-W} + y_{+W}; acc = acc z + (y_{-d} + y_{+d}) ... ; c = sqrt3 (acc z + y_0)
    const double* tc = tile + (size_t)(threadIdx.x + PF_W) * order + o;
    double acc = 0.0;
    for (int d = PF_W; d >= 1; d--) acc = (acc + (tc[-d * order] + tc[d * order])) * z;
    coef[(size_t)(i + 1) * CSTRIDE + o] = r3 * (acc + tc[0]);
  }
}

// closes one end (or a whole short table): unknown nodes lo+1 .. hi-1 with c_lo, c_hi already known; then the padding nodes
__global__ void prefilter_ends_kernel(const double* __restrict__ y, int N, int order, double* __restrict__ coef, int whole) {
  const int o = threadIdx.x % order, end = threadIdx.x / order;      // end 0: left, 1: right
  if (end > 1 || (whole && end == 1)) return;
  auto C = [&](int i) -> double& { return coef[(size_t)(i + 1) * CSTRIDE + o]; };
  auto Y = [&](int i) { return y[(size_t)i * order + o]; };
  int lo, hi;
  if (whole) { lo = 0; hi = N - 1; C(0) = Y(0); C(N - 1) = Y(N - 1); }
  else if (end == 0) { lo = 0; hi = PF_W; C(0) = Y(0); }
  else { lo = N - 1 - PF_W; hi = N - 1; C(N - 1) = Y(N - 1); }
  // Thomas on (1,4,1)/6 with Dirichlet ends; the modified super-diagonal cp_j converges to 2 - sqrt(3) within 1e-23 by j = 40
  double cp[48];
  cp[0] = 0.25;
  for (int j = 1; j < 48; j++) cp[j] = 1.0 / (4.0 - cp[j - 1]);
  const int n = hi - lo - 1;
  double prev = 0.0;
  for (int j = 0; j < n; j++) {      // forward sweep, d'_j kept in the coefficient array
    const int i = lo + 1 + j;
    double rhs = 6.0 * Y(i);
    if (j == 0) rhs -= C(lo);
    if (j == n - 1) rhs -= C(hi);
    const double den = (j == 0) ? 4.0 : 4.0 - cp[min(j - 1, 47)];
    prev = (rhs - (j == 0 ? 0.0 : prev)) / den;
    C(i) = prev;
  }
  for (int j = n - 2; j >= 0; j--) {
    const int i = lo + 1 + j;
    C(i) = C(i) - cp[min(j, 47)] * C(i + 1);
  }
  if (whole || end == 0) C(-1) = 2.0 * C(0) - C(1);
  if (whole || end == 1) C(N) = 2.0 * C(N - 1) - C(N - 2);
}

__global__ void table_eval_kernel(MomentSpec sp, TableView T, const double* __restrict__ x, int n, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r[MAXO];
  table_call(sp, T, x[i], r);
  for (int o = 0; o < sp.order; o++) out[(size_t)i * sp.order + o] = r[o];
}

__device__ inline double filon_piece(double f, double f1, double f2, double k, double a, const double* Ia, const double* Ib) {
  const double c2 = 0.5 * f2, af2 = a * f2, c1 = f1 - af2, c0 = f - a * (f1 - 0.5 * af2);
  const double ki = 1.0 / k, ki2 = ki * ki, ki3 = ki2 * ki;
  return (c0 * ki) * (Ib[0] - Ia[0]) + (c1 * ki2) * (Ib[1] - Ia[1]) + (c2 * ki3) * (Ib[2] - Ia[2]);
}

// independent pieces (integrator.jl:7-20): one thread per piece
__global__ void filon_pieces_kernel(MomentSpec sp, TableView T, int n, const double* __restrict__ f, const double* __restrict__ f1,
                                    const double* __restrict__ f2, const double* __restrict__ k, const double* __restrict__ a,
                                    const double* __restrict__ b, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double Ia[MAXO], Ib[MAXO];
  table_call(sp, T, k[i] * a[i], Ia);
  table_call(sp, T, k[i] * b[i], Ib);
  out[i] = filon_piece(f[i], f1[i], f2[i], k[i], a[i], Ia, Ib);
}

// a chain of pieces over consecutive nodes for every k (the loop form, integrator.jl:25-38: I(k a_{i+1}) is re-used as the
// next piece's I(k a_i)): one block per k, the node moments staged in shared memory, a fixed-order block reduction
template <int BT, int NPT>
__global__ void __launch_bounds__(BT) filon_chain_kernel(MomentSpec sp, TableView T, int n_nodes, const double* __restrict__ nodes,
                                   const double* __restrict__ f, const double* __restrict__ f1, const double* __restrict__ f2,
                                   const double* __restrict__ k, double* __restrict__ out) {
  constexpr int TILE = BT * NPT;      // nodes per tile -> TILE-1 pieces; NPT independent gathers in flight per thread
  __shared__ double I[TILE][3];
  __shared__ double red[BT];
  const int ik = blockIdx.x, t = threadIdx.x;
  const double kk = k[ik];
  const double *F = f + (size_t)ik * n_nodes, *F1 = f1 + (size_t)ik * n_nodes, *F2 = f2 + (size_t)ik * n_nodes;
  double acc = 0.0;
  for (int base = 0; base < n_nodes - 1; base += TILE - 1) {
    double xa[NPT], fa[NPT], f1a[NPT], f2a[NPT];
#pragma unroll
    for (int j = 0; j < NPT; j++) {
      const int i = base + t + j * BT;
      const bool live = i < n_nodes;
      xa[j] = live ? nodes[i] : 0.0;
      fa[j] = live ? F[i] : 0.0;
      f1a[j] = live ? F1[i] : 0.0;
      f2a[j] = live ? F2[i] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < NPT; j++) {
      if (base + t + j * BT < n_nodes) {
        double r[MAXO];
        table_call(sp, T, kk * xa[j], r);
        I[t + j * BT][0] = r[0]; I[t + j * BT][1] = r[1]; I[t + j * BT][2] = r[2];
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NPT; j++) {
      const int l = t + j * BT;
      if (l < TILE - 1 && base + l + 1 < n_nodes) acc += filon_piece(fa[j], f1a[j], f2a[j], kk, xa[j], I[l], I[l + 1]);
    }
    __syncthreads();
  }
  red[t] = acc;
  __syncthreads();
  for (int s = BT / 2; s > 0; s >>= 1) {
    if (t < s) red[t] += red[t + s];
    __syncthreads();
  }
  if (t == 0) out[ik] = red[0];
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
int make_spec(int nu, int order, const double* powers, MomentSpec& sp) {
  if (nu < 1 || nu > 3) { g_bm_err = "nu must be 1, 2 or 3 (the reference tabulates nu = 2, 3: interpolator.jl:13-16)"; return BOLT_ERR_ARG; }
  if (order < 1 || order > MAXO) { g_bm_err = "order must be 1..4"; return BOLT_ERR_ARG; }
  sp.nu = nu;
  sp.order = order;
  sp.c2 = (nu + 0.5) * std::log(2.0) + std::lgamma(nu + 1.5);
  for (int o = 0; o < MAXO; o++) { sp.power[o] = 0.0; sp.pref[o] = 0.0; }
  for (int o = 0; o < order; o++) {
    const double m = powers ? powers[o] : (double)o;
    if (!(m > -(double)nu - 1.0)) { g_bm_err = "power must exceed -(nu+1) for the moment to exist"; return BOLT_ERR_ARG; }
    sp.power[o] = m;
    // int_0^inf t^m j_nu = sqrt(pi) 2^{m-1} Gamma((nu+m+1)/2) / Gamma((nu-m+2)/2)   (moments.jl:37-38,57-58); 0 at a pole of the denominator
    const double den = std::tgamma(0.5 * (nu - m + 2.0));
    sp.pref[o] = std::isfinite(den) ? std::sqrt(PI) * std::exp2(m - 1.0) * std::tgamma(0.5 * (nu + m + 1.0)) / den : 0.0;
  }
  return BOLT_OK;
}

struct DevBuf {
  double* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
};

// builds the fine prefix table covering [X0_SMALL, xtop]
int build_prefix(const MomentSpec& sp, double xtop, DevBuf& hi, DevBuf& lo, Prefix& P, cudaStream_t st) {
  int n = 0;
  if (xtop > X0_SMALL) n = (int)std::ceil((xtop - X0_SMALL) / H_FINE) + 1;
  if (n > (1 << 24)) { g_bm_err = "small-argument range too long (weniger_cut > 2.6e5)"; return BOLT_ERR_ARG; }
  BM_CHECK(cudaMalloc(&hi.p, sizeof(double) * (size_t)(n + 1) * sp.order));
  BM_CHECK(cudaMalloc(&lo.p, sizeof(double) * (size_t)(n + 1) * sp.order));
  prefix_kernel<<<1, 256, 0, st>>>(sp, n, hi.p, lo.p);
  BM_CHECK(cudaGetLastError());
  P.hi = hi.p;
  P.lo = lo.p;
  P.n = n;
  return BOLT_OK;
}

MomentSpec spec_of(const bolt_moment_table* t) {
  MomentSpec sp;
  sp.nu = t->nu;
  sp.order = t->order;
  sp.c2 = t->c2;
  for (int o = 0; o < MAXO; o++) { sp.power[o] = (double)o; sp.pref[o] = t->pref[o]; }
  return sp;
}

TableView view_of(const bolt_moment_table* t) {
  TableView T;
  T.coef = t->d_coef;
  T.N = t->N;
  T.xmin = t->xmin;
  T.xmax = t->xmax;
  T.inv_h = (double)(t->N - 1) / (t->xmax - t->xmin);
  return T;
}

}  // namespace

extern "C" {

const char* bolt_moments_last_error(void) { return g_bm_err.c_str(); }

int bolt_sph_j_moments(int device_ordinal, int nu, int n_powers, const double* powers, int method, const double* x, int n, double* out) {
  if (n < 0 || method < 0 || method > 2) { g_bm_err = "bad argument"; return BOLT_ERR_ARG; }
  MomentSpec sp;
  int rc = make_spec(nu, n_powers, powers, sp);
  if (rc != BOLT_OK) return rc;
  if (n == 0) return BOLT_OK;
  if (!x || !out) { g_bm_err = "null buffer"; return BOLT_ERR_ARG; }
  BM_CHECK(cudaSetDevice(device_ordinal));
  double xtop = 0.0;
  for (int i = 0; i < n; i++) {
    if (!(x[i] >= 0.0) || !std::isfinite(x[i])) { g_bm_err = "x must be finite and >= 0"; return BOLT_ERR_ARG; }
    xtop = std::fmax(xtop, x[i]);
  }
  if (method == METHOD_ASYMP) for (int i = 0; i < n; i++) if (!(x[i] > 0.0)) { g_bm_err = "the asymptotic form needs x > 0"; return BOLT_ERR_ARG; }
  DevBuf hi, lo, dx, dout;
  Prefix P{nullptr, nullptr, 0};
  if (method == METHOD_SMALL) { rc = build_prefix(sp, xtop, hi, lo, P, 0); if (rc != BOLT_OK) return rc; }
  BM_CHECK(cudaMalloc(&dx.p, sizeof(double) * n));
  BM_CHECK(cudaMalloc(&dout.p, sizeof(double) * (size_t)n * sp.order));
  BM_CHECK(cudaMemcpy(dx.p, x, sizeof(double) * n, cudaMemcpyHostToDevice));
  direct_kernel<<<(n + 127) / 128, 128>>>(sp, P, method, dx.p, n, dout.p);
  BM_CHECK(cudaGetLastError());
  BM_CHECK(cudaMemcpy(out, dout.p, sizeof(double) * (size_t)n * sp.order, cudaMemcpyDeviceToHost));
  return BOLT_OK;
}

int bolt_moment_table_create(int device_ordinal, int nu, int order, double keta_min, double keta_max, int N, double weniger_cut,
                             bolt_moment_table** out) {
  if (!out) { g_bm_err = "null output"; return BOLT_ERR_ARG; }
  *out = nullptr;
  if (!(keta_min >= 0.0) || !(keta_max > keta_min) || N < 4 || !(weniger_cut >= 0.0)) {
    g_bm_err = "need 0 <= keta_min < keta_max, N >= 4, weniger_cut >= 0";
    return BOLT_ERR_ARG;
  }
  if (nu != 2 && nu != 3) { g_bm_err = "moment tables exist for nu = 2, 3 (interpolator.jl:13-16)"; return BOLT_ERR_ARG; }
  MomentSpec sp;
  int rc = make_spec(nu, order, nullptr, sp);
  if (rc != BOLT_OK) return rc;
  BM_CHECK(cudaSetDevice(device_ordinal));
  bolt_moment_table* t = new bolt_moment_table();
  t->device = device_ordinal; t->nu = nu; t->order = order; t->N = N;
  t->xmin = keta_min; t->xmax = keta_max; t->cut = weniger_cut; t->c2 = sp.c2;
  for (int o = 0; o < MAXO; o++) t->pref[o] = sp.pref[o];
  t->d_coef = nullptr;
  t->stream = nullptr;
  auto fail = [&](int code) { if (t->d_coef) cudaFree(t->d_coef); if (t->stream) cudaStreamDestroy(t->stream); delete t; return code; };
  if (cudaStreamCreate(&t->stream) != cudaSuccess) { g_bm_err = "cudaStreamCreate failed"; return fail(BOLT_ERR_CUDA); }
  if (cudaMalloc(&t->d_coef, sizeof(double) * (size_t)(N + 2) * CSTRIDE) != cudaSuccess) { g_bm_err = "cudaMalloc(table) failed"; return fail(BOLT_ERR_CUDA); }
  cudaMemsetAsync(t->d_coef, 0, sizeof(double) * (size_t)(N + 2) * CSTRIDE, t->stream);
  DevBuf hi, lo, y;
  Prefix P{nullptr, nullptr, 0};
  rc = build_prefix(sp, std::fmin(weniger_cut, keta_max), hi, lo, P, t->stream);
  if (rc != BOLT_OK) return fail(rc);
  if (cudaMalloc(&y.p, sizeof(double) * (size_t)N * order) != cudaSuccess) { g_bm_err = "cudaMalloc(nodes) failed"; return fail(BOLT_ERR_CUDA); }
  fill_kernel<<<(N + 127) / 128, 128, 0, t->stream>>>(sp, P, keta_min, keta_max, N, weniger_cut, y.p);
  const int whole = (N <= 2 * PF_W + 2);
  if (!whole) {
    const int bt = 256, n_int = N - 2 * PF_W;
    prefilter_interior_kernel<<<(n_int + bt - 1) / bt, bt, sizeof(double) * (bt + 2 * PF_W) * order, t->stream>>>(y.p, N, order, t->d_coef);
  }
  prefilter_ends_kernel<<<1, 2 * order, 0, t->stream>>>(y.p, N, order, t->d_coef, whole);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(t->stream);
  if (e != cudaSuccess) { g_bm_err = std::string("table build: ") + cudaGetErrorString(e); return fail(BOLT_ERR_CUDA); }
  *out = t;
  return BOLT_OK;
}

void bolt_moment_table_free(bolt_moment_table* t) {
  if (!t) return;
  cudaSetDevice(t->device);
  cudaFree(t->d_coef);
  cudaStreamDestroy(t->stream);
  delete t;
}

int bolt_moment_table_eval(bolt_moment_table* t, const double* x, int n, double* out) {
  if (!t || n < 0) { g_bm_err = "bad argument"; return BOLT_ERR_ARG; }
  if (n == 0) return BOLT_OK;
  if (!x || !out) { g_bm_err = "null buffer"; return BOLT_ERR_ARG; }
  for (int i = 0; i < n; i++) if (!(x[i] >= 0.0) || !std::isfinite(x[i])) { g_bm_err = "x must be finite and >= 0"; return BOLT_ERR_ARG; }
  BM_CHECK(cudaSetDevice(t->device));
  DevBuf dx, dout;
  BM_CHECK(cudaMalloc(&dx.p, sizeof(double) * n));
  BM_CHECK(cudaMalloc(&dout.p, sizeof(double) * (size_t)n * t->order));
  BM_CHECK(cudaMemcpyAsync(dx.p, x, sizeof(double) * n, cudaMemcpyHostToDevice, t->stream));
  table_eval_kernel<<<(n + 127) / 128, 128, 0, t->stream>>>(spec_of(t), view_of(t), dx.p, n, dout.p);
  BM_CHECK(cudaGetLastError());
  BM_CHECK(cudaMemcpyAsync(out, dout.p, sizeof(double) * (size_t)n * t->order, cudaMemcpyDeviceToHost, t->stream));
  BM_CHECK(cudaStreamSynchronize(t->stream));
  return BOLT_OK;
}

int bolt_filon_pieces(bolt_moment_table* t, int n, const double* f, const double* f1, const double* f2, const double* k, const double* a,
                      const double* b, double* out) {
  if (!t || n < 0) { g_bm_err = "bad argument"; return BOLT_ERR_ARG; }
  if (t->order < 3) { g_bm_err = "the Filon rule needs a table of order >= 3 (quadratic pieces)"; return BOLT_ERR_ARG; }
  if (n == 0) return BOLT_OK;
  if (!f || !f1 || !f2 || !k || !a || !b || !out) { g_bm_err = "null buffer"; return BOLT_ERR_ARG; }
  for (int i = 0; i < n; i++)
    if (!(k[i] > 0.0) || !(a[i] >= 0.0) || !(b[i] >= 0.0)) { g_bm_err = "need k > 0 and a, b >= 0"; return BOLT_ERR_ARG; }
  BM_CHECK(cudaSetDevice(t->device));
  DevBuf in, dout;
  BM_CHECK(cudaMalloc(&in.p, sizeof(double) * 6 * (size_t)n));
  BM_CHECK(cudaMalloc(&dout.p, sizeof(double) * n));
  const double* src[6] = {f, f1, f2, k, a, b};
  for (int j = 0; j < 6; j++) BM_CHECK(cudaMemcpyAsync(in.p + (size_t)j * n, src[j], sizeof(double) * n, cudaMemcpyHostToDevice, t->stream));
  filon_pieces_kernel<<<(n + 127) / 128, 128, 0, t->stream>>>(spec_of(t), view_of(t), n, in.p, in.p + n, in.p + 2 * (size_t)n, in.p + 3 * (size_t)n,
                                                            in.p + 4 * (size_t)n, in.p + 5 * (size_t)n, dout.p);
  BM_CHECK(cudaGetLastError());
  BM_CHECK(cudaMemcpyAsync(out, dout.p, sizeof(double) * n, cudaMemcpyDeviceToHost, t->stream));
  BM_CHECK(cudaStreamSynchronize(t->stream));
  return BOLT_OK;
}

int bolt_filon_chain(bolt_moment_table* t, int n_k, int n_nodes, const double* nodes, const double* f, const double* f1, const double* f2,
                     const double* k, double* out, float* kernel_ms) {
  if (!t || n_k < 0 || n_nodes < 2) { g_bm_err = "bad argument"; return BOLT_ERR_ARG; }
  if (t->order < 3) { g_bm_err = "the Filon rule needs a table of order >= 3 (quadratic pieces)"; return BOLT_ERR_ARG; }
  if (n_k == 0) return BOLT_OK;
  if (!nodes || !f || !f1 || !f2 || !k || !out) { g_bm_err = "null buffer"; return BOLT_ERR_ARG; }
  for (int i = 0; i < n_nodes; i++) if (!(nodes[i] >= 0.0)) { g_bm_err = "nodes must be >= 0"; return BOLT_ERR_ARG; }
  for (int i = 0; i < n_k; i++) if (!(k[i] > 0.0)) { g_bm_err = "need k > 0"; return BOLT_ERR_ARG; }
  BM_CHECK(cudaSetDevice(t->device));
  const size_t ns = (size_t)n_k * n_nodes;
  DevBuf src, dn, dk, dout;
  BM_CHECK(cudaMalloc(&src.p, sizeof(double) * 3 * ns));
  BM_CHECK(cudaMalloc(&dn.p, sizeof(double) * n_nodes));
  BM_CHECK(cudaMalloc(&dk.p, sizeof(double) * n_k));
  BM_CHECK(cudaMalloc(&dout.p, sizeof(double) * n_k));
  BM_CHECK(cudaMemcpyAsync(src.p, f, sizeof(double) * ns, cudaMemcpyHostToDevice, t->stream));
  BM_CHECK(cudaMemcpyAsync(src.p + ns, f1, sizeof(double) * ns, cudaMemcpyHostToDevice, t->stream));
  BM_CHECK(cudaMemcpyAsync(src.p + 2 * ns, f2, sizeof(double) * ns, cudaMemcpyHostToDevice, t->stream));
  BM_CHECK(cudaMemcpyAsync(dn.p, nodes, sizeof(double) * n_nodes, cudaMemcpyHostToDevice, t->stream));
  BM_CHECK(cudaMemcpyAsync(dk.p, k, sizeof(double) * n_k, cudaMemcpyHostToDevice, t->stream));
  cudaEvent_t e0, e1;
  BM_CHECK(cudaEventCreate(&e0));
  BM_CHECK(cudaEventCreate(&e1));
  const int reps = kernel_ms ? 3 : 1;      // timed callers get the third (warm) launch
  for (int r = 0; r < reps; r++) {
    if (r == reps - 1) cudaEventRecord(e0, t->stream);
    filon_chain_kernel<256, 4><<<n_k, 256, 0, t->stream>>>(spec_of(t), view_of(t), n_nodes, dn.p, src.p, src.p + ns, src.p + 2 * ns, dk.p, dout.p);
  }
  cudaEventRecord(e1, t->stream);
  BM_CHECK(cudaGetLastError());
  BM_CHECK(cudaMemcpyAsync(out, dout.p, sizeof(double) * n_k, cudaMemcpyDeviceToHost, t->stream));
  BM_CHECK(cudaStreamSynchronize(t->stream));
  if (kernel_ms) cudaEventElapsedTime(kernel_ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return BOLT_OK;
}

}  // extern "C"
