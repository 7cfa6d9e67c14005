// FFTLog on the device (src/util.jl:33-108: plan_fftlog, U_mu, u_m, k0r0_low_ringing, mul!, ldiv!) -- SURVEY 8f row n4.
// The reference builds the coefficients u_m with SpecialFunctions.loggamma of a complex argument and applies them between an
// FFT and an inverse FFT (FFTW).  Here: one CTA, the log-gamma by Stirling's series after an upward shift of the argument, a
// radix-2 FFT in shared memory (N a power of two, N <= 4096).  Not called by any spectrum function of the reference (nor of
// this library): a utility row, kept small.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace bolt {

struct cplx { double re, im; };
__host__ __device__ inline cplx cmul(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__host__ __device__ inline cplx cdiv(cplx a, cplx b) { const double d = b.re * b.re + b.im * b.im; return {(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d}; }
__host__ __device__ inline cplx clog(cplx a) { return {0.5 * log(a.re * a.re + a.im * a.im), atan2(a.im, a.re)}; }
__host__ __device__ inline cplx cexp(cplx a) { const double e = exp(a.re); double s, c; sincos(a.im, &s, &c); return {e * c, e * s}; }

// log Gamma(z) for Re z > 0, up to an integer multiple of 2 pi i in the imaginary part (every use below is insensitive to it:
// exp(...) in U_mu, `arg - round(arg)` in the low-ringing condition).  Shift z up until |z| >= 10, Stirling with 8 Bernoulli terms.
__host__ __device__ inline cplx cloggamma(cplx z) {
  cplx shift = {0.0, 0.0};
  while (z.re * z.re + z.im * z.im < 100.0) { const cplx l = clog(z); shift.re += l.re; shift.im += l.im; z.re += 1.0; }
  const cplx lz = clog(z);
  cplx r = cmul({z.re - 0.5, z.im}, lz);
  r.re += -z.re + 0.9189385332046727418;      // ln(2 pi)/2
  r.im += -z.im;
  const double B[8] = {1.0 / 12.0, -1.0 / 360.0, 1.0 / 1260.0, -1.0 / 1680.0, 1.0 / 1188.0, -691.0 / 360360.0, 1.0 / 156.0, -3617.0 / 122400.0};
  const cplx iz = cdiv({1.0, 0.0}, z), iz2 = cmul(iz, iz);
  cplx p = iz;
  for (int k = 0; k < 8; k++) { r.re += B[k] * p.re; r.im += B[k] * p.im; p = cmul(p, iz2); }
  return {r.re - shift.re, r.im - shift.im};
}

// U_mu(mu, x) = 2^x Gamma((mu+1+x)/2) / Gamma((mu+1-x)/2)        (util.jl:76)
__host__ __device__ inline cplx fftlog_U(double mu, cplx x) {
  const cplx a = cloggamma({0.5 * (mu + 1.0 - x.re), -0.5 * x.im}), b = cloggamma({0.5 * (mu + 1.0 + x.re), 0.5 * x.im});
  const double ln2 = 0.69314718055994530942;
  return cexp({x.re * ln2 - a.re + b.re, x.im * ln2 - a.im + b.im});
}
// u_m = (k0 r0)^(-2 pi i m/(dlnr N)) U_mu(mu, q + 2 pi i m/(dlnr N))        (util.jl:77)
__host__ __device__ inline cplx fftlog_um(double m, double mu, double q, double dlnr, double k0r0, int N) {
  const double w = 2.0 * 3.14159265358979323846 * m / (dlnr * N);
  const cplx ph = cexp({0.0, -w * log(k0r0)});
  return cmul(ph, fftlog_U(mu, {q, w}));
}
// util.jl:79-89 (from pyfftlog)
__host__ __device__ inline double fftlog_k0r0_low_ringing(int N, double mu, double q, double L, double k0r0) {
  const double pi = 3.14159265358979323846;
  const double dlnr = L / (N - 1);
  const double xp = (mu + 1.0 + q) / 2.0, xm = (mu + 1.0 - q) / 2.0, y = pi / 2.0 / dlnr;
  const cplx zp = cloggamma({xp, y}), zm = cloggamma({xm, y});
  const double arg = log(2.0 / k0r0) / dlnr + (zp.im + zm.im) / pi;
  return k0r0 * exp((arg - rint(arg)) * dlnr);
}

// in-place radix-2 FFT of N = 2^lg points in shared memory; sign = -1 forward, +1 inverse (unnormalised)
__device__ inline void fft_smem(double2* a, int N, int lg, int sign) {
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const int j = (int)(__brev((unsigned)i) >> (32 - lg));
    if (j > i) { const double2 t = a[i]; a[i] = a[j]; a[j] = t; }
  }
  __syncthreads();
  for (int len = 2; len <= N; len <<= 1) {
    const int half = len >> 1;
    for (int t = threadIdx.x; t < N / 2; t += blockDim.x) {
      const int blk = t / half, pos = t - blk * half;
      const int i0 = blk * len + pos, i1 = i0 + half;
      double s, c; sincospi((double)sign * 2.0 * (double)pos / (double)len, &s, &c);
      const double2 u = a[i0], v = a[i1];
      const double2 w = {v.x * c - v.y * s, v.x * s + v.y * c};
      a[i0] = {u.x + w.x, u.y + w.y}; a[i1] = {u.x - w.x, u.y - w.y};
    }
    __syncthreads();
  }
}

// mul! / ldiv! (util.jl:91-107).  r[N] the log-spaced abscissae, a the samples (a_im may be null), y [N] complex out.
__global__ void fftlog_kernel(const double* __restrict__ r, int N, int lg, double mu, double q, double dlnr, double k0r0, int inverse,
                              const double* __restrict__ a_re, const double* __restrict__ a_im, double2* __restrict__ y) {
  extern __shared__ double2 buf[];
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const double s = pow(r[i], -q);
    buf[i] = {a_re[i] * s, (a_im ? a_im[i] : 0.0) * s};
  }
  __syncthreads();
  fft_smem(buf, N, lg, -1);
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const double m = (i < N / 2) ? (double)i : (double)(i - N);            // fftfreq(N, N)
    cplx u = fftlog_um(m, mu, q, dlnr, k0r0, N);
    if (i == N / 2) u.im = 0.0;                                          // eq. 19: the Nyquist coefficient is real
    const cplx v = {buf[i].x, buf[i].y};
    const cplx w = inverse ? cdiv(v, u) : cmul(v, u);
    buf[i] = {w.re, w.im};
  }
  __syncthreads();
  fft_smem(buf, N, lg, +1);
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const double s = pow(r[i], q) / (double)N;
    y[i] = {buf[i].x * s, buf[i].y * s};
  }
}

}  // namespace bolt
