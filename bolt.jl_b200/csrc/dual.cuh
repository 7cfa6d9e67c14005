// Forward-mode dual numbers for the device kernels: value + NP partials, same memory order as
// ForwardDiff.Dual{Tag,Float64,NP} (value first).  The reference differentiates the whole pipeline by running it on
// Dual-typed CosmoParams (examples/plot_deriv_cl.jl:28-33); here the partials ride along in registers through K1
// and K2 so that value and gradient come from one pass.
#pragma once
#include <cuda_runtime.h>

namespace bolt {

constexpr int MAX_NP = 8;

template <int NP>
struct Dual {
  double v;
  double d[NP];
  __host__ __device__ Dual() {}
  __host__ __device__ Dual(double x) : v(x) {
#pragma unroll
    for (int i = 0; i < NP; i++) d[i] = 0.0;
  }
};

#define BOLT_HD __host__ __device__ __forceinline__
#define BOLT_FOR_NP _Pragma("unroll") for (int i = 0; i < NP; i++)

template <int NP> BOLT_HD double val(const Dual<NP>& a) { return a.v; }
BOLT_HD double val(double a) { return a; }

template <int NP> BOLT_HD Dual<NP> operator-(const Dual<NP>& a) { Dual<NP> r; r.v = -a.v; BOLT_FOR_NP r.d[i] = -a.d[i]; return r; }
template <int NP> BOLT_HD Dual<NP> operator+(const Dual<NP>& a, const Dual<NP>& b) { Dual<NP> r; r.v = a.v + b.v; BOLT_FOR_NP r.d[i] = a.d[i] + b.d[i]; return r; }
template <int NP> BOLT_HD Dual<NP> operator-(const Dual<NP>& a, const Dual<NP>& b) { Dual<NP> r; r.v = a.v - b.v; BOLT_FOR_NP r.d[i] = a.d[i] - b.d[i]; return r; }
template <int NP> BOLT_HD Dual<NP> operator*(const Dual<NP>& a, const Dual<NP>& b) { Dual<NP> r; r.v = a.v * b.v; BOLT_FOR_NP r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int NP> BOLT_HD Dual<NP> operator/(const Dual<NP>& a, const Dual<NP>& b) {
  Dual<NP> r; const double ib = 1.0 / b.v; r.v = a.v * ib; BOLT_FOR_NP r.d[i] = (a.d[i] - r.v * b.d[i]) * ib; return r;
}
template <int NP> BOLT_HD Dual<NP> operator+(const Dual<NP>& a, double b) { Dual<NP> r = a; r.v += b; return r; }
template <int NP> BOLT_HD Dual<NP> operator+(double b, const Dual<NP>& a) { Dual<NP> r = a; r.v += b; return r; }
template <int NP> BOLT_HD Dual<NP> operator-(const Dual<NP>& a, double b) { Dual<NP> r = a; r.v -= b; return r; }
template <int NP> BOLT_HD Dual<NP> operator-(double b, const Dual<NP>& a) { Dual<NP> r; r.v = b - a.v; BOLT_FOR_NP r.d[i] = -a.d[i]; return r; }
template <int NP> BOLT_HD Dual<NP> operator*(const Dual<NP>& a, double b) { Dual<NP> r; r.v = a.v * b; BOLT_FOR_NP r.d[i] = a.d[i] * b; return r; }
template <int NP> BOLT_HD Dual<NP> operator*(double b, const Dual<NP>& a) { return a * b; }
template <int NP> BOLT_HD Dual<NP> operator/(const Dual<NP>& a, double b) { return a * (1.0 / b); }
template <int NP> BOLT_HD Dual<NP> operator/(double a, const Dual<NP>& b) {
  Dual<NP> r; const double ib = 1.0 / b.v; r.v = a * ib; BOLT_FOR_NP r.d[i] = -r.v * b.d[i] * ib; return r;
}
template <int NP> BOLT_HD Dual<NP>& operator+=(Dual<NP>& a, const Dual<NP>& b) { a = a + b; return a; }
template <int NP> BOLT_HD Dual<NP>& operator-=(Dual<NP>& a, const Dual<NP>& b) { a = a - b; return a; }
template <int NP> BOLT_HD Dual<NP>& operator+=(Dual<NP>& a, double b) { a.v += b; return a; }

template <int NP> BOLT_HD Dual<NP> dsqrt(const Dual<NP>& a) { Dual<NP> r; r.v = sqrt(a.v); const double h = 0.5 / r.v; BOLT_FOR_NP r.d[i] = a.d[i] * h; return r; }
BOLT_HD double dsqrt(double a) { return sqrt(a); }
template <int NP> BOLT_HD Dual<NP> dexp(const Dual<NP>& a) { Dual<NP> r; r.v = exp(a.v); BOLT_FOR_NP r.d[i] = a.d[i] * r.v; return r; }
BOLT_HD double dexp(double a) { return exp(a); }
template <int NP> BOLT_HD Dual<NP> dlog(const Dual<NP>& a) { Dual<NP> r; r.v = log(a.v); const double ia = 1.0 / a.v; BOLT_FOR_NP r.d[i] = a.d[i] * ia; return r; }
BOLT_HD double dlog(double a) { return log(a); }
// a^b for a > 0
template <int NP> BOLT_HD Dual<NP> dpow(const Dual<NP>& a, const Dual<NP>& b) { return dexp(b * dlog(a)); }
template <int NP> BOLT_HD Dual<NP> dpow(const Dual<NP>& a, double b) { Dual<NP> r; r.v = pow(a.v, b); const double f = b * r.v / a.v; BOLT_FOR_NP r.d[i] = a.d[i] * f; return r; }
template <int NP> BOLT_HD Dual<NP> dpow(double a, const Dual<NP>& b) { Dual<NP> r; r.v = pow(a, b.v); const double f = r.v * log(a); BOLT_FOR_NP r.d[i] = b.d[i] * f; return r; }
BOLT_HD double dpow(double a, double b) { return pow(a, b); }

// number of doubles per element
template <class T> struct NumComp { static constexpr int value = 1; };
template <int NP> struct NumComp<Dual<NP>> { static constexpr int value = 1 + NP; };

}  // namespace bolt
