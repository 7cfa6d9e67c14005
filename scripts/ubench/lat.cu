// Latency microbenchmarks on one warp (B200): dependent DFMA / DADD / DMUL, LDS, SHFL, rcp.approx.f64 + Newton, exp, sqrt.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double fast_rcp(double x) {
  double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0); y = fma(y, e, y); e = fma(-x, y, 1.0); y = fma(y, e, y); return y;
}
__global__ void k(double* out, long long* cyc, double a, double b, int n) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (double)((i * 7 + 1) % 1024);
  __syncthreads();
  double x = threadIdx.x * 1e-3 + 1.0;
  long long t0, t1;
  // 0: dependent DFMA
  t0 = clock64(); for (int i = 0; i < n; i++) x = fma(x, a, b); t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // 1: dependent DADD
  t0 = clock64(); for (int i = 0; i < n; i++) x = x + b; t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
  // 2: 4 independent DFMA chains
  double y0 = x, y1 = x + 1, y2 = x + 2, y3 = x + 3;
  t0 = clock64(); for (int i = 0; i < n; i++) { y0 = fma(y0, a, b); y1 = fma(y1, a, b); y2 = fma(y2, a, b); y3 = fma(y3, a, b); } t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
  x = y0 + y1 + y2 + y3;
  // 3: dependent LDS (pointer chase)
  int idx = threadIdx.x;
  t0 = clock64(); for (int i = 0; i < n; i++) idx = (int)sm[idx & 1023]; t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
  x += idx;
  // 4: dependent SHFL (64-bit)
  t0 = clock64(); for (int i = 0; i < n; i++) x = __shfl_xor_sync(0xffffffffu, x, 1) + b; t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
  // 5: dependent fast_rcp
  x = fabs(x) + 1.5;
  t0 = clock64(); for (int i = 0; i < n; i++) x = fast_rcp(x) + 1.0; t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
  // 6: dependent exp
  t0 = clock64(); for (int i = 0; i < n; i++) x = exp(-x) + 0.5; t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
  // 7: dependent sqrt
  t0 = clock64(); for (int i = 0; i < n; i++) x = sqrt(x + 1.0); t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
  // 8: dependent 1/x (IEEE division)
  t0 = clock64(); for (int i = 0; i < n; i++) x = 1.0 / (x + 1.0); t1 = clock64(); if (threadIdx.x == 0) cyc[8] = t1 - t0;
  // 9: warp_sum (5 rounds)
  t0 = clock64(); for (int i = 0; i < n; i++) { for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o); x *= 0.03; } t1 = clock64(); if (threadIdx.x == 0) cyc[9] = t1 - t0;
  // 10: dependent FFMA (fp32) for reference
  float f = (float)x;
  t0 = clock64(); for (int i = 0; i < n; i++) f = fmaf(f, (float)a, (float)b); t1 = clock64(); if (threadIdx.x == 0) cyc[10] = t1 - t0;
  // 11: log
  x = fabs(x) + 1.1;
  t0 = clock64(); for (int i = 0; i < n; i++) x = log(x + 2.0); t1 = clock64(); if (threadIdx.x == 0) cyc[11] = t1 - t0;
  out[threadIdx.x + blockIdx.x * blockDim.x] = x + f;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8 * 1024 * 64); cudaMalloc(&cyc, 8 * 16);
  const int n = 4096;
  const char* names[12] = {"DFMA dep", "DADD dep", "4xDFMA indep (per 4)", "LDS dep", "SHFL64+DADD dep", "fast_rcp+DADD dep", "exp+DADD dep", "sqrt(+DADD) dep", "1/x dep", "warp_sum+DMUL", "FFMA dep", "log dep"};
  for (int warps = 1; warps <= 32; warps *= 2) {
    k<<<1, 32 * warps>>>(out, cyc, 0.999999, 1e-9, n); cudaDeviceSynchronize();
    k<<<1, 32 * warps>>>(out, cyc, 0.999999, 1e-9, n); cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, cyc, 8 * 16, cudaMemcpyDeviceToHost);
    printf("warps per CTA = %d (one SM; warp 0's cycles per iteration)\n", warps);
    for (int i = 0; i < 12; i++) printf("  %-24s %.1f\n", names[i], (double)h[i] / n);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
