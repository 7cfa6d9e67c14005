// Separate translation unit of K1 with partials, one CTA per mode (hierarchy_dual_cta.cuh): compiled in parallel with bolt_capi.cu.
#include <algorithm>
#include "hierarchy_dual_cta.cuh"

namespace bolt {

int k1_dual_cta_init_constants() {     // this unit's copy of the l/(2l+1) tables
  double rl[MAX_L + 1], rl1[MAX_L + 1];
  for (int l = 0; l <= MAX_L; l++) { rl[l] = (double)l / (double)(2 * l + 1); rl1[l] = 1.0 - rl[l]; }
  if (cudaMemcpyToSymbol(c_rl, rl, sizeof(rl)) != cudaSuccess) return 1;
  if (cudaMemcpyToSymbol(c_rl1, rl1, sizeof(rl1)) != cudaSuccess) return 1;
  return 0;
}

template <class TR, int NP>
static cudaError_t launch_t(const SolveParams& p, int num_sms, cudaStream_t st) {
  auto kern = hierarchy_dual_cta_kernel<TR, NP>;
  const size_t smem = k1_dual_cta_smem_doubles<TR, NP>() * sizeof(double);
  const int threads = 32 * (1 + NP);
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
  int occ = 0;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem)) != cudaSuccess) return e;
  if (occ < 1) return cudaErrorInvalidConfiguration;
  const int grid = std::max(1, std::min(p.nk, occ * num_sms));
  kern<<<grid, threads, smem, st>>>(p);
  return cudaGetLastError();
}

// l_gamma = 8, l_nu = 8, l_mnu = 10, nq = 15 (source_grid's truncations, src/spectra.jl:11) with np carried partials
cudaError_t k1_dual_cta_launch(const SolveParams& p, int np, int num_sms, cudaStream_t st) {
  typedef Trunc<8, 8, 10, 15, 19> TR;
  switch (np) {
    case 1: return launch_t<TR, 1>(p, num_sms, st);
    case 2: return launch_t<TR, 2>(p, num_sms, st);
    case 3: return launch_t<TR, 3>(p, num_sms, st);
    case 4: return launch_t<TR, 4>(p, num_sms, st);
    case 6: return launch_t<TR, 6>(p, num_sms, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace bolt
