// Separate translation unit of the pipelined K1 (hierarchy_pipe.cuh): compiled in parallel with bolt_capi.cu.
#include <algorithm>
#include "hierarchy_pipe.cuh"

namespace bolt {

int k1_pipe_init_constants() {     // this unit's copy of the l/(2l+1) tables
  double rl[MAX_L + 1], rl1[MAX_L + 1];
  for (int l = 0; l <= MAX_L; l++) { rl[l] = (double)l / (double)(2 * l + 1); rl1[l] = 1.0 - rl[l]; }
  if (cudaMemcpyToSymbol(c_rl, rl, sizeof(rl)) != cudaSuccess) return 1;
  if (cudaMemcpyToSymbol(c_rl1, rl1, sizeof(rl1)) != cudaSuccess) return 1;
  return 0;
}

template <class TR>
static cudaError_t launch_t(const SolveParams& p, int num_sms, cudaStream_t st, int* grid_out) {
  auto kern = hierarchy_pipe_kernel<TR>;
  const size_t smem = (size_t)PipeLayout<TR>::TOTAL * sizeof(double);
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
  int occ = 0;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, PIPE_THREADS, smem)) != cudaSuccess) return e;
  if (occ < 1) return cudaErrorInvalidConfiguration;
  const int grid = std::max(1, std::min(p.nk, occ * num_sms));
  if (grid_out) *grid_out = grid;
  kern<<<grid, PIPE_THREADS, smem, st>>>(p);
  return cudaGetLastError();
}

// l_gamma = 8 or 10 with l_nu = 8, l_mnu = 10, nq = 15 (source_grid's truncations, src/spectra.jl:11)
cudaError_t k1_pipe_launch(const SolveParams& p, int num_sms, cudaStream_t st, int* grid_out) {
  if (p.L == 8) return launch_t<Trunc<8, 8, 10, 15, 19>>(p, num_sms, st, grid_out);
  if (p.L == 10) return launch_t<Trunc<10, 8, 10, 15, 19>>(p, num_sms, st, grid_out);
  return cudaErrorInvalidValue;
}

}  // namespace bolt
