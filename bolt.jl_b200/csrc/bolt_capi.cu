// libbolt_cuda.so -- C ABI (include/bolt_cuda.h) over the sm_100a kernels.
// There is deliberately no CPU fallback: every entry point fails with BOLT_ERR_CUDA if no device is usable.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>
#include <memory>
#include <dlfcn.h>

// shared-memory layout of the register-resident value kernel: 19 = nq + 4 (compact + shared all-zero column for the idle
// lanes, flat stage assembly: 3 % faster), 32 = one private column per lane
#ifndef K1_NCH
#define K1_NCH 19
#endif
#include "hierarchy_kernel.cuh"
#include "hierarchy_dual.cuh"
#include "hierarchy_dual_reg.cuh"
#include "projection_kernel.cuh"
#include "fftlog.cuh"

using namespace bolt;

namespace bolt {      // kernels with their own translation units (compiled in parallel)
int k1_dual_cta_init_constants();  // k1_dual_cta.cu: K1 with partials, one CTA per mode (value warp + one warp per sensitivity system)
cudaError_t k1_dual_cta_launch(const SolveParams& p, int np, int num_sms, cudaStream_t st);
int k1_pipe_init_constants();      // k1_pipe.cu: the pipelined K1 (one CTA per k-mode, six warps by role)
cudaError_t k1_pipe_launch(const SolveParams& p, int num_sms, cudaStream_t st, int* grid_out);
}

struct PoolBlock { void* p; size_t bytes; bool used; };

struct bolt_ctx {
  std::vector<PoolBlock> pool;   // grow-only device workspace cache (cudaMalloc/cudaFree are synchronous and slow)
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;    // j_l tables are independent of K1: they are built concurrently with the hierarchy solve
  cudaEvent_t ev_tab = nullptr;
  cudaEvent_t ev[8];
  std::string err;
  double timing[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int* d_counter = nullptr;
  double bessel_xmax = 0.0;  // 0: kgrid[end]*eta0 (src/spectra.jl:85); > 0: caller-fixed table range (bolt_set_bessel_xmax)
  double* d_dbg = nullptr;   // step log (only when BOLT_DEBUG_STEPS is set)
  void* comm = nullptr;      // ncclComm_t of this rank (bolt_comm_init); collectives run on `stream`
  int rank = 0, nranks = 1;
};
constexpr int DBG_CAP = 1 << 16;

struct bolt_cosmo {
  DevCosmo h;              // host copy (table pointers are device pointers)
  DevCosmo* d = nullptr;   // device copy
  double* d_tables = nullptr;
  double* d_dtables = nullptr;         // partial tables [NTABLES][np][n_x+2] (nd > 1)
  const DevCosmo** d_list = nullptr;   // device array {d}: the 1-cosmology work list of K1
  // K1 view of the partials: parameters that never reach the hierarchy (A, n: they enter through the primordial weight of
  // K2 only) have identically zero sensitivities and are not carried through the ODE solve
  DevCosmo* d_k1 = nullptr; const DevCosmo** d_list_k1 = nullptr; double* d_dtables_k1 = nullptr;
  int np_k1 = 0; int map_k1[MAX_NP] = {0};
  // single-partial views of the K1 cosmology (partial j moved to slot 0, np = 1): what a sensitivity warp of the CTA kernel
  // with partials evaluates its Dual<1> background and G_j rows on (hierarchy_dual_cta.cuh)
  DevCosmo h_k1; DevCosmo* d_views = nullptr; const DevCosmo** d_view_list = nullptr;
};

// DFMA throughput microbenchmark: 8 independent FMA chains per thread (the FP64 roofline denominator;
// MEASURED_PEAKS.json carries no FP64 figure).
__global__ void fp64_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// eta(x_grid[end]) evaluated with the device's own spline arithmetic, so that y = k(eta_end - eta(x)) is exactly 0 in the
// last row of S_P, as in the reference (perturbations.jl:401-403).
__global__ void eta_end_kernel(DevCosmo* c) {
  c->eta_end = spline_eval(c->tab[BOLT_T_eta], c->n_x, c->x0, c->dx, c->x0 + c->dx * (c->n_x - 1));
}

namespace {

// NCCL entry points, bound at run time (bolt_comm_*): see the multi-GPU section below
struct NcclId128 { char b[128]; };          // ncclUniqueId (nccl.h: 128 opaque bytes, passed by value)
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId128, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mutex;
constexpr int NCCL_INT32 = 2, NCCL_INT64 = 4, NCCL_FLOAT64 = 8, NCCL_SUM = 0;      // ncclDataType_t / ncclRedOp_t (nccl.h, stable ABI)

int fail(bolt_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}
#define CUDA_OK(call)                                                                              \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(ctx, BOLT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
  } while (0)

void* pool_get(bolt_ctx* ctx, size_t bytes) {
  PoolBlock* best = nullptr;
  for (auto& b : ctx->pool) if (!b.used && b.bytes >= bytes && (!best || b.bytes < best->bytes)) best = &b;
  if (best && best->bytes <= 2 * bytes + (1 << 20)) { best->used = true; return best->p; }
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) {   // out of memory: drop the idle cache and retry once
    for (auto it = ctx->pool.begin(); it != ctx->pool.end();) { if (!it->used) { cudaFree(it->p); it = ctx->pool.erase(it); } else ++it; }
    if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
  }
  ctx->pool.push_back({p, bytes, true});
  return p;
}
void pool_put(bolt_ctx* ctx, void* p) { for (auto& b : ctx->pool) if (b.p == p) { b.used = false; return; } }

// Device scratch buffer leased from the context's pool for the duration of one call.
template <class T>
struct DevBuf {
  bolt_ctx* ctx = nullptr; T* p = nullptr; size_t n = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  ~DevBuf() { if (p) pool_put(ctx, p); }
  cudaError_t alloc(bolt_ctx* c, size_t count) {
    ctx = c; n = count;
    if (!count) return cudaSuccess;
    p = (T*)pool_get(c, count * sizeof(T));
    return p ? cudaSuccess : cudaErrorMemoryAllocation;
  }
};

bool g_const_init[64] = {false};
std::mutex g_const_mutex;      // contexts may be created from many host threads (one per thread x device)

int init_constants(bolt_ctx* ctx) {
  std::lock_guard<std::mutex> lock(g_const_mutex);
  if (ctx->device >= 64) return fail(ctx, BOLT_ERR_ARG, "device ordinal out of range");
  if (g_const_init[ctx->device]) return BOLT_OK;
  double rl[MAX_L + 1];
  for (int l = 0; l <= MAX_L; l++) rl[l] = (double)l / (double)(2 * l + 1);
  CUDA_OK(cudaMemcpyToSymbol(c_rl, rl, sizeof(rl)));
  double rl1[MAX_L + 1];
  for (int l = 0; l <= MAX_L; l++) rl1[l] = 1.0 - rl[l];
  CUDA_OK(cudaMemcpyToSymbol(c_rl1, rl1, sizeof(rl1)));
  if (k1_pipe_init_constants()) return fail(ctx, BOLT_ERR_CUDA, "constant tables of k1_pipe.cu");
  if (k1_dual_cta_init_constants()) return fail(ctx, BOLT_ERR_CUDA, "constant tables of k1_dual_cta.cu");
  g_const_init[ctx->device] = true;
  return BOLT_OK;
}

int check_opts(bolt_ctx* ctx, const bolt_cosmo* c, const bolt_opts* o) {
  if (!o) return fail(ctx, BOLT_ERR_ARG, "opts is null");
  if (o->l_gamma < 3 || o->l_nu < 3 || o->l_mnu < 3)
    return fail(ctx, BOLT_ERR_ARG, "truncations must be >= 3 (source_function needs Theta_3, perturbations.jl:378)");
  if (o->l_gamma > MAX_L || o->l_nu > MAX_L || o->l_mnu > MAX_L) return fail(ctx, BOLT_ERR_ARG, "truncation too large");
  if (o->mode == BOLT_MODE_FIXED && !(o->fixed_dt > 0)) return fail(ctx, BOLT_ERR_ARG, "fixed_dt must be > 0");
  if (o->mode == BOLT_MODE_ADAPTIVE && !(o->reltol > 0 && o->abstol > 0)) return fail(ctx, BOLT_ERR_ARG, "tolerances must be > 0");
  return BOLT_OK;
}

// the one-warp-per-mode kernel with WPB warps per block in lockstep (hierarchy_kernel_t's header)
template <class TR, int WPB>
int launch_k1_lockstep(bolt_ctx* ctx, const SolveParams& p) {
  auto kern = hierarchy_kernel_t<TR, WPB>;
  const size_t smem = WPB * ((size_t)k1_num_arrays<TR>() * k1_array_len<TR>(p.n) + k1_extra_doubles<TR>()) * sizeof(double);
  CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  int occ = 0;
  CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * WPB, smem));
  if (occ < 1) return fail(ctx, BOLT_ERR_UNSUPPORTED, "state does not fit in shared memory");
  const int grid = std::max(1, std::min((p.nk + WPB - 1) / WPB, occ * ctx->num_sms));
  CUDA_OK(cudaMemsetAsync(ctx->d_counter, 0, sizeof(int), ctx->stream));
  CUDA_OK(cudaEventRecord(ctx->ev[0], ctx->stream));
  kern<<<grid, 32 * WPB, smem, ctx->stream>>>(p);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(ctx->ev[1], ctx->stream));
  ctx->timing[4] += 1;
  return BOLT_OK;
}

template <class TR>
int launch_k1(bolt_ctx* ctx, const SolveParams& p) {
  auto kern = hierarchy_kernel_t<TR>;
  const size_t smem = ((size_t)k1_num_arrays<TR>() * k1_array_len<TR>(p.n) + k1_extra_doubles<TR>()) * sizeof(double);
  CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  int occ = 0;
  CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32, smem));
  if (occ < 1) return fail(ctx, BOLT_ERR_UNSUPPORTED, "state does not fit in shared memory");
  const int grid = std::max(1, std::min(p.nk, occ * ctx->num_sms));
  CUDA_OK(cudaMemsetAsync(ctx->d_counter, 0, sizeof(int), ctx->stream));
  CUDA_OK(cudaEventRecord(ctx->ev[0], ctx->stream));
  kern<<<grid, 32, smem, ctx->stream>>>(p);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(ctx->ev[1], ctx->stream));
  ctx->timing[4] += 1;
  return BOLT_OK;
}

// One CTA per k-mode, six warps by role (hierarchy_pipe.cuh)
int launch_k1_pipe(bolt_ctx* ctx, const SolveParams& p) {
  CUDA_OK(cudaMemsetAsync(ctx->d_counter, 0, 257 * sizeof(int), ctx->stream));
  CUDA_OK(cudaEventRecord(ctx->ev[0], ctx->stream));
  CUDA_OK(k1_pipe_launch(p, ctx->num_sms, ctx->stream, nullptr));
  CUDA_OK(cudaEventRecord(ctx->ev[1], ctx->stream));
  ctx->timing[4] += 1;
  return BOLT_OK;
}

template <int NP>
int launch_k1_dual(bolt_ctx* ctx, const SolveParams& p) {
  auto kern = hierarchy_dual_kernel<NP>;
  const size_t smem = k1_dual_smem_doubles<NP>(p.n) * sizeof(double);
  if (smem > 227 * 1024) return fail(ctx, BOLT_ERR_UNSUPPORTED, "state x partials does not fit in shared memory");
  CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  int occ = 0;
  CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32, smem));
  if (occ < 1) return fail(ctx, BOLT_ERR_UNSUPPORTED, "state x partials does not fit in shared memory");
  const int grid = std::max(1, std::min(p.nk, occ * ctx->num_sms));
  CUDA_OK(cudaMemsetAsync(ctx->d_counter, 0, sizeof(int), ctx->stream));
  CUDA_OK(cudaEventRecord(ctx->ev[0], ctx->stream));
  kern<<<grid, 32, smem, ctx->stream>>>(p);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(ctx->ev[1], ctx->stream));
  ctx->timing[4] += 1;
  return BOLT_OK;
}

template <class TR, int NP>
int launch_k1_dual_reg(bolt_ctx* ctx, const SolveParams& p) {
  auto kern = hierarchy_dual_reg_kernel<TR, NP>;
  const size_t smem = k1_dualreg_smem_doubles<TR, NP>() * sizeof(double);
  CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  int occ = 0;
  CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32, smem));
  if (occ < 1) return fail(ctx, BOLT_ERR_UNSUPPORTED, "dual state does not fit in shared memory");
  const int grid = std::max(1, std::min(p.nk, occ * ctx->num_sms));
  CUDA_OK(cudaMemsetAsync(ctx->d_counter, 0, sizeof(int), ctx->stream));
  CUDA_OK(cudaEventRecord(ctx->ev[0], ctx->stream));
  kern<<<grid, 32, smem, ctx->stream>>>(p);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(ctx->ev[1], ctx->stream));
  ctx->timing[4] += 1;
  return BOLT_OK;
}

// Launch K1 on device buffers.  cos_list: device array of ncos cosmology pointers; work item g (0 <= g < nk) belongs to
// cosmology g / nk_per.
int launch_hierarchy(bolt_ctx* ctx, const DevCosmo* const* cos_list, int nq, int np, int out_nd, const int* comp_map, int nk_per, const double* d_k, const int* d_order, int nk,
                     const bolt_opts* o, double* d_ST, double* d_SP, double* d_hist, double* d_final, int* d_status,
                     long long* d_nsteps, long long* d_nreject, const DevCosmo* const* view_list = nullptr) {
  SolveParams p;
  p.sync_mask = getenv("BOLT_K1_SYNCMASK") ? (int)strtol(getenv("BOLT_K1_SYNCMASK"), nullptr, 0) : 0x02;     // lockstep kernel: meet once per step (before stage 1)
  p.view_list = view_list;
  p.cos_list = cos_list; p.nk_per = nk_per; p.k = d_k; p.order = d_order; p.nk = nk;
  p.L = o->l_gamma; p.Lnu = o->l_nu; p.Lm = o->l_mnu;
  p.n = bolt_state_dim(p.L, p.Lnu, p.Lm, nq);
  p.mode = o->mode; p.reltol = o->reltol; p.abstol = o->abstol; p.fixed_dt = o->fixed_dt;
  p.max_steps = o->max_steps; p.ix_first = o->ix_first;
  p.S_T = d_ST; p.S_P = d_SP; p.u_hist = d_hist; p.u_final = d_final;
  p.status = d_status; p.nsteps = d_nsteps; p.nreject = d_nreject; p.counter = ctx->d_counter;
  p.dbg = ctx->d_dbg; p.dbg_cap = ctx->d_dbg ? DBG_CAP : 0;
  p.out_nd = out_nd;
  for (int j = 0; j < MAX_NP; j++) p.comp_map[j] = (comp_map && j < np) ? comp_map[j] : 1 + j;
  if (np > 0) {     // value + gradient in one pass
    if (!getenv("BOLT_K1_GENERIC") && nq == 15 && p.L == 8 && p.Lnu == 8 && p.Lm == 10) {     // register-resident (hierarchy_dual_reg.cuh)
      typedef Trunc<8, 8, 10, 15, 19> TRD;
      // one CTA per mode: the value system and every sensitivity system on a warp of their own, sharing the stage factorisation
      // (hierarchy_dual_cta.cuh).  BOLT_K1_DUAL_WARP=1: the one-warp-per-mode kernel below.
      if (p.view_list && !getenv("BOLT_K1_DUAL_WARP") && (np == 1 || np == 2 || np == 3 || np == 4 || np == 6)) {
        CUDA_OK(cudaMemsetAsync(ctx->d_counter, 0, sizeof(int), ctx->stream));
        CUDA_OK(cudaEventRecord(ctx->ev[0], ctx->stream));
        CUDA_OK(k1_dual_cta_launch(p, np, ctx->num_sms, ctx->stream));
        CUDA_OK(cudaEventRecord(ctx->ev[1], ctx->stream));
        ctx->timing[4] += 1;
        return BOLT_OK;
      }
      switch (np) {
        case 1: return launch_k1_dual_reg<TRD, 1>(ctx, p);
        case 2: return launch_k1_dual_reg<TRD, 2>(ctx, p);
        case 3: return launch_k1_dual_reg<TRD, 3>(ctx, p);
        case 4: return launch_k1_dual_reg<TRD, 4>(ctx, p);
        case 6: return launch_k1_dual_reg<TRD, 6>(ctx, p);
        default: break;
      }
    }
    switch (np) {     // any truncation (hierarchy_dual.cuh)
#define BOLT_DUAL_CASE(N) case N: return launch_k1_dual<N>(ctx, p);
      BOLT_DUAL_CASE(1) BOLT_DUAL_CASE(2) BOLT_DUAL_CASE(3) BOLT_DUAL_CASE(4) BOLT_DUAL_CASE(6)
#undef BOLT_DUAL_CASE
      default: return fail(ctx, BOLT_ERR_UNSUPPORTED, "this build carries 1, 2, 3, 4 or 6 partials per call");
    }
  }
  const bool force_generic = getenv("BOLT_K1_GENERIC") != nullptr;     // development switch
  if (!force_generic && nq == 15 && p.Lnu == 8 && p.Lm == 10) {          // source_grid's truncations (src/spectra.jl:11)
    // Two kernels for source_grid's truncations.  The pipelined CTA-per-mode kernel (hierarchy_pipe.cuh) has 2.1-2.6x lower
    // per-mode latency (one mode: 12.9 vs 33.1 ms; 296 modes: 15.4 vs 33.1 ms) but holds only 2 modes per SM; the one-warp-per-mode
    // kernel holds 8 and wins once the modes outnumber the resident CTAs several times (500 modes: 16.0 vs 34.1 ms, 1000: 30.5 vs 42.3,
    // 1400: 41.5 vs 44.4, 2000: 59.6 vs 46.8 ms).
    // BOLT_K1_PIPE=1 / BOLT_K1_WARP=1 force one.
    if (p.L == 8 || p.L == 10) {
      if (getenv("BOLT_K1_PIPE") || (!getenv("BOLT_K1_WARP") && p.nk <= 9 * ctx->num_sms)) return launch_k1_pipe(ctx, p);
    }
    // Beyond that the one-warp-per-mode kernel, four warps per block in lockstep at stage granularity (shared instruction-cache
    // fills: 592 / 1184 / 2000 modes 36.4 / 42.4 / 45.9 ms against 36.7 / 46.2 / 49.9 ms for independent warps, same box;
    // 2 warps 35.3 / 42.5 / 48.7, 8 warps 44.9 / 44.8 / 49.5).  BOLT_K1_WPB=1: independent warps.
    {
      const int wpb = getenv("BOLT_K1_WPB") ? atoi(getenv("BOLT_K1_WPB")) : 4;
      if (p.L == 8 && wpb == 4) return launch_k1_lockstep<Trunc<8, 8, 10, 15, K1_NCH>, 4>(ctx, p);
      if (p.L == 10 && wpb == 4) return launch_k1_lockstep<Trunc<10, 8, 10, 15, K1_NCH>, 4>(ctx, p);
      if (p.L == 8 && wpb == 2) return launch_k1_lockstep<Trunc<8, 8, 10, 15, K1_NCH>, 2>(ctx, p);
      if (p.L == 8 && wpb == 8) return launch_k1_lockstep<Trunc<8, 8, 10, 15, K1_NCH>, 8>(ctx, p);
    }
    if (p.L == 8) return launch_k1<Trunc<8, 8, 10, 15, K1_NCH>>(ctx, p);         // l_gamma = 8: the reference default
    if (p.L == 10) return launch_k1<Trunc<10, 8, 10, 15, K1_NCH>>(ctx, p);       // l_gamma = 10: BASELINE config 1
  }
  // any other truncation (plin: 50/50/20, C4: 50/8/10): the runtime-truncation path needs l_max >= 2 on every chain
  if (!force_generic && p.L >= 2 && p.Lnu >= 2 && p.Lm >= 2) {
    if (getenv("BOLT_K1_RT_WPB")) {      // development: lockstep blocks for the runtime-truncation path
      const int wpb = atoi(getenv("BOLT_K1_RT_WPB"));
      if (wpb == 2) return launch_k1_lockstep<TruncRT, 2>(ctx, p);
      if (wpb == 4) return launch_k1_lockstep<TruncRT, 4>(ctx, p);
    }
    return launch_k1<TruncRT>(ctx, p);
  }
  return launch_k1<Trunc<0, 0, 0, 0>>(ctx, p);
}
int launch_hierarchy(bolt_ctx* ctx, const bolt_cosmo* c, const double* d_k, const int* d_order, int nk, const bolt_opts* o,
                     double* d_ST, double* d_SP, double* d_hist, double* d_final, int* d_status, long long* d_nsteps,
                     long long* d_nreject) {
  const bool compact = c->d_list_k1 != nullptr;
  return launch_hierarchy(ctx, compact ? c->d_list_k1 : c->d_list, c->h.nq, c->np_k1, c->h.nd, c->map_k1, nk, d_k, d_order, nk, o, d_ST, d_SP, d_hist, d_final, d_status, d_nsteps, d_nreject,
                          c->d_view_list);
}

int upload_k_sorted(bolt_ctx* ctx, const double* k, int nk, DevBuf<double>& d_k, DevBuf<int>& d_order) {
  std::vector<int> order(nk);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return k[a] > k[b]; });   // longest solves first
  CUDA_OK(d_k.alloc(ctx, nk)); CUDA_OK(d_order.alloc(ctx, nk));
  CUDA_OK(cudaMemcpyAsync(d_k.p, k, nk * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_OK(cudaMemcpyAsync(d_order.p, order.data(), nk * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));   // `order` is a local
  return BOLT_OK;
}

int collect_timing(bolt_ctx* ctx) {
  float ms = 0;
  if (ctx->timing[4] > 0) { CUDA_OK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1])); ctx->timing[0] = ms; }
  if (ctx->timing[5] > 0) { CUDA_OK(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3])); ctx->timing[1] = ms; }
  if (ctx->timing[6] > 0) { CUDA_OK(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5])); ctx->timing[2] = ms; }
  CUDA_OK(cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7])); ctx->timing[3] = ms;
  return BOLT_OK;
}
void reset_timing(bolt_ctx* ctx) { for (int i = 0; i < 8; i++) ctx->timing[i] = 0; }

#ifndef K2_NL
#define K2_NL 4
#define K2_NT 512
#endif
constexpr int PROJ_NL = K2_NL, PROJ_NT = K2_NT;

constexpr int PROJD_NL = 2, PROJD_NT = 384;

template <int NP>
int launch_project_dual(bolt_ctx* ctx, const ProjectParamsD& pd, int groups, int nsplit, size_t smem) {
  auto kern = project_kernel_dual<PROJD_NL, PROJD_NT, NP>;
  CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3(groups, nsplit), PROJD_NT, smem, ctx->stream>>>(pd);
  CUDA_OK(cudaGetLastError());
  return BOLT_OK;
}

// K2 with partials.  Sources are [nk][n_x][nd]; d_cl = [3][nell][nd].
int project_device_dual(bolt_ctx* ctx, const bolt_cosmo* c, const double* d_ST, const double* d_SP, int nell, int ix_start, int nrows,
                        int nkd1, int ld, double dg, const int* d_ell, const double* d_Cf, const double* d_ks, const double* d_wk,
                        const int* d_jlo, const double* d_wl, double* d_cl) {
  const DevCosmo& h = c->h;
  const int np = h.np, nd = 1 + np;
  const size_t cstr = (size_t)nrows * ld;
  DevBuf<double> d_chi, d_dwk, d_SDT, d_SDP, d_part, d_partd;
  CUDA_OK(d_chi.alloc(ctx, (size_t)nd * nrows)); CUDA_OK(d_dwk.alloc(ctx, (size_t)np * nkd1));
  chi_kernel_nd<<<(nrows + 127) / 128, 128, 0, ctx->stream>>>(c->d, ix_start, nrows, d_chi.p);
  CUDA_OK(cudaGetLastError());
  dense_k_partials_kernel<<<(nkd1 + 127) / 128, 128, 0, ctx->stream>>>(c->d, d_ks, dg, d_wk, nkd1, d_dwk.p);
  CUDA_OK(cudaGetLastError());
  dim3 gsd((nkd1 + 127) / 128, nrows);
  if (d_ST) { CUDA_OK(d_SDT.alloc(ctx, (size_t)nd * cstr));
    for (int q = 0; q < nd; q++) dense_source_kernel_nd<<<gsd, 128, 0, ctx->stream>>>(d_ST, h.n_x, nd, q, ix_start, nrows, d_jlo, d_wl, nkd1, ld, h.x0, h.dx, d_SDT.p + q * cstr);
    CUDA_OK(cudaGetLastError()); ctx->timing[6] += nd; }
  if (d_SP) { CUDA_OK(d_SDP.alloc(ctx, (size_t)nd * cstr));
    for (int q = 0; q < nd; q++) dense_source_kernel_nd<<<gsd, 128, 0, ctx->stream>>>(d_SP, h.n_x, nd, q, ix_start, nrows, d_jlo, d_wl, nkd1, ld, h.x0, h.dx, d_SDP.p + q * cstr);
    CUDA_OK(cudaGetLastError()); ctx->timing[6] += nd; }
  const int groups = (nell + PROJD_NL - 1) / PROJD_NL;
  int nsplit = (8 * ctx->num_sms + groups - 1) / groups;
  nsplit = std::max(1, std::min(nsplit, std::max(1, nkd1 / PROJD_NT)));
  CUDA_OK(d_part.alloc(ctx, (size_t)nell * nsplit * 3)); CUDA_OK(d_partd.alloc(ctx, (size_t)nell * nsplit * 3 * np));
  ProjectParamsD pd;
  pd.v.Cf = d_Cf; pd.v.ells = d_ell; pd.v.nell = nell; pd.v.chi = d_chi.p; pd.v.nrows = nrows; pd.v.kscaled = d_ks; pd.v.wk = d_wk;
  pd.v.nkd1 = nkd1; pd.v.ld = ld; pd.v.SD_T = d_SDT.p; pd.v.SD_P = d_SDP.p; pd.v.nsplit = nsplit; pd.v.partial = d_part.p;
  pd.chi_d = d_chi.p + nrows; pd.dwk = d_dwk.p; pd.SD_T_d = d_ST ? d_SDT.p + cstr : nullptr; pd.SD_P_d = d_SP ? d_SDP.p + cstr : nullptr;
  pd.partial_d = d_partd.p;
  const size_t smem = ((size_t)PROJD_NL * BESSEL_NC + (size_t)nd * nrows) * sizeof(double);
  int rc;
  switch (np) {
    case 1: rc = launch_project_dual<1>(ctx, pd, groups, nsplit, smem); break;
    case 2: rc = launch_project_dual<2>(ctx, pd, groups, nsplit, smem); break;
    case 3: rc = launch_project_dual<3>(ctx, pd, groups, nsplit, smem); break;
    case 4: rc = launch_project_dual<4>(ctx, pd, groups, nsplit, smem); break;
    case 6: rc = launch_project_dual<6>(ctx, pd, groups, nsplit, smem); break;
    default: return fail(ctx, BOLT_ERR_UNSUPPORTED, "this build carries 1, 2, 3, 4 or 6 partials per call");
  }
  if (rc) return rc;
  cl_finalize_kernel_nd<<<(nell + 127) / 128, 128, 0, ctx->stream>>>(d_part.p, d_partd.p, d_ell, nell, nsplit, np, d_ST ? d_cl : nullptr,
                                                                     (d_ST && d_SP) ? d_cl + (size_t)nell * nd : nullptr,
                                                                     d_SP ? d_cl + 2 * (size_t)nell * nd : nullptr);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(ctx->ev[5], ctx->stream));
  ctx->timing[6] += 5;
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return BOLT_OK;
}

// j_l tables + B-spline prefilter for a set of multipoles (bessel_interpolator, spectra.jl:49-58).  They depend on the
// cosmology only through the table range, not on the perturbation solve, so bolt_spectra builds them on a second stream
// while K1 runs.
struct BesselTabs {
  DevBuf<int> d_ell; DevBuf<double> d_J, d_Cf, d_cp, d_iden;
  double dg = 0.0;
};
// An error after work has been enqueued: drain both streams before the DevBufs of the caller return to the pool.
int bail(bolt_ctx* ctx, int rc) {
  cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->stream2);
  return rc;
}
int check_ells(bolt_ctx* ctx, const int32_t* ell, int nell) {
  for (int i = 0; i < nell; i++) {
    if (ell[i] < 0) return fail(ctx, BOLT_ERR_ARG, "negative multipole");
    if (i > 0 && ell[i] <= ell[i - 1]) return fail(ctx, BOLT_ERR_ARG, "multipoles must be strictly increasing");
  }
  return BOLT_OK;
}
int check_projection_args(bolt_ctx* ctx, const bolt_cosmo* c, const int32_t* ell, int nell, int n_kd, int ix_start) {
  int rc = check_ells(ctx, ell, nell); if (rc) return rc;
  if (n_kd < 2 || ix_start < 0 || ix_start >= c->h.n_x - 1) return fail(ctx, BOLT_ERR_ARG, "bad dense grid / ix_start");
  return BOLT_OK;
}
int bessel_prepare(bolt_ctx* ctx, const bolt_cosmo* c, const int32_t* ell, int nell, double kd_max, cudaStream_t st, BesselTabs& bt) {
  int rc = check_ells(ctx, ell, nell); if (rc) return rc;
  // kgrid[end]*eta0 (spectra.jl:85; quadratic_k ends exactly at kmax).  The reference strips partials from this range
  // (assume_nondual, spectra.jl:52): the table GRID is not differentiated.
  const double xmax = ctx->bessel_xmax > 0.0 ? ctx->bessel_xmax : kd_max * c->h.s[BOLT_S_eta0];
  bt.dg = xmax / 5000.0;
  CUDA_OK(bt.d_ell.alloc(ctx, nell)); CUDA_OK(bt.d_J.alloc(ctx, (size_t)nell * BESSEL_NB)); CUDA_OK(bt.d_Cf.alloc(ctx, (size_t)nell * BESSEL_NC));
  const int m = BESSEL_NB - 2;
  CUDA_OK(bt.d_cp.alloc(ctx, m)); CUDA_OK(bt.d_iden.alloc(ctx, m));
  {  // Thomas multipliers of the (1/6, 2/3, 1/6) interior system
    std::vector<double> cp(m), iden(m);
    const double a = 1.0 / 6.0, b = 2.0 / 3.0;
    for (int i = 0; i < m; i++) { const double den = (i == 0) ? b : b - a * cp[i - 1]; cp[i] = a / den; iden[i] = 1.0 / den; }
    CUDA_OK(cudaMemcpyAsync(bt.d_ell.p, ell, nell * sizeof(int), cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(bt.d_cp.p, cp.data(), m * 8, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(bt.d_iden.p, iden.data(), m * 8, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaStreamSynchronize(st));      // host staging vectors go out of scope
  }
  CUDA_OK(cudaEventRecord(ctx->ev[2], st));
  bessel_table_kernel<<<(BESSEL_NB + 63) / 64, 64, 0, st>>>(bt.d_ell.p, nell, bt.dg, bt.d_J.p);
  CUDA_OK(cudaGetLastError());
  bessel_prefilter_kernel<<<(nell + 31) / 32, 32, 0, st>>>(bt.d_J.p, nell, bt.d_cp.p, bt.d_iden.p, bt.d_Cf.p);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(ctx->ev[3], st));
  CUDA_OK(cudaEventRecord(ctx->ev_tab, st));
  ctx->timing[5] += 2;
  return BOLT_OK;
}

// K2 pipeline on device buffers (tables already prepared, possibly on another stream).  d_cl = [3][nell][nd] (tt, te, ee).
int project_device(bolt_ctx* ctx, const bolt_cosmo* c, const double* d_ST, const double* d_SP, const double* d_kc, int nk,
                   BesselTabs& bt, int nell, double kd_min, double kd_max, int n_kd, int ix_start, double* d_cl) {
  const DevCosmo& h = c->h;
  if (n_kd < 2 || ix_start < 0 || ix_start >= h.n_x - 1) return fail(ctx, BOLT_ERR_ARG, "bad dense grid / ix_start");
  const int nrows = h.n_x - 1 - ix_start;           // x_grid[ix_start .. n_x-2] (spectra.jl:70-76)
  const int nkd1 = n_kd - 1;
  const int ld = (nkd1 + 31) / 32 * 32;
  const double dg = bt.dg;
  DevBuf<int>& d_ell = bt.d_ell; DevBuf<double>& d_Cf = bt.d_Cf;
  DevBuf<int> d_jlo; DevBuf<double> d_ks, d_wk, d_wl, d_chi, d_SDT, d_SDP, d_part;
  CUDA_OK(cudaStreamWaitEvent(ctx->stream, ctx->ev_tab, 0));
  CUDA_OK(d_ks.alloc(ctx, nkd1)); CUDA_OK(d_wk.alloc(ctx, nkd1)); CUDA_OK(d_wl.alloc(ctx, nkd1)); CUDA_OK(d_jlo.alloc(ctx, nkd1)); CUDA_OK(d_chi.alloc(ctx, nrows));
  CUDA_OK(cudaEventRecord(ctx->ev[4], ctx->stream));
  dense_k_kernel<<<(nkd1 + 127) / 128, 128, 0, ctx->stream>>>(d_kc, nk, kd_min, kd_max, n_kd, h.s[BOLT_S_A], h.s[BOLT_S_n], dg,
                                                              d_ks.p, d_wk.p, d_jlo.p, d_wl.p);
  CUDA_OK(cudaGetLastError());
  if (h.np > 0) return project_device_dual(ctx, c, d_ST, d_SP, nell, ix_start, nrows, nkd1, ld, dg, d_ell.p, d_Cf.p, d_ks.p, d_wk.p,
                                           d_jlo.p, d_wl.p, d_cl);
  chi_kernel<<<(nrows + 127) / 128, 128, 0, ctx->stream>>>(c->d, ix_start, nrows, d_chi.p);
  CUDA_OK(cudaGetLastError());
  dim3 gsd((nkd1 + 127) / 128, nrows);
  if (d_ST) { CUDA_OK(d_SDT.alloc(ctx, (size_t)nrows * ld));
    dense_source_kernel<<<gsd, 128, 0, ctx->stream>>>(d_ST, h.n_x, ix_start, nrows, d_jlo.p, d_wl.p, nkd1, ld, h.x0, h.dx, d_SDT.p);
    CUDA_OK(cudaGetLastError()); ctx->timing[6] += 1; }
  if (d_SP) { CUDA_OK(d_SDP.alloc(ctx, (size_t)nrows * ld));
    dense_source_kernel<<<gsd, 128, 0, ctx->stream>>>(d_SP, h.n_x, ix_start, nrows, d_jlo.p, d_wl.p, nkd1, ld, h.x0, h.dx, d_SDP.p);
    CUDA_OK(cudaGetLastError()); ctx->timing[6] += 1; }
  const int groups = (nell + PROJ_NL - 1) / PROJ_NL;
  int nsplit = (8 * ctx->num_sms + groups - 1) / groups;
  nsplit = std::max(1, std::min(nsplit, std::max(1, nkd1 / PROJ_NT)));
  CUDA_OK(d_part.alloc(ctx, (size_t)nell * nsplit * 3));
  ProjectParams pp;
  pp.Cf = d_Cf.p; pp.ells = d_ell.p; pp.nell = nell; pp.chi = d_chi.p; pp.nrows = nrows; pp.kscaled = d_ks.p; pp.wk = d_wk.p;
  pp.nkd1 = nkd1; pp.ld = ld; pp.SD_T = d_SDT.p; pp.SD_P = d_SDP.p; pp.nsplit = nsplit; pp.partial = d_part.p;
  const size_t smem = ((size_t)PROJ_NL * BESSEL_NC + nrows) * sizeof(double);
  CUDA_OK(cudaFuncSetAttribute(project_kernel<PROJ_NL, PROJ_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  project_kernel<PROJ_NL, PROJ_NT><<<dim3(groups, nsplit), PROJ_NT, smem, ctx->stream>>>(pp);
  CUDA_OK(cudaGetLastError());
  cl_finalize_kernel<<<(nell + 127) / 128, 128, 0, ctx->stream>>>(d_part.p, d_ell.p, nell, nsplit, d_ST ? d_cl : nullptr,
                                                                  (d_ST && d_SP) ? d_cl + nell : nullptr,
                                                                  d_SP ? d_cl + 2 * (size_t)nell : nullptr);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(ctx->ev[5], ctx->stream));
  ctx->timing[6] += 4;
  // the DevBufs above are freed on return: wait for the stream first
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return BOLT_OK;
}

}  // namespace

extern "C" {

int bolt_abi_version(void) { return BOLT_ABI_VERSION; }

int bolt_state_dim(int l_gamma, int l_nu, int l_mnu, int nq) { return 2 * (l_gamma + 1) + (l_nu + 1) + (l_mnu + 1) * nq + 5; }

int bolt_init(int device_ordinal, bolt_ctx** out) {
  if (!out) return BOLT_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return BOLT_ERR_CUDA;   // no CPU fallback, by design
  if (device_ordinal < 0 || device_ordinal >= ndev) return BOLT_ERR_ARG;
  bolt_ctx* ctx = new bolt_ctx();
  ctx->device = device_ordinal;
  if (cudaSetDevice(device_ordinal) != cudaSuccess) { delete ctx; return BOLT_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess) { delete ctx; return BOLT_ERR_CUDA; }
  ctx->num_sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return BOLT_ERR_CUDA; }
  if (cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return BOLT_ERR_CUDA; }
  if (cudaEventCreateWithFlags(&ctx->ev_tab, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return BOLT_ERR_CUDA; }
  for (auto& e : ctx->ev) if (cudaEventCreate(&e) != cudaSuccess) { delete ctx; return BOLT_ERR_CUDA; }
  // work-queue head + per-SM resident-CTA counters (k1_pipe)
  if (cudaMalloc(&ctx->d_counter, 257 * sizeof(int)) != cudaSuccess) { delete ctx; return BOLT_ERR_ALLOC; }
  if (init_constants(ctx) != BOLT_OK) { delete ctx; return BOLT_ERR_CUDA; }
  if (getenv("BOLT_DEBUG_STEPS")) { cudaMalloc(&ctx->d_dbg, sizeof(double) * 4 * DBG_CAP); cudaMemset(ctx->d_dbg, 0, sizeof(double) * 4 * DBG_CAP); }
  *out = ctx;
  return BOLT_OK;
}

int bolt_finalize(bolt_ctx* ctx) {
  if (!ctx) return BOLT_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->comm && g_nccl.CommDestroy) { g_nccl.CommDestroy(ctx->comm); ctx->comm = nullptr; }
  for (auto& e : ctx->ev) cudaEventDestroy(e);
  for (auto& b : ctx->pool) cudaFree(b.p);
  cudaFree(ctx->d_counter);
  cudaStreamDestroy(ctx->stream); cudaStreamDestroy(ctx->stream2); cudaEventDestroy(ctx->ev_tab);
  delete ctx;
  return BOLT_OK;
}

const char* bolt_last_error(const bolt_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context (no CUDA device?)"; }

int bolt_last_timing(const bolt_ctx* ctx, double* out8) {
  if (!ctx || !out8) return BOLT_ERR_ARG;
  for (int i = 0; i < 8; i++) out8[i] = ctx->timing[i];
  return BOLT_OK;
}

int bolt_cosmo_upload(bolt_ctx* ctx, const bolt_cosmo_desc* d, bolt_cosmo** out) {
  if (!ctx || !d || !out) return BOLT_ERR_ARG;
  *out = nullptr;
  if (d->abi_version != BOLT_ABI_VERSION) return fail(ctx, BOLT_ERR_ARG, "ABI version mismatch");
  if (d->nq < 1 || d->nq > MAX_NQ) return fail(ctx, BOLT_ERR_ARG, "nq out of range (1..29)");
  if (d->n_x < 4 || d->nd < 1) return fail(ctx, BOLT_ERR_ARG, "bad grid");
  CUDA_OK(cudaSetDevice(ctx->device));
  bolt_cosmo* c = new bolt_cosmo();
  DevCosmo& h = c->h;
  h.n_x = d->n_x; h.nq = d->nq; h.nd = d->nd; h.x0 = d->x0; h.dx = d->dx; h.inv_dx = 1.0 / d->dx;
  const int nd = d->nd, nc = d->n_x + 2;
  for (int i = 0; i < BOLT_NSCALARS; i++) h.s[i] = d->scalars[(size_t)i * nd];
  // value tables, contiguous [NTABLES][n_x+2]
  std::vector<double> tabs((size_t)BOLT_NTABLES * nc);
  for (int t = 0; t < BOLT_NTABLES; t++)
    for (int i = 0; i < nc; i++) tabs[(size_t)t * nc + i] = d->tables[((size_t)t * nc + i) * nd];
  if (cudaMalloc(&c->d_tables, tabs.size() * sizeof(double)) != cudaSuccess) { delete c; return fail(ctx, BOLT_ERR_ALLOC, "cudaMalloc tables"); }
  cudaMemcpy(c->d_tables, tabs.data(), tabs.size() * sizeof(double), cudaMemcpyHostToDevice);
  for (int t = 0; t < BOLT_NTABLES; t++) h.tab[t] = c->d_tables + (size_t)t * nc;
  // momentum grid constants: T_nu (perturbations.jl:164), q_i (:165-166, util.jl:24-27), f0, dlnf0dlnq (background.jl:21-30)
  const double N_nu = h.s[BOLT_S_N_nu], Om_r = h.s[BOLT_S_Omega_r], rho_crit = h.s[BOLT_S_rho_crit];
  const double Tnu = std::pow(N_nu / 3.0, 0.25) * std::pow(4.0 / 11.0, 1.0 / 3.0) * std::pow(15.0 / (M_PI * M_PI) * rho_crit * Om_r, 0.25);
  const double lqmi = std::log10(Tnu / 30.0), lqma = std::log10(Tnu * 30.0);
  for (int i = 0; i < h.nq; i++) {
    const double lq = lqmi + (lqma - lqmi) / 2.0 * (d->quad_pts[i] + 1.0);
    const double q = std::pow(10.0, lq);
    const double dxdq = (2.0 / (lqma - lqmi)) / (q * std::log(10.0));
    const double f0 = 2.0 / std::pow(2.0 * M_PI, 3) / (std::exp(q / Tnu) + 1.0);
    h.q[i] = q;
    h.wq[i] = 4.0 * M_PI * q * q * (f0 / dxdq * d->quad_wts[i]);
    h.df0[i] = -q / Tnu / (1.0 + std::exp(-q / Tnu));
  }
  h.Omega_nu = 7.0 * (2.0 / 3.0) * N_nu / 8.0 * std::pow(4.0 / 11.0, 4.0 / 3.0) * Om_r;
  {  // eta(x_grid[end]) with the same spline arithmetic as the device
    const double x_end = h.x0 + h.dx * (h.n_x - 1);
    const double* ce = tabs.data() + (size_t)BOLT_T_eta * nc;
    double t = (x_end - h.x0) / h.dx; int i = (int)std::floor(t); i = std::max(0, std::min(i, h.n_x - 2));
    double dd = t - i, e = 1.0 - dd;
    h.eta_end = ce[i] * (e * e * e / 6.0) + ce[i + 1] * (2.0 / 3.0 - dd * dd + dd * dd * dd / 2.0) +
                ce[i + 2] * (2.0 / 3.0 - e * e + e * e * e / 2.0) + ce[i + 3] * (dd * dd * dd / 6.0);
  }
  // forward-mode partials (nd > 1): partial tables component-major, and the log-derivatives of the momentum-grid constants
  // (T_nu ~ (N_nu rho_crit Om_r)^(1/4); q_i ~ T_nu; wq_i ~ q_i^3; dlnf0dlnq(q_i) depends on q_i/T_nu only: no partials)
  h.np = nd - 1;
  if (h.np > MAX_NP) { cudaFree(c->d_tables); delete c; return fail(ctx, BOLT_ERR_UNSUPPORTED, "more than 8 partials"); }
  for (int t = 0; t < BOLT_NTABLES; t++) h.dtab[t] = nullptr;
  if (h.np > 0) {
    const int np = h.np;
    std::vector<double> dt((size_t)BOLT_NTABLES * np * nc);
    for (int t = 0; t < BOLT_NTABLES; t++)
      for (int j = 0; j < np; j++)
        for (int i = 0; i < nc; i++) dt[((size_t)t * np + j) * nc + i] = d->tables[((size_t)t * nc + i) * nd + 1 + j];
    if (cudaMalloc(&c->d_dtables, dt.size() * sizeof(double)) != cudaSuccess) { cudaFree(c->d_tables); delete c; return fail(ctx, BOLT_ERR_ALLOC, "cudaMalloc partial tables"); }
    cudaMemcpy(c->d_dtables, dt.data(), dt.size() * sizeof(double), cudaMemcpyHostToDevice);
    for (int t = 0; t < BOLT_NTABLES; t++) h.dtab[t] = c->d_dtables + (size_t)t * np * nc;
    const double* ce = dt.data() + (size_t)BOLT_T_eta * np * nc;
    for (int j = 0; j < np; j++) {
      for (int i = 0; i < BOLT_NSCALARS; i++) h.ds[i][j] = d->scalars[(size_t)i * nd + 1 + j];
      const double dlnT = 0.25 * (h.ds[BOLT_S_N_nu][j] / N_nu + h.ds[BOLT_S_rho_crit][j] / rho_crit + h.ds[BOLT_S_Omega_r][j] / Om_r);
      for (int i = 0; i < h.nq; i++) { h.dq[i][j] = h.q[i] * dlnT; h.dwq[i][j] = 3.0 * h.wq[i] * dlnT; }
      h.dOmega_nu[j] = h.Omega_nu * (h.ds[BOLT_S_N_nu][j] / N_nu + h.ds[BOLT_S_Omega_r][j] / Om_r);
      h.deta_end[j] = ce[(size_t)j * nc + h.n_x];     // eta spline at the last knot: c[n] = y[n-1] (Line(OnGrid()) boundary)
    }
  }
  if (cudaMalloc(&c->d, sizeof(DevCosmo)) != cudaSuccess) { cudaFree(c->d_tables); delete c; return fail(ctx, BOLT_ERR_ALLOC, "cudaMalloc cosmo"); }
  cudaMemcpy(c->d, &h, sizeof(DevCosmo), cudaMemcpyHostToDevice);
  eta_end_kernel<<<1, 1, 0, ctx->stream>>>(c->d);
  cudaStreamSynchronize(ctx->stream);
  // compact K1 view: keep only the partials with a non-zero table or (K1-relevant) scalar partial
  c->np_k1 = 0;
  if (h.np > 0) {
    const int np = h.np;
    for (int j = 0; j < np; j++) {
      bool active = false;
      for (int i = 0; i < BOLT_NSCALARS && !active; i++) if (i != BOLT_S_A && i != BOLT_S_n && h.ds[i][j] != 0.0) active = true;
      for (size_t i = 0; i < (size_t)BOLT_NTABLES * nc && !active; i++) if (d->tables[i * nd + 1 + j] != 0.0) active = true;
      if (active) c->map_k1[c->np_k1++] = 1 + j;
    }
    static const int supported[] = {0, 1, 2, 3, 4, 6};
    bool ok = false; for (int v : supported) ok |= (v == c->np_k1);
    if (c->np_k1 < np && ok) {
      DevCosmo k1 = h;
      k1.np = c->np_k1; k1.nd = 1 + c->np_k1;
      const int na = c->np_k1;
      if (na > 0) {
        std::vector<double> dt((size_t)BOLT_NTABLES * na * nc);
        for (int t = 0; t < BOLT_NTABLES; t++)
          for (int a = 0; a < na; a++)
            for (int i = 0; i < nc; i++) dt[((size_t)t * na + a) * nc + i] = d->tables[((size_t)t * nc + i) * nd + c->map_k1[a]];
        if (cudaMalloc(&c->d_dtables_k1, dt.size() * sizeof(double)) != cudaSuccess ||
            cudaMemcpy(c->d_dtables_k1, dt.data(), dt.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
          bolt_cosmo_free(ctx, c); return fail(ctx, BOLT_ERR_ALLOC, "cudaMalloc/cudaMemcpy compact partial tables");
        }
        for (int t = 0; t < BOLT_NTABLES; t++) k1.dtab[t] = c->d_dtables_k1 + (size_t)t * na * nc;
        for (int a = 0; a < na; a++) {
          const int j = c->map_k1[a] - 1;
          for (int i = 0; i < BOLT_NSCALARS; i++) k1.ds[i][a] = h.ds[i][j];
          for (int i = 0; i < h.nq; i++) { k1.dq[i][a] = h.dq[i][j]; k1.dwq[i][a] = h.dwq[i][j]; }
          k1.dOmega_nu[a] = h.dOmega_nu[j]; k1.deta_end[a] = h.deta_end[j];
        }
      }
      if (cudaMalloc(&c->d_k1, sizeof(DevCosmo)) != cudaSuccess ||
          cudaMemcpy(c->d_k1, &k1, sizeof(DevCosmo), cudaMemcpyHostToDevice) != cudaSuccess) {
        bolt_cosmo_free(ctx, c); return fail(ctx, BOLT_ERR_ALLOC, "cudaMalloc/cudaMemcpy compact cosmology");
      }
      c->h_k1 = k1;
      eta_end_kernel<<<1, 1, 0, ctx->stream>>>(c->d_k1);
      if (cudaStreamSynchronize(ctx->stream) != cudaSuccess ||
          cudaMalloc(&c->d_list_k1, sizeof(DevCosmo*)) != cudaSuccess ||
          cudaMemcpy(c->d_list_k1, &c->d_k1, sizeof(DevCosmo*), cudaMemcpyHostToDevice) != cudaSuccess) {
        bolt_cosmo_free(ctx, c); return fail(ctx, BOLT_ERR_CUDA, "compact cosmology upload");
      }
    } else {
      c->np_k1 = np;
      for (int j = 0; j < np; j++) c->map_k1[j] = 1 + j;
      c->h_k1 = h;
    }
    if (c->np_k1 > 0) {
      const int na = c->np_k1;
      std::vector<DevCosmo> views(na, c->h_k1);
      for (int a = 0; a < na; a++) {
        DevCosmo& v = views[a];
        v.np = 1; v.nd = 2;
        for (int t = 0; t < BOLT_NTABLES; t++) v.dtab[t] = c->h_k1.dtab[t] + (size_t)a * nc;
        for (int i = 0; i < BOLT_NSCALARS; i++) v.ds[i][0] = c->h_k1.ds[i][a];
        for (int i = 0; i < h.nq; i++) { v.dq[i][0] = c->h_k1.dq[i][a]; v.dwq[i][0] = c->h_k1.dwq[i][a]; }
        v.dOmega_nu[0] = c->h_k1.dOmega_nu[a]; v.deta_end[0] = c->h_k1.deta_end[a];
      }
      std::vector<const DevCosmo*> vl(na);
      if (cudaMalloc(&c->d_views, na * sizeof(DevCosmo)) != cudaSuccess ||
          cudaMemcpy(c->d_views, views.data(), na * sizeof(DevCosmo), cudaMemcpyHostToDevice) != cudaSuccess) {
        bolt_cosmo_free(ctx, c); return fail(ctx, BOLT_ERR_ALLOC, "cudaMalloc/cudaMemcpy single-partial views");
      }
      for (int a = 0; a < na; a++) vl[a] = c->d_views + a;
      if (cudaMalloc(&c->d_view_list, na * sizeof(DevCosmo*)) != cudaSuccess ||
          cudaMemcpy(c->d_view_list, vl.data(), na * sizeof(DevCosmo*), cudaMemcpyHostToDevice) != cudaSuccess) {
        bolt_cosmo_free(ctx, c); return fail(ctx, BOLT_ERR_ALLOC, "cudaMalloc/cudaMemcpy view list");
      }
    }
  }
  if (cudaMalloc(&c->d_list, sizeof(DevCosmo*)) != cudaSuccess) { cudaFree(c->d); cudaFree(c->d_tables); delete c; return fail(ctx, BOLT_ERR_ALLOC, "cudaMalloc list"); }
  cudaMemcpy(c->d_list, &c->d, sizeof(DevCosmo*), cudaMemcpyHostToDevice);
  *out = c;
  return BOLT_OK;
}

int bolt_cosmo_free(bolt_ctx* ctx, bolt_cosmo* c) {
  if (!c) return BOLT_OK;
  if (ctx) cudaSetDevice(ctx->device);
  cudaFree(c->d); cudaFree(c->d_tables); cudaFree(c->d_dtables); cudaFree(c->d_list);
  cudaFree(c->d_k1); cudaFree(c->d_list_k1); cudaFree(c->d_dtables_k1); cudaFree(c->d_views); cudaFree(c->d_view_list);
  delete c;
  return BOLT_OK;
}

int bolt_solve(bolt_ctx* ctx, const bolt_cosmo* c, const double* k, int nk, const bolt_opts* o,
               double* S_T, double* S_P, double* u_hist, double* u_final,
               int32_t* status, int64_t* nsteps, int64_t* nreject) {
  if (!ctx) return BOLT_ERR_ARG;
  if (!c || !k || nk < 1) return fail(ctx, BOLT_ERR_ARG, "bad arguments");
  int rc = check_opts(ctx, c, o); if (rc) return rc;
  CUDA_OK(cudaSetDevice(ctx->device));
  reset_timing(ctx);
  CUDA_OK(cudaEventRecord(ctx->ev[6], ctx->stream));
  const int n = bolt_state_dim(o->l_gamma, o->l_nu, o->l_mnu, c->h.nq), n_x = c->h.n_x * c->h.nd;   // n_x: doubles per source column
  if (u_hist && c->h.nd > 1) return fail(ctx, BOLT_ERR_UNSUPPORTED, "u_hist is value-only (nd = 1)");
  DevBuf<double> d_k, d_ST, d_SP, d_hist, d_final; DevBuf<int> d_order, d_status; DevBuf<long long> d_ns, d_nr;
  rc = upload_k_sorted(ctx, k, nk, d_k, d_order); if (rc) return rc;
  if (S_T) { CUDA_OK(d_ST.alloc(ctx, (size_t)nk * n_x)); CUDA_OK(cudaMemsetAsync(d_ST.p, 0, d_ST.n * 8, ctx->stream)); }
  if (S_P) { CUDA_OK(d_SP.alloc(ctx, (size_t)nk * n_x)); CUDA_OK(cudaMemsetAsync(d_SP.p, 0, d_SP.n * 8, ctx->stream)); }
  if (u_hist) { CUDA_OK(d_hist.alloc(ctx, (size_t)nk * n_x * n)); CUDA_OK(cudaMemsetAsync(d_hist.p, 0, d_hist.n * 8, ctx->stream)); }
  if (u_final) { CUDA_OK(d_final.alloc(ctx, (size_t)nk * n * c->h.nd)); CUDA_OK(cudaMemsetAsync(d_final.p, 0, d_final.n * 8, ctx->stream)); }
  CUDA_OK(d_status.alloc(ctx, nk)); CUDA_OK(d_ns.alloc(ctx, nk)); CUDA_OK(d_nr.alloc(ctx, nk));
  rc = launch_hierarchy(ctx, c, d_k.p, d_order.p, nk, o, d_ST.p, d_SP.p, d_hist.p, d_final.p, d_status.p, d_ns.p, d_nr.p);
  if (rc) return rc;
  if (S_T) CUDA_OK(cudaMemcpyAsync(S_T, d_ST.p, d_ST.n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (S_P) CUDA_OK(cudaMemcpyAsync(S_P, d_SP.p, d_SP.n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (u_hist) CUDA_OK(cudaMemcpyAsync(u_hist, d_hist.p, d_hist.n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (u_final) CUDA_OK(cudaMemcpyAsync(u_final, d_final.p, d_final.n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (status) CUDA_OK(cudaMemcpyAsync(status, d_status.p, nk * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (nsteps) CUDA_OK(cudaMemcpyAsync(nsteps, d_ns.p, nk * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  if (nreject) CUDA_OK(cudaMemcpyAsync(nreject, d_nr.p, nk * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaEventRecord(ctx->ev[7], ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  if (ctx->d_dbg) {   // development aid: per-step log of mode 0
    std::vector<double> h((size_t)4 * DBG_CAP);
    cudaMemcpy(h.data(), ctx->d_dbg, h.size() * 8, cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(getenv("BOLT_DEBUG_STEPS"), "w")) {
      for (int i = 0; i < DBG_CAP && h[4 * i + 1] != 0.0; i++) fprintf(f, "x=%.10g dt=%.4g EEst=%.17g acc=%d\n", h[4 * i], h[4 * i + 1], h[4 * i + 2], (int)h[4 * i + 3]);
      fclose(f);
    }
  }
  return collect_timing(ctx);
}

int bolt_project(bolt_ctx* ctx, const bolt_cosmo* c, const double* S_T, const double* S_P, const double* k, int nk,
                 const int32_t* ell, int nell, double kd_min, double kd_max, int n_kd, int ix_start,
                 double* cl_tt, double* cl_te, double* cl_ee) {
  if (!ctx) return BOLT_ERR_ARG;
  if (!c || !k || nk < 2 || !ell || nell < 1 || (!S_T && !S_P)) return fail(ctx, BOLT_ERR_ARG, "bad arguments");
  CUDA_OK(cudaSetDevice(ctx->device));
  reset_timing(ctx);
  CUDA_OK(cudaEventRecord(ctx->ev[6], ctx->stream));
  const int nd = c->h.nd, n_x = c->h.n_x * nd;     // doubles per source column
  DevBuf<double> d_k, d_ST, d_SP, d_cl;
  CUDA_OK(d_k.alloc(ctx, nk));
  CUDA_OK(cudaMemcpyAsync(d_k.p, k, nk * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (S_T) { CUDA_OK(d_ST.alloc(ctx, (size_t)nk * n_x)); CUDA_OK(cudaMemcpyAsync(d_ST.p, S_T, d_ST.n * 8, cudaMemcpyHostToDevice, ctx->stream)); }
  if (S_P) { CUDA_OK(d_SP.alloc(ctx, (size_t)nk * n_x)); CUDA_OK(cudaMemcpyAsync(d_SP.p, S_P, d_SP.n * 8, cudaMemcpyHostToDevice, ctx->stream)); }
  const size_t ncl = (size_t)nell * nd;
  CUDA_OK(d_cl.alloc(ctx, 3 * ncl));
  BesselTabs bt;
  int rc = bessel_prepare(ctx, c, ell, nell, kd_max, ctx->stream, bt); if (rc) return rc;
  rc = project_device(ctx, c, d_ST.p, d_SP.p, d_k.p, nk, bt, nell, kd_min, kd_max, n_kd, ix_start, d_cl.p);
  if (rc) return rc;
  if (cl_tt && S_T) CUDA_OK(cudaMemcpyAsync(cl_tt, d_cl.p, ncl * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (cl_te && S_T && S_P) CUDA_OK(cudaMemcpyAsync(cl_te, d_cl.p + ncl, ncl * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (cl_ee && S_P) CUDA_OK(cudaMemcpyAsync(cl_ee, d_cl.p + 2 * ncl, ncl * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaEventRecord(ctx->ev[7], ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return collect_timing(ctx);
}

int bolt_spectra(bolt_ctx* ctx, const bolt_cosmo* c, const double* k, int nk, const bolt_opts* o,
                 const int32_t* ell, int nell, double kd_min, double kd_max, int n_kd, int ix_start,
                 double* cl_tt, double* cl_te, double* cl_ee, int32_t* status, int64_t* nsteps, int64_t* nreject) {
  if (!ctx) return BOLT_ERR_ARG;
  if (!c || !k || nk < 2 || !ell || nell < 1) return fail(ctx, BOLT_ERR_ARG, "bad arguments");
  int rc = check_opts(ctx, c, o); if (rc) return rc;
  rc = check_projection_args(ctx, c, ell, nell, n_kd, ix_start); if (rc) return rc;     // before any work is enqueued
  CUDA_OK(cudaSetDevice(ctx->device));
  reset_timing(ctx);
  CUDA_OK(cudaEventRecord(ctx->ev[6], ctx->stream));
  const int nd = c->h.nd, n_x = c->h.n_x * nd;
  const size_t ncl = (size_t)nell * nd;
  DevBuf<double> d_k, d_ST, d_SP, d_cl; DevBuf<int> d_order, d_status; DevBuf<long long> d_ns, d_nr;
  rc = upload_k_sorted(ctx, k, nk, d_k, d_order); if (rc) return rc;
  CUDA_OK(d_ST.alloc(ctx, (size_t)nk * n_x)); CUDA_OK(d_SP.alloc(ctx, (size_t)nk * n_x));
  // Always cleared: a mode that stops early (status 1..3) leaves its remaining rows unwritten, and the pool hands out
  // recycled memory -- K2 must never read another cosmology's sources there.  Partials the hierarchy does not carry
  // (A, n) stay exactly zero for the same reason.
  CUDA_OK(cudaMemsetAsync(d_ST.p, 0, d_ST.n * 8, ctx->stream)); CUDA_OK(cudaMemsetAsync(d_SP.p, 0, d_SP.n * 8, ctx->stream));
  CUDA_OK(d_status.alloc(ctx, nk)); CUDA_OK(d_ns.alloc(ctx, nk)); CUDA_OK(d_nr.alloc(ctx, nk)); CUDA_OK(d_cl.alloc(ctx, 3 * ncl));
  bolt_opts oo = *o;
  oo.ix_first = std::max(oo.ix_first, ix_start);    // the LOS sum only reads rows >= ix_start (spectra.jl:86)
  BesselTabs bt;
  rc = launch_hierarchy(ctx, c, d_k.p, d_order.p, nk, &oo, d_ST.p, d_SP.p, nullptr, nullptr, d_status.p, d_ns.p, d_nr.p);
  if (rc) return bail(ctx, rc);
  // The j_l tables do not depend on K1.  Enqueued AFTER it on a second stream, their blocks are scheduled as K1's persistent
  // warps retire, i.e. they fill the low-occupancy tail of the hierarchy solve instead of delaying its start.
  rc = bessel_prepare(ctx, c, ell, nell, kd_max, ctx->stream2, bt); if (rc) return bail(ctx, rc);
  rc = project_device(ctx, c, d_ST.p, d_SP.p, d_k.p, nk, bt, nell, kd_min, kd_max, n_kd, ix_start, d_cl.p);
  if (rc) return bail(ctx, rc);
  if (cl_tt) CUDA_OK(cudaMemcpyAsync(cl_tt, d_cl.p, ncl * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (cl_te) CUDA_OK(cudaMemcpyAsync(cl_te, d_cl.p + ncl, ncl * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (cl_ee) CUDA_OK(cudaMemcpyAsync(cl_ee, d_cl.p + 2 * ncl, ncl * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (status) CUDA_OK(cudaMemcpyAsync(status, d_status.p, nk * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (nsteps) CUDA_OK(cudaMemcpyAsync(nsteps, d_ns.p, nk * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  if (nreject) CUDA_OK(cudaMemcpyAsync(nreject, d_nr.p, nk * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaEventRecord(ctx->ev[7], ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return collect_timing(ctx);
}

int bolt_spectra_batch(bolt_ctx* ctx, const bolt_cosmo* const* cosmos, int ncos, const double* k, int nk, const bolt_opts* o,
                       const int32_t* ell, int nell, const double* kd_min, const double* kd_max, int n_kd, int ix_start,
                       double* cl_tt, double* cl_te, double* cl_ee, int32_t* status, int64_t* nsteps) {
  if (!ctx) return BOLT_ERR_ARG;
  if (!cosmos || ncos < 1 || ncos > BOLT_MAX_BATCH || !k || nk < 2 || !ell || nell < 1 || !kd_min || !kd_max)
    return fail(ctx, BOLT_ERR_ARG, "bad arguments (1 <= ncos <= BOLT_MAX_BATCH)");
  for (int i = 0; i < ncos; i++) {
    const bolt_cosmo* c = cosmos[i];
    if (!c) return fail(ctx, BOLT_ERR_ARG, "null cosmology in batch");
    if (c->h.nd != 1) return fail(ctx, BOLT_ERR_UNSUPPORTED, "bolt_spectra_batch is value-only (nd = 1): call bolt_spectra per cosmology for partials");
    if (c->h.n_x != cosmos[0]->h.n_x || c->h.nq != cosmos[0]->h.nq || c->h.x0 != cosmos[0]->h.x0 || c->h.dx != cosmos[0]->h.dx)
      return fail(ctx, BOLT_ERR_ARG, "cosmologies of a batch must share x_grid and nq");
    int rc = check_opts(ctx, c, o); if (rc) return rc;
    rc = check_projection_args(ctx, c, ell, nell, n_kd, ix_start); if (rc) return rc;
  }
  CUDA_OK(cudaSetDevice(ctx->device));
  reset_timing(ctx);
  CUDA_OK(cudaEventRecord(ctx->ev[6], ctx->stream));
  const bolt_cosmo* c0 = cosmos[0];
  const int n_x = c0->h.n_x, nkt = ncos * nk;
  DevBuf<double> d_k, d_ST, d_SP, d_cl; DevBuf<int> d_order, d_status; DevBuf<long long> d_ns; DevBuf<const DevCosmo*> d_list;
  int rc = upload_k_sorted(ctx, k, nkt, d_k, d_order); if (rc) return rc;      // ONE queue over all cosmologies, longest solves first
  {
    std::vector<const DevCosmo*> list(ncos);
    for (int i = 0; i < ncos; i++) list[i] = cosmos[i]->d;
    CUDA_OK(d_list.alloc(ctx, ncos));
    CUDA_OK(cudaMemcpyAsync(d_list.p, list.data(), ncos * sizeof(const DevCosmo*), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_OK(cudaStreamSynchronize(ctx->stream));   // `list` is a local
  }
  CUDA_OK(d_ST.alloc(ctx, (size_t)nkt * n_x)); CUDA_OK(d_SP.alloc(ctx, (size_t)nkt * n_x));
  CUDA_OK(cudaMemsetAsync(d_ST.p, 0, d_ST.n * 8, ctx->stream)); CUDA_OK(cudaMemsetAsync(d_SP.p, 0, d_SP.n * 8, ctx->stream));   // see bolt_spectra
  CUDA_OK(d_status.alloc(ctx, nkt)); CUDA_OK(d_ns.alloc(ctx, nkt)); CUDA_OK(d_cl.alloc(ctx, (size_t)3 * nell * ncos));
  bolt_opts oo = *o;
  oo.ix_first = std::max(oo.ix_first, ix_start);
  // K1: one persistent launch over ncos x nk modes (mode ik belongs to cosmology ik / nk): the tail of the launch is paid once
  rc = launch_hierarchy(ctx, d_list.p, c0->h.nq, 0, 1, nullptr, nk, d_k.p, d_order.p, nkt, &oo, d_ST.p, d_SP.p, nullptr, nullptr,
                        d_status.p, d_ns.p, nullptr);
  if (rc) return bail(ctx, rc);
  // K2 per cosmology (the j_l table range k_max eta_0 differs); tables on the second stream as in bolt_spectra
  // All tables are allocated and enqueued BEFORE the first projection: project_device returns its temporaries to the pool
  // while its kernels are still in flight on the main stream, and a table built on the second stream must never land in one.
  std::vector<std::unique_ptr<BesselTabs>> tabs;
  for (int i = 0; i < ncos; i++) {
    tabs.emplace_back(new BesselTabs());
    rc = bessel_prepare(ctx, cosmos[i], ell, nell, kd_max[i], ctx->stream2, *tabs.back()); if (rc) return bail(ctx, rc);
  }
  for (int i = 0; i < ncos; i++) {
    rc = project_device(ctx, cosmos[i], d_ST.p + (size_t)i * nk * n_x, d_SP.p + (size_t)i * nk * n_x, d_k.p + (size_t)i * nk, nk, *tabs[i],
                        nell, kd_min[i], kd_max[i], n_kd, ix_start, d_cl.p + (size_t)i * 3 * nell);
    if (rc) return bail(ctx, rc);
    double* base = d_cl.p + (size_t)i * 3 * nell;
    if (cl_tt) CUDA_OK(cudaMemcpyAsync(cl_tt + (size_t)i * nell, base, nell * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (cl_te) CUDA_OK(cudaMemcpyAsync(cl_te + (size_t)i * nell, base + nell, nell * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (cl_ee) CUDA_OK(cudaMemcpyAsync(cl_ee + (size_t)i * nell, base + 2 * nell, nell * 8, cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (status) CUDA_OK(cudaMemcpyAsync(status, d_status.p, nkt * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (nsteps) CUDA_OK(cudaMemcpyAsync(nsteps, d_ns.p, nkt * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaEventRecord(ctx->ev[7], ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));     // the tables in `tabs` are released after the last projection has run
  return collect_timing(ctx);
}

// P(k) from the final state of every mode (the reference's plin epilogue, src/spectra.jl:174-198)
static int launch_plin_epilogue(bolt_ctx* ctx, const bolt_cosmo* c, const double* d_k, int nk, const double* d_final, const bolt_opts* o, double* d_pk) {
  const int gb = (nk + 127) / 128;
  switch (c->h.np) {
    case 0: plin_kernel<<<gb, 128, 0, ctx->stream>>>(c->d, d_k, nk, d_final, o->l_gamma, o->l_nu, o->l_mnu, d_pk); break;
    case 1: plin_kernel_d<1><<<gb, 128, 0, ctx->stream>>>(c->d, d_k, nk, d_final, o->l_gamma, o->l_nu, o->l_mnu, d_pk); break;
    case 2: plin_kernel_d<2><<<gb, 128, 0, ctx->stream>>>(c->d, d_k, nk, d_final, o->l_gamma, o->l_nu, o->l_mnu, d_pk); break;
    case 3: plin_kernel_d<3><<<gb, 128, 0, ctx->stream>>>(c->d, d_k, nk, d_final, o->l_gamma, o->l_nu, o->l_mnu, d_pk); break;
    case 4: plin_kernel_d<4><<<gb, 128, 0, ctx->stream>>>(c->d, d_k, nk, d_final, o->l_gamma, o->l_nu, o->l_mnu, d_pk); break;
    case 6: plin_kernel_d<6><<<gb, 128, 0, ctx->stream>>>(c->d, d_k, nk, d_final, o->l_gamma, o->l_nu, o->l_mnu, d_pk); break;
    default: return fail(ctx, BOLT_ERR_UNSUPPORTED, "this build carries 1, 2, 3, 4 or 6 partials per call");
  }
  CUDA_OK(cudaGetLastError());
  return BOLT_OK;
}

int bolt_plin(bolt_ctx* ctx, const bolt_cosmo* c, const double* k, int nk, const bolt_opts* o, double* pk, int32_t* status,
              int64_t* nsteps) {
  if (!ctx) return BOLT_ERR_ARG;
  if (!c || !k || nk < 1 || !pk) return fail(ctx, BOLT_ERR_ARG, "bad arguments");
  int rc = check_opts(ctx, c, o); if (rc) return rc;
  CUDA_OK(cudaSetDevice(ctx->device));
  reset_timing(ctx);
  CUDA_OK(cudaEventRecord(ctx->ev[6], ctx->stream));
  const int nd = c->h.nd;
  const int n = bolt_state_dim(o->l_gamma, o->l_nu, o->l_mnu, c->h.nq);
  DevBuf<double> d_k, d_final, d_pk; DevBuf<int> d_order, d_status; DevBuf<long long> d_ns;
  rc = upload_k_sorted(ctx, k, nk, d_k, d_order); if (rc) return rc;
  CUDA_OK(d_final.alloc(ctx, (size_t)nk * n * nd)); CUDA_OK(d_pk.alloc(ctx, (size_t)nk * nd));
  CUDA_OK(cudaMemsetAsync(d_final.p, 0, d_final.n * 8, ctx->stream));
  CUDA_OK(d_status.alloc(ctx, nk)); CUDA_OK(d_ns.alloc(ctx, nk));
  bolt_opts oo = *o; oo.ix_first = c->h.n_x;   // plin only needs perturb(0): no source sampling
  rc = launch_hierarchy(ctx, c, d_k.p, d_order.p, nk, &oo, nullptr, nullptr, nullptr, d_final.p, d_status.p, d_ns.p, nullptr);
  if (rc) return rc;
  rc = launch_plin_epilogue(ctx, c, d_k.p, nk, d_final.p, o, d_pk.p); if (rc) return rc;
  CUDA_OK(cudaMemcpyAsync(pk, d_pk.p, (size_t)nk * nd * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (status) CUDA_OK(cudaMemcpyAsync(status, d_status.p, nk * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (nsteps) CUDA_OK(cudaMemcpyAsync(nsteps, d_ns.p, nk * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaEventRecord(ctx->ev[7], ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return collect_timing(ctx);
}

int bolt_solve_device(bolt_ctx* ctx, const bolt_cosmo* c, const double* d_k, int nk, const bolt_opts* o,
                      double* d_S_T, double* d_S_P, double* d_u_final, int32_t* d_status, int64_t* d_nsteps, int64_t* d_nreject) {
  if (!ctx) return BOLT_ERR_ARG;
  if (!c || !d_k || nk < 1) return fail(ctx, BOLT_ERR_ARG, "bad arguments");
  int rc = check_opts(ctx, c, o); if (rc) return rc;
  CUDA_OK(cudaSetDevice(ctx->device));
  reset_timing(ctx);
  CUDA_OK(cudaEventRecord(ctx->ev[6], ctx->stream));
  std::vector<double> hk(nk);
  CUDA_OK(cudaMemcpyAsync(hk.data(), d_k, nk * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  DevBuf<double> d_k2; DevBuf<int> d_order;
  rc = upload_k_sorted(ctx, hk.data(), nk, d_k2, d_order); if (rc) return rc;
  static_assert(sizeof(long long) == sizeof(int64_t), "int64");
  {   // the kernels write only the rows >= ix_first and the partial slots they carry: everything else must read as zero
    const size_t ncol = (size_t)nk * c->h.n_x * c->h.nd;
    if (d_S_T) CUDA_OK(cudaMemsetAsync(d_S_T, 0, ncol * 8, ctx->stream));
    if (d_S_P) CUDA_OK(cudaMemsetAsync(d_S_P, 0, ncol * 8, ctx->stream));
    if (d_u_final) CUDA_OK(cudaMemsetAsync(d_u_final, 0, (size_t)nk * bolt_state_dim(o->l_gamma, o->l_nu, o->l_mnu, c->h.nq) * c->h.nd * 8, ctx->stream));
  }
  rc = launch_hierarchy(ctx, c, d_k, d_order.p, nk, o, d_S_T, d_S_P, nullptr, d_u_final, d_status, (long long*)d_nsteps, (long long*)d_nreject);
  if (rc) return rc;
  CUDA_OK(cudaEventRecord(ctx->ev[7], ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return collect_timing(ctx);
}

int bolt_project_device(bolt_ctx* ctx, const bolt_cosmo* c, const double* d_S_T, const double* d_S_P, const double* d_k, int nk,
                        const int32_t* ell, int nell, double kd_min, double kd_max, int n_kd, int ix_start, double* d_cl) {
  if (!ctx) return BOLT_ERR_ARG;
  if (!c || !d_k || nk < 2 || !ell || nell < 1 || (!d_S_T && !d_S_P) || !d_cl) return fail(ctx, BOLT_ERR_ARG, "bad arguments");
  CUDA_OK(cudaSetDevice(ctx->device));
  reset_timing(ctx);
  CUDA_OK(cudaEventRecord(ctx->ev[6], ctx->stream));
  BesselTabs bt;
  int rc = bessel_prepare(ctx, c, ell, nell, kd_max, ctx->stream, bt); if (rc) return rc;
  rc = project_device(ctx, c, d_S_T, d_S_P, d_k, nk, bt, nell, kd_min, kd_max, n_kd, ix_start, d_cl);
  if (rc) return rc;
  CUDA_OK(cudaEventRecord(ctx->ev[7], ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return collect_timing(ctx);
}

// ---------------------------------------------------------------------------------------------------
// Multi-GPU: the k-modes of ONE cosmology sharded over the ranks of a communicator (the reference's fan-out over
// k is the threaded map of src/spectra.jl:10, its fan-out over l the qmap of :149).  One process per GPU, one context per
// process; the host (MPI.jl / torch.distributed / a file) only carries the 128-byte unique id from rank 0 to the others.
// NCCL is bound at run time (dlopen) so that single-GPU users do not need it; a process that already loaded an NCCL (e.g.
// torch's) shares that copy.
// ---------------------------------------------------------------------------------------------------
namespace {
int nccl_load(bolt_ctx* ctx) {
  std::lock_guard<std::mutex> lock(g_nccl_mutex);
  if (g_nccl.lib) return BOLT_OK;
  const char* names[] = {getenv("BOLT_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* nm : names) if (nm && (h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
  if (!h) return fail(ctx, BOLT_ERR_UNSUPPORTED, "NCCL not found (libnccl.so.2); set BOLT_NCCL_LIB");
#define BOLT_NCCL_SYM(field, name) *(void**)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) return fail(ctx, BOLT_ERR_UNSUPPORTED, "NCCL symbol missing: " name);
  BOLT_NCCL_SYM(GetUniqueId, "ncclGetUniqueId") BOLT_NCCL_SYM(CommInitRank, "ncclCommInitRank") BOLT_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  BOLT_NCCL_SYM(AllGather, "ncclAllGather") BOLT_NCCL_SYM(AllReduce, "ncclAllReduce") BOLT_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef BOLT_NCCL_SYM
  g_nccl.lib = h;
  return BOLT_OK;
}
int nccl_fail(bolt_ctx* ctx, int rc, const char* what) {
  return fail(ctx, BOLT_ERR_CUDA, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error"));
}
#define NCCL_OK(call, what) do { int rc_ = (call); if (rc_ != 0) return nccl_fail(ctx, rc_, what); } while (0)

// gathered [R][2][per][row] -> S_T, S_P [nk][row] in the caller's k order: shard r, slot j holds mode order[r + j*R]
__global__ void unshard_sources_kernel(const double* __restrict__ g, const int* __restrict__ order, int nk, int R, int per, int row,
                                       double* __restrict__ S_T, double* __restrict__ S_P) {
  const int pos = blockIdx.x;                 // position in the descending-k order
  const int r = pos % R, j = pos / R;
  const int ik = order[pos];
  const double* src = g + ((size_t)r * 2 * per + j) * row;
  for (int i = threadIdx.x; i < row; i += blockDim.x) {
    S_T[(size_t)ik * row + i] = src[i];
    S_P[(size_t)ik * row + i] = src[(size_t)per * row + i];
  }
}
// local [3][nl][nd] (multipoles rank, rank+R, ...) -> full [3][nell][nd], zero elsewhere (the all-reduce then concatenates)
__global__ void scatter_cl_kernel(const double* __restrict__ loc, int nl, int nell, int nd, int rank, int R, double* __restrict__ full) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * nl * nd) return;
  const int d = i % nd, il = (i / nd) % nl, s = i / (nd * nl);
  full[((size_t)s * nell + (rank + (size_t)il * R)) * nd + d] = loc[i];
}
}  // namespace

int bolt_shard_plan(const double* k, int nk, int rank, int nranks, int32_t* idx, int32_t* n_local) {
  if (!k || nk < 1 || nranks < 1 || rank < 0 || rank >= nranks || !idx || !n_local) return BOLT_ERR_ARG;
  std::vector<int> order(nk);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return k[a] > k[b]; });
  int n = 0;
  for (int pos = rank; pos < nk; pos += nranks) idx[n++] = order[pos];
  *n_local = n;
  return BOLT_OK;
}

int bolt_comm_unique_id(bolt_ctx* ctx, void* id128) {
  if (!ctx || !id128) return BOLT_ERR_ARG;
  int rc = nccl_load(ctx); if (rc) return rc;
  NCCL_OK(g_nccl.GetUniqueId(id128), "ncclGetUniqueId");
  return BOLT_OK;
}

int bolt_comm_init(bolt_ctx* ctx, int rank, int nranks, const void* id128) {
  if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return ctx ? fail(ctx, BOLT_ERR_ARG, "bad communicator arguments") : BOLT_ERR_ARG;
  if (ctx->comm) return fail(ctx, BOLT_ERR_ARG, "communicator already initialised");
  int rc = nccl_load(ctx); if (rc) return rc;
  CUDA_OK(cudaSetDevice(ctx->device));
  NcclId128 id; memcpy(id.b, id128, 128);
  NCCL_OK(g_nccl.CommInitRank(&ctx->comm, nranks, id, rank), "ncclCommInitRank");
  ctx->rank = rank; ctx->nranks = nranks;
  return BOLT_OK;
}

int bolt_comm_free(bolt_ctx* ctx) {
  if (!ctx) return BOLT_ERR_ARG;
  if (ctx->comm) { cudaStreamSynchronize(ctx->stream); g_nccl.CommDestroy(ctx->comm); ctx->comm = nullptr; ctx->rank = 0; ctx->nranks = 1; }
  return BOLT_OK;
}

// bolt_spectra with the k-modes (K1) and the multipoles (K2) sharded over the communicator's ranks.  Every rank passes the
// SAME arguments and receives the SAME complete results.  K1 on the rank's cyclic shard of the descending-k order -> one
// ncclAllGather of the source columns (C_l is quadratic in the k-interpolated source, src/spectra.jl:91-93: both bracketing
// coarse columns must be present everywhere) -> K2 on multipoles rank, rank+R, ... -> ONE ncclAllReduce(sum) of the C_l vector
// (disjoint supports) -- all on the context's stream.  Without a communicator (or with one rank) it is bolt_spectra.
int bolt_spectra_sharded(bolt_ctx* ctx, const bolt_cosmo* c, const double* k, int nk, const bolt_opts* o,
                         const int32_t* ell, int nell, double kd_min, double kd_max, int n_kd, int ix_start,
                         double* cl_tt, double* cl_te, double* cl_ee, int32_t* status, int64_t* nsteps, int64_t* nreject) {
  if (!ctx) return BOLT_ERR_ARG;
  if (!ctx->comm || ctx->nranks == 1)
    return bolt_spectra(ctx, c, k, nk, o, ell, nell, kd_min, kd_max, n_kd, ix_start, cl_tt, cl_te, cl_ee, status, nsteps, nreject);
  if (!c || !k || nk < 2 || !ell || nell < 1) return fail(ctx, BOLT_ERR_ARG, "bad arguments");
  int rc = check_opts(ctx, c, o); if (rc) return rc;
  rc = check_projection_args(ctx, c, ell, nell, n_kd, ix_start); if (rc) return rc;
  CUDA_OK(cudaSetDevice(ctx->device));
  reset_timing(ctx);
  CUDA_OK(cudaEventRecord(ctx->ev[6], ctx->stream));
  const int R = ctx->nranks, rank = ctx->rank;
  const int nd = c->h.nd, row = c->h.n_x * nd;
  const int per = (nk + R - 1) / R;
  // descending-k order; this rank's shard = positions rank, rank+R, ...
  std::vector<int> order(nk);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return k[a] > k[b]; });
  std::vector<double> kloc; std::vector<int> ident;
  for (int pos = rank; pos < nk; pos += R) { kloc.push_back(k[order[pos]]); ident.push_back((int)ident.size()); }
  const int nloc = (int)kloc.size();
  DevBuf<double> d_k, d_kloc, d_loc, d_gath, d_ST, d_SP, d_cl_loc, d_cl; DevBuf<int> d_order, d_ident, d_st_loc, d_st_all;
  DevBuf<long long> d_cnt_loc, d_cnt_all;
  CUDA_OK(d_k.alloc(ctx, nk)); CUDA_OK(d_order.alloc(ctx, nk)); CUDA_OK(d_kloc.alloc(ctx, std::max(nloc, 1))); CUDA_OK(d_ident.alloc(ctx, std::max(nloc, 1)));
  CUDA_OK(cudaMemcpyAsync(d_k.p, k, nk * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_OK(cudaMemcpyAsync(d_order.p, order.data(), nk * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  if (nloc) {
    CUDA_OK(cudaMemcpyAsync(d_kloc.p, kloc.data(), nloc * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_OK(cudaMemcpyAsync(d_ident.p, ident.data(), nloc * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  }
  CUDA_OK(cudaStreamSynchronize(ctx->stream));     // host vectors are locals
  CUDA_OK(d_loc.alloc(ctx, (size_t)2 * per * row)); CUDA_OK(d_gath.alloc(ctx, (size_t)R * 2 * per * row));
  CUDA_OK(d_ST.alloc(ctx, (size_t)nk * row)); CUDA_OK(d_SP.alloc(ctx, (size_t)nk * row));
  CUDA_OK(d_st_loc.alloc(ctx, per)); CUDA_OK(d_st_all.alloc(ctx, (size_t)R * per));
  CUDA_OK(d_cnt_loc.alloc(ctx, (size_t)2 * per)); CUDA_OK(d_cnt_all.alloc(ctx, (size_t)R * 2 * per));
  CUDA_OK(cudaMemsetAsync(d_loc.p, 0, d_loc.n * 8, ctx->stream));          // see bolt_spectra: early-stopped modes, uncarried partials
  CUDA_OK(cudaMemsetAsync(d_st_loc.p, 0, per * sizeof(int), ctx->stream));
  CUDA_OK(cudaMemsetAsync(d_cnt_loc.p, 0, (size_t)2 * per * sizeof(long long), ctx->stream));
  bolt_opts oo = *o;
  oo.ix_first = std::max(oo.ix_first, ix_start);
  // K1 on the local shard (already in descending k: identity work order)
  if (nloc) {
    rc = launch_hierarchy(ctx, c, d_kloc.p, d_ident.p, nloc, &oo, d_loc.p, d_loc.p + (size_t)per * row, nullptr, nullptr, d_st_loc.p,
                          d_cnt_loc.p, d_cnt_loc.p + per);
    if (rc) return bail(ctx, rc);
  }
  // K2 prerequisites that do not depend on K1: the j_l tables of the local multipoles, on the second stream (fills K1's tail)
  std::vector<int32_t> ell_loc;
  for (int i = rank; i < nell; i += R) ell_loc.push_back(ell[i]);
  const int nl = (int)ell_loc.size();
  BesselTabs bt;
  if (nl) { rc = bessel_prepare(ctx, c, ell_loc.data(), nl, kd_max, ctx->stream2, bt); if (rc) return bail(ctx, rc); }
  // exchange 1: source columns (+ per-mode status and step counts, a few KB)
  NCCL_OK(g_nccl.AllGather(d_loc.p, d_gath.p, (size_t)2 * per * row, NCCL_FLOAT64, ctx->comm, ctx->stream), "ncclAllGather(sources)");
  NCCL_OK(g_nccl.AllGather(d_st_loc.p, d_st_all.p, (size_t)per, NCCL_INT32, ctx->comm, ctx->stream), "ncclAllGather(status)");
  NCCL_OK(g_nccl.AllGather(d_cnt_loc.p, d_cnt_all.p, (size_t)2 * per, NCCL_INT64, ctx->comm, ctx->stream), "ncclAllGather(step counts)");
  unshard_sources_kernel<<<nk, 256, 0, ctx->stream>>>(d_gath.p, d_order.p, nk, R, per, row, d_ST.p, d_SP.p);
  CUDA_OK(cudaGetLastError());
  // K2 on the local multipoles, scattered into the full (zeroed) C_l vector; exchange 2: one all-reduce
  const size_t ncl = (size_t)nell * nd;
  CUDA_OK(d_cl.alloc(ctx, 3 * ncl));
  CUDA_OK(cudaMemsetAsync(d_cl.p, 0, 3 * ncl * 8, ctx->stream));
  if (nl) {
    CUDA_OK(d_cl_loc.alloc(ctx, (size_t)3 * nl * nd));
    rc = project_device(ctx, c, d_ST.p, d_SP.p, d_k.p, nk, bt, nl, kd_min, kd_max, n_kd, ix_start, d_cl_loc.p);
    if (rc) return bail(ctx, rc);
    scatter_cl_kernel<<<(3 * nl * nd + 255) / 256, 256, 0, ctx->stream>>>(d_cl_loc.p, nl, nell, nd, rank, R, d_cl.p);
    CUDA_OK(cudaGetLastError());
  }
  NCCL_OK(g_nccl.AllReduce(d_cl.p, d_cl.p, 3 * ncl, NCCL_FLOAT64, NCCL_SUM, ctx->comm, ctx->stream), "ncclAllReduce(C_l)");
  if (cl_tt) CUDA_OK(cudaMemcpyAsync(cl_tt, d_cl.p, ncl * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (cl_te) CUDA_OK(cudaMemcpyAsync(cl_te, d_cl.p + ncl, ncl * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (cl_ee) CUDA_OK(cudaMemcpyAsync(cl_ee, d_cl.p + 2 * ncl, ncl * 8, cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<int> h_st((size_t)R * per); std::vector<long long> h_cnt((size_t)R * 2 * per);
  CUDA_OK(cudaMemcpyAsync(h_st.data(), d_st_all.p, h_st.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaMemcpyAsync(h_cnt.data(), d_cnt_all.p, h_cnt.size() * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaEventRecord(ctx->ev[7], ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  for (int pos = 0; pos < nk; pos++) {
    const int r = pos % R, j = pos / R, ik = order[pos];
    if (status) status[ik] = h_st[(size_t)r * per + j];
    if (nsteps) nsteps[ik] = h_cnt[((size_t)r * 2) * per + j];
    if (nreject) nreject[ik] = h_cnt[((size_t)r * 2 + 1) * per + j];
  }
  return collect_timing(ctx);
}

// plin over several GPUs ("P(k): no reduction, ncclAllGather of [nd][n_k]"): K1 + the epilogue on this rank's cyclic shard
// of the descending-k order, one all-gather of the P(k) rows (and of status / step counts), un-sharded on the host.
int bolt_plin_sharded(bolt_ctx* ctx, const bolt_cosmo* c, const double* k, int nk, const bolt_opts* o, double* pk, int32_t* status,
                      int64_t* nsteps) {
  if (!ctx) return BOLT_ERR_ARG;
  if (!ctx->comm || ctx->nranks == 1) return bolt_plin(ctx, c, k, nk, o, pk, status, nsteps);
  if (!c || !k || nk < 1 || !pk) return fail(ctx, BOLT_ERR_ARG, "bad arguments");
  int rc = check_opts(ctx, c, o); if (rc) return rc;
  CUDA_OK(cudaSetDevice(ctx->device));
  reset_timing(ctx);
  CUDA_OK(cudaEventRecord(ctx->ev[6], ctx->stream));
  const int R = ctx->nranks, rank = ctx->rank, nd = c->h.nd;
  const int n = bolt_state_dim(o->l_gamma, o->l_nu, o->l_mnu, c->h.nq);
  const int per = (nk + R - 1) / R;
  std::vector<int> order(nk);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return k[a] > k[b]; });
  std::vector<double> kloc; std::vector<int> ident;
  for (int pos = rank; pos < nk; pos += R) { kloc.push_back(k[order[pos]]); ident.push_back((int)ident.size()); }
  const int nloc = (int)kloc.size();
  DevBuf<double> d_kloc, d_final, d_pk_loc, d_pk_all; DevBuf<int> d_ident, d_st_loc, d_st_all; DevBuf<long long> d_ns_loc, d_ns_all;
  CUDA_OK(d_kloc.alloc(ctx, std::max(nloc, 1))); CUDA_OK(d_ident.alloc(ctx, std::max(nloc, 1)));
  if (nloc) {
    CUDA_OK(cudaMemcpyAsync(d_kloc.p, kloc.data(), nloc * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_OK(cudaMemcpyAsync(d_ident.p, ident.data(), nloc * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  }
  CUDA_OK(cudaStreamSynchronize(ctx->stream));     // host vectors are locals
  CUDA_OK(d_final.alloc(ctx, (size_t)std::max(nloc, 1) * n * nd));
  CUDA_OK(d_pk_loc.alloc(ctx, (size_t)per * nd)); CUDA_OK(d_pk_all.alloc(ctx, (size_t)R * per * nd));
  CUDA_OK(d_st_loc.alloc(ctx, per)); CUDA_OK(d_st_all.alloc(ctx, (size_t)R * per));
  CUDA_OK(d_ns_loc.alloc(ctx, per)); CUDA_OK(d_ns_all.alloc(ctx, (size_t)R * per));
  CUDA_OK(cudaMemsetAsync(d_final.p, 0, d_final.n * 8, ctx->stream));
  CUDA_OK(cudaMemsetAsync(d_pk_loc.p, 0, d_pk_loc.n * 8, ctx->stream));
  CUDA_OK(cudaMemsetAsync(d_st_loc.p, 0, per * sizeof(int), ctx->stream));
  CUDA_OK(cudaMemsetAsync(d_ns_loc.p, 0, per * sizeof(long long), ctx->stream));
  if (nloc) {
    bolt_opts oo = *o; oo.ix_first = c->h.n_x;
    rc = launch_hierarchy(ctx, c, d_kloc.p, d_ident.p, nloc, &oo, nullptr, nullptr, nullptr, d_final.p, d_st_loc.p, d_ns_loc.p, nullptr);
    if (rc) return bail(ctx, rc);
    rc = launch_plin_epilogue(ctx, c, d_kloc.p, nloc, d_final.p, o, d_pk_loc.p); if (rc) return bail(ctx, rc);
  }
  NCCL_OK(g_nccl.AllGather(d_pk_loc.p, d_pk_all.p, (size_t)per * nd, NCCL_FLOAT64, ctx->comm, ctx->stream), "ncclAllGather(P(k))");
  NCCL_OK(g_nccl.AllGather(d_st_loc.p, d_st_all.p, (size_t)per, NCCL_INT32, ctx->comm, ctx->stream), "ncclAllGather(status)");
  NCCL_OK(g_nccl.AllGather(d_ns_loc.p, d_ns_all.p, (size_t)per, NCCL_INT64, ctx->comm, ctx->stream), "ncclAllGather(step counts)");
  std::vector<double> h_pk((size_t)R * per * nd); std::vector<int> h_st((size_t)R * per); std::vector<long long> h_ns((size_t)R * per);
  CUDA_OK(cudaMemcpyAsync(h_pk.data(), d_pk_all.p, h_pk.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaMemcpyAsync(h_st.data(), d_st_all.p, h_st.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaMemcpyAsync(h_ns.data(), d_ns_all.p, h_ns.size() * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaEventRecord(ctx->ev[7], ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  for (int pos = 0; pos < nk; pos++) {
    const int r = pos % R, j = pos / R, ik = order[pos];
    for (int d = 0; d < nd; d++) pk[(size_t)ik * nd + d] = h_pk[((size_t)r * per + j) * nd + d];
    if (status) status[ik] = h_st[(size_t)r * per + j];
    if (nsteps) nsteps[ik] = h_ns[(size_t)r * per + j];
  }
  return collect_timing(ctx);
}

// FFTLog (src/util.jl:33-108; pinned by the reference's test/runtests.jl:11-35 against test/data/fftlog_example.txt)
int bolt_fftlog(bolt_ctx* ctx, const double* r, int N, double mu, double q, double k0r0, int kropt, int inverse,
                const double* a_re, const double* a_im, double* y, double* k_out, double* k0r0_out) {
  if (!ctx) return BOLT_ERR_ARG;
  if (!r || !a_re || !y || N < 2 || N > 4096 || (N & (N - 1)) != 0) return fail(ctx, BOLT_ERR_ARG, "bolt_fftlog: N must be a power of two, 2 <= N <= 4096");
  if (!(r[0] > 0.0) || !(r[N - 1] > r[0])) return fail(ctx, BOLT_ERR_ARG, "bolt_fftlog: r must be positive and increasing");
  CUDA_OK(cudaSetDevice(ctx->device));
  const double logrmin = log(r[0]), logrmax = log(r[N - 1]);
  const double L = logrmax - logrmin, dlnr = L / (N - 1), r0 = exp((logrmin + logrmax) / 2.0);       // plan_fftlog, util.jl:47-54
  if (kropt) k0r0 = fftlog_k0r0_low_ringing(N, mu, q, L, k0r0);
  if (k0r0_out) *k0r0_out = k0r0;
  if (k_out) {       // k = reverse(k0 exp(n L / N)), n = range(-N/2, N/2, length = N)        (util.jl:58-60)
    const double k0 = k0r0 / r0, nh = (double)(N / 2);
    for (int i = 0; i < N; i++) { const double n = -nh + (2.0 * nh) * (double)i / (double)(N - 1); k_out[N - 1 - i] = k0 * exp(n * L / N); }
  }
  int lg = 0; while ((1 << lg) < N) lg++;
  DevBuf<double> d_r, d_are, d_aim, d_y;
  CUDA_OK(d_r.alloc(ctx, N)); CUDA_OK(d_are.alloc(ctx, N)); CUDA_OK(d_y.alloc(ctx, 2 * (size_t)N));
  CUDA_OK(cudaMemcpyAsync(d_r.p, r, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_OK(cudaMemcpyAsync(d_are.p, a_re, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (a_im) { CUDA_OK(d_aim.alloc(ctx, N)); CUDA_OK(cudaMemcpyAsync(d_aim.p, a_im, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream)); }
  CUDA_OK(cudaFuncSetAttribute(fftlog_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * (int)sizeof(double2)));
  fftlog_kernel<<<1, std::min(N, 512), (size_t)N * sizeof(double2), ctx->stream>>>(d_r.p, N, lg, mu, q, dlnr, k0r0, inverse, d_are.p,
                                                                                  a_im ? d_aim.p : nullptr, reinterpret_cast<double2*>(d_y.p));
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpyAsync(y, d_y.p, 2 * (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return BOLT_OK;
}

int bolt_fp64_peak(bolt_ctx* ctx, double* tflops) {
  if (!ctx || !tflops) return BOLT_ERR_ARG;
  CUDA_OK(cudaSetDevice(ctx->device));
  const int blocks = ctx->num_sms * 8, threads = 256, iters = 1 << 15;
  DevBuf<double> d_out; CUDA_OK(d_out.alloc(ctx, (size_t)blocks * threads));
  double best = 0.0;
  for (int rep = 0; rep < 4; rep++) {
    CUDA_OK(cudaEventRecord(ctx->ev[6], ctx->stream));
    fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d_out.p, iters, 0.999999, 1e-9);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventRecord(ctx->ev[7], ctx->stream));
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    float ms = 0; CUDA_OK(cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]));
    const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
    if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  *tflops = best;
  return BOLT_OK;
}

int bolt_set_bessel_xmax(bolt_ctx* ctx, double xmax) {
  if (!ctx || xmax < 0.0) return BOLT_ERR_ARG;
  ctx->bessel_xmax = xmax;
  return BOLT_OK;
}

}  // extern "C"
