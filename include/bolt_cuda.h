/* libbolt_cuda.so -- C ABI of the B200-native Boltzmann hot path.
 *
 * The reference (xzackli/Bolt.jl) has no FFI boundary: its "operator interface" for this
 * path is the exported Julia function set (src/Bolt.jl:8-9).  This header defines the C ABI
 * those functions bind through `ccall`; each entry point cites the reference function it
 * replaces.  Everything crossing the boundary is plain C: pointers, sizes, POD structs.
 *
 * Dual numbers: a Julia `Vector{ForwardDiff.Dual{Tag,Float64,N}}` is bit-identical to a
 * column-major (1+N) x len Float64 matrix, value first.  `nd = 1+N` is therefore the leading
 * stride of every "dual-capable" array below; nd = 1 means plain Float64.
 *
 * Threading: a context is owned by one host thread; distinct contexts are independent.
 * All calls are synchronous (they return after the context's stream has drained).
 * Errors: 0 = success, negative = bolt_status; text via bolt_last_error().  Per-k solver
 * outcomes are reported in status[] because the reference never inspects the ODE retcode
 * (src/perturbations.jl:29-32).
 */
#ifndef BOLT_CUDA_H
#define BOLT_CUDA_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define BOLT_ABI_VERSION 3

/* order of the scalar block (each entry nd doubles, value first) */
enum bolt_scalar {
  BOLT_S_h = 0, BOLT_S_Omega_r, BOLT_S_Omega_b, BOLT_S_Omega_c, BOLT_S_A, BOLT_S_n,
  BOLT_S_Y_p, BOLT_S_N_nu, BOLT_S_Sum_m_nu,          /* CosmoParams, src/Bolt.jl:56-66        */
  BOLT_S_H0, BOLT_S_eta0, BOLT_S_rho_crit, BOLT_S_Omega_L, /* Background, src/background.jl:85-89 */
  BOLT_NSCALARS
};

/* order of the cubic-B-spline coefficient tables (each (n_x+2) x nd doubles):
 * bg.ℋ, bg.ℋ′, bg.ℋ′′, bg.η, bg.ρ₀ℳ (src/background.jl:95-101) and
 * ih.τ, ih.τ′, ih.τ′′, ih.g̃, ih.g̃′, ih.g̃′′, ih.csb² (src/ionization/recfast.jl:7-19). */
enum bolt_table {
  BOLT_T_H = 0, BOLT_T_Hp, BOLT_T_Hpp, BOLT_T_eta, BOLT_T_rho0M,
  BOLT_T_tau, BOLT_T_taup, BOLT_T_taupp, BOLT_T_g, BOLT_T_gp, BOLT_T_gpp, BOLT_T_csb2,
  BOLT_NTABLES
};

typedef struct bolt_cosmo_desc {
  int32_t abi_version;     /* BOLT_ABI_VERSION */
  int32_t nd;              /* 1 + number of ForwardDiff partials */
  int32_t n_x;             /* samples on bg.x_grid (tables hold n_x+2 coefficients) */
  int32_t nq;              /* momentum quadrature nodes (bg.quad_pts) */
  double  x0, dx;          /* bg.x_grid = x0 + dx*(0..n_x-1)  (src/background.jl:104) */
  const double* scalars;   /* [BOLT_NSCALARS][nd] */
  const double* quad_pts;  /* [nq]  Gauss-Legendre nodes on [-1,1] */
  const double* quad_wts;  /* [nq] */
  const double* tables;    /* [BOLT_NTABLES][n_x+2][nd] spline coefficients (itp.itp.coefs) */
} bolt_cosmo_desc;

enum bolt_mode { BOLT_MODE_ADAPTIVE = 0, BOLT_MODE_FIXED = 1 };

/* options of one batch of k-mode solves: Hierarchy(...) truncations (src/perturbations.jl:20-21)
 * and boltsolve keyword arguments (src/perturbations.jl:25). */
typedef struct bolt_opts {
  int32_t l_gamma, l_nu, l_mnu;  /* ℓᵧ, ℓ_ν, ℓ_mν */
  int32_t mode;                  /* bolt_mode */
  double  reltol, abstol;        /* adaptive mode */
  double  fixed_dt;              /* fixed mode: step in x; (0 - x0)/fixed_dt steps are taken */
  int64_t max_steps;             /* per k-mode; 0 = library default (1e6) */
  int32_t ix_first;              /* first x_grid row for which sources / history are sampled */
  int32_t reserved;
} bolt_opts;

enum bolt_status {
  BOLT_OK = 0,
  BOLT_ERR_ARG = -1, BOLT_ERR_CUDA = -2, BOLT_ERR_ALLOC = -3, BOLT_ERR_UNSUPPORTED = -4
};
/* per-k status[] values */
enum bolt_k_status { BOLT_K_OK = 0, BOLT_K_MAXSTEPS = 1, BOLT_K_DT_UNDERFLOW = 2,
                     BOLT_K_NONFINITE = 3, BOLT_K_RSA_TRIGGERED = 4 };

typedef struct bolt_ctx bolt_ctx;       /* opaque: device, stream, scratch */
typedef struct bolt_cosmo bolt_cosmo;   /* opaque: device-resident tables of one cosmology */

/* lifecycle ------------------------------------------------------------------------------ */
int  bolt_init(int device_ordinal, bolt_ctx** ctx);
int  bolt_finalize(bolt_ctx* ctx);
const char* bolt_last_error(const bolt_ctx* ctx);
int  bolt_abi_version(void);

/* Upload what Background(par) and IonizationHistory(𝕣, par, bg) computed on the host
 * (src/background.jl:104-128, src/ionization/recfast.jl:674-726).  Copies; the caller keeps
 * ownership of every host buffer. */
int  bolt_cosmo_upload(bolt_ctx* ctx, const bolt_cosmo_desc* desc, bolt_cosmo** out);
int  bolt_cosmo_free(bolt_ctx* ctx, bolt_cosmo* c);

/* State dimension n = 2(ℓᵧ+1)+(ℓ_ν+1)+(ℓ_mν+1)nq+5 (src/perturbations.jl:281). */
int  bolt_state_dim(int l_gamma, int l_nu, int l_mnu, int nq);

/* boltsolve / source_grid / source_grid_P for a batch of k-modes
 * (src/perturbations.jl:25-33, src/spectra.jl:6-42).  One solve per k feeds both sources.
 * Any output pointer may be NULL.  Host buffers, caller-owned:
 *   S_T, S_P   [nk][n_x][nd]   (i.e. Julia Matrix{T}(n_x, nk), column-major)
 *   u_hist     [nk][n_x][n][nd]  solution sampled on bg.x_grid (what perturb(x) returns there)
 *   u_final    [nk][n][nd]     perturb(0)
 *   status     [nk], nsteps [nk] (accepted), nreject [nk] */
int  bolt_solve(bolt_ctx* ctx, const bolt_cosmo* c, const double* k, int nk, const bolt_opts* o,
                double* S_T, double* S_P, double* u_hist, double* u_final,
                int32_t* status, int64_t* nsteps, int64_t* nreject);

/* cltt / clte / clee for a vector of multipoles (src/spectra.jl:84-160).
 * S_T,S_P are source grids on the coarse k grid k[nk] as returned by bolt_solve; the dense
 * integration grid is quadratic_k(kd_min, kd_max, n_kd) (src/spectra.jl:60-63,133); ix_start is
 * the 0-based index of the first x_grid point > -8 (src/spectra.jl:86).
 * cl_* are [nell][nd]; any may be NULL. */
int  bolt_project(bolt_ctx* ctx, const bolt_cosmo* c, const double* S_T, const double* S_P,
                  const double* k, int nk, const int32_t* ell, int nell,
                  double kd_min, double kd_max, int n_kd, int ix_start,
                  double* cl_tt, double* cl_te, double* cl_ee);

/* Fused source_grid + source_grid_P + cltt/clte/clee with the source grids kept in HBM
 * (no host round trip).  Same meaning as bolt_solve followed by bolt_project. */
int  bolt_spectra(bolt_ctx* ctx, const bolt_cosmo* c, const double* k, int nk, const bolt_opts* o,
                  const int32_t* ell, int nell, double kd_min, double kd_max, int n_kd, int ix_start,
                  double* cl_tt, double* cl_te, double* cl_ee,
                  int32_t* status, int64_t* nsteps, int64_t* nreject);

/* bolt_spectra for a batch of cosmologies (the emulator / MCMC workload: SURVEY 8d C5; the reference maps source_grid over
 * parameter sets one at a time).  All ncos x nk hierarchy solves share ONE launch (one work queue, longest solves first), so
 * the low-occupancy tail of the persistent kernel is paid once per batch instead of once per cosmology; the projections run
 * per cosmology.  k is [ncos][nk] (each cosmology its own grid), kd_min/kd_max are [ncos]; cl_* are [ncos][nell], status and
 * nsteps [ncos][nk].  Value-only (nd = 1), the cosmologies must share x_grid and nq, 1 <= ncos <= BOLT_MAX_BATCH. */
#define BOLT_MAX_BATCH 16
int  bolt_spectra_batch(bolt_ctx* ctx, const bolt_cosmo* const* cosmos, int ncos, const double* k, int nk,
                        const bolt_opts* o, const int32_t* ell, int nell, const double* kd_min, const double* kd_max,
                        int n_kd, int ix_start, double* cl_tt, double* cl_te, double* cl_ee,
                        int32_t* status, int64_t* nsteps);

/* plin for a vector of k (src/spectra.jl:163-198; x = 0). pk is [nk][nd]. */
int  bolt_plin(bolt_ctx* ctx, const bolt_cosmo* c, const double* k, int nk, const bolt_opts* o,
               double* pk, int32_t* status, int64_t* nsteps);

/* Timing of the last call's kernels on the context's stream (CUDA events), in ms:
 * out[0] = hierarchy kernel, out[1] = Bessel tables, out[2] = projection kernel, out[3] = total.
 * out[4..7] = launch counts of the same. */
int  bolt_last_timing(const bolt_ctx* ctx, double* out8);

/* Fix the range [0, xmax] of the 5001-point j_l tables for subsequent projections on this context (0 restores the default
 * kgrid[end]*eta0, src/spectra.jl:85).  The reference treats that range as non-differentiable (assume_nondual,
 * src/spectra.jl:46-52); finite-difference checks of gradients must therefore hold it fixed across the +/- runs. */
int  bolt_set_bessel_xmax(bolt_ctx* ctx, double xmax);

/* Measured DFMA throughput of the device (TFLOP/s): the FP64 roofline denominator. */
int  bolt_fp64_peak(bolt_ctx* ctx, double* tflops);

/* Device-pointer variants for callers that own HBM buffers (e.g. torch tensors exchanged with NCCL when the
 * k-modes of ONE cosmology are sharded over GPUs: solve local k -> all-gather S_T,S_P -> project local l ->
 * all-reduce C_l).  Same semantics as bolt_solve / bolt_project, every array argument except `ell` is a
 * DEVICE pointer on the context's device; status/nsteps/nreject use int32/int64 device arrays.  The calls
 * run on the context's stream and return after it has drained. */
int  bolt_solve_device(bolt_ctx* ctx, const bolt_cosmo* c, const double* d_k, int nk, const bolt_opts* o,
                       double* d_S_T, double* d_S_P, double* d_u_final,
                       int32_t* d_status, int64_t* d_nsteps, int64_t* d_nreject);
int  bolt_project_device(bolt_ctx* ctx, const bolt_cosmo* c, const double* d_S_T, const double* d_S_P,
                         const double* d_k, int nk, const int32_t* ell, int nell,
                         double kd_min, double kd_max, int n_kd, int ix_start,
                         double* d_cl /* [3][nell]: tt, te, ee */);

/* Multi-GPU (one process per GPU, one context per process): the k-modes of ONE cosmology sharded over the ranks of a
 * communicator -- what replaces the reference's threaded fan-out over k (src/spectra.jl:10, `tmap`) and over l
 * (src/spectra.jl:149, `qmap`) when one GPU is not enough (BASELINE config 4).  NCCL is bound at run time.
 *   bolt_comm_unique_id   rank 0 creates the 128-byte id; the HOST carries it to the other ranks (MPI.jl bcast,
 *                         torch.distributed, a file): the only thing that crosses between processes outside NCCL.
 *   bolt_comm_init        collective over all ranks.
 *   bolt_spectra_sharded  same arguments and results as bolt_spectra on EVERY rank.  K1 on the rank's cyclic shard of the
 *                         descending-k order, one ncclAllGather of the source columns, K2 on multipoles rank, rank+R, ...,
 *                         ONE ncclAllReduce(sum, double) of the C_l vector; all on the context's stream.
 *   bolt_plin_sharded     same arguments and results as bolt_plin on EVERY rank: K1 and the P(k) epilogue on the rank's shard, ONE
 *                         ncclAllGather of the [n_local][nd] rows (+ status and step counts); no reduction.
 *   bolt_shard_plan       (host only, no device needed) the indices into k[] that rank `rank` of `nranks` solves, in work order. */
int  bolt_comm_unique_id(bolt_ctx* ctx, void* id128);
int  bolt_comm_init(bolt_ctx* ctx, int rank, int nranks, const void* id128);
int  bolt_comm_free(bolt_ctx* ctx);
int  bolt_spectra_sharded(bolt_ctx* ctx, const bolt_cosmo* c, const double* k, int nk, const bolt_opts* o,
                          const int32_t* ell, int nell, double kd_min, double kd_max, int n_kd, int ix_start,
                          double* cl_tt, double* cl_te, double* cl_ee,
                          int32_t* status, int64_t* nsteps, int64_t* nreject);
int  bolt_plin_sharded(bolt_ctx* ctx, const bolt_cosmo* c, const double* k, int nk, const bolt_opts* o,
                       double* pk, int32_t* status, int64_t* nsteps);
int  bolt_shard_plan(const double* k, int nk, int rank, int nranks, int32_t* idx, int32_t* n_local);

/* FFTLog (src/util.jl:33-108: plan_fftlog + mul! / ldiv!), SURVEY 8f row n4: the biased Hankel-type transform of a[N] sampled on
 * the log-spaced grid r[N] (N a power of two <= 4096), order mu, bias q.  kropt != 0 applies k0r0_low_ringing (util.jl:79-89).
 * inverse = 0: mul! (multiply by u_m), 1: ldiv! (divide).  a_im may be NULL.  y is [N][2] (re, im); k_out [N] (the output
 * abscissae, may be NULL); k0r0_out the value actually used (may be NULL).  Not used by any spectrum function. */
int  bolt_fftlog(bolt_ctx* ctx, const double* r, int N, double mu, double q, double k0r0, int kropt, int inverse,
                 const double* a_re, const double* a_im, double* y, double* k_out, double* k0r0_out);

/* Batched input tables on the device: Background + RECFAST + reionization + optical depth for ncos cosmologies
 * at once -- what the reference computes on the host one cosmology at a time (src/background.jl:5-128,
 * src/ionization/recfast.jl:22-536, 674-726, src/ionization/ionization.jl:107-137).  params is [ncos][9] in CosmoParams order
 * (h, Omega_r, Omega_b, Omega_c, A, n, Y_p, N_nu, Sum_m_nu; src/Bolt.jl:56-66); the x grid is x0 + dx*(0..n_x-1); quad_pts/quad_wts
 * the Gauss-Legendre rule of Background (background.jl:105).  Output, caller-owned HOST buffers in the layout bolt_cosmo_desc
 * points at (nd = 1): tables_out [ncos][BOLT_NTABLES][n_x+2], scalars_out [ncos][BOLT_NSCALARS]; status_out [ncos] (0 = ok,
 * 1 = the recombination integrator gave up) may be NULL.  Value-only.  No context: the call selects `device_ordinal` itself. */
int  bolt_hostgen_batch(int device_ordinal, const double* params, int ncos, double x0, double dx, int n_x,
                        const double* quad_pts, const double* quad_wts, int nq,
                        double* tables_out, double* scalars_out, int32_t* status_out);
const char* bolt_hostgen_last_error(void);

/* Bessel-moment tables and the Filon line-of-sight rule.  Replaces, on the device,
 *   src/bessel/moments.jl:56-112      sph_j_moment_{weniger_1F2, asymp, maclaurin_1F2} (and, through J_nu = sqrt(2t/pi) j_{nu-1/2},
 *                                     the J_moment_* building blocks of moments.jl:24-51: pass half-integer powers)
 *   src/bessel/interpolator.jl:26-110 sph_bessel_interpolator(nu, order, keta_min, keta_max, N; weniger_cut) and the MomentTable call
 *   src/bessel/integrator.jl:7-38     integrate_sph_bessel_filon and its loop form.
 * All pointers are HOST buffers.  No context: calls select `device_ordinal` themselves (a table remembers its device).
 *   bolt_sph_j_moments   out[n][n_powers] = int_0^x t^power j_nu(t) dt at x[n]; powers NULL means 0,1,..,n_powers-1 (n_powers <= 4);
 *                        method 0: the small-argument evaluator (the role of the reference's Double64 Weniger sum; here Maclaurin
 *                        below x = 4, Gauss-Legendre pieces on a double-double prefix table above), 1: Lommel asymptotic form,
 *                        2: Maclaurin series.  nu = 1, 2, 3.
 *   bolt_moment_table_create  the N-node cubic-B-spline table of the first `order` (3 or 4; 1..4 accepted) moments on
 *                        [keta_min, keta_max], nu = 2 or 3; nodes below weniger_cut from method 0, above from method 1.
 *   bolt_moment_table_eval    out[n][order]: the table inside its range, Maclaurin below, asymptotic above (interpolator.jl:26-34).
 *   bolt_filon_pieces    out[i] = int_a^b (f + f'(x-a) + f''(x-a)^2/2) j_nu(k x) dx for n independent pieces (order >= 3).
 *   bolt_filon_chain     for every k[n_k]: the sum of the pieces between consecutive nodes[n_nodes] with f, f', f'' given at the
 *                        nodes as [n_k][n_nodes]; each node's moments are evaluated once (the loop form, integrator.jl:25-38).
 *                        kernel_ms (may be NULL): device time of one warm launch. */
typedef struct bolt_moment_table bolt_moment_table;
int  bolt_sph_j_moments(int device_ordinal, int nu, int n_powers, const double* powers, int method, const double* x, int n, double* out);
int  bolt_moment_table_create(int device_ordinal, int nu, int order, double keta_min, double keta_max, int N, double weniger_cut,
                              bolt_moment_table** out);
void bolt_moment_table_free(bolt_moment_table* t);
int  bolt_moment_table_eval(bolt_moment_table* t, const double* x, int n, double* out);
int  bolt_filon_pieces(bolt_moment_table* t, int n, const double* f, const double* f1, const double* f2, const double* k,
                       const double* a, const double* b, double* out);
int  bolt_filon_chain(bolt_moment_table* t, int n_k, int n_nodes, const double* nodes, const double* f, const double* f1,
                      const double* f2, const double* k, double* out, float* kernel_ms);
const char* bolt_moments_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
