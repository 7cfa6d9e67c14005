"""ctypes binding of libbolt_cuda.so (include/bolt_cuda.h) -- the Python twin of julia/BoltCUDA.jl.

There is no CPU fallback: loading fails loudly if the shared library is missing, and `Context()`
raises if no CUDA device is usable.
"""
import ctypes as C
import os
import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BOLT_CUDA_LIB", os.path.join(_HERE, "csrc", "libbolt_cuda.so"))
_LIB = None

EXPORTS = ["bolt_abi_version", "bolt_init", "bolt_finalize", "bolt_last_error", "bolt_last_timing",
           "bolt_cosmo_upload", "bolt_cosmo_free", "bolt_state_dim", "bolt_solve", "bolt_project",
           "bolt_spectra", "bolt_spectra_batch", "bolt_plin", "bolt_solve_device", "bolt_project_device", "bolt_fp64_peak", "bolt_set_bessel_xmax",
           "bolt_comm_unique_id", "bolt_comm_init", "bolt_comm_free", "bolt_spectra_sharded", "bolt_plin_sharded", "bolt_shard_plan", "bolt_fftlog",
           "bolt_hostgen_batch", "bolt_hostgen_last_error",
           "bolt_sph_j_moments", "bolt_moment_table_create", "bolt_moment_table_free", "bolt_moment_table_eval",
           "bolt_filon_pieces", "bolt_filon_chain", "bolt_moments_last_error"]


class BoltError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise BoltError(f"{LIB_PATH} is missing: build it with `make -C {os.path.dirname(LIB_PATH)}` "
                            "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp, dp, ip, lp = C.c_void_p, abi.c_double_p, abi.c_int32_p, abi.c_int64_p
        L.bolt_abi_version.restype = C.c_int
        L.bolt_init.argtypes = [C.c_int, C.POINTER(vp)]
        L.bolt_finalize.argtypes = [vp]
        L.bolt_last_error.argtypes = [vp]; L.bolt_last_error.restype = C.c_char_p
        L.bolt_last_timing.argtypes = [vp, dp]
        L.bolt_cosmo_upload.argtypes = [vp, C.POINTER(abi.CosmoDesc), C.POINTER(vp)]
        L.bolt_cosmo_free.argtypes = [vp, vp]
        L.bolt_state_dim.argtypes = [C.c_int] * 4
        L.bolt_solve.argtypes = [vp, vp, dp, C.c_int, C.POINTER(abi.Opts), dp, dp, dp, dp, ip, lp, lp]
        L.bolt_project.argtypes = [vp, vp, dp, dp, dp, C.c_int, ip, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                   dp, dp, dp]
        L.bolt_spectra.argtypes = [vp, vp, dp, C.c_int, C.POINTER(abi.Opts), ip, C.c_int, C.c_double, C.c_double,
                                   C.c_int, C.c_int, dp, dp, dp, ip, lp, lp]
        L.bolt_spectra_batch.argtypes = [vp, C.POINTER(vp), C.c_int, dp, C.c_int, C.POINTER(abi.Opts), ip, C.c_int, dp, dp,
                                         C.c_int, C.c_int, dp, dp, dp, ip, lp]
        L.bolt_plin.argtypes = [vp, vp, dp, C.c_int, C.POINTER(abi.Opts), dp, ip, lp]
        L.bolt_fp64_peak.argtypes = [vp, dp]
        L.bolt_set_bessel_xmax.argtypes = [vp, C.c_double]
        L.bolt_solve_device.argtypes = [vp, vp, vp, C.c_int, C.POINTER(abi.Opts), vp, vp, vp, vp, vp, vp]
        L.bolt_project_device.argtypes = [vp, vp, vp, vp, vp, C.c_int, ip, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, vp]
        L.bolt_comm_unique_id.argtypes = [vp, vp]
        L.bolt_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
        L.bolt_comm_free.argtypes = [vp]
        L.bolt_spectra_sharded.argtypes = L.bolt_spectra.argtypes
        L.bolt_plin_sharded.argtypes = L.bolt_plin.argtypes
        L.bolt_shard_plan.argtypes = [dp, C.c_int, C.c_int, C.c_int, ip, ip]
        L.bolt_hostgen_batch.argtypes = [C.c_int, dp, C.c_int, C.c_double, C.c_double, C.c_int, dp, dp, C.c_int, dp, dp, ip]
        L.bolt_hostgen_last_error.restype = C.c_char_p
        L.bolt_fftlog.argtypes = [vp, dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, dp, dp, dp, dp, dp]
        L.bolt_sph_j_moments.argtypes = [C.c_int, C.c_int, C.c_int, dp, C.c_int, dp, C.c_int, dp]
        L.bolt_moment_table_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.POINTER(vp)]
        L.bolt_moment_table_free.argtypes = [vp]; L.bolt_moment_table_free.restype = None
        L.bolt_moment_table_eval.argtypes = [vp, dp, C.c_int, dp]
        L.bolt_filon_pieces.argtypes = [vp, C.c_int, dp, dp, dp, dp, dp, dp, dp]
        L.bolt_filon_chain.argtypes = [vp, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, C.POINTER(C.c_float)]
        L.bolt_moments_last_error.restype = C.c_char_p
        _LIB = L
    return _LIB


class Context:
    """bolt_ctx: one per (host thread x device)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().bolt_init(device, C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            raise BoltError(f"bolt_init(device={device}) failed with {rc}: no usable CUDA device "
                            "(libbolt_cuda has no CPU fallback)")

    def check(self, rc):
        if rc != 0:
            raise BoltError(f"libbolt_cuda error {rc}: {lib().bolt_last_error(self._h).decode()}")

    def timing(self):
        t = np.zeros(8)
        lib().bolt_last_timing(self._h, abi.ptr(t))
        return dict(hierarchy_ms=t[0], bessel_ms=t[1], project_ms=t[2], total_ms=t[3],
                    hierarchy_launches=int(t[4]), bessel_launches=int(t[5]), project_launches=int(t[6]))

    def set_bessel_xmax(self, xmax):
        self.check(lib().bolt_set_bessel_xmax(self._h, float(xmax)))

    def fp64_peak_tflops(self):
        t = np.zeros(1)
        self.check(lib().bolt_fp64_peak(self._h, abi.ptr(t)))
        return float(t[0])

    # ---- multi-GPU: one context per rank; the host only carries the 128-byte id from rank 0 to the others -------------
    def comm_unique_id(self):
        buf = (C.c_char * 128)()
        self.check(lib().bolt_comm_unique_id(self._h, C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def comm_init(self, rank, nranks, unique_id):
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        self.check(lib().bolt_comm_init(self._h, int(rank), int(nranks), C.cast(buf, C.c_void_p)))
        self.rank, self.nranks = int(rank), int(nranks)

    def comm_init_torch(self, group=None):
        """Bootstrap the library's communicator from an initialised torch.distributed process group (any backend)."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [self.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        self.comm_init(rank, world, box[0])

    def comm_free(self):
        if self._h:
            lib().bolt_comm_free(self._h)

    def close(self):
        if self._h:
            lib().bolt_finalize(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def spectra_batch(ctx, dcs, ks, opts, ells, kd_min, kd_max, n_kd, ix_start):
    """bolt_spectra_batch: dcs = list of DeviceCosmo on ctx, ks = [ncos][nk], kd_min/kd_max = [ncos].
    Returns tt, te, ee [ncos][nell], status, nsteps [ncos][nk]."""
    ncos = len(dcs)
    ks = np.ascontiguousarray(ks, dtype=np.float64); assert ks.shape[0] == ncos
    ells = np.ascontiguousarray(ells, dtype=np.int32)
    kd_min = np.ascontiguousarray(kd_min, dtype=np.float64); kd_max = np.ascontiguousarray(kd_max, dtype=np.float64)
    nk, nell = ks.shape[1], len(ells)
    tt, te, ee = np.zeros((ncos, nell)), np.zeros((ncos, nell)), np.zeros((ncos, nell))
    st = np.zeros((ncos, nk), dtype=np.int32); ns = np.zeros((ncos, nk), dtype=np.int64)
    handles = (C.c_void_p * ncos)(*[d._h for d in dcs])
    ctx.check(lib().bolt_spectra_batch(ctx._h, handles, ncos, abi.ptr(ks), nk, C.byref(opts), abi.ptr(ells, abi.c_int32_p), nell,
                                       abi.ptr(kd_min), abi.ptr(kd_max), n_kd, ix_start, abi.ptr(tt), abi.ptr(te), abi.ptr(ee),
                                       abi.ptr(st, abi.c_int32_p), abi.ptr(ns, abi.c_int64_p)))
    return tt, te, ee, st, ns


class DeviceCosmo:
    """bolt_cosmo: the device-resident tables of one cosmology."""

    SUPPORTED_NP = (0, 1, 2, 3, 4, 6)       # kernel instantiations of this build (partials per call)

    def __init__(self, ctx, host_cosmo):
        # any number of partials up to 6: a count without its own instantiation (5) is padded with all-zero partials to the
        # next one; `nd_user` components are handed back (the padded ones are identically zero and dropped)
        self.nd_user = host_cosmo.nd
        np_ = host_cosmo.nd - 1
        if np_ not in self.SUPPORTED_NP:
            bigger = [v for v in self.SUPPORTED_NP if v > np_]
            if not bigger:
                raise BoltError(f"{np_} partials per call: this build carries at most {self.SUPPORTED_NP[-1]}")
            host_cosmo = host_cosmo.padded(1 + bigger[0])
        self.ctx, self.hc = ctx, host_cosmo
        self._h = C.c_void_p()
        ctx.check(lib().bolt_cosmo_upload(ctx._h, C.byref(host_cosmo.desc), C.byref(self._h)))

    def _user(self, a):
        """Drop padded partial components from a dual-capable result."""
        if a is None or self.hc.nd == self.nd_user or a.ndim == 0 or a.shape[-1] != self.hc.nd:
            return a
        return np.ascontiguousarray(a[..., :self.nd_user]) if self.nd_user > 1 else np.ascontiguousarray(a[..., 0])

    def close(self):
        if self._h and self.ctx._h:
            lib().bolt_cosmo_free(self.ctx._h, self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self, k, opts, want=("S_T", "S_P")):
        k = np.ascontiguousarray(k, dtype=np.float64)
        nk, n_x = len(k), self.hc.n_x
        n = abi.state_dim(opts.l_gamma, opts.l_nu, opts.l_mnu, self.hc.nq)
        out = {}
        nd = self.hc.nd
        tail = () if nd == 1 else (nd,)      # dual-capable arrays carry a trailing (value, partials...) axis
        out["S_T"] = np.zeros((nk, n_x) + tail) if "S_T" in want else None
        out["S_P"] = np.zeros((nk, n_x) + tail) if "S_P" in want else None
        out["u_hist"] = np.zeros((nk, n_x, n)) if "u_hist" in want else None
        out["u_final"] = np.zeros((nk, n) + tail) if "u_final" in want else None
        out["status"] = np.zeros(nk, dtype=np.int32)
        out["nsteps"] = np.zeros(nk, dtype=np.int64)
        out["nreject"] = np.zeros(nk, dtype=np.int64)
        self.ctx.check(lib().bolt_solve(self.ctx._h, self._h, abi.ptr(k), nk, C.byref(opts), abi.ptr(out["S_T"]),
                                        abi.ptr(out["S_P"]), abi.ptr(out["u_hist"]), abi.ptr(out["u_final"]),
                                        abi.ptr(out["status"], abi.c_int32_p), abi.ptr(out["nsteps"], abi.c_int64_p),
                                        abi.ptr(out["nreject"], abi.c_int64_p)))
        for key in ("S_T", "S_P", "u_final"):
            out[key] = self._user(out[key])
        return out

    def project(self, S_T, S_P, k, ells, kd_min, kd_max, n_kd, ix_start):
        k = np.ascontiguousarray(k, dtype=np.float64)
        ells = np.ascontiguousarray(ells, dtype=np.int32)
        def widen(S):       # source grids coming back from the caller carry nd_user components: re-attach the zero padding
            if S is None:
                return None
            S = np.ascontiguousarray(S, dtype=np.float64)
            if self.hc.nd != self.nd_user and S.shape[-1] == self.nd_user:
                S = np.concatenate([S, np.zeros(S.shape[:-1] + (self.hc.nd - self.nd_user,))], axis=-1)
            return np.ascontiguousarray(S)
        S_T, S_P = widen(S_T), widen(S_P)
        shp = (len(ells),) if self.hc.nd == 1 else (len(ells), self.hc.nd)
        tt = np.zeros(shp) if S_T is not None else None
        ee = np.zeros(shp) if S_P is not None else None
        te = np.zeros(shp) if (S_T is not None and S_P is not None) else None
        self.ctx.check(lib().bolt_project(self.ctx._h, self._h, abi.ptr(S_T), abi.ptr(S_P), abi.ptr(k), len(k),
                                          abi.ptr(ells, abi.c_int32_p), len(ells), kd_min, kd_max, n_kd, ix_start,
                                          abi.ptr(tt), abi.ptr(te), abi.ptr(ee)))
        return self._user(tt), self._user(te), self._user(ee)

    def spectra(self, k, opts, ells, kd_min, kd_max, n_kd, ix_start):
        k = np.ascontiguousarray(k, dtype=np.float64)
        ells = np.ascontiguousarray(ells, dtype=np.int32)
        shp = (len(ells),) if self.hc.nd == 1 else (len(ells), self.hc.nd)
        tt, te, ee = np.zeros(shp), np.zeros(shp), np.zeros(shp)
        st = np.zeros(len(k), dtype=np.int32); ns = np.zeros(len(k), dtype=np.int64); nr = np.zeros(len(k), dtype=np.int64)
        self.ctx.check(lib().bolt_spectra(self.ctx._h, self._h, abi.ptr(k), len(k), C.byref(opts),
                                          abi.ptr(ells, abi.c_int32_p), len(ells), kd_min, kd_max, n_kd, ix_start,
                                          abi.ptr(tt), abi.ptr(te), abi.ptr(ee), abi.ptr(st, abi.c_int32_p),
                                          abi.ptr(ns, abi.c_int64_p), abi.ptr(nr, abi.c_int64_p)))
        self.last_nreject = nr          # rejected steps per mode of the last call (cost as much as accepted ones)
        return self._user(tt), self._user(te), self._user(ee), st, ns

    def spectra_sharded(self, k, opts, ells, kd_min, kd_max, n_kd, ix_start):
        """bolt_spectra_sharded: collective over the ranks of the context's communicator (Context.comm_init)."""
        k = np.ascontiguousarray(k, dtype=np.float64)
        ells = np.ascontiguousarray(ells, dtype=np.int32)
        shp = (len(ells),) if self.hc.nd == 1 else (len(ells), self.hc.nd)
        tt, te, ee = np.zeros(shp), np.zeros(shp), np.zeros(shp)
        st = np.zeros(len(k), dtype=np.int32); ns = np.zeros(len(k), dtype=np.int64); nr = np.zeros(len(k), dtype=np.int64)
        self.ctx.check(lib().bolt_spectra_sharded(self.ctx._h, self._h, abi.ptr(k), len(k), C.byref(opts),
                                                  abi.ptr(ells, abi.c_int32_p), len(ells), kd_min, kd_max, n_kd, ix_start,
                                                  abi.ptr(tt), abi.ptr(te), abi.ptr(ee), abi.ptr(st, abi.c_int32_p),
                                                  abi.ptr(ns, abi.c_int64_p), abi.ptr(nr, abi.c_int64_p)))
        self.last_nreject = nr
        return self._user(tt), self._user(te), self._user(ee), st, ns

    def plin(self, k, opts):
        k = np.ascontiguousarray(k, dtype=np.float64)
        pk = np.zeros(len(k) if self.hc.nd == 1 else (len(k), self.hc.nd))
        st = np.zeros(len(k), dtype=np.int32); ns = np.zeros(len(k), dtype=np.int64)
        self.ctx.check(lib().bolt_plin(self.ctx._h, self._h, abi.ptr(k), len(k), C.byref(opts), abi.ptr(pk),
                                       abi.ptr(st, abi.c_int32_p), abi.ptr(ns, abi.c_int64_p)))
        return self._user(pk), st, ns

    def plin_sharded(self, k, opts):
        """bolt_plin_sharded: collective over the ranks of the context's communicator; same results as plin on every rank."""
        k = np.ascontiguousarray(k, dtype=np.float64)
        pk = np.zeros(len(k) if self.hc.nd == 1 else (len(k), self.hc.nd))
        st = np.zeros(len(k), dtype=np.int32); ns = np.zeros(len(k), dtype=np.int64)
        self.ctx.check(lib().bolt_plin_sharded(self.ctx._h, self._h, abi.ptr(k), len(k), C.byref(opts), abi.ptr(pk),
                                               abi.ptr(st, abi.c_int32_p), abi.ptr(ns, abi.c_int64_p)))
        return self._user(pk), st, ns

    # ---- device-pointer variants (torch tensors own the HBM buffers) ---------------------------------
    def solve_device(self, k_t, opts, want_final=False):
        """k_t: float64 CUDA tensor [nk].  Returns CUDA tensors S_T, S_P [nk][n_x], status, nsteps."""
        import torch
        nk, n_x = k_t.numel(), self.hc.n_x
        dev = k_t.device
        if self.hc.nd != 1:
            raise BoltError("solve_device: value-only cosmologies (nd = 1); the kernels with partials write [nk][n_x][nd]")
        S_T = torch.zeros((nk, n_x), dtype=torch.float64, device=dev)
        S_P = torch.zeros((nk, n_x), dtype=torch.float64, device=dev)
        status = torch.zeros(nk, dtype=torch.int32, device=dev)
        nsteps = torch.zeros(nk, dtype=torch.int64, device=dev)
        n = abi.state_dim(opts.l_gamma, opts.l_nu, opts.l_mnu, self.hc.nq)
        u_final = torch.zeros((nk, n), dtype=torch.float64, device=dev) if want_final else None
        torch.cuda.current_stream(dev).synchronize()
        self.ctx.check(lib().bolt_solve_device(self.ctx._h, self._h, k_t.data_ptr(), nk, C.byref(opts), S_T.data_ptr(), S_P.data_ptr(),
                                               u_final.data_ptr() if want_final else None, status.data_ptr(), nsteps.data_ptr(), None))
        return S_T, S_P, status, nsteps, u_final

    def project_device(self, S_T, S_P, k_t, ells, kd_min, kd_max, n_kd, ix_start):
        """S_T, S_P: float64 CUDA tensors [nk][n_x]; returns a CUDA tensor [3][nell] (tt, te, ee)."""
        import torch
        ells = np.ascontiguousarray(ells, dtype=np.int32)
        cl = torch.zeros((3, len(ells)), dtype=torch.float64, device=k_t.device)
        torch.cuda.current_stream(k_t.device).synchronize()
        self.ctx.check(lib().bolt_project_device(self.ctx._h, self._h, S_T.data_ptr(), S_P.data_ptr(), k_t.data_ptr(), k_t.numel(),
                                                 abi.ptr(ells, abi.c_int32_p), len(ells), kd_min, kd_max, n_kd, ix_start, cl.data_ptr()))
        return cl


def shard_plan(k, rank, nranks):
    """bolt_shard_plan (host only): indices into k that `rank` of `nranks` solves, in work order (descending k, cyclic)."""
    k = np.ascontiguousarray(k, dtype=np.float64)
    idx = np.zeros(len(k), dtype=np.int32); n = np.zeros(1, dtype=np.int32)
    rc = lib().bolt_shard_plan(abi.ptr(k), len(k), int(rank), int(nranks), abi.ptr(idx, abi.c_int32_p), abi.ptr(n, abi.c_int32_p))
    if rc != 0:
        raise BoltError(f"bolt_shard_plan failed with {rc}")
    return idx[:int(n[0])].copy()


def fftlog(ctx, r, a, mu, q, k0r0=1.0, kropt=True, inverse=False):
    """bolt_fftlog: Bolt.plan_fftlog(r, mu, q, k0r0; kropt) followed by mul! (inverse=False) or ldiv! (src/util.jl:33-108).
    Returns (y complex [N], k [N], k0r0 used)."""
    r = np.ascontiguousarray(r, dtype=np.float64); a = np.asarray(a)
    are = np.ascontiguousarray(a.real, dtype=np.float64)
    aim = np.ascontiguousarray(a.imag, dtype=np.float64) if np.iscomplexobj(a) else None
    y = np.zeros((len(r), 2)); k = np.zeros(len(r)); used = np.zeros(1)
    ctx.check(lib().bolt_fftlog(ctx._h, abi.ptr(r), len(r), float(mu), float(q), float(k0r0), int(bool(kropt)), int(bool(inverse)),
                                abi.ptr(are), abi.ptr(aim), abi.ptr(y), abi.ptr(k), abi.ptr(used)))
    return y[:, 0] + 1j * y[:, 1], k, float(used[0])


def hostgen_batch(pars, x0=-20.0, dx=0.01, n_x=2001, nq=15, device=0):
    """bolt_hostgen_batch: the input tables of a BATCH of cosmologies computed on the device (background, RECFAST, reionization,
    optical depth, visibility, baryon sound speed and their spline coefficients).  pars: list of CosmoParams.
    Returns (list of abi.HostCosmo ready for DeviceCosmo, status [ncos])."""
    names = ["h", "Ω_r", "Ω_b", "Ω_c", "A", "n", "Y_p", "N_ν", "Σm_ν"]
    P = np.ascontiguousarray([[getattr(p, nm) for nm in names] for p in pars], dtype=np.float64)
    ncos = len(pars)
    pts, wts = np.polynomial.legendre.leggauss(nq)           # FastGaussQuadrature.gausslegendre(nq), background.jl:105
    pts = np.ascontiguousarray(pts); wts = np.ascontiguousarray(wts)
    tabs = np.zeros((ncos, abi.NTABLES, n_x + 2)); sc = np.zeros((ncos, abi.NSCALARS)); st = np.zeros(ncos, dtype=np.int32)
    rc = lib().bolt_hostgen_batch(int(device), abi.ptr(P), ncos, float(x0), float(dx), int(n_x), abi.ptr(pts), abi.ptr(wts), int(nq),
                                  abi.ptr(tabs), abi.ptr(sc), abi.ptr(st, abi.c_int32_p))
    if rc != 0:
        raise BoltError(f"bolt_hostgen_batch failed with {rc}: {lib().bolt_hostgen_last_error().decode()}")
    return [abi.HostCosmo(sc[i][:, None], pts, wts, tabs[i][:, :, None], x0, dx) for i in range(ncos)], st
