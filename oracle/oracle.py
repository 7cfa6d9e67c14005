"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY -- see bolt_oracle.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libbolt_oracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        from bolt_b200 import abi
        try:
            L = C.CDLL(build())
        except OSError:
            L = C.CDLL(build(force=True))
        dp, ip, lp = abi.c_double_p, abi.c_int32_p, abi.c_int64_p
        L.oracle_cosmo_create.restype = C.c_void_p
        L.oracle_cosmo_create.argtypes = [C.POINTER(abi.CosmoDesc)]
        L.oracle_cosmo_free.argtypes = [C.c_void_p]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_initial_conditions.argtypes = [C.c_void_p, C.c_double, C.POINTER(abi.Opts), dp]
        L.oracle_hierarchy.argtypes = [C.c_void_p, C.c_double, C.POINTER(abi.Opts), C.c_double, dp, dp]
        L.oracle_hierarchy.restype = C.c_int
        L.oracle_source_functions.argtypes = [C.c_void_p, C.c_double, C.POINTER(abi.Opts), C.c_double, dp, dp, dp, dp]
        L.oracle_spline_eval.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.oracle_spline_eval.restype = C.c_double
        L.oracle_sph_bessel_j.argtypes = [C.c_int, C.c_double]
        L.oracle_sph_bessel_j.restype = C.c_double
        L.oracle_solve.argtypes = [C.c_void_p, dp, C.c_int, C.POINTER(abi.Opts), C.c_int, dp, dp, dp, dp, ip, lp, lp]
        L.oracle_project.argtypes = [C.c_void_p, dp, dp, dp, C.c_int, ip, C.c_int, C.c_double, C.c_double,
                                     C.c_int, C.c_int, dp, dp, dp]
        L.oracle_plin.argtypes = [C.c_void_p, dp, C.c_int, C.POINTER(abi.Opts), C.c_int, dp, ip, lp]
        L.oracle_set_num_threads.argtypes = [C.c_int]
        L.oracle_solve_sens.argtypes = [C.c_void_p, dp, C.c_int, C.POINTER(abi.Opts), C.c_int, dp, dp, dp, ip, lp, lp]
        L.oracle_project_sens.argtypes = [C.c_void_p, dp, dp, dp, C.c_int, ip, C.c_int, C.c_double, C.c_double,
                                          C.c_int, C.c_int, C.c_double, dp, dp, dp]
        L.oracle_plin_sens.argtypes = [C.c_void_p, dp, C.c_int, C.POINTER(abi.Opts), C.c_int, dp, ip, lp]
        _LIB = L
    return _LIB


class OracleCosmo:
    def __init__(self, host_cosmo):
        self.hc = host_cosmo
        self.h = lib().oracle_cosmo_create(C.byref(host_cosmo.desc))

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_cosmo_free(self.h)
            self.h = None

    def solve(self, k, opts, want=("S_T", "S_P"), lu_mode=1):
        from bolt_b200 import abi
        k = np.ascontiguousarray(k, dtype=np.float64)
        nk, n_x = len(k), self.hc.n_x
        n = abi.state_dim(opts.l_gamma, opts.l_nu, opts.l_mnu, self.hc.nq)
        out = {}
        out["S_T"] = np.zeros((nk, n_x)) if "S_T" in want else None
        out["S_P"] = np.zeros((nk, n_x)) if "S_P" in want else None
        out["u_hist"] = np.zeros((nk, n_x, n)) if "u_hist" in want else None
        out["u_final"] = np.zeros((nk, n)) if "u_final" in want else None
        out["status"] = np.zeros(nk, dtype=np.int32)
        out["nsteps"] = np.zeros(nk, dtype=np.int64)
        out["nreject"] = np.zeros(nk, dtype=np.int64)
        lib().oracle_solve(self.h, abi.ptr(k), nk, C.byref(opts), lu_mode, abi.ptr(out["S_T"]), abi.ptr(out["S_P"]),
                           abi.ptr(out["u_hist"]), abi.ptr(out["u_final"]), abi.ptr(out["status"], abi.c_int32_p),
                           abi.ptr(out["nsteps"], abi.c_int64_p), abi.ptr(out["nreject"], abi.c_int64_p))
        return out

    def project(self, S_T, S_P, k, ells, kd_min, kd_max, n_kd, ix_start):
        from bolt_b200 import abi
        k = np.ascontiguousarray(k, dtype=np.float64)
        ells = np.ascontiguousarray(ells, dtype=np.int32)
        tt = np.zeros(len(ells)) if S_T is not None else None
        ee = np.zeros(len(ells)) if S_P is not None else None
        te = np.zeros(len(ells)) if (S_T is not None and S_P is not None) else None
        lib().oracle_project(self.h, abi.ptr(S_T), abi.ptr(S_P), abi.ptr(k), len(k), abi.ptr(ells, abi.c_int32_p),
                             len(ells), kd_min, kd_max, n_kd, ix_start, abi.ptr(tt), abi.ptr(te), abi.ptr(ee))
        return tt, te, ee

    def plin(self, k, opts, lu_mode=1):
        from bolt_b200 import abi
        k = np.ascontiguousarray(k, dtype=np.float64)
        pk = np.zeros(len(k)); st = np.zeros(len(k), dtype=np.int32); ns = np.zeros(len(k), dtype=np.int64)
        lib().oracle_plin(self.h, abi.ptr(k), len(k), C.byref(opts), lu_mode, abi.ptr(pk), abi.ptr(st, abi.c_int32_p),
                          abi.ptr(ns, abi.c_int64_p))
        return pk, st, ns

    # ---- gradient oracle (host cosmology with partials, nd > 1): forward sensitivities through the same stepper ----
    def solve_sens(self, k, opts, want=("S_T", "S_P"), lu_mode=1):
        from bolt_b200 import abi
        assert self.hc.nd > 1
        k = np.ascontiguousarray(k, dtype=np.float64)
        nk, n_x, nd = len(k), self.hc.n_x, self.hc.nd
        n = abi.state_dim(opts.l_gamma, opts.l_nu, opts.l_mnu, self.hc.nq)
        out = {}
        out["S_T"] = np.zeros((nk, n_x, nd)) if "S_T" in want else None
        out["S_P"] = np.zeros((nk, n_x, nd)) if "S_P" in want else None
        out["u_final"] = np.zeros((nk, n, nd)) if "u_final" in want else None
        out["status"] = np.zeros(nk, dtype=np.int32)
        out["nsteps"] = np.zeros(nk, dtype=np.int64)
        out["nreject"] = np.zeros(nk, dtype=np.int64)
        rc = lib().oracle_solve_sens(self.h, abi.ptr(k), nk, C.byref(opts), lu_mode, abi.ptr(out["S_T"]), abi.ptr(out["S_P"]),
                                     abi.ptr(out["u_final"]), abi.ptr(out["status"], abi.c_int32_p),
                                     abi.ptr(out["nsteps"], abi.c_int64_p), abi.ptr(out["nreject"], abi.c_int64_p))
        assert rc == 0
        return out

    def project_sens(self, S_T, S_P, k, ells, kd_min, kd_max, n_kd, ix_start, xmax_fixed=0.0):
        from bolt_b200 import abi
        k = np.ascontiguousarray(k, dtype=np.float64)
        ells = np.ascontiguousarray(ells, dtype=np.int32)
        nd = self.hc.nd
        S_T = None if S_T is None else np.ascontiguousarray(S_T, dtype=np.float64)
        S_P = None if S_P is None else np.ascontiguousarray(S_P, dtype=np.float64)
        tt = np.zeros((len(ells), nd)) if S_T is not None else None
        ee = np.zeros((len(ells), nd)) if S_P is not None else None
        te = np.zeros((len(ells), nd)) if (S_T is not None and S_P is not None) else None
        rc = lib().oracle_project_sens(self.h, abi.ptr(S_T), abi.ptr(S_P), abi.ptr(k), len(k), abi.ptr(ells, abi.c_int32_p),
                                       len(ells), kd_min, kd_max, n_kd, ix_start, xmax_fixed, abi.ptr(tt), abi.ptr(te), abi.ptr(ee))
        assert rc == 0
        return tt, te, ee

    def plin_sens(self, k, opts, lu_mode=1):
        from bolt_b200 import abi
        k = np.ascontiguousarray(k, dtype=np.float64)
        pk = np.zeros((len(k), self.hc.nd)); st = np.zeros(len(k), dtype=np.int32); ns = np.zeros(len(k), dtype=np.int64)
        rc = lib().oracle_plin_sens(self.h, abi.ptr(k), len(k), C.byref(opts), lu_mode, abi.ptr(pk), abi.ptr(st, abi.c_int32_p),
                                    abi.ptr(ns, abi.c_int64_p))
        assert rc == 0
        return pk, st, ns

    def initial_conditions(self, k, opts):
        from bolt_b200 import abi
        n = abi.state_dim(opts.l_gamma, opts.l_nu, opts.l_mnu, self.hc.nq)
        u = np.zeros(n)
        lib().oracle_initial_conditions(self.h, k, C.byref(opts), abi.ptr(u))
        return u

    def hierarchy(self, k, opts, x, u):
        from bolt_b200 import abi
        u = np.array(u, dtype=np.float64); du = np.zeros_like(u)
        rsa = lib().oracle_hierarchy(self.h, k, C.byref(opts), x, abi.ptr(u), abi.ptr(du))
        return du, u, bool(rsa)
