"""ctypes mirror of include/bolt_cuda.h (structs and enums only; no library is loaded here)."""
import ctypes as C
import numpy as np

BOLT_ABI_VERSION = 3

SCALARS = ["h", "Ω_r", "Ω_b", "Ω_c", "A", "n", "Y_p", "N_ν", "Σm_ν", "H0", "η0", "ρ_crit", "Ω_Λ"]
TABLES = ["H", "Hp", "Hpp", "η", "ρ0M", "τ", "τp", "τpp", "g", "gp", "gpp", "csb2"]
S = {name: i for i, name in enumerate(SCALARS)}
T = {name: i for i, name in enumerate(TABLES)}
NSCALARS, NTABLES = len(SCALARS), len(TABLES)

MODE_ADAPTIVE, MODE_FIXED = 0, 1
K_OK, K_MAXSTEPS, K_DT_UNDERFLOW, K_NONFINITE, K_RSA_TRIGGERED = 0, 1, 2, 3, 4

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class CosmoDesc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("nd", C.c_int32), ("n_x", C.c_int32), ("nq", C.c_int32),
                ("x0", C.c_double), ("dx", C.c_double),
                ("scalars", c_double_p), ("quad_pts", c_double_p), ("quad_wts", c_double_p),
                ("tables", c_double_p)]


class Opts(C.Structure):
    _fields_ = [("l_gamma", C.c_int32), ("l_nu", C.c_int32), ("l_mnu", C.c_int32), ("mode", C.c_int32),
                ("reltol", C.c_double), ("abstol", C.c_double), ("fixed_dt", C.c_double),
                ("max_steps", C.c_int64), ("ix_first", C.c_int32), ("reserved", C.c_int32)]


def make_opts(l_gamma=8, l_nu=8, l_mnu=10, reltol=1e-6, abstol=1e-6, fixed_dt=0.0, max_steps=0, ix_first=0):
    mode = MODE_FIXED if fixed_dt > 0 else MODE_ADAPTIVE
    return Opts(l_gamma, l_nu, l_mnu, mode, reltol, abstol, fixed_dt, max_steps, ix_first, 0)


def state_dim(l_gamma, l_nu, l_mnu, nq):
    return 2 * (l_gamma + 1) + (l_nu + 1) + (l_mnu + 1) * nq + 5


def ptr(a, typ=c_double_p):
    return None if a is None else a.ctypes.data_as(typ)


class HostCosmo:
    """The host-side arrays a bolt_cosmo_desc points at (kept alive here)."""

    def __init__(self, scalars, quad_pts, quad_wts, tables, x0, dx):
        self.scalars = np.ascontiguousarray(scalars, dtype=np.float64)     # [NSCALARS][nd]
        self.tables = np.ascontiguousarray(tables, dtype=np.float64)       # [NTABLES][n_x+2][nd]
        self.quad_pts = np.ascontiguousarray(quad_pts, dtype=np.float64)
        self.quad_wts = np.ascontiguousarray(quad_wts, dtype=np.float64)
        assert self.scalars.ndim == 2 and self.scalars.shape[0] == NSCALARS
        assert self.tables.ndim == 3 and self.tables.shape[0] == NTABLES
        self.nd = self.scalars.shape[1]
        assert self.tables.shape[2] == self.nd
        self.n_x = self.tables.shape[1] - 2
        self.nq = self.quad_pts.shape[0]
        self.x0, self.dx = float(x0), float(dx)
        self.desc = CosmoDesc(BOLT_ABI_VERSION, self.nd, self.n_x, self.nq, self.x0, self.dx,
                              ptr(self.scalars), ptr(self.quad_pts), ptr(self.quad_wts), ptr(self.tables))

    def scalar(self, name):
        return self.scalars[S[name], 0]

    def padded(self, nd):
        """The same cosmology with all-zero partials appended up to `nd` components (the library instantiates its kernels for
        1, 2, 3, 4 and 6 partials; 5 is served as 6 with a zero partial whose results are dropped by the binding)."""
        assert nd >= self.nd
        sc = np.zeros((NSCALARS, nd)); tb = np.zeros(self.tables.shape[:2] + (nd,))
        sc[:, :self.nd] = self.scalars; tb[:, :, :self.nd] = self.tables
        return HostCosmo(sc, self.quad_pts, self.quad_wts, tb, self.x0, self.dx)

    @staticmethod
    def from_host(par, bg, ih):
        """Pack Background + IonizationHistory (value only, nd = 1)."""
        sc = np.array([[par.h, par.Ω_r, par.Ω_b, par.Ω_c, par.A, par.n, par.Y_p, par.N_ν, par.Σm_ν,
                        bg.H0, bg.η0, bg.ρ_crit, bg.Ω_Λ]], dtype=np.float64).T
        tabs = [bg.H, bg.Hp, bg.Hpp, bg.η, bg.ρ0M, ih.τ, ih.τp, ih.τpp, ih.g, ih.gp, ih.gpp, ih.csb2]
        tb = np.stack([t.coefs for t in tabs])[:, :, None]
        return HostCosmo(sc, bg.quad_pts, bg.quad_wts, tb, bg.x0, bg.dx)

    @staticmethod
    def with_partials(base, plus_minus, steps):
        """Attach partials by central differences of whole host pipelines.
        plus_minus: list of (HostCosmo_plus, HostCosmo_minus); steps: list of Δ."""
        nd = 1 + len(plus_minus)
        sc = np.zeros((NSCALARS, nd)); tb = np.zeros(base.tables.shape[:2] + (nd,))
        sc[:, 0] = base.scalars[:, 0]; tb[:, :, 0] = base.tables[:, :, 0]
        for j, ((p, m), d) in enumerate(zip(plus_minus, steps)):
            sc[:, 1 + j] = (p.scalars[:, 0] - m.scalars[:, 0]) / (2 * d)
            tb[:, :, 1 + j] = (p.tables[:, :, 0] - m.tables[:, :, 0]) / (2 * d)
        return HostCosmo(sc, base.quad_pts, base.quad_wts, tb, base.x0, base.dx)
