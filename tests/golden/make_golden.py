"""Regenerates tests/golden/*.npz.  Run in the build container (needs /root/reference/test/data):

    python tests/golden/make_golden.py

1. Reference fixtures, subsampled so they stay small (provenance: xzackli/Bolt.jl test/data):
   recfast_xe.npz      test_recfast_1.dat (Fortran RECFAST z, Xe)          -> test/runtests.jl:38-48
   class_px.npz        zack_N_class_px_k{p03,p1}_nofluid_nonu.dat (x, phi, d_b) -> test/runtests.jl:83-147
   camb_cl.npz         camb_rough_ttteee_unlensed.dat at l = 10:10:2500   -> test/runtests.jl:149-185
2. Oracle outputs for the BASELINE C1 configuration (default CosmoParams, 100 quadratic k-modes, l_gamma = 8,
   reltol 1e-11): source grids on the LOS rows and C_l -- the vectors the GPU tests compare against without
   re-running the slow CPU solve.
"""
import os, sys, time
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = "/root/reference/test/data"


def reference_fixtures():
    d = np.loadtxt(f"{REF}/test_recfast_1.dat", delimiter=",", skiprows=1, usecols=(0, 1))
    d = d[:-5][::7]                                   # runtests.jl:41 drops the last 5 rows
    np.savez_compressed(f"{HERE}/recfast_xe.npz", z=d[:, 0], Xe=d[:, 1])
    out = {}
    for tag in ("p001", "p01", "p03", "p1", "p3", "p5", "1p0"):     # only p03 is used by a reference test; the rest are extra pins
        c = np.loadtxt(f"{REF}/zack_N_class_px_k{tag}_nofluid_nonu.dat")
        sel = np.arange(0, c.shape[1], 4 if tag in ("p03", "p1") else 8)
        out[f"x_{tag}"] = c[0, sel]; out[f"k_{tag}"] = c[1, 0]
        out[f"d_b_{tag}"] = c[3, sel]; out[f"phi_{tag}"] = c[7, sel]
    np.savez_compressed(f"{HERE}/class_px.npz", **out)
    # massive-neutrino run with reionization (16 rows: ..., d_ncdm[0] at row 6, phi at row 8); consumed only by
    # scripts/plot_perts_x.jl in the reference: an extra pin for the massive-neutrino hierarchy and rho_sigma
    c = np.loadtxt(f"{REF}/class_px_kp03_nofluid_re.dat")
    sel = np.arange(0, c.shape[1], 6)
    np.savez_compressed(f"{HERE}/class_px_mnu.npz", x=c[0, sel], k=c[1, 0], d_b=c[3, sel], d_cdm=c[4, sel], d_ncdm=c[6, sel], phi=c[8, sel])
    pk = np.loadtxt(f"{REF}/camb_pk_x0.dat")                  # k [h/Mpc], P(k) [(Mpc/h)^3] at z = 0; no reference test consumes it
    sel = np.arange(0, pk.shape[1], 20)
    np.savez_compressed(f"{HERE}/camb_pk.npz", k_h=pk[0, sel], pk_h3=pk[1, sel])
    ff = np.loadtxt(f"{REF}/fftlog_example.txt")              # FFTLog known-answer vector (pyfftlog), test/runtests.jl:11-35
    np.savez_compressed(f"{HERE}/fftlog_example.npz", k=ff[:, 0], f=ff[:, 1])
    camb = np.loadtxt(f"{REF}/camb_rough_ttteee_unlensed.dat")
    ells = np.arange(10, 2501, 10)
    np.savez_compressed(f"{HERE}/camb_cl.npz", ell=ells, tt=np.interp(ells, camb[0], camb[1]),
                        te=np.interp(ells, camb[0], camb[2]), ee=np.interp(ells, camb[0], camb[3]))


def oracle_vectors():
    import bolt_b200 as B
    import hostgen as HG
    from bolt_b200 import abi
    from oracle.oracle import OracleCosmo
    par = B.CosmoParams(); bg = HG.Background(par)
    ih = HG.IonizationHistory(HG.RECFAST(bg, OmegaB=par.Ω_b, Yp=par.Y_p, OmegaG=par.Ω_r), par, bg)
    hc = abi.HostCosmo.from_host(par, bg, ih); oc = OracleCosmo(hc)
    kg = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, 100)
    o = abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6)
    t = time.time(); out = oc.solve(kg, o, want=("S_T", "S_P")); print("oracle solve %.1fs" % (time.time() - t))
    ix_start = int(np.argmax(bg.x_grid > -8))
    ells = np.arange(10, 2501, 10)
    tt, te, ee = oc.project(out["S_T"], out["S_P"], kg, ells, 0.01 * bg.H0, 1000 * bg.H0, 5000, ix_start)
    np.savez_compressed(f"{HERE}/oracle_c1.npz", k=kg, ix_start=ix_start, S_T=out["S_T"][:, ix_start:], S_P=out["S_P"][:, ix_start:],
                        nsteps=out["nsteps"], nreject=out["nreject"], ell=ells, tt=tt, te=te, ee=ee,
                        tables=hc.tables, scalars=hc.scalars, quad_pts=hc.quad_pts, quad_wts=hc.quad_wts, x0=hc.x0, dx=hc.dx)


# ------------------------------------------------------------------------------------------------------------------
# 3. Oracle outputs for the BENCHMARKED configurations (bench.py's synthetic cosmology #0), value and gradient:
#    oracle_c3.npz            C3 value: all 2000 quadratic k-modes, reltol 1e-11, C_l at every 25th multipole
#    oracle_c3grad.npz        C3 with 6 forward-mode partials (nd = 7: the K1 NP=4 / K2 NP=6 kernels the bench times), full size
#    oracle_c3grad_small.npz  the same on a 200-mode k grid (minutes instead of an hour of CPU)
#    oracle_c2.npz            plin, 500 log10_k modes, n = 473, reltol 1e-5; partials on every 10th mode
#    oracle_c4mini.npz        l_gamma = 50 (n = 281) adaptive, 128 quadratic k-modes
#    The inputs (host tables WITH partials) are stored in the files: the GPU tests feed bit-identical tables to the device.
# ------------------------------------------------------------------------------------------------------------------
GRAD_NAMES = ["Ω_b", "Ω_c", "h", "n", "A", "Σm_ν"]      # bench.py gradient workload (tau is not a reference parameter)


def bench_cosmology():
    """bench.py's synthetic cosmology #0 with partials (host generator + central differences, api.host_cosmo_with_partials)."""
    import bench
    import bolt_b200 as B
    import hostgen as HG
    from hostgen import host_cosmo_with_partials
    par = bench.synthetic_params(0)
    dual, base, bg, ih, pm, steps = host_cosmo_with_partials(par, GRAD_NAMES, rel_step=1e-3)
    return par, bg, dual, base


def _inputs(hc):
    return dict(tables=hc.tables, scalars=hc.scalars, quad_pts=hc.quad_pts, quad_wts=hc.quad_wts, x0=hc.x0, dx=hc.dx)


def c3_value(par, bg, dual, base):
    import bolt_b200 as B
    import hostgen as HG
    from bolt_b200 import abi
    from oracle.oracle import OracleCosmo
    oc = OracleCosmo(base)
    kg = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, 2000)
    ix0 = int(np.argmax(bg.x_grid > -8))
    o = abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6, ix_first=ix0)
    t = time.time(); out = oc.solve(kg, o, want=("S_T", "S_P")); print("c3 value solve %.0fs" % (time.time() - t), flush=True)
    ells = np.unique(np.r_[np.arange(2, 2501, 25), 2500]).astype(np.int32)
    tt, te, ee = oc.project(out["S_T"], out["S_P"], kg, ells, 0.01 * bg.H0, 1000 * bg.H0, 5000, ix0)
    sel = np.arange(49, 2000, 100)
    np.savez_compressed(f"{HERE}/oracle_c3.npz", k=kg, ix_start=ix0, ell=ells, tt=tt, te=te, ee=ee, nsteps=out["nsteps"], nreject=out["nreject"],
                        status=out["status"], sel=sel, S_T=out["S_T"][sel][:, ix0:], S_P=out["S_P"][sel][:, ix0:], **_inputs(base))


def c3_grad(par, bg, dual, base, nk, name):
    import bolt_b200 as B
    import hostgen as HG
    from bolt_b200 import abi
    from oracle.oracle import OracleCosmo
    od = OracleCosmo(dual)
    kg = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, nk)
    ix0 = int(np.argmax(bg.x_grid > -8))
    o = abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6, ix_first=ix0)
    t = time.time(); out = od.solve_sens(kg, o, want=("S_T", "S_P")); print("%s sens solve %.0fs" % (name, time.time() - t), flush=True)
    ells = np.unique(np.r_[np.arange(2, 2501, 25 if nk >= 2000 else 100), 2500]).astype(np.int32)
    xmax = 1000 * bg.H0 * bg.η0
    tt, te, ee = od.project_sens(out["S_T"], out["S_P"], kg, ells, 0.01 * bg.H0, 1000 * bg.H0, 5000, ix0, xmax)
    sel = np.arange(nk // 40, nk, nk // 20)
    np.savez_compressed(f"{HERE}/{name}.npz", k=kg, ix_start=ix0, ell=ells, tt=tt, te=te, ee=ee, nsteps=out["nsteps"], nreject=out["nreject"],
                        status=out["status"], sel=sel, S_T=out["S_T"][sel][:, ix0:], S_P=out["S_P"][sel][:, ix0:], bessel_xmax=xmax,
                        names=np.array(GRAD_NAMES), **_inputs(dual))


def c2_plin(par, bg, dual, base):
    import bolt_b200 as B
    import hostgen as HG
    from bolt_b200 import abi
    from oracle.oracle import OracleCosmo
    oc = OracleCosmo(base); od = OracleCosmo(dual)
    ks = B.log10_k(10 * bg.H0, 5000 * bg.H0, 500)
    o = abi.make_opts(50, 50, 20, reltol=1e-5, abstol=1e-6)
    t = time.time(); pk, st, ns = oc.plin(ks, o); print("c2 plin %.0fs" % (time.time() - t), flush=True)
    gsel = np.arange(5, 500, 10)
    t = time.time(); pkg, stg, nsg = od.plin_sens(ks[gsel], o); print("c2 plin sens %.0fs" % (time.time() - t), flush=True)
    np.savez_compressed(f"{HERE}/oracle_c2.npz", k=ks, pk=pk, status=st, nsteps=ns, gsel=gsel, pk_grad=pkg, nsteps_grad=nsg, status_grad=stg,
                        names=np.array(GRAD_NAMES), **_inputs(dual))


def c4_mini(par, bg, dual, base):
    import bolt_b200 as B
    import hostgen as HG
    from bolt_b200 import abi
    from oracle.oracle import OracleCosmo
    oc = OracleCosmo(base)
    kg = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, 128)
    ix0 = int(np.argmax(bg.x_grid > -8))
    o = abi.make_opts(50, 8, 10, reltol=1e-11, abstol=1e-6, ix_first=ix0)
    t = time.time(); out = oc.solve(kg, o, want=("S_T", "S_P")); print("c4mini solve %.0fs" % (time.time() - t), flush=True)
    ells = np.array([2, 10, 30, 100, 220, 400, 540, 800, 1000, 1500, 2000, 2500], dtype=np.int32)
    tt, te, ee = oc.project(out["S_T"], out["S_P"], kg, ells, 0.01 * bg.H0, 1000 * bg.H0, 5000, ix0)
    sel = np.arange(3, 128, 8)
    np.savez_compressed(f"{HERE}/oracle_c4mini.npz", k=kg, ix_start=ix0, ell=ells, tt=tt, te=te, ee=ee, nsteps=out["nsteps"], nreject=out["nreject"],
                        status=out["status"], sel=sel, S_T=out["S_T"][sel][:, ix0:], S_P=out["S_P"][sel][:, ix0:], **_inputs(base))


def bench_goldens(which):
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    cos = bench_cosmology()
    if "c4mini" in which: c4_mini(*cos)
    if "c3grad_small" in which: c3_grad(*cos, 200, "oracle_c3grad_small")
    if "c3" in which: c3_value(*cos)
    if "c2" in which: c2_plin(*cos)
    if "c3grad" in which: c3_grad(*cos, 2000, "oracle_c3grad")


if __name__ == "__main__":
    stages = [a for a in sys.argv[1:] if not a.startswith("--")]
    if stages:                       # e.g. `make_golden.py c3 c2 c4mini c3grad_small c3grad` (the last one is ~1 h of CPU)
        bench_goldens(stages)
    else:
        reference_fixtures()
        if "--fixtures-only" not in sys.argv:
            oracle_vectors()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
