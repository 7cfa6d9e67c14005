"""Development check (GPU): K1 with partials, CTA-per-mode kernel (hierarchy_dual_cta.cuh) against the one-warp kernel
(BOLT_K1_DUAL_WARP=1): every component of S_T, S_P, u_final; step counts; timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bolt_b200 as B
from bolt_b200 import abi, capi
from hostgen import host_cosmo_with_partials

names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["Ω_b", "Ω_c", "h", "Σm_ν"]
sizes = [int(a) for a in sys.argv[2:]] or [200, 2000]
par = B.CosmoParams()
dual, base, bg, ih, pm, steps = host_cosmo_with_partials(par, names, rel_step=1e-3)
ctx = capi.Context(0); dc = capi.DeviceCosmo(ctx, dual)


def run(which, ks, o, want):
    os.environ.pop("BOLT_K1_DUAL_WARP", None)
    if which == "warp":
        os.environ["BOLT_K1_DUAL_WARP"] = "1"
    t0 = time.perf_counter(); out = dc.solve(ks, o, want=want); dt = time.perf_counter() - t0
    return out, ctx.timing()["hierarchy_ms"]


def relmax(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


ks = np.array([0.5, 3.0, 30.0, 100.0, 300.0, 900.0]) * bg.H0
o = abi.make_opts(8, 8, 10, fixed_dt=0.01)
a, _ = run("cta", ks, o, ("S_T", "S_P", "u_final")); b, _ = run("warp", ks, o, ("S_T", "S_P", "u_final"))
print("fixed: status", a["status"], b["status"], flush=True)
for comp in range(1 + len(names)):
    print("  comp %d: S_T %.2e S_P %.2e u_final %.2e" % (comp, relmax(a["S_T"][..., comp], b["S_T"][..., comp]),
                                                          relmax(a["S_P"][:, :-1, comp], b["S_P"][:, :-1, comp]), relmax(a["u_final"][..., comp], b["u_final"][..., comp])), flush=True)
o = abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6, ix_first=1201)
for nk in sizes:
    k = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, nk)
    res = {}
    for which in ("cta", "warp"):
        best = 1e9
        for rep in range(2):
            out, ms = run(which, k, o, ("S_T", "S_P")); best = min(best, ms)
        res[which] = out
        print("nk=%5d %-4s K1 %.1f ms  steps %d max %d rej %d bad %d" % (nk, which, best, out["nsteps"].sum(), out["nsteps"].max(), out["nreject"].sum(),
                                                                       (out["status"] != 0).sum()), flush=True)
    a, b = res["cta"], res["warp"]
    print("   nsteps equal %.3f; S_T value %.2e partials %.2e; S_P value %.2e partials %.2e" % (
        (a["nsteps"] == b["nsteps"]).mean(), relmax(a["S_T"][..., 0], b["S_T"][..., 0]), max(relmax(a["S_T"][..., c], b["S_T"][..., c]) for c in range(1, 1 + len(names))),
        relmax(a["S_P"][:, :-1, 0], b["S_P"][:, :-1, 0]), max(relmax(a["S_P"][:, :-1, c], b["S_P"][:, :-1, c]) for c in range(1, 1 + len(names)))), flush=True)
