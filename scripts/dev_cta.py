"""Development check (GPU): the pipelined CTA-per-mode K1 kernel (DEV_NEW=pipe, hierarchy_pipe.cuh; DEV_NEW=auto: the library's dispatch) against the one-warp kernel: parity + timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bolt_b200 as B
import hostgen as HG
from bolt_b200 import abi, capi

par = B.CosmoParams(); bg = HG.Background(par)
ih = HG.IonizationHistory(HG.RECFAST(bg, OmegaB=par.Ω_b, Yp=par.Y_p, OmegaG=par.Ω_r), par, bg)
ctx = capi.Context(0); dc = capi.DeviceCosmo(ctx, abi.HostCosmo.from_host(par, bg, ih))


def run(which, ks, o, want):
    for nm in ("WARP", "CTA", "PIPE"):
        os.environ.pop("BOLT_K1_" + nm, None)
    if which != "auto":          # auto: the library's own dispatch (pipe up to 9 x SMs modes, both kernels side by side beyond)
        os.environ["BOLT_K1_" + which.upper()] = "1"
    out = dc.solve(ks, o, want=want)
    return out, ctx.timing()["hierarchy_ms"]


def relmax(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


NEW = os.environ.get("DEV_NEW", "pipe")
sizes = [int(a) for a in sys.argv[1:]] or [296, 2000, 8000]
ks = np.array([0.5, 3.0, 30.0, 100.0, 300.0, 900.0]) * bg.H0
for lg in (8, 10):
    o = abi.make_opts(lg, 8, 10, fixed_dt=0.01)
    a, _ = run(NEW, ks, o, ("S_T", "S_P", "u_hist", "u_final")); b, _ = run("warp", ks, o, ("S_T", "S_P", "u_hist", "u_final"))
    print("fixed lg=%d status" % lg, a["status"], b["status"], "S_T %.2e S_P %.2e u_hist %.2e u_final %.2e" % (
        relmax(a["S_T"], b["S_T"]), relmax(a["S_P"][:, :-1], b["S_P"][:, :-1]), relmax(a["u_hist"], b["u_hist"]), relmax(a["u_final"], b["u_final"])), flush=True)
o = abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6, ix_first=1201)
k = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, 200)
a, _ = run(NEW, k, o, ("S_T", "S_P")); b, _ = run("warp", k, o, ("S_T", "S_P"))
print("adaptive 200: status", np.unique(a["status"]), "nsteps equal %.3f maxdiff %d nreject equal %.3f" % (
    (a["nsteps"] == b["nsteps"]).mean(), np.abs(a["nsteps"] - b["nsteps"]).max(), (a["nreject"] == b["nreject"]).mean()),
    "S_T %.2e S_P %.2e" % (relmax(a["S_T"], b["S_T"]), relmax(a["S_P"][:, :-1], b["S_P"][:, :-1])), flush=True)
for nk in sizes:
    k = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, nk)
    for which in (NEW, "warp"):
        best = 1e9
        for rep in range(3):
            out, ms = run(which, k, o, ("S_T", "S_P")); best = min(best, ms)
        print("nk=%5d %-5s K1 %.2f ms  steps %d max %d rej %d bad %d" % (nk, which, best, out["nsteps"].sum(), out["nsteps"].max(), out["nreject"].sum(),
                                                                       (out["status"] != 0).sum()), flush=True)
