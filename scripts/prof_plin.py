"""Timing driver for the long-chain truncations: plin (50,50,20), C4 (50,8,10) and the CLASS-pin solve."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bolt_b200 as B
import hostgen as HG
from bolt_b200 import abi, capi
which = sys.argv[1] if len(sys.argv) > 1 else "plin"
par = B.CosmoParams(); bg = HG.Background(par)
ih = HG.IonizationHistory(HG.RECFAST(bg, OmegaB=par.Ω_b, Yp=par.Y_p, OmegaG=par.Ω_r), par, bg)
ctx = capi.Context(0); dc = capi.DeviceCosmo(ctx, abi.HostCosmo.from_host(par, bg, ih))
for rep in range(2):
    if which == "plin":
        k = B.log10_k(10 * bg.H0, 5000 * bg.H0, 500)
        o = abi.make_opts(50, 50, 20, reltol=1e-5, abstol=1e-6)
        pk, st, ns = dc.plin(k, o)
        print("plin", ctx.timing()["hierarchy_ms"], ns.sum(), ns.max(), int((st != 0).sum()), float(np.log(pk).sum()))
    else:
        nk = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
        k = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, nk)
        o = abi.make_opts(50, 8, 10, reltol=1e-11, abstol=1e-6, ix_first=1201)
        out = dc.solve(k, o, want=("S_T", "S_P"))
        print("c4", nk, ctx.timing()["hierarchy_ms"], out["nsteps"].sum(), out["nsteps"].max(), float(np.abs(out["S_T"][:, 1201:]).sum()))
