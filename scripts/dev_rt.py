"""Development (GPU): the runtime-truncation K1 path (C4: l_gamma = 50; plin: 50/50/20) with independent warps vs lockstep blocks."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bolt_b200 as B
import hostgen as HG
from bolt_b200 import abi, capi
par = B.CosmoParams(); bg = HG.Background(par)
ih = HG.IonizationHistory(HG.RECFAST(bg, OmegaB=par.Ω_b, Yp=par.Y_p, OmegaG=par.Ω_r), par, bg)
ctx = capi.Context(0); dc = capi.DeviceCosmo(ctx, abi.HostCosmo.from_host(par, bg, ih))
for wpb in ("", "2", "4"):
    os.environ.pop("BOLT_K1_RT_WPB", None)
    if wpb: os.environ["BOLT_K1_RT_WPB"] = wpb
    for (trunc, nk, rt) in (((50, 8, 10), 4000, 1e-11), ((50, 8, 10), 10000, 1e-11), ((50, 50, 20), 500, 1e-5)):
        k = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, nk) if nk != 500 else B.log10_k(10 * bg.H0, 5000 * bg.H0, 500)
        o = abi.make_opts(*trunc, reltol=rt, abstol=1e-6, ix_first=1201)
        best = 1e9
        for rep in range(2):
            out = dc.solve(k, o, want=("S_T",)); best = min(best, ctx.timing()["hierarchy_ms"])
        print("WPB %-2s trunc %s nk %5d: K1 %.1f ms steps %d bad %d" % (wpb or "1", trunc, nk, best, out["nsteps"].sum(), (out["status"] != 0).sum()), flush=True)
