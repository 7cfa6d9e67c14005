"""Development (GPU): the share of source sampling in the value K1 kernels (sources on / off)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bolt_b200 as B
import hostgen as HG
from bolt_b200 import abi, capi
par = B.CosmoParams(); bg = HG.Background(par)
ih = HG.IonizationHistory(HG.RECFAST(bg, OmegaB=par.Ω_b, Yp=par.Y_p, OmegaG=par.Ω_r), par, bg)
ctx = capi.Context(0); dc = capi.DeviceCosmo(ctx, abi.HostCosmo.from_host(par, bg, ih))
o = abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6, ix_first=1201)
for nk in (296, 2000):
    k = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, nk)
    for want in (("S_T", "S_P"), ("u_final",)):
        best = 1e9
        for rep in range(3):
            out = dc.solve(k, o, want=want); best = min(best, ctx.timing()["hierarchy_ms"])
        print("nk %4d %-16s K1 %.2f ms steps %d" % (nk, "+".join(want), best, out["nsteps"].sum()), flush=True)
