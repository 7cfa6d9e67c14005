"""Development (GPU): one K1 launch of the selected kernel (BOLT_K1_PIPE / BOLT_K1_WARP; default: the library's dispatch) for ncu captures."""
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
nk = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 1
k = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, 2000)[-nk:] if nk < 2000 else B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, nk)
for rep in range(nrep):
    out = dc.solve(k, o, want=("S_T", "S_P"))
print("nk", nk, "K1 ms", ctx.timing()["hierarchy_ms"], "steps", out["nsteps"].sum(), "rej", out["nreject"].sum())
