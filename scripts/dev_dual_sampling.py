"""Development (GPU): how much of K1 with partials (CTA per mode) is the sampling of the source functions on warp 0."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bolt_b200 as B
from bolt_b200 import abi, capi
from hostgen import host_cosmo_with_partials
names = ["Ω_b", "Ω_c", "h", "Σm_ν"]
par = B.CosmoParams()
dual, base, bg, ih, pm, steps = host_cosmo_with_partials(par, names, rel_step=1e-3)
ctx = capi.Context(0); dc = capi.DeviceCosmo(ctx, dual)
nk = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
k = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, nk)
o = abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6, ix_first=1201)
for want in (("S_T", "S_P"), ("u_final",), ("S_T", "S_P")):
    out = dc.solve(k, o, want=want)
    print(want, "K1 %.1f ms steps %d bad %d" % (ctx.timing()["hierarchy_ms"], out["nsteps"].sum(), (out["status"] != 0).sum()), flush=True)
