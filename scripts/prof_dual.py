"""Profiling driver: K1 with partials (register-resident dual kernel) on a configurable number of k-modes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bolt_b200 as B
from bolt_b200 import abi, capi
from hostgen import host_cosmo_with_partials
nk = int(sys.argv[1]) if len(sys.argv) > 1 else 296
names = ["Ω_b", "Ω_c", "h", "Σm_ν"]
par = B.CosmoParams()
dual, base, bg, ih, pm, steps = host_cosmo_with_partials(par, names, rel_step=1e-3)
ctx = capi.Context(0); dc = capi.DeviceCosmo(ctx, dual)
k = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, nk)
o = abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6, ix_first=1201)
for rep in range(2):
    out = dc.solve(k, o, want=("S_T", "S_P"))
    print(ctx.timing()["hierarchy_ms"], out["nsteps"].sum(), out["nsteps"].max(), out["nreject"].sum())
