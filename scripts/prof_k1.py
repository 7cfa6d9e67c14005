"""Profiling driver: one adaptive K1 launch (+ K2) on a configurable number of k-modes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bolt_b200 as B
import hostgen as HG
from bolt_b200 import abi, capi
nk = int(sys.argv[1]) if len(sys.argv) > 1 else 296
nell = int(sys.argv[2]) if len(sys.argv) > 2 else 0
par = B.CosmoParams(); bg = HG.Background(par)
ih = HG.IonizationHistory(HG.RECFAST(bg, OmegaB=par.Ω_b, Yp=par.Y_p, OmegaG=par.Ω_r), par, bg)
ctx = capi.Context(0); dc = capi.DeviceCosmo(ctx, abi.HostCosmo.from_host(par, bg, ih))
k = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, nk)
o = abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6, ix_first=1201)
for rep in range(2):
    if nell:
        ells = np.unique(np.linspace(2, 2500, nell).astype(np.int32))
        out = dc.spectra(k, o, ells, 0.01 * bg.H0, 1000 * bg.H0, 5000, 1201)
        print(ctx.timing(), out[4].sum())
    else:
        out = dc.solve(k, o, want=("S_T", "S_P"))
        print(ctx.timing(), out["nsteps"].sum(), out["nsteps"].max(), out["nreject"].sum())
