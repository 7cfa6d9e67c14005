"""Development (GPU): one C3-value spectra call (for ncu captures of K2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bolt_b200 as B
import hostgen as HG
from bolt_b200 import abi, capi
par = B.CosmoParams(); bg = HG.Background(par)
ih = HG.IonizationHistory(HG.RECFAST(bg, OmegaB=par.Ω_b, Yp=par.Y_p, OmegaG=par.Ω_r), par, bg)
ctx = capi.Context(0); dc = capi.DeviceCosmo(ctx, abi.HostCosmo.from_host(par, bg, ih))
k = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, 2000)
o = abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6)
ells = np.arange(2, 2501, dtype=np.int32); ix0 = int(np.argmax(bg.x_grid > -8))
for rep in range(2):
    r = dc.spectra(k, o, ells, 0.01 * bg.H0, 1000 * bg.H0, 5000, ix0)
    print(ctx.timing())
