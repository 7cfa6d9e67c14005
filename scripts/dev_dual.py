"""Development check (GPU): K1 with partials vs (a) the value-only kernel, (b) central finite differences."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bolt_b200 as B
import hostgen as HG
from bolt_b200 import abi, capi

def host(par):
    bg = HG.Background(par)
    ih = HG.IonizationHistory(HG.RECFAST(bg, OmegaB=par.Ω_b, Yp=par.Y_p, OmegaG=par.Ω_r), par, bg)
    return abi.HostCosmo.from_host(par, bg, ih), bg

par = B.CosmoParams()
names = ["Ω_b", "h"]; rel = 1e-3
base, bg = host(par)
pm, steps = [], []
for nm in names:
    d = rel * getattr(par, nm)
    pm.append((host(par.replace(**{nm: getattr(par, nm) + d}))[0], host(par.replace(**{nm: getattr(par, nm) - d}))[0])); steps.append(d)
dual = abi.HostCosmo.with_partials(base, pm, steps)
ctx = capi.Context(0)
ks = np.array([2.0, 40.0, 300.0]) * bg.H0       # fixed k (the reference strips partials from k grids)
o = abi.make_opts(8, 8, 10, fixed_dt=0.01)
t = time.time(); g = capi.DeviceCosmo(ctx, dual).solve(ks, o, want=("S_T", "S_P", "u_final")); print("dual solve %.2fs" % (time.time() - t), g["status"], ctx.timing())
v = capi.DeviceCosmo(ctx, base).solve(ks, o, want=("S_T", "S_P", "u_final"))
for key in ("S_T", "S_P", "u_final"):
    a, b = g[key][..., 0], v[key]
    if key == "S_P": a, b = a[:, :-1], b[:, :-1]
    print(key, "value vs value-only kernel: max rel diff %.2e" % (np.abs(a - b).max() / np.abs(b).max()))
for j, nm in enumerate(names):
    vp = capi.DeviceCosmo(ctx, pm[j][0]).solve(ks, o, want=("S_T", "S_P", "u_final"))
    vm = capi.DeviceCosmo(ctx, pm[j][1]).solve(ks, o, want=("S_T", "S_P", "u_final"))
    for key in ("S_T", "S_P", "u_final"):
        fd = (vp[key] - vm[key]) / (2 * steps[j]); ad = g[key][..., 1 + j]
        if key == "S_P": fd, ad = fd[:, :-1], ad[:, :-1]
        for i in range(len(ks)):
            sc = np.abs(fd[i]).max()
            print("  d%s/d%s k/H0=%g: max |AD-FD|/max|FD| = %.2e" % (key, nm, ks[i] / bg.H0, np.abs(ad[i] - fd[i]).max() / sc))
