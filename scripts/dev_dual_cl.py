import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bolt_b200 as B
from bolt_b200 import abi, capi
from hostgen import host_cosmo_with_partials
par = B.CosmoParams(); ctx = capi.Context(0)
for rel in (1e-5, 1e-4, 1e-3):
    dual, base, bg, ih, pm, steps = host_cosmo_with_partials(par, ["Ω_b"], rel_step=rel)
    ks = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, 60)
    ells = np.array([2, 10, 30, 100, 220, 400, 650, 1000, 1500, 2000], dtype=np.int32)
    o = abi.make_opts(8, 8, 10, fixed_dt=0.01)
    args = (ks, o, ells, 0.01 * bg.H0, 1000 * bg.H0, 5000, 1201)
    ad = capi.DeviceCosmo(ctx, dual).spectra(*args)
    p = capi.DeviceCosmo(ctx, pm[0][0]).spectra(*args); m = capi.DeviceCosmo(ctx, pm[0][1]).spectra(*args)
    # K2 in isolation: project the SAME dual source grids
    g = capi.DeviceCosmo(ctx, dual).solve(ks, o, want=("S_T", "S_P"))
    gp = capi.DeviceCosmo(ctx, pm[0][0]).solve(ks, o, want=("S_T", "S_P")); gm = capi.DeviceCosmo(ctx, pm[0][1]).solve(ks, o, want=("S_T", "S_P"))
    print("rel", rel)
    for nm, i in (("tt", 0), ("te", 1), ("ee", 2)):
        fd = (p[i] - m[i]) / (2 * steps[0]); a = ad[i][:, 1]
        print(" ", nm, "AD/FD-1:", np.array2string(a / fd - 1, precision=2), " dlnC/dlnp:", np.array2string(fd * par.Ω_b / ad[i][:, 0], precision=2))
    for key in ("S_T", "S_P"):
        fd = (gp[key] - gm[key]) / (2 * steps[0]); a = g[key][..., 1]
        e = np.abs(a - fd)[:, 1201:1995].max(axis=1) / np.abs(fd)[:, 1201:1995].max(axis=1)
        print(" ", key, "rows 1201:1995 max err per k: max %.2e at k idx %d" % (e.max(), e.argmax()))
