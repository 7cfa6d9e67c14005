"""Development check (GPU): K1 against the oracle, fixed-step and adaptive."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hostgen.background import CosmoParams, Background
from hostgen.recfast import RECFAST, IonizationHistory
from bolt_b200 import abi, capi
from oracle.oracle import OracleCosmo

par = CosmoParams(); bg = Background(par); r = RECFAST(bg, OmegaB=par.Ω_b, Yp=par.Y_p, OmegaG=par.Ω_r)
ih = IonizationHistory(r, par, bg)
hc = abi.HostCosmo.from_host(par, bg, ih)
oc = OracleCosmo(hc)
ctx = capi.Context(0); dc = capi.DeviceCosmo(ctx, hc)

def rel(a, b):
    sc = np.abs(b).max(axis=1, keepdims=True) if a.ndim == 3 else np.abs(b).max()
    return np.abs(a - b) / (sc + 1e-300)

ks = np.array([1.0, 30.0, 300.0]) * bg.H0
# fixed-step histories
o = abi.make_opts(8, 8, 10, fixed_dt=0.005)
t = time.time(); go = dc.solve(ks, o, want=("S_T", "S_P", "u_hist", "u_final")); tg = time.time() - t
t = time.time(); oo = oc.solve(ks, o, want=("S_T", "S_P", "u_hist", "u_final")); to = time.time() - t
print("fixed: gpu %.2fs oracle %.2fs status" % (tg, to), go["status"], oo["status"], go["nsteps"], oo["nsteps"])
for i in range(len(ks)):
    uh_g, uh_o = go["u_hist"][i], oo["u_hist"][i]
    sc = np.abs(uh_o).max(axis=0) + 1e-300
    e = np.abs(uh_g - uh_o) / sc
    print(" k/H0=%g  max hist rel err (per-variable scale) %.3e at var %d ; S_T %.3e S_P %.3e" % (
        ks[i] / bg.H0, e.max(), np.unravel_index(e.argmax(), e.shape)[1],
        np.abs(go["S_T"][i] - oo["S_T"][i]).max() / np.abs(oo["S_T"][i]).max(),
        np.abs(go["S_P"][i, :-1] - oo["S_P"][i, :-1]).max() / np.abs(oo["S_P"][i, :-1]).max()))
# adaptive
o = abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6)
t = time.time(); go = dc.solve(ks, o, want=("S_T", "S_P", "u_final")); tg = time.time() - t
t = time.time(); oo = oc.solve(ks, o, want=("S_T", "S_P", "u_final")); to = time.time() - t
print("adaptive: gpu %.2fs oracle %.2fs" % (tg, to), go["status"], go["nsteps"], oo["nsteps"], go["nreject"], oo["nreject"])
for i in range(len(ks)):
    print(" k/H0=%g S_T %.3e S_P %.3e u_final %.3e" % (
        ks[i] / bg.H0, np.abs(go["S_T"][i] - oo["S_T"][i]).max() / np.abs(oo["S_T"][i]).max(),
        np.abs(go["S_P"][i, :-1] - oo["S_P"][i, :-1]).max() / np.abs(oo["S_P"][i, :-1]).max(),
        np.abs(go["u_final"][i] - oo["u_final"][i]).max() / np.abs(oo["u_final"][i]).max()))
print(ctx.timing())
