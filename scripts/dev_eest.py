import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
os.environ["BOLT_DEBUG_STEPS"] = "gpurun_out/gpu_steps.txt"
import bolt_b200 as B
import hostgen as HG
from bolt_b200 import abi, capi
from oracle.oracle import OracleCosmo
par = B.CosmoParams(); bg = HG.Background(par)
ih = HG.IonizationHistory(HG.RECFAST(bg, OmegaB=par.Ω_b, Yp=par.Y_p, OmegaG=par.Ω_r), par, bg)
hc = abi.HostCosmo.from_host(par, bg, ih)
ctx = capi.Context(0); dc = capi.DeviceCosmo(ctx, hc)
k = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, 100)[int(sys.argv[1]):int(sys.argv[1]) + 1]
o = abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6)
g = dc.solve(k, o, want=("S_T",))
print("gpu", g["nsteps"], g["nreject"])
code = f"""
import sys, os; sys.path.insert(0, '{os.getcwd()}')
import numpy as np, pickle
import bolt_b200 as B
from bolt_b200 import abi
from oracle.oracle import OracleCosmo
par = B.CosmoParams(); bg = HG.Background(par)
ih = HG.IonizationHistory(HG.RECFAST(bg, OmegaB=par.Ω_b, Yp=par.Y_p, OmegaG=par.Ω_r), par, bg)
oc = OracleCosmo(abi.HostCosmo.from_host(par, bg, ih))
o = abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6)
r = oc.solve(np.array([{float(k[0])!r}]), o, want=("S_T",)); print("oracle", r["nsteps"], r["nreject"])
"""
env = dict(os.environ, ORACLE_DEBUG="1")
with open("gpurun_out/oracle_steps.txt", "w") as f:
    subprocess.run([sys.executable, "-c", code], env=env, stderr=f)
a = [l.split() for l in open("gpurun_out/gpu_steps.txt")]; b = [l.split() for l in open("gpurun_out/oracle_steps.txt")]
for i, (x, y) in enumerate(zip(a, b)):
    ea, eb = float(x[2][5:]), float(y[2][5:])
    if abs(ea / eb - 1) > 1e-7 or x[3] != y[3]:
        print("first divergence at step", i, x, y); break
worst = max(abs(float(x[2][5:]) / float(y[2][5:]) - 1) for x, y in list(zip(a, b))[:i])
print("max rel EEst diff before divergence", worst, "steps", len(a), len(b))
for j in range(max(0, i - 3), min(len(a), i + 3)): print(a[j], b[j])
