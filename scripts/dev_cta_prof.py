"""Development (GPU): cycle counters of the pipelined K1 kernel's solver warp (k1_pipe.cu built with -DK1P_PROF into a library
selected by BOLT_CUDA_LIB, BOLT_K1_PIPE=1).  The rows of the first CTA kernel (removed) are kept for reading old logs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BOLT_DEBUG_STEPS"] = "/tmp/k1c_prof.txt"
import numpy as np
import bolt_b200 as B
import hostgen as HG
from bolt_b200 import abi, capi

par = B.CosmoParams(); bg = HG.Background(par)
ih = HG.IonizationHistory(HG.RECFAST(bg, OmegaB=par.Ω_b, Yp=par.Y_p, OmegaG=par.Ω_r), par, bg)
ctx = capi.Context(0); dc = capi.DeviceCosmo(ctx, abi.HostCosmo.from_host(par, bg, ih))
o = abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6, ix_first=1201)
nk = int(sys.argv[1]) if len(sys.argv) > 1 else 1
k = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, 2000)[-nk:]
for rep in range(2):
    out = dc.solve(k, o, want=("S_T", "S_P"))
print("nk", nk, "K1 ms", ctx.timing()["hierarchy_ms"], "steps", out["nsteps"].max(), "rej", out["nreject"].max())
if os.environ.get("BOLT_K1_PIPE"):
    names = {0: "acquire factors+post", 1: "stage: wait P_s + rhs", 2: "stage: wait W_s", 3: "stage: back-solve", 4: "stage: z + publish",
             5: "stage total", 6: "error norm", 7: "accept/controller", 8: "whole step", 9: "solve: loads+down sweep",
             10: "solve: a0..a2 + reductions", 11: "solve: border rhs + y", 12: "solve: scalars + U0..2", 13: "solve: up sweep"}
else:
  names = {0: "solver assemble", 1: "solver wait full[s]", 2: "solver solve_slot", 3: "solver zout", 4: "solver err pass", 5: "solver norm+ctrl",
           6: "solver accept/sample", 7: "solver whole step", 8: "f0 wait start", 9: "f0 bg eval", 10: "f0 factor_reg", 11: "f0 lu4+store",
           13: "f1 wait start", 14: "f1 bg eval", 15: "f1 factor_reg", 16: "f1 lu4+store"}
for line in open("/tmp/k1c_prof.txt"):
    f = dict(t.split("=") for t in line.split())
    cat, cyc, cnt = int(float(f["x"])), float(f["dt"]), float(f["EEst"])
    if cnt > 0:
        print("%-24s total %12.0f cyc  count %8.0f  per call %8.1f" % (names.get(cat, str(cat)), cyc, cnt, cyc / cnt))
