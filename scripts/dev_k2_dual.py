"""Development (GPU): one C3 spectra call with six partials (for ncu captures of the dual K2 kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bolt_b200 as B
from bolt_b200 import abi, capi
from hostgen import host_cosmo_with_partials
names = ["Ω_b", "Ω_c", "h", "n", "A", "Σm_ν"]
par = B.CosmoParams()
dual, base, bg, ih, pm, steps = host_cosmo_with_partials(par, names, rel_step=1e-3)
ctx = capi.Context(0); dc = capi.DeviceCosmo(ctx, dual)
nk = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
k = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, nk)
o = abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6)
ells = np.arange(2, 2501, dtype=np.int32); ix0 = int(np.argmax(bg.x_grid > -8))
for rep in range(int(os.environ.get("REPS", "2"))):
    r = dc.spectra(k, o, ells, 0.01 * bg.H0, 1000 * bg.H0, 5000, ix0)
    print(ctx.timing(), flush=True)
