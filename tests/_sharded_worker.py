"""Worker of tests/test_gpu_multi.py (one process per GPU, launched by torch.distributed.run): the in-library NCCL paths
bolt_spectra_sharded and bolt_plin_sharded against the single-GPU entry points, on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bolt_b200 as B  # noqa: E402
import hostgen as HG  # noqa: E402
from bolt_b200 import abi, capi  # noqa: E402

rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
par = B.CosmoParams()
bg = HG.Background(par)
ih = HG.IonizationHistory(HG.RECFAST(bg, OmegaB=par.Ω_b, Yp=par.Y_p, OmegaG=par.Ω_r), par, bg)
ctx = capi.Context(lrank)
saved = os.dup(1); os.dup2(2, 1)          # NCCL's banner goes to stderr
try:
    ctx.comm_init_torch()
finally:
    os.dup2(saved, 1); os.close(saved)
dc = capi.DeviceCosmo(ctx, abi.HostCosmo.from_host(par, bg, ih))

# C_l: 301 modes (ragged over the ranks), every 7th multipole
k = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, 301)
o = abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6)
ells = np.arange(2, 2501, 7, dtype=np.int32)
ix0 = int(np.argmax(bg.x_grid > -8))
args = (k, o, ells, 0.01 * bg.H0, 1000 * bg.H0, 5000, ix0)
os.environ["BOLT_K1_WARP"] = "1"          # same K1 kernel on both sides: results must then agree to rounding of the reduction order
tt, te, ee, st, ns = dc.spectra_sharded(*args)
r = dc.spectra(*args)
assert np.array_equal(st, r[3]) and np.array_equal(ns, r[4]), "status / step counts differ"
for nm, a, b in (("tt", tt, r[0]), ("te", te, r[1]), ("ee", ee, r[2])):
    err = float(np.abs(a - b).max() / np.abs(b).max())
    assert err < 1e-13, (nm, err)

# P(k): 37 modes of plin's truncation, bit-identical (no reduction on this path)
kp = B.log10_k(10 * bg.H0, 5000 * bg.H0, 37)
op = abi.make_opts(50, 50, 20, reltol=1e-5, abstol=1e-6)
pk, stp, nsp = dc.plin_sharded(kp, op)
pk1, st1, ns1 = dc.plin(kp, op)
assert np.array_equal(pk, pk1) and np.array_equal(stp, st1) and np.array_equal(nsp, ns1), "plin_sharded differs"

# fewer modes than ranks x 2 and a single multipole: empty shards must not hang or corrupt
k3 = B.quadratic_k(0.1 * bg.H0, 100 * bg.H0, 3)
one = dc.spectra_sharded(k3, o, np.array([10], dtype=np.int32), 0.01 * bg.H0, 1000 * bg.H0, 5000, ix0)
ref = dc.spectra(k3, o, np.array([10], dtype=np.int32), 0.01 * bg.H0, 1000 * bg.H0, 5000, ix0)
assert abs(one[0][0] / ref[0][0] - 1) < 1e-13
dist.barrier()
ctx.comm_free()
dist.destroy_process_group()
print(f"rank {rank}: sharded paths OK", flush=True)
