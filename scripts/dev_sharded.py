"""Development (N GPUs, torchrun): bolt_spectra_sharded (k- and l-sharded, NCCL inside the library) against the single-GPU bolt_spectra.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dev_sharded.py [nk] [lgamma]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import bolt_b200 as B
import hostgen as HG
from bolt_b200 import abi, capi

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lrank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lrank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
nk = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
lg = int(sys.argv[2]) if len(sys.argv) > 2 else 8
par = B.CosmoParams(); bg = HG.Background(par)
ih = HG.IonizationHistory(HG.RECFAST(bg, OmegaB=par.Ω_b, Yp=par.Y_p, OmegaG=par.Ω_r), par, bg)
ctx = capi.Context(lrank)
if world > 1:
    ctx.comm_init_torch()
dc = capi.DeviceCosmo(ctx, abi.HostCosmo.from_host(par, bg, ih))
k = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, nk)
o = abi.make_opts(lg, 8, 10, reltol=1e-11, abstol=1e-6)
ells = np.arange(2, 2501, dtype=np.int32); ix0 = int(np.argmax(bg.x_grid > -8))
args = (k, o, ells, 0.01 * bg.H0, 1000 * bg.H0, 5000, ix0)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for rep in range(3):
    barrier(); t0 = time.perf_counter()
    tt, te, ee, st, ns = dc.spectra_sharded(*args)
    barrier(); dt = time.perf_counter() - t0
    tm = ctx.timing()
    if rank == 0:
        print("sharded x%d: %.2f ms  (K1 %.2f, K2 %.2f, total dev %.2f)  bad %d steps %d" % (world, 1e3 * dt, tm["hierarchy_ms"], tm["project_ms"], tm["total_ms"],
                                                                                             (st != 0).sum(), ns.sum()), flush=True)
if rank == 0:
    for rep in range(2):
        t0 = time.perf_counter(); r = dc.spectra(*args); dt1 = time.perf_counter() - t0
    tm = ctx.timing()
    print("single GPU: %.2f ms (K1 %.2f, K2 %.2f)" % (1e3 * dt1, tm["hierarchy_ms"], tm["project_ms"]))
    for nm, a, b in (("tt", tt, r[0]), ("te", te, r[1]), ("ee", ee, r[2])):
        print(nm, "max rel diff sharded vs single: %.3e" % np.abs(a / b - 1).max())
    print("status equal", np.array_equal(st, r[3]), "nsteps equal", np.array_equal(ns, r[4]))
# plin (BASELINE config 2): 500 modes, 50/50/20, sharded with one all-gather
kp = B.log10_k(10 * bg.H0, 5000 * bg.H0, 500)
op = abi.make_opts(50, 50, 20, reltol=1e-5, abstol=1e-6)
for rep in range(3):
    barrier(); t0 = time.perf_counter()
    pk, stp, nsp = dc.plin_sharded(kp, op)
    barrier(); dtp = time.perf_counter() - t0
if rank == 0:
    for rep in range(2):
        t0 = time.perf_counter(); pk1, st1, ns1 = dc.plin(kp, op); dt1 = time.perf_counter() - t0
    print("plin sharded x%d: %.2f ms; single GPU %.2f ms; bit-identical %s, status equal %s, nsteps equal %s" % (
        world, 1e3 * dtp, 1e3 * dt1, np.array_equal(pk, pk1), np.array_equal(stp, st1), np.array_equal(nsp, ns1)), flush=True)
barrier()
if world > 1:
    ctx.comm_free()
    dist.destroy_process_group()
