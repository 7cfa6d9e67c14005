"""2+ GPU check (torchrun): k-sharded spectra of ONE cosmology vs the single-GPU result; NCCL all-gather + all-reduce."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bolt_b200 as B
import hostgen as HG
from bolt_b200 import abi, capi
from bolt_b200.parallel import device_spectra_k_sharded
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
par = B.CosmoParams(); bg = HG.Background(par)
ih = HG.IonizationHistory(HG.RECFAST(bg, OmegaB=par.Ω_b, Yp=par.Y_p, OmegaG=par.Ω_r), par, bg)
ctx = capi.Context(lr); dc = capi.DeviceCosmo(ctx, abi.HostCosmo.from_host(par, bg, ih))
k = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, 2000)
ells = np.arange(2, 2501, dtype=np.int32)
o = abi.make_opts(8, 8, 10, reltol=1e-11, abstol=1e-6)
dev = torch.device("cuda", lr)
for rep in range(3):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    cl = device_spectra_k_sharded(dc, k, o, ells, 0.01 * bg.H0, 1000 * bg.H0, 5000, 1201, dev)
    torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
if rank == 0:
    t0 = time.perf_counter(); tt, te, ee, st, ns = dc.spectra(k, o, ells, 0.01 * bg.H0, 1000 * bg.H0, 5000, 1201); t1 = time.perf_counter() - t0
    c = cl.cpu().numpy()
    print("world %d: k-sharded %.1f ms vs single-GPU %.1f ms; max rel diff TT %.2e EE %.2e TE %.2e" % (
        world, 1e3 * dt, 1e3 * t1, np.abs(c[0] / tt - 1).max(), np.abs(c[2] / ee - 1).max(), np.abs(c[1] - te).max() / np.sqrt(tt * ee).max()))
dist.barrier(); dist.destroy_process_group()
