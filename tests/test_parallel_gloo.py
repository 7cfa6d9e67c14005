"""world_size-2 test of the multi-GPU sharding logic on CPU (gloo); the compute callables are stand-ins."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _fake_solve(k_loc, n_x=7):
    x = torch.arange(n_x, dtype=torch.float64)
    return torch.sin(k_loc[:, None] * (1 + x[None, :])), torch.cos(k_loc[:, None] + x[None, :])


def _fake_project(S_T, S_P, k_all, ells):
    l = torch.from_numpy(ells.astype(np.float64))
    w = k_all / k_all.sum()
    a = (S_T.sum(1) * w).sum() * l; b = (S_P.sum(1) * w).sum() * l ** 2
    return torch.stack([a * a, a * b, b * b])


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bolt_b200.parallel import spectra_k_sharded, k_shard, shard_round_robin
    k = np.linspace(0.3, 9.0, 11); ells = np.arange(2, 15, dtype=np.int32)
    cl = spectra_k_sharded(k, ells, _fake_solve, _fake_project, 7, torch.device("cpu"))
    shards = [k_shard(k, r, world) for r in range(world)]
    rr = [shard_round_robin(5, r, world) for r in range(world)]
    if rank == 0:
        torch.save(dict(cl=cl, shards=shards, rr=rr), out)
    dist.barrier()
    dist.destroy_process_group()


def test_k_sharded_spectra_matches_single_rank(tmp_path):
    from bolt_b200.parallel import spectra_k_sharded
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    k = np.linspace(0.3, 9.0, 11); ells = np.arange(2, 15, dtype=np.int32)
    ref = spectra_k_sharded(k, ells, _fake_solve, _fake_project, 7, torch.device("cpu"))
    assert torch.allclose(got["cl"], ref, rtol=1e-13, atol=0)
    # the k shards partition the modes, alternate the largest k between ranks, and stay index-sorted
    s0, s1 = got["shards"]
    assert sorted(np.concatenate([s0, s1]).tolist()) == list(range(11))
    assert 10 in s0 and 9 in s1 and np.all(np.diff(s0) > 0)
    assert got["rr"][0].tolist() == [0, 2, 4] and got["rr"][1].tolist() == [1, 3]
