"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed (NCCL over NVLink) as plumbing.

Two ways the path shards:
  * batches of cosmologies: independent, sharded round-robin over ranks, NO collective on the data path
    (`shard_round_robin`); this is what bench.py runs at N > 1 (weak scaling);
  * the k-modes of ONE cosmology: `spectra_k_sharded` solves the local k-modes, all-gathers the source
    grids (C_l is quadratic in the k-interpolated source, spectra.jl:91-93, so both bracketing coarse
    columns must be present), projects the local multipoles and combines the partial C_l vector with one
    all-reduce.
The compute callables are injected so that the sharding logic is testable on CPU with the gloo backend.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_round_robin(n_items, rank, world):
    """Indices of the items rank `rank` owns."""
    return np.arange(rank, n_items, world)


def k_shard(k, rank, world):
    """Cyclic shard of the k-modes after sorting by descending k (step count grows with k: load balance)."""
    order = np.argsort(-np.asarray(k), kind="stable")
    return np.sort(order[rank::world])


def spectra_k_sharded(k, ells, solve_fn, project_fn, n_x, device, group=None):
    """k: float64 array [nk]; ells: int array (increasing).
    solve_fn(k_local: tensor) -> (S_T, S_P) tensors [nk_local][n_x] on `device`;
    project_fn(S_T, S_P, k_all: tensor, ells_local: ndarray) -> tensor [3][nell_local].
    Returns a tensor [3][nell] (tt, te, ee), identical on every rank."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    k = np.ascontiguousarray(k, dtype=np.float64)
    ells = np.ascontiguousarray(ells, dtype=np.int32)
    nk = len(k)
    mine = k_shard(k, rank, world)
    k_all = torch.from_numpy(k).to(device)
    S_T_loc, S_P_loc = solve_fn(k_all[torch.from_numpy(mine).to(device)])
    if world > 1:
        per = (nk + world - 1) // world
        pad = torch.zeros((2, per, n_x), dtype=torch.float64, device=device)
        pad[0, :len(mine)] = S_T_loc; pad[1, :len(mine)] = S_P_loc
        gathered = torch.empty((world, 2, per, n_x), dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(gathered.view(-1), pad.view(-1), group=group)
        S_T = torch.empty((nk, n_x), dtype=torch.float64, device=device); S_P = torch.empty_like(S_T)
        for r in range(world):
            idx = torch.from_numpy(k_shard(k, r, world)).to(device)
            S_T[idx] = gathered[r, 0, :len(idx)]; S_P[idx] = gathered[r, 1, :len(idx)]
    else:
        S_T, S_P = S_T_loc, S_P_loc
    my_l = np.arange(rank, len(ells), world)
    cl = torch.zeros((3, len(ells)), dtype=torch.float64, device=device)
    if len(my_l):
        cl[:, torch.from_numpy(my_l).to(device)] = project_fn(S_T, S_P, k_all, ells[my_l])
    if world > 1:
        dist.all_reduce(cl, op=dist.ReduceOp.SUM, group=group)   # disjoint supports: the sum is a concatenation
    return cl


def device_spectra_k_sharded(dc, k, opts, ells, kd_min, kd_max, n_kd, ix_start, device, group=None):
    """spectra_k_sharded bound to libbolt_cuda's device-pointer entry points."""
    import copy
    o = copy.copy(opts); o.ix_first = max(o.ix_first, ix_start)

    def solve_fn(k_loc):
        S_T, S_P, status, nsteps, _ = dc.solve_device(k_loc.contiguous(), o)
        return S_T, S_P

    def project_fn(S_T, S_P, k_all, ells_loc):
        return dc.project_device(S_T, S_P, k_all, ells_loc, kd_min, kd_max, n_kd, ix_start)

    return spectra_k_sharded(k, ells, solve_fn, project_fn, dc.hc.n_x, device, group)
