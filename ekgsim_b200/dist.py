"""ekgsim_b200/dist.py -- multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs,
gloo in the CPU tests) for the little that has to cross ranks.

The path shards in two ways:

* individuals (BASELINE configs 3 and 5): every rank holds the whole model and evaluates a
  contiguous share of the parameter vectors.  No data-path collective; the per-individual results
  (criteria, or ECGs) are gathered once at the end -- exactly what the reference does with one MPI
  message per evaluation (ParallelFramework.h:171-178, :432-438).
* one large model (config 4): every rank takes a z-slab of the voxels (`ekg_model_set_slab`), balanced
  by occupied voxels.  The action potential is a closed form of static per-voxel data and the
  lead-field coefficient of a voxel only needs the *occupancy* of its neighbours, which every rank
  knows from the full layer map, so no halo of potentials is exchanged; the partial ECGs
  [B][L][T] (f64, 6.4 kB per simulation on model_24) are summed with one all-reduce.
"""
from __future__ import annotations

import os

import numpy as np


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced [begin, end) share of n items for `rank` (first n % world ranks get one more)."""
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def slab_ranges(occupied_per_z, world: int):
    """Cuts [0, Z) into `world` contiguous z-slabs with (nearly) equal numbers of occupied voxels.
    Returns a list of (z_begin, z_end); slabs partition [0, Z) and may be empty for tiny models."""
    occ = np.asarray(occupied_per_z, dtype=np.int64)
    Z = len(occ)
    cum = np.concatenate([[0], np.cumsum(occ)])
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        z = int(np.searchsorted(cum, target, side="left"))
        # the cut that leaves the prefix closest to the target
        if z > 0 and abs(cum[z - 1] - target) <= abs(cum[min(z, Z)] - target):
            z -= 1
        cuts.append(min(max(z, cuts[-1]), Z))
    cuts.append(Z)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend=None):
    """Initialises torch.distributed from the torchrun environment (no-op for a single process)."""
    import torch
    import torch.distributed as dist
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def allreduce_sum_(t):
    """In-place sum over ranks of a tensor of partial ECGs (f64)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def gather_rows(t, counts):
    """All-gathers row blocks of different lengths (`counts[r]` rows on rank r) into one tensor, rank order."""
    import torch
    import torch.distributed as dist
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return t
    world = dist.get_world_size()
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([out[r][: counts[r]] for r in range(world)], dim=0)


def max_over_ranks(x: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
