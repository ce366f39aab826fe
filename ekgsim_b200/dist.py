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
* the activation automaton on such a sharded model: the one place with a real halo exchange.  `linked_activation` is the
  NVLink form -- every rank maps its neighbours' grids (CUDA IPC) and ONE kernel per rank writes improved face voxels
  into the neighbour's grid and queues the neighbour's bricks itself; torch.distributed only carries the 256-byte link
  records once and three barriers per run.  `sharded_activation` is the host-driven form (per round one plane each way
  through point-to-point messages plus an all-reduce of the "anything improved?" counts), the fallback where the devices
  have no native peer atomics.
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


class ModelPlanes:
    """Adapter between a device-resident `Model` (capi.py) and torch tensors: the sharded automaton's plane
    exchange needs tensors for the collectives, the model only speaks raw device pointers."""

    def __init__(self, model, device):
        import torch
        self.model, self.device, self.torch = model, device, torch
        self.plane_elems = model.plane_elems
        self.async_merge = True   # merges add into the driver's device counter

    def begin(self):
        self.model.activation_begin()

    def relax(self, max_visits=0):
        """-> (brick visits, bricks still queued)"""
        return self.model.activation_relax_bounded(max_visits)

    def export(self, z_begin, z_end):
        t = self.torch.empty((max(z_end - z_begin, 0), self.plane_elems), dtype=self.torch.float64, device=self.device)
        if t.numel():
            self.model.activation_export(z_begin, z_end, t.data_ptr(), self.torch.cuda.current_stream(self.device).cuda_stream)
        return t

    def merge(self, z_begin, planes, counter=None):
        """counter: an int64 device tensor the number of improved cells is added to (no read-back, no synchronisation);
        None: the count is returned"""
        if planes.numel() == 0:
            return 0
        stream = self.torch.cuda.current_stream(self.device)   # the collective that filled `planes` was ordered on it (req.wait())
        if counter is not None:
            self.model.activation_merge_async(z_begin, z_begin + planes.shape[0], planes.data_ptr(), counter.data_ptr(), stream.cuda_stream)
            return 0
        stream.synchronize()
        return self.model.activation_merge(z_begin, z_begin + planes.shape[0], planes.data_ptr(), stream.cuda_stream)

    def end(self, download=True):
        return self.model.activation_end(download=download)


def sharded_activation(planes, slabs, rank=None, world=None, max_rounds=10000, timings=None, download=True, visits_per_round=None):
    """The activation automaton of a model sharded into z-slabs.

    `planes` offers begin / relax(max_visits) -> (visits, bricks left) / export(z0, z1) -> tensor[z1-z0, plane_elems] /
    merge(z0, tensor) -> improved cells / end (ModelPlanes for the CUDA model); `slabs[r] = (z_begin, z_end)` is rank r's
    slab, already set on the model.  Per round every rank relaxes its slab for at most `visits_per_round` brick visits
    (None: a bound derived from the slab size; 0: to the fixed point of its current halo), sends its first and last
    own plane to the rank below / above (one point-to-point message each way over NCCL / NVLink), merges what it
    receives (elementwise minimum) and the ranks agree through one all-reduce whether anything improved or is still
    queued anywhere.  Bounding the relaxation hands the wave to the neighbouring slabs early: an excitation wave
    crosses the slabs one after the other, and with unbounded rounds the ranks would work one after the other, too.
    Min-merging never overshoots the least fixed point, so the result has the bits of the single-GPU run whatever the
    bound.  At the end every rank broadcasts its slab so that all ranks hold the whole map, like after
    `ekg_model_activation`.
    Returns (delay[Z, Y, X] as numpy, rounds, brick visits of this rank); `timings` (a dict) receives the wall seconds
    of the three phases: rounds, gather, publish.  download=False leaves the map on the device (delay is None): the
    ECG entry points only need it there."""
    import time
    import torch
    import torch.distributed as dist
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    assert len(slabs) == world
    live = [r for r in range(world) if slabs[r][1] > slabs[r][0]]     # ranks with an empty slab only take part in the collectives
    below = max([r for r in live if r < rank], default=None) if rank in live else None
    above = min([r for r in live if r > rank], default=None) if rank in live else None
    z0, z1 = slabs[rank]
    if visits_per_round is None:
        # Two slabs: the wave starts next to the cut, both ranks are busy from the first round on -- unbounded rounds are
        # best (measured on the 4x heart, 2 GPUs: 22 ms in 2-3 rounds; 25 ms / 5 rounds bounded to 400 k visits, 35 ms /
        # 18 rounds bounded to 100 k: a round costs 0.5-1 ms of launches, exchange and agreement).  More slabs: the wave
        # would cross them one after the other (8 GPUs, unbounded: 4 rounds of ~10 ms = 43 ms, no faster than one GPU);
        # a bound hands it on after one round.  Too small a bound multiplies the rounds AND the work (stale halos are
        # relaxed again: 8 GPUs, 25 k visits: 49 rounds, 41 ms; 50 k: 20 rounds, 27 ms) -- about a third of the slab's
        # brick cells, at least 50 k.
        visits_per_round = 0 if world <= 2 else max(50000, (z1 - z0) * planes.plane_elems // (64 * 3))
    t_begin = time.perf_counter()
    planes.begin()
    visits, rounds = 0, 0
    while True:
        rounds += 1
        left = 0
        if rank in live:
            v, left = planes.relax(visits_per_round)
            visits += v
        improved = left
        if world > 1:
            ops, recv_below, recv_above = [], None, None
            if below is not None:
                send_dn = planes.export(z0, z0 + 1)                     # my first plane is the halo of the rank below
                recv_below = torch.empty_like(send_dn)
                ops += [dist.P2POp(dist.isend, send_dn, below), dist.P2POp(dist.irecv, recv_below, below)]
            if above is not None:
                send_up = planes.export(z1 - 1, z1)
                recv_above = torch.empty_like(send_up)
                ops += [dist.P2POp(dist.isend, send_up, above), dist.P2POp(dist.irecv, recv_above, above)]
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
            # improved cells of both merges + bricks still queued, summed over the ranks; on the GPU the merges add into the
            # device counter (no read-back): the only host synchronisation of the exchange is the .item() below
            flag = torch.full((1,), improved, dtype=torch.int64, device=send_device(planes))
            dev_counter = flag if getattr(planes, "async_merge", False) else None
            if recv_below is not None:
                improved += planes.merge(slabs[below][1] - 1, recv_below, dev_counter) if dev_counter is not None else planes.merge(slabs[below][1] - 1, recv_below)   # = its last own plane
            if recv_above is not None:
                improved += planes.merge(slabs[above][0], recv_above, dev_counter) if dev_counter is not None else planes.merge(slabs[above][0], recv_above)
            if dev_counter is None:
                flag.fill_(improved)
            dist.all_reduce(flag, op=dist.ReduceOp.SUM)
            improved = int(flag.item())
        if improved == 0:
            break
        if rounds >= max_rounds:
            raise RuntimeError("sharded activation automaton did not converge in %d rounds" % max_rounds)
    t_rounds = time.perf_counter()
    # every rank ends up with the whole map: rank s broadcasts its slab, the others min-merge it (what they hold
    # outside their own slab are upper bounds)
    if world > 1:
        for s in live:
            a, b = slabs[s]
            buf = planes.export(a, b) if s == rank else torch.empty((b - a, planes.plane_elems), dtype=torch.float64, device=send_device(planes))
            dist.broadcast(buf, src=s)
            if s != rank:
                planes.merge(a, buf)
    t_gather = time.perf_counter()
    delay = planes.end() if download else planes.end(download=False)
    if timings is not None:
        timings.update(rounds_s=t_rounds - t_begin, gather_s=t_gather - t_rounds, publish_s=time.perf_counter() - t_gather)
    return delay, rounds, visits


def barrier(device=None):
    import torch
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        if device is not None and getattr(device, "type", "cpu") == "cuda":
            dist.barrier(device_ids=[device.index])
            torch.cuda.synchronize(device)
        else:
            dist.barrier()


def exchange_link_infos(info: bytes, device=None):
    """All-gathers the ranks' link records (ekg_model_activation_link_info), rank order."""
    import torch
    import torch.distributed as dist
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return [info]
    world = dist.get_world_size()
    dev = device if device is not None and getattr(device, "type", "cpu") == "cuda" else "cpu"
    mine = torch.frombuffer(bytearray(info), dtype=torch.uint8).to(dev)
    out = torch.empty(world * len(info), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, mine)
    blob = out.cpu().numpy().tobytes()
    return [blob[r * len(info):(r + 1) * len(info)] for r in range(world)]


def link_model(model, slabs, rank=None, world=None, device=None):
    """Maps the other ranks' grids into `model` (once per model / slab layout).  Raises EkgError (code -7) where the
    devices offer no native peer atomics; callers then use sharded_activation."""
    import torch.distributed as dist
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    infos = exchange_link_infos(model.activation_link_info(), device)
    assert len(infos) == world == len(slabs)
    model.activation_link(rank, infos, slabs)


def linked_activation(model, slabs, rank=None, world=None, device=None, timings=None, download=True, gather=True, info=None):
    """The activation automaton of a model sharded into z-slabs, peer-linked (include/ekgsim_b200.h): after
    `link_model` every rank launches ONE frontier kernel; improved voxels on the slab faces go into the neighbouring
    rank's grid and the neighbour's bricks into the neighbour's ring from inside the kernel (NVLink atomics), idle ranks
    wait in the kernel, rank 0's first warp detects global termination.  torch.distributed contributes barriers only:
    after begin (rings ready before anyone pushes into them), after the kernels (slabs final before anyone copies them),
    after the gather.  gather=False leaves the map sharded -- outside its slab a rank then holds upper bounds only, which
    is all the slab's ECG needs, but `end` (range of the activation times) wants the whole map, so gather stays on unless
    the caller knows better.  Returns (delay or None, brick visits of this rank); `timings` receives wall seconds of the
    phases and `info` (a dict) the message counters."""
    import time
    import torch.distributed as dist
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    assert len(slabs) == world
    model.activation_begin()
    barrier(device)
    t0 = time.perf_counter()
    model.activation_linked_launch()
    visits, remote = model.activation_linked_wait()
    t_local = time.perf_counter()
    barrier(device)
    t1 = time.perf_counter()
    if gather and world > 1:
        model.activation_linked_gather()
        barrier(device)
    t2 = time.perf_counter()
    delay = model.activation_end(download=download)
    if timings is not None:
        timings.update(run_s=t1 - t0, local_s=t_local - t0, gather_s=t2 - t1, publish_s=time.perf_counter() - t2)
    if info is not None:
        info.update(bricks_queued_at_neighbours=remote[0], bricks_queued_here_by_neighbours=remote[1], cells_written_to_neighbours=remote[2],
                    kernel_ms=model.activation_ms)
    return delay, visits


def send_device(planes):
    return getattr(planes, "device", "cpu")
