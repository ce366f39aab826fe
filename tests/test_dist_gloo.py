"""World-size-2 gloo tests (CPU) of the multi-GPU plumbing in ekgsim_b200/dist.py.  The per-rank compute
is played by the oracle (allowed in tests); what is under test is the sharding arithmetic and the
collectives: individuals sharded + gathered, z-slabs summed by all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synth
from ekgsim_b200 import dist as ekdist
from oracle import oracle


def test_shard_range_partitions():
    for n in (0, 1, 7, 256, 1001):
        for world in (1, 2, 3, 8):
            r = [ekdist.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1


def test_slab_ranges_balance(model24):
    occ = ((model24["layers"] & 0x0FFF) > 0).sum(axis=(1, 2))
    for world in (1, 2, 4, 8):
        slabs = ekdist.slab_ranges(occ, world)
        assert slabs[0][0] == 0 and slabs[-1][1] == len(occ)
        assert all(slabs[i][1] == slabs[i + 1][0] for i in range(world - 1))
        counts = [int(occ[a:b].sum()) for a, b in slabs]
        assert sum(counts) == 555868
        assert max(counts) <= 1.1 * 555868 / world + occ.max()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w, _ = ekdist.init("gloo")
    layers, transfer, leads = synth.small_heart(seed=5, shape=(12, 13, 14), n_layers=4)
    nl = 4
    k = synth.layer_params(nl, seed=9, batch=5)
    delay = oracle.activation(layers, transfer)
    # (1) individuals sharded over ranks, gathered in rank order
    b0, b1 = ekdist.shard_range(5, r, w)
    mine = np.stack([oracle.run_factored(layers, delay, k[b], leads, "3D4", 0.0, 1.0, 24.0) for b in range(b0, b1)]) if b1 > b0 \
        else np.zeros((0, 2, 24))
    counts = [ekdist.shard_range(5, i, w)[1] - ekdist.shard_range(5, i, w)[0] for i in range(w)]
    allv = ekdist.gather_rows(torch.from_numpy(mine), counts).numpy()
    # (2) one model split into z-slabs, partial ECGs summed
    occ = ((layers & 0x0FFF) > 0).sum(axis=(1, 2))
    z0, z1 = ekdist.slab_ranges(occ, w)[r]
    part = torch.from_numpy(oracle.run_factored_slab(layers, delay, k[0], leads, z0, z1, "3D4", 0.0, 1.0, 24.0))
    ekdist.allreduce_sum_(part)
    tmax = ekdist.max_over_ranks(float(r + 1))
    if r == 0:
        ref = np.stack([oracle.run_factored(layers, delay, k[b], leads, "3D4", 0.0, 1.0, 24.0) for b in range(5)])
        q.put((float(np.abs(allv - ref).max()), float(np.abs(part.numpy() - ref[0]).max() / np.abs(ref[0]).max()), tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_sharding_and_collectives(built):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    gathered_err, slab_err, tmax = q.get()
    assert gathered_err == 0.0       # gather is a pure permutation
    assert slab_err < 1e-12          # two slabs add up to the whole model
    assert tmax == 2.0


class NumpySlabPlanes:
    """CPU stand-in for ekgsim_b200.dist.ModelPlanes: the same begin / relax / export / merge / end contract on a
    zero-bordered numpy grid, relaxing only the voxels of its z-slab (vectorised label-correcting sweeps with the
    reference's edge weights fl(T[exciting][excited] * sqrt(|dif|^2)) and fl(+))."""

    def __init__(self, layers, transfer, z0, z1):
        self.z0, self.z1 = z0, z1
        self.lay = np.pad((layers & 0x0FFF).astype(np.int64), 1)
        self.start = np.pad((layers & 0x1000) != 0, 1)
        self.T = transfer
        self.shape = layers.shape
        self.plane_elems = self.lay.shape[1] * self.lay.shape[2]
        self.device = "cpu"

    def begin(self):
        self.t = np.full(self.lay.shape, np.inf)
        self.t[self.start] = 1.0

    def relax(self, max_visits=0):
        """-> (sweeps, 1 if the bound stopped it before the slab's fixed point else 0); the bound counts sweeps here"""
        Z, Y, X = self.shape
        own = (slice(self.z0 + 1, self.z1 + 1), slice(1, Y + 1), slice(1, X + 1))
        lv = self.lay[own]
        sweeps = 0
        while True:
            best = self.t[own].copy()
            for dz in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    for dx in (-1, 0, 1):
                        if not (dz or dy or dx):
                            continue
                        src = (slice(self.z0 + 1 + dz, self.z1 + 1 + dz), slice(1 + dy, Y + 1 + dy), slice(1 + dx, X + 1 + dx))
                        lu = self.lay[src]
                        w = self.T[lu, lv] * np.sqrt(float(dz * dz + dy * dy + dx * dx))
                        cand = np.where((lu > 0) & (lv > 0), self.t[src] + w, np.inf)
                        best = np.minimum(best, cand)
            sweeps += 1
            if not (best < self.t[own]).any():
                return sweeps, 0
            self.t[own] = best
            if max_visits and sweeps >= max_visits:
                return sweeps, 1

    def export(self, z_begin, z_end):
        return torch.from_numpy(self.t[z_begin + 1:z_end + 1].reshape(z_end - z_begin, -1).copy())

    def merge(self, z_begin, planes):
        cur = self.t[z_begin + 1:z_begin + 1 + planes.shape[0]]
        new = planes.numpy().reshape(cur.shape)
        improved = int((new < cur).sum())
        np.minimum(cur, new, out=cur)
        return improved

    def end(self, download=True):
        if not download:      # like ekg_model_activation_end with delay_out = NULL: the map stays where it is
            return None
        out = self.t[1:-1, 1:-1, 1:-1].copy()
        out[~np.isfinite(out) | (self.lay[1:-1, 1:-1, 1:-1] == 0)] = 0.0
        return out


def _automaton_worker(rank, world, port, q, cuts):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w, _ = ekdist.init("gloo")
    layers, transfer, _ = synth.small_heart(seed=6, shape=(14, 13, 12), n_layers=4)
    slabs = [(cuts[i], cuts[i + 1]) for i in range(w)]
    planes = NumpySlabPlanes(layers, transfer, *slabs[r])
    delay, rounds, _ = ekdist.sharded_activation(planes, slabs, r, w)
    ref = oracle.activation(layers, transfer)
    ok = delay.tobytes() == ref.tobytes()
    # bounded relaxation (the wave is handed over before a slab is finished): more rounds, the same bits
    delay_b, rounds_b, _ = ekdist.sharded_activation(planes, slabs, r, w, visits_per_round=2)
    ok = ok and delay_b.tobytes() == ref.tobytes() and rounds_b >= rounds
    # again without the host copy (what the slab ECG needs): same rounds, nothing returned, the rank's state is the map
    none, rounds2, _ = ekdist.sharded_activation(planes, slabs, r, w, download=False)
    ok = ok and none is None and rounds2 == rounds and planes.end().tobytes() == ref.tobytes()
    q.put((r, ok, rounds))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cuts", [[0, 7, 14], [0, 5, 5, 14], [0, 1, 13, 14]])
def test_sharded_automaton_driver_gloo(built, cuts):
    """ekgsim_b200.dist.sharded_activation (plane exchange, global termination, final gather) over gloo with a numpy
    stand-in for the per-rank relaxation: every rank must end with the oracle's bits -- also with an empty slab in the
    middle and with one-plane slabs."""
    world = len(cuts) - 1
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_automaton_worker, args=(r, world, port, q, cuts)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    got = sorted(q.get() for _ in range(world))
    assert all(ok for _, ok, _ in got), got
    assert got[0][2] >= 2 and len({rounds for _, _, rounds in got}) == 1      # every rank leaves the loop in the same round
