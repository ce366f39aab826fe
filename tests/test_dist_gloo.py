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


class LinkedNumpyModel(NumpySlabPlanes):
    """CPU stand-in for the peer-linked contract of capi.Model (activation_link_info / _link / _begin / _linked_launch /
    _linked_wait / _linked_gather / _end): the ranks' grids and counters are files mapped into every process (the role
    CUDA IPC plays on the GPUs).  A rank relaxes its slab, writes its first / last plane into the neighbour's grid and
    bumps the neighbour's inbox counter; rank 0 detects termination from all ranks' (idle, sent, received) like
    link_detector in csrc/automaton.cu does (two consecutive passes: everyone idle, sum of sent now == sum of received
    in the previous pass)."""
    INFO = 256

    def __init__(self, layers, transfer, tmpdir, rank):
        super().__init__(layers, transfer, 0, layers.shape[0])
        self.rank, self.tmpdir = rank, tmpdir
        self.grid_path = os.path.join(tmpdir, "grid_%d.bin" % rank)
        self.cnt_path = os.path.join(tmpdir, "cnt_%d.bin" % rank)
        np.full(self.lay.shape, np.inf).tofile(self.grid_path)
        np.zeros(8, dtype=np.int64).tofile(self.cnt_path)
        self.t = np.memmap(self.grid_path, dtype=np.float64, mode="r+", shape=self.lay.shape)
        self.cnt = np.memmap(self.cnt_path, dtype=np.int64, mode="r+", shape=(8,))   # working, sent, inbox_dn, inbox_up, done_dn, done_up, verdict
        self.activation_ms = 0.0

    def set_slab(self, z0, z1):
        self.z0, self.z1 = z0, z1

    def activation_link_info(self):
        return self.tmpdir.encode().ljust(self.INFO, b"\0")

    def activation_link(self, rank, infos, slabs):
        assert rank == self.rank and len(infos) == len(slabs) and all(len(i) == self.INFO for i in infos)
        dirs = [i.rstrip(b"\0").decode() for i in infos]
        self.slabs = slabs
        self.grids = [np.memmap(os.path.join(d, "grid_%d.bin" % r), dtype=np.float64, mode="r+", shape=self.lay.shape) for r, d in enumerate(dirs)]
        self.cnts = [np.memmap(os.path.join(d, "cnt_%d.bin" % r), dtype=np.int64, mode="r+", shape=(8,)) for r, d in enumerate(dirs)]
        live = [r for r in range(len(slabs)) if slabs[r][1] > slabs[r][0]]
        self.below = max([r for r in live if r < rank], default=None) if rank in live else None
        self.above = min([r for r in live if r > rank], default=None) if rank in live else None

    def activation_begin(self):
        self.t[:] = np.inf
        self.t[self.start] = 1.0
        self.cnt[:] = 0

    def activation_linked_launch(self, max_ctas=0):
        pass

    def _send(self, peer, z, slot):
        plane = np.array(self.t[z + 1])
        if not (plane < self.grids[peer][z + 1]).any():
            return
        self.cnt[1] += 1                                   # sent, before the neighbour can see the work
        np.minimum(self.grids[peer][z + 1], plane, out=self.grids[peer][z + 1])
        self.cnts[peer][slot] += 1                         # its inbox: received

    def activation_linked_wait(self):
        import time
        first, r_prev, visits = True, None, 0
        deadline = time.time() + 120
        while not self.cnt[6]:
            assert time.time() < deadline
            inbox = (int(self.cnt[2]), int(self.cnt[3]))
            if (first or inbox != (int(self.cnt[4]), int(self.cnt[5]))) and self.z1 > self.z0:
                self.cnt[0] = 1
                self.cnt[4], self.cnt[5] = inbox
                visits += self.relax()[0]
                if self.below is not None:
                    self._send(self.below, self.z0, 3)     # my first plane arrives from above, seen from the rank below
                if self.above is not None:
                    self._send(self.above, self.z1 - 1, 2)
                self.cnt[0] = 0
            else:
                time.sleep(0.0005)
            first = False
            if self.rank == 0:                             # one detector pass
                idle, sent, recv = True, 0, 0
                for c in self.cnts:
                    done = (int(c[4]), int(c[5]))
                    inbox = (int(c[2]), int(c[3]))
                    idle = idle and done == inbox and int(c[0]) == 0
                    sent += int(c[1])
                    recv += inbox[0] + inbox[1]
                if idle and r_prev is not None and sent == r_prev:
                    for c in self.cnts:
                        c[6] = 1
                r_prev = recv
        return visits, (int(self.cnt[1]), int(self.cnt[2] + self.cnt[3]), 0)

    def activation_linked_gather(self):
        for r, (a, b) in enumerate(self.slabs):
            if r != self.rank and b > a:
                self.t[a + 1:b + 1] = self.grids[r][a + 1:b + 1]

    def activation_end(self, download=True):
        return self.end(download)


def _linked_worker(rank, world, port, q, cuts, tmpdir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w, _ = ekdist.init("gloo")
    layers, transfer, _ = synth.small_heart(seed=6, shape=(14, 13, 12), n_layers=4)
    slabs = [(cuts[i], cuts[i + 1]) for i in range(w)]
    model = LinkedNumpyModel(layers, transfer, tmpdir, r)
    model.set_slab(*slabs[r])
    dist.barrier()                                         # every rank's files exist
    ekdist.link_model(model, slabs, r, w)
    ref = oracle.activation(layers, transfer)
    ok = True
    for _ in range(2):                                     # the second run starts from the first one's leftovers
        info, tm = {}, {}
        delay, visits = ekdist.linked_activation(model, slabs, r, w, timings=tm, info=info)
        ok = ok and delay.tobytes() == ref.tobytes() and set(tm) == {"run_s", "local_s", "gather_s", "publish_s"}
    q.put((r, ok, info["bricks_queued_at_neighbours"], info["bricks_queued_here_by_neighbours"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cuts", [[0, 7, 14], [0, 5, 5, 14], [0, 0, 13, 14]])
def test_linked_automaton_driver_gloo(built, cuts, tmp_path):
    """ekgsim_b200.dist.link_model / linked_activation over gloo: the link records are all-gathered in rank order, the
    barriers sit where the contract wants them (rings ready before anyone pushes, slabs final before anyone copies), and
    the termination scheme of the linked kernel -- restated on memory-mapped files -- ends every rank with the oracle's
    bits, also with empty slabs (rank 0, the detector, included)."""
    world = len(cuts) - 1
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_linked_worker, args=(r, world, port, q, cuts, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    got = sorted(q.get() for _ in range(world))
    assert all(ok for _, ok, _, _ in got), got
    assert sum(s for _, _, s, _ in got) == sum(r for _, _, _, r in got) > 0      # every plane sent was received
