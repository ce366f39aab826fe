#!/usr/bin/env python
"""tools/bench_heart4x.py -- BASELINE config 4: synthetic 4x-resolution voxel heart
(np.repeat(model_24, 4) along z, y, x -> 496 x 496 x 372 grid, 35.6 M occupied voxels, one start
voxel, lead positions x4), ONE simulation, voxels sharded over the ranks as z-slabs, partial ECGs
summed with one all-reduce.  Run with torchrun (one process per GPU) or as a single process.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_heart4x.py [--check] [--steps K]

--check compares the automaton with the oracle (bit-exact; ~2 min of CPU on rank 0) and the ECG with
the oracle's class-factored loop on the first 8 samples.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import ekgio  # noqa: E402
import ekgsim_b200 as ek  # noqa: E402
from ekgsim_b200 import dist as ekdist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--factor", type=int, default=4)
    ap.add_argument("--mode", default="direct")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--sharded-automaton", action="store_true",
                    help="also run the automaton sharded over the z-slabs (plane exchange over NCCL) and compare with the replicated run")
    ap.add_argument("--linked-automaton", action="store_true",
                    help="also run the peer-linked sharded automaton (in-kernel exchange over NVLink peer memory, no host rounds)")
    ap.add_argument("--visits-per-round", type=int, default=-1, help="bound of the sharded automaton's relaxation per round (-1: the driver's default, 0: none)")
    a = ap.parse_args()
    rank, world, local = ekdist.init()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(local)
    layers, transfer, leads = ekgio.scaled_heart(a.factor)
    n_occ = int(((layers & 0x0FFF) > 0).sum())
    t0 = time.time()
    model = ek.Model(layers, transfer, device=local)
    t_create = time.time() - t0
    t0 = time.perf_counter()
    _, sweeps = model.activation(download=False)     # what the facade / CLI does: the map stays on the device
    t_auto_call = time.perf_counter() - t0
    auto_ms = model.activation_ms
    t0 = time.perf_counter()
    delay = model.get_activation()                   # the optional raster host copy (chunked through pinned buffers)
    t_download = time.perf_counter() - t0
    occ_z = ((layers & 0x0FFF) > 0).sum(axis=(1, 2))
    z0, z1 = ekdist.slab_ranges(occ_z, world)[rank]
    model.set_slab(z0, z1)
    sharded = None
    if a.sharded_automaton:
        slabs = ekdist.slab_ranges(occ_z, world)
        planes = ekdist.ModelPlanes(model, dev)
        for _ in range(2):
            torch.cuda.synchronize()
            if world > 1:
                torch.distributed.barrier()
            t0 = time.perf_counter()
            tm = {}
            _, rounds, visits = ekdist.sharded_activation(planes, slabs, rank, world, timings=tm, download=False,
                                                          visits_per_round=None if a.visits_per_round < 0 else a.visits_per_round)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        d2 = model.get_activation()
        same = bool(d2.tobytes() == delay.tobytes())
        sharded = {"ms_total_incl_final_gather": ekdist.max_over_ranks(dt * 1e3, dev),
                   "ms_rounds": ekdist.max_over_ranks(tm["rounds_s"] * 1e3, dev), "ms_gather": ekdist.max_over_ranks(tm["gather_s"] * 1e3, dev),
                   "ms_publish_on_device": ekdist.max_over_ranks(tm["publish_s"] * 1e3, dev), "rounds": rounds, "brick_visits_rank0": visits,
                   "visits_per_round": a.visits_per_round,
                   "bit_identical_to_replicated_run": same}
        assert same, "sharded automaton differs from the replicated run"
    linked = None
    if a.linked_automaton:
        slabs = ekdist.slab_ranges(occ_z, world)
        t0 = time.perf_counter()
        ekdist.link_model(model, slabs, rank, world, dev)
        t_link = time.perf_counter() - t0
        runs = []
        for _ in range(4):
            tm, info = {}, {}
            _, visits = ekdist.linked_activation(model, slabs, rank, world, dev, timings=tm, download=False, info=info)
            runs.append((ekdist.max_over_ranks(tm["run_s"] * 1e3, dev), ekdist.max_over_ranks(tm["gather_s"] * 1e3, dev),
                         ekdist.max_over_ranks(info["kernel_ms"], dev), ekdist.max_over_ranks(tm["publish_s"] * 1e3, dev)))
        d2 = model.get_activation()
        same = bool(d2.tobytes() == delay.tobytes())
        best = min(runs)
        linked = {"ms_run_launch_to_all_ranks_done": best[0], "ms_gather_over_links": best[1], "kernel_ms_max_over_ranks": best[2],
                  "ms_publish_on_device": best[3], "all_runs_ms": [r[0] for r in runs], "link_setup_ms": t_link * 1e3,
                  "brick_visits_this_rank": visits, "counters_rank0": info, "bit_identical_to_replicated_run": same}
        assert same, "linked automaton differs from the replicated run"
        model.activation_unlink()
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_glue256.npz"))
    k = np.ascontiguousarray(g["layer_k"][:1])
    lead_b = np.ascontiguousarray(leads[None])
    T = 400
    mode = {"direct": ek.MODE_DIRECT, "hoisted": ek.MODE_HOISTED, "separable": ek.MODE_SEPARABLE}[a.mode]
    d_k = torch.from_numpy(k).to(dev)
    d_l = torch.from_numpy(lead_b).to(dev)
    d_e = torch.empty((1, 2, T), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        model.simulate_device(d_k.data_ptr(), d_l.data_ptr(), 1, 2, d_e.data_ptr(), "3D4", 100.0, 1.0, float(T), mode=mode, stream=stream)
        ekdist.allreduce_sum_(d_e)   # per-lead partial sums of all slabs -> ECG (6.4 kB)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = ekdist.max_over_ranks(e0.elapsed_time(e1) / a.steps, dev)
    ecg = d_e.cpu().numpy()[0]
    out = {"workload": "configs[3]: %dx heart, %d occupied voxels, z-slab sharded over %d GPU(s)" % (a.factor, n_occ, world),
           "n_gpus": world, "ms_per_sim": ms, "voxel_timesteps_per_s": n_occ * T / (ms * 1e-3), "mode": a.mode,
           "automaton_ms": auto_ms, "automaton_call_ms_map_stays_on_device": t_auto_call * 1e3, "activation_host_copy_ms": t_download * 1e3,
           "automaton_sweeps": sweeps, "sharded_automaton": sharded, "linked_automaton": linked, "model_create_s": t_create,
           "slab_voxels_rank0": model.num_voxels, "ecg_peak": np.abs(ecg).max(axis=1).tolist()}
    if a.check and rank == 0:
        from oracle import oracle
        t0 = time.time()
        ref_delay = oracle.activation(layers, transfer)
        out["automaton_bit_exact"] = bool(ref_delay.tobytes() == delay.tobytes())
        out["oracle_automaton_s"] = time.time() - t0
        t0 = time.time()
        ref = oracle.run_factored(layers, ref_delay, k[0], leads, "3D4", 100.0, 1.0, 8.0)
        out["oracle_ecg8_s"] = time.time() - t0
        full_peak = np.abs(ecg).max(axis=1, keepdims=True)
        out["ecg_err_of_peak_first8"] = float((np.abs(ecg[:, :8] - ref) / full_peak).max())
    if world > 1:
        torch.distributed.barrier()
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
