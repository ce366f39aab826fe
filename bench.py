#!/usr/bin/env python
"""bench.py -- headline benchmark of the EkgSim hot path on B200 (contract: see task prompt).

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): the reference's
testRun model_24 (124x124x93 grid, 555 868 occupied voxels, 24 layers, 2 leads, 3D4 stencil,
T = 400 samples from 100 ms) evaluated for a batch of 256 seeded parameter vectors per GPU
(weak scaling: every rank evaluates its own 256 individuals against its resident copy of the
model; there is no data-path collective -- the reference farms out individuals the same way,
README.md:126-135).  A "step" is one pass of the hot path over one batch:
    value  = voxel-timesteps/s, inputs (24x9 layer coefficients + lead positions per individual)
             already resident in HBM, ECGs left in HBM            (ekg_simulate_device, DIRECT kernel)
    e2e    = the same through the host-buffer C-ABI call ekg_simulate(): pinned-host -> device copy
             of the step's inputs and device -> host copy of all ECGs inside the timed region.
The layer coefficients are the reference glue's own output for the 256 vectors
(tests/golden/golden_glue256.npz), so the kernel sees exactly what EkgSim::run would.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--mode direct|hoisted]
  python bench.py --impl reference ...   times the reference's own CPU path (oracle/_ref, the
                                         unmodified reference compiled by oracle/Makefile) on a
                                         bounded sample of the same workload, all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_VOX = 555868
T_FULL = 400
SFU_OPS_PER_VTS = 7          # 5 ex2 + 1 lg2 + 1 rcp per voxel-timestep
SFU_LANES_PER_SM = 16        # MUFU ops / clk / SM on sm_100
BYTES_PER_VOXEL = 16         # pos u32 + mask u32 + activation f64, read once per (segment, individual)
METRIC = "voxel_timesteps_per_s"
UNIT = "voxel-timesteps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--mode", default="direct", choices=["direct", "hoisted"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true")
    ap.add_argument("--ref-length", type=int, default=24, help="time samples per reference sample run")
    ap.add_argument("--ref-full", action="store_true",
                    help="cpu_baseline: additionally time ONE process of the reference at the full 400 samples (~210 s)")
    ap.add_argument("--no-heart-cli", action="store_true", help="skip the C++-host leg of the config-4 block (ekgSim -slabs N on a text model)")
    ap.add_argument("--no-heart", action="store_true", help="skip the config-4 block (synthetic finer heart, z-slabs + all-reduce)")
    ap.add_argument("--heart-factor", type=int, default=4)
    ap.add_argument("--heart-steps", type=int, default=3)
    return ap.parse_args()


def workload_config(args, n_gpus):
    return {
        "workload": "configs[2]: batched evaluation of %d parameter vectors per GPU on testRun model_24 "
                    "(555868 voxels x 400 samples x 2 leads, 3D4 stencil)" % args.batch,
        "heart": "model_24", "voxels": N_VOX, "time_samples": T_FULL, "leads": 2,
        "batch_per_gpu": args.batch, "global_batch": args.batch * n_gpus,
        "parallelism": "individuals sharded over %d GPU(s), model replicated, no collective" % n_gpus,
    }


# ---- clocks sampling -----------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # samples under load = upper half (idle samples before/after the region pull the median down)
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}



# ---- ncu evidence: numbers quoted in the JSON line come from the committed summaries, not from constants ----
PROFILE_DIRECT = "profiles/r02_ecg_direct_b256_ncu_full.txt"
PROFILE_FIT = "profiles/r02_fit_b256_ncu_full.txt"


def profile_metrics(rel_path, kernel_substr):
    """Metrics of the first kernel whose name contains `kernel_substr` in a summary written by tools/ncu_summary.py
    (`kernel: <name> ...` followed by `  <metric>  <unit>  <value>` lines).  Byte counts are returned in bytes."""
    path = os.path.join(ROOT, rel_path)
    if not os.path.exists(path):
        return None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    out, on = {}, False
    for ln in open(path):
        if ln.startswith("kernel:"):
            if on:
                break
            on = kernel_substr in ln
            continue
        if on:
            f = ln.split()
            if len(f) >= 2:
                try:
                    v = float(f[-1].replace(",", ""))
                except ValueError:
                    continue
                out[f[0]] = v * scale.get(f[1], 1.0) if len(f) >= 3 else v
    return out or None


# ---- the reference arm / cpu baseline --------------------------------------------------------------
def reference_sample(length, n_proc, vec_lines, keep_dir=None):
    """Runs n_proc concurrent processes of the UNMODIFIED reference CLI (oracle/_ref/ekgSim_ref,
    `test -sim <16 params> -out result`, main.cpp:367-384) in a materialised testRun directory with
    `length = <length>`; these stand in for the reference's MPI worker ranks, which never
    communicate during an evaluation (README.md:128).  Returns (voxel-timesteps/s, seconds list)."""
    import ekgio
    exe = os.path.join(ROOT, "oracle", "_ref", "ekgSim_ref")
    if not os.path.exists(exe):
        raise FileNotFoundError("oracle/_ref/ekgSim_ref is not built (make -C oracle ref needs /root/reference)")
    base = tempfile.mkdtemp(prefix="ekg_ref_")
    try:
        ekgio.materialise_testrun(os.path.join(base, "p0"), length=length)
        procs = []
        for i in range(n_proc):
            d = os.path.join(base, "p%d" % i)
            if i:
                shutil.copytree(os.path.join(base, "p0"), d)
        t0 = time.time()
        for i in range(n_proc):
            d = os.path.join(base, "p%d" % i)
            cmd = [exe, "test", "-sim", vec_lines[i % len(vec_lines)].strip(), "-out", "result"]
            procs.append(subprocess.Popen(cmd, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True))
        secs = []
        for p in procs:
            out = p.communicate()[0]
            m = re.search(r"simulation done in ([0-9.eE+-]+) seconds", out)
            if not m:
                raise RuntimeError("reference run failed: " + out[-300:])
            secs.append(float(m.group(1)))
        wall = time.time() - t0
        vts = n_proc * N_VOX * length / max(secs)
        return vts, secs, wall
    finally:
        shutil.rmtree(base, ignore_errors=True)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vec_lines = open(os.path.join(ROOT, "tests", "golden", "vectors256.txt")).read().strip().split("\n")
    cores = host_cores()
    if args.warmup > 0:  # one short warm-up sample pages the binary and the inputs in
        reference_sample(2, cores, vec_lines)
    vals, times = [], []
    for _ in range(args.steps):
        t0 = time.time()
        v, secs, wall = reference_sample(args.ref_length, cores, vec_lines)
        vals.append(v)
        times.append(time.time() - t0)
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "reference testRun model_24 (compact fixture) + seeded vectors",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": "%d concurrent processes of the unmodified reference CLI (oracle/_ref/ekgSim_ref), one parameter "
                                   "vector each, simulation length cut to %d of 400 samples; throughput from the reference's own "
                                   "'simulation done in' timer" % (cores, args.ref_length)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sims_per_s": value / (N_VOX * T_FULL),
        "gpu_launches": 0,
    }
    print(json.dumps(line))



# ---- BASELINE configs[3]: synthetic finer heart, one simulation, z-slab sharded ---------------------------
def heart_block(args, rank, world, local):
    """ONE simulation of ekgio.scaled_heart(f) (f = 4: 496 x 496 x 372 grid, 35.6 M occupied voxels), the voxels split
    into z-slabs over the ranks (ekg_model_set_slab), the partial ECGs [L][T] summed with ONE all-reduce per simulation
    (no per-timestep halo is needed because V is a closed form of static voxel data).  Strong scaling:
    the same simulation is first timed on this rank's whole model (the N = 1 value, same run, same box).  Parity in the
    run: sha256 of the activation map and the first 16 ECG samples against the oracle's goldens
    (tests/golden/golden_heart<f>x.json).  The automaton runs replicated and -- for N > 1 -- sharded over the slabs."""
    import hashlib
    import numpy as np
    import torch
    import torch.distributed as dist
    import ekgio
    import ekgsim_b200 as ek
    from ekgsim_b200 import dist as ekdist

    f = args.heart_factor
    dev = torch.device("cuda", local)
    layers, transfer, _ = ekgio.scaled_heart(f)
    gpath = os.path.join(ROOT, "tests", "golden", "golden_heart%dx.json" % f)
    gold = json.load(open(gpath)) if os.path.exists(gpath) else None
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_glue256.npz"))
    k = np.ascontiguousarray(g["layer_k"][:1])
    leads = np.ascontiguousarray((g["leads_zyx"][0] * f)[None])
    occ = (layers & 0x0FFF) > 0
    n_occ = int(occ.sum())
    occ_z = occ.sum(axis=(1, 2))
    del occ
    t0 = time.perf_counter()
    model = ek.Model(layers, transfer, device=local)
    create_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    _, visits = model.activation(download=False)
    auto_call_ms = 1e3 * (time.perf_counter() - t0)
    auto_first_ms = model.activation_ms
    auto_all = []
    for _ in range(3):
        model.activation(download=False)
        auto_all.append(model.activation_ms)
    auto_ms = sorted(auto_all)[1]
    out = {"workload": "configs[3]: %dx heart, %d occupied voxels, one simulation, z-slabs over %d GPU(s) + one all-reduce of the "
                       "[2][400] partial ECGs per simulation" % (f, n_occ, world),
           "factor": f, "voxels": n_occ, "n_gpus": world, "model_create_s": create_s,
           "automaton": {"ms": auto_ms, "ms_first_call": auto_first_ms, "first_call_wall_ms": auto_call_ms, "ms_all": auto_all, "brick_visits": visits,
                         "replicated": True, "queue": "time buckets (csrc/automaton.cu)", "timing": "median of 3 calls after the first; device time of the whole call"}}
    if rank == 0 and gold is not None:
        delay = model.get_activation()
        out["automaton_bit_exact"] = bool(hashlib.sha256(delay.tobytes()).hexdigest() == gold["sha256_f64_raster"])
        del delay
    T = T_FULL
    d_k = torch.from_numpy(k).to(dev)
    d_l = torch.from_numpy(leads).to(dev)
    d_e = torch.empty((1, 2, T), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    modes = (("direct", ek.MODE_DIRECT), ("default", ek.MODE_DEFAULT))

    def timed(step, steps):
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        return ekdist.max_over_ranks(e0.elapsed_time(e1) / steps, dev)

    # N = 1 value: the whole model on this rank, no collective
    k1_min, decay_max = ek.coefficient_hints(k)

    def simulate(md):
        if md == ek.MODE_DEFAULT:   # the coefficients came from this host: no read-back of their rates inside the call
            model.simulate_device_hinted(d_k.data_ptr(), d_l.data_ptr(), 1, 2, d_e.data_ptr(), k1_min, decay_max, "3D4", 100.0, 1.0, float(T),
                                         mode=md, stream=stream)
        else:
            model.simulate_device(d_k.data_ptr(), d_l.data_ptr(), 1, 2, d_e.data_ptr(), "3D4", 100.0, 1.0, float(T), mode=md, stream=stream)

    one = {}
    for nm, md in modes:
        one[nm] = timed(lambda: simulate(md), args.heart_steps)
    slabs = ekdist.slab_ranges(occ_z, world)
    z0, z1 = slabs[rank]
    model.set_slab(z0, z1)
    for nm, md in modes:
        def step():
            simulate(md)
            ekdist.allreduce_sum_(d_e)
        ms = timed(step, args.heart_steps)
        ecg = d_e.cpu().numpy()[0]
        r = {"ms_per_sim": ms, "ms_per_sim_1gpu_same_run": one[nm], "efficiency": one[nm] / (world * ms),
             "voxel_timesteps_per_s": n_occ * T / (ms * 1e-3), "kernel": model.last_kernel_name,
             "launches_per_sim": int(model.last_launch_count), "collectives_per_sim": 1 if world > 1 else 0}
        if gold is not None:
            peak = np.array(gold["peak_full"])[:, None]
            err = float((np.abs(ecg[:, :16] - np.array(gold["ecg16"])) / peak).max())
            for t, want in gold["ecg_at_t"].items():
                err = max(err, float((np.abs(ecg[:, int(t)] - np.array(want)) / peak[:, 0]).max()))
            r["ecg_err_of_peak"] = err
        if nm == "direct":
            r["roofline_frac_7op"] = SFU_OPS_PER_VTS * n_occ * T / (ms * 1e-3) / (world * 148 * SFU_LANES_PER_SM * 1.965e9)
        out[nm] = r
    out["ms_per_sim"] = out["direct"]["ms_per_sim"]
    out["efficiency"] = out["direct"]["efficiency"]
    if gold is not None:
        out["ecg_err_of_peak"] = max(out["direct"]["ecg_err_of_peak"], out["default"]["ecg_err_of_peak"])
    if world > 1:
        # the automaton on the sharded model, bits as in the replicated run.  (a) peer-linked: one
        # kernel per rank, face planes and brick pushes go through the neighbours' memory over NVLink, no host round;
        # (b) the host-driven rounds (plane exchange through NCCL point-to-point messages + an all-reduce per round).
        def check_bits(info):
            if rank == 0 and gold is not None:
                d2 = model.get_activation()
                info["bit_exact"] = bool(hashlib.sha256(d2.tobytes()).hexdigest() == gold["sha256_f64_raster"])
        try:
            ekdist.link_model(model, slabs, rank, world, dev)
            linked_ok = True
        except ek.EkgError as e:
            linked_ok = False
            out["automaton_sharded"] = {"linked_unavailable": str(e)}
        ok_all = ekdist.max_over_ranks(0.0 if linked_ok else 1.0, dev) == 0.0
        if ok_all:
            best = None
            for _ in range(3):
                tm, cnt = {}, {}
                _, v = ekdist.linked_activation(model, slabs, rank, world, dev, timings=tm, download=False, info=cnt)
                run_ms = ekdist.max_over_ranks(tm["run_s"] * 1e3, dev)
                if best is None or run_ms < best["ms"]:
                    best = {"ms": run_ms, "kernel_ms_max_over_ranks": ekdist.max_over_ranks(cnt["kernel_ms"], dev),
                            "ms_gather_over_links": ekdist.max_over_ranks(tm["gather_s"] * 1e3, dev), "brick_visits_rank0": v,
                            "bricks_queued_at_neighbours_rank0": cnt["bricks_queued_at_neighbours"],
                            "cells_written_to_neighbours_rank0": cnt["cells_written_to_neighbours"], "host_rounds": 0, "collectives": 0,
                            "how": "peer-linked: in-kernel plane exchange and brick pushes through NVLink peer memory (CUDA IPC), "
                                   "global termination detected by rank 0's first warp"}
            check_bits(best)
            best["speedup_vs_replicated"] = auto_ms / best["ms"]
            out["automaton_sharded"] = best
        if linked_ok:
            model.activation_unlink()
        planes = ekdist.ModelPlanes(model, dev)
        best, info = None, None
        for _ in range(2):
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            tm = {}
            _, rounds, v = ekdist.sharded_activation(planes, slabs, rank, world, timings=tm, download=False)
            torch.cuda.synchronize()
            dt = ekdist.max_over_ranks(1e3 * (time.perf_counter() - t0), dev)
            if best is None or dt < best:
                best, info = dt, {"ms": dt, "ms_rounds": ekdist.max_over_ranks(tm["rounds_s"] * 1e3, dev),
                                  "ms_gather": ekdist.max_over_ranks(tm["gather_s"] * 1e3, dev), "rounds": rounds, "brick_visits_rank0": v}
        check_bits(info)
        info["speedup_vs_replicated"] = auto_ms / info["ms"]
        out["automaton_sharded_host_rounds"] = info
        if "automaton_sharded" not in out or "ms" not in out["automaton_sharded"]:
            out["automaton_sharded"] = dict(info, **out.get("automaton_sharded", {}))
    model.close()
    return out


def heart_cli_block(f, n_gpus):
    """Config 4 as a user of the reference runs it: `ekgSim test -sim <16 params> -out result` in a directory whose
    simulator.ini names the f-times finer heart (a 208 MB text .matrix at f = 4), with `-slabs N`: the product's C++
    host spreads the one model over the N GPUs (z-slabs, peer-linked automaton, slab ECGs added).  The reference needs
    ~114 s for the excitation sequence and ~3.6 h for the simulation of this model on one core."""
    import ekgio
    cli = os.path.join(ROOT, "ekgsim_b200", "bin", "ekgSim")
    vec = "0.00035813,0.0890636,0.0632915,226.183,0.000369406,0.0965625,0.0523254,232.278,0.000710767,0.0720323,0.0187579,200.93,23,22,15,13"
    out = {"command": "ekgSim test -sim <README vector> -out result -slabs %d" % n_gpus, "factor": f}
    with tempfile.TemporaryDirectory(prefix="ekg_heart_cli_") as d:
        t0 = time.time()
        ekgio.materialise_heart(d, f)
        out["write_text_model_s"] = time.time() - t0
        env = dict(os.environ)
        for k in ("EKGSIM_B200_DEVICE", "EKGSIM_B200_DEVICES", "EKGSIM_B200_SLABS"):
            env.pop(k, None)
        env["EKGSIM_B200_CACHE"] = "1"     # the first run leaves a binary side-car of the shape file, the second one uses it
        runs = []
        for slabs, label in ((n_gpus, "first_run_text_model"), (n_gpus, "second_run_binary_side_car"), (1, "one_gpu_second_run")):
            if label == "one_gpu_second_run" and n_gpus == 1:
                continue
            t0 = time.time()
            r = subprocess.run([cli, "test", "-sim", vec, "-out", "result"] + (["-slabs", str(slabs)] if slabs > 1 else []), cwd=d,
                               capture_output=True, text=True, env=env)
            wall = time.time() - t0
            def timer(label_re):
                m = re.search(label_re + r"\s+done \(([0-9.e+-]+)s\)", r.stderr)
                return float(m.group(1)) if m else None
            m = re.search(r" simulation done in ([0-9.e+-]+) seconds", r.stdout)
            c = re.search(r" criteria = <([0-9.e+-]+),([0-9.e+-]+)>, violation = ([0-9.e+-]+)", r.stdout)
            k = re.search(r"one model on (\d+) z-slabs \((\S+) excitation sequence\)", r.stderr)
            runs.append({"run": label, "gpus": slabs, "process_wall_s": wall, "loading_shape_s": timer("loading shape"),
                         "excitation_sequence_s": timer("calculating excitation sequence"),
                         "eval_s_fit_plus_simulation_plus_criteria": float(m.group(1)) if m else None,
                         "criteria": [float(c.group(1)), float(c.group(2))] if c else None,
                         "slabs": int(k.group(1)) if k else 1, "automaton": k.group(2) if k else "single device",
                         "ok": bool(r.returncode == 0 and m and c)})
            if not runs[-1]["ok"]:
                runs[-1]["tail"] = (r.stdout + r.stderr)[-400:]
        out["runs"] = runs
        cs = [x["criteria"] for x in runs if x["criteria"]]
        if len(cs) > 1:
            out["criteria_max_abs_diff_between_runs"] = max(abs(a - b) for c in cs[1:] for a, b in zip(c, cs[0]))
    return out



# ---- BASELINE configs[4]: an optimizer generation through the evaluation boundary ------------------------
def generation_block(n_gpus):
    """Population 100 against target_ecg_v2_v6.column on the N GPUs of the box, through the
    reference-facing boundaries: (a) `ekgSim -batch pop.txt -devices N`: the generation as ONE batch split over the GPUs;
    (b) the reference's UNMODIFIED stand-alone optimizer (oracle/_ref/DEMO_ref, AMS-DEMO/main.cpp, sequential build: one
    ExternalEvaluation at a time) driving `ekgSim -extern` against a resident `ekgSim -serve -devices N`.  The reference
    itself needs 100 x 205 s of CPU per generation."""
    import ekgio
    out = {"population": 100, "n_gpus": n_gpus, "reference_cpu_core_seconds_per_generation": 100 * 205.28}
    cli = os.path.join(ROOT, "ekgsim_b200", "bin", "ekgSim")
    d = tempfile.mkdtemp(prefix="ekg_gen_")
    try:
        ekgio.materialise_testrun(d, targets="target_ecg_v2_v6.column")
        vec = open(os.path.join(ROOT, "tests", "golden", "vectors256.txt")).read().strip().split("\n")[:100]
        open(os.path.join(d, "pop.txt"), "w").write("\n".join(vec) + "\n")
        env = dict(os.environ)
        env.pop("EKGSIM_B200_DEVICE", None)
        t0 = time.time()
        r = subprocess.run([cli, "-batch", "pop.txt", "-batchout", "crit.txt", "-devices", str(n_gpus)], cwd=d, capture_output=True, text=True, env=env)
        dt = time.time() - t0
        m = re.search(r"batch of (\d+) simulations done in ([0-9.e+-]+) seconds on (\d+) GPU", r.stdout)
        out["batched_cli"] = {"process_wall_s": dt, "s_per_generation_cold": float(m.group(2)) if m else None, "devices": int(m.group(3)) if m else None,
                              "note": "a fresh process: the one batch it evaluates also pays the first-use allocations (pinned staging, scratch) on every device"}
        # the same generation inside a resident evaluator (what `ekgSim -serve` or an in-process optimizer sees from the second
        # generation on): Evaluator::evalBatch on all N devices, best of 5
        import numpy as np
        import hostlib
        genes = np.array([[float(x) for x in ln.replace(",", " ").split()] for ln in vec])
        os.environ.pop("EKGSIM_B200_DEVICE", None)
        ev = hostlib.Evaluator(d, devices=str(n_gpus))
        ev.eval_batch(genes)
        best = 1e9
        for _ in range(5):
            t0 = time.perf_counter()
            ev.eval_batch(genes)
            best = min(best, time.perf_counter() - t0)
        ev.close()
        out["resident_evaluator"] = {"s_per_generation": best, "sims_per_s": len(genes) / best, "devices": n_gpus,
                                     "api": "Evaluator::evalBatch (100 individuals, v6 target) on %d device(s), warm" % n_gpus}
        demo = os.path.join(ROOT, "oracle", "_ref", "DEMO_ref")
        if os.path.exists(demo):
            open(os.path.join(d, "settings.ini"), "w").write("[evaluation]\ncommand line = %s -extern\ninput file name = input.txt\n"
                "output file name = output.txt\nchromosome vector length = 16\ncriteria vector length = 2\nproperties vector length = 0\n\n"
                "[optimization]\nrandom seed = 11\npopulation size = 100\nmax number of generations = 2\nDE schema = rand/1/bin\n"
                "p crossover = 0.3\nscaling factors = 0.5\nqueue length = 1\n\n[initial population]\n"
                "gene min = 0.0003, 0.01, 0.01, 200, 0.0003, 0.01, 0.01, 200, 0.0003, 0.01, 0.01, 200, -50, -50, -50, -50\n"
                "gene max = 0.001, 0.1, 0.1, 400, 0.001, 0.1, 0.1, 400, 0.001, 0.1, 0.1, 400, 50, 50, 50, 50\n" % cli)
            sock = os.path.join(d, "ekg.sock")
            srv = subprocess.Popen([cli, "-serve", sock, "-devices", str(n_gpus)], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=env)
            for _ in range(900):
                if os.path.exists(sock) or srv.poll() is not None:
                    break
                time.sleep(0.1)
            t0 = time.time()
            r = subprocess.run([demo], cwd=d, capture_output=True, text=True, env=dict(env, EKGSIM_B200_SERVER=sock), timeout=300)
            dt = time.time() - t0
            subprocess.run([cli, "-shutdown", sock], cwd=d, capture_output=True)
            err = srv.communicate(timeout=60)[1]
            n_eval = len([ln for ln in open(os.path.join(d, "evaluations.txt")) if ln.strip() and not ln.startswith("#")])
            out["reference_ams_demo_via_extern_server"] = {"evaluations": n_eval, "wall_s": dt, "s_per_generation_of_100": dt * 100 / max(n_eval, 1),
                                                           "ms_per_evaluation": 1e3 * dt / max(n_eval, 1), "ok": "front.txt" in r.stdout,
                                                           "note": "the sequential optimizer build evaluates one individual at a time (no MPI runtime in "
                                                                   "this image): a process start + socket round trip per evaluation, one GPU busy at a time",
                                                           "server": err.strip().split("\n")[-1] if err.strip() else ""}
    finally:
        shutil.rmtree(d, ignore_errors=True)
    return out


# ---- our arm ---------------------------------------------------------------------------------------
def run_b200_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import ekgio
    import ekgsim_b200 as ek

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # libraries (NCCL's version banner, the facade's console lines) write to fd 1; keep stdout for the
    # ONE JSON line by pointing fd 1 at stderr until the result is printed
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    m24 = ekgio.load_model24()
    model = ek.Model(m24["layers"], m24["transfer"], device=local)
    delay, sweeps = model.activation()
    automaton_first_ms = model.activation_ms      # the first launch of the process (module load, cold instruction cache)
    automaton_all = []
    for _ in range(5):
        model.activation(download=False)
        automaton_all.append(model.activation_ms)
    automaton_ms = sorted(automaton_all)[len(automaton_all) // 2]
    fp = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_activation.json")))
    import hashlib
    act_ok = hashlib.sha256(delay.tobytes()).hexdigest() == fp["sha256_f64_raster"]

    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_glue256.npz"))
    B = args.batch
    idx = (np.arange(B) + rank * B) % g["layer_k"].shape[0]
    layer_k = np.ascontiguousarray(g["layer_k"][idx])          # [B,24,9]
    leads = np.ascontiguousarray(g["leads_zyx"][idx])          # [B,2,3]
    L = leads.shape[1]
    mode = ek.MODE_DIRECT if args.mode == "direct" else ek.MODE_HOISTED

    d_k = torch.from_numpy(layer_k).to(dev)
    d_leads = torch.from_numpy(leads).to(dev)
    d_ecg = torch.empty((B, L, T_FULL), dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream().cuda_stream

    def step_resident(timed_kernel=False):
        flush.fill_(1)  # L2 flush between iterations (inputs are smaller than L2)
        model.simulate_device(d_k.data_ptr(), d_leads.data_ptr(), B, L, d_ecg.data_ptr(), "3D4", 100.0, 1.0, float(T_FULL),
                              mode=mode | (ek.FLAG_TIME_KERNEL if timed_kernel else 0), stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # While rank 0 alone drives ALL GPUs of the box (the C++ multi-device evaluator, the generation block) the other ranks
    # must not wait inside an NCCL barrier: that is a kernel spinning on their GPU, and kernels of two processes on one
    # GPU are time-sliced, not run side by side.  They wait on the CPU (gloo) instead.
    cpu_group = dist.new_group(backend="gloo") if world > 1 else None

    def idle_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)

    # -- resident timing: K steps, CUDA events on the launching stream, barrier+sync on both sides
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step_resident(timed_kernel=True)
        # the kernel events are only READ after the loop for all but the last step would be lost, so
        # collect per step (cudaEventSynchronize on the kernel's own end event; the stream keeps going)
        kernel_ms.append(model.last_kernel_ms)
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    direct_kernel_name = model.last_kernel_name
    direct_launches = int(model.last_launch_count)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    vts_step_global = world * B * N_VOX * T_FULL
    value = vts_step_global / (ms_per_step * 1e-3)

    # -- e2e: host buffers through the C-ABI entry point EkgSim::run maps to
    for _ in range(2):
        model.simulate(layer_k, leads, "3D4", 100.0, 1.0, float(T_FULL), mode=mode)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        ecg_host = model.simulate(layer_k, leads, "3D4", 100.0, 1.0, float(T_FULL), mode=mode)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = vts_step_global * args.steps / e2e_s
    h2d = layer_k.nbytes + leads.nbytes
    d2h = ecg_host.nbytes
    clocks = sampler.stop() if rank == 0 else None

    # -- the library's default kernel variant (HOISTED: voxel- and time-invariant AP factors hoisted,
    #    sigmoid skipped where it is exactly 1 in fp32), same workload, resident inputs
    fast = None
    if mode == ek.MODE_DIRECT:
        def step_fast():
            flush.fill_(1)
            model.simulate_device(d_k.data_ptr(), d_leads.data_ptr(), B, L, d_ecg.data_ptr(), "3D4", 100.0, 1.0, float(T_FULL),
                                  mode=ek.MODE_HOISTED, stream=stream)
        for _ in range(3):
            step_fast()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            step_fast()
        f1.record()
        barrier()
        fms = f0.elapsed_time(f1) / args.steps
        t = torch.tensor([fms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        fms = float(t.item())
        fast = {"kernel": "ecg_kernel<HOISTED>", "ms_per_step": fms, "value": vts_step_global / (fms * 1e-3), "unit": UNIT,
                "sims_per_s": vts_step_global / (fms * 1e-3) / (N_VOX * T_FULL),
                "note": "same results within tolerance; executes 2 (or 0) MUFU ops per voxel-timestep instead of 7, so the "
                        "7-op roofline accounting does not apply to it"}

    # -- the library's DEFAULT path (SEPARABLE): the run starts after the last depolarisation, so the sum over
    #    voxels leaves the time loop (per-layer moments of the lead field); same workload, same results within
    #    tolerance, O(voxels) instead of O(voxels x samples) work -> reported as sims/s
    separable = None
    if mode == ek.MODE_DIRECT:
        def step_sep(timed=False):
            flush.fill_(1)
            model.simulate_device(d_k.data_ptr(), d_leads.data_ptr(), B, L, d_ecg.data_ptr(), "3D4", 100.0, 1.0, float(T_FULL),
                                  mode=ek.MODE_SEPARABLE | (ek.FLAG_TIME_KERNEL if timed else 0), stream=stream)
        for _ in range(3):
            step_sep()
        barrier()
        # A/B of the moment kernel: interior voxels through the direct 8-corner sum (EKG_FLAG_CORNER_SUM) instead of the series
        corner_sum_ms = []
        for i in range(3 + args.steps):
            flush.fill_(1)
            model.simulate_device(d_k.data_ptr(), d_leads.data_ptr(), B, L, d_ecg.data_ptr(), "3D4", 100.0, 1.0, float(T_FULL),
                                  mode=ek.MODE_SEPARABLE | ek.FLAG_TIME_KERNEL | ek.FLAG_CORNER_SUM, stream=stream)
            if i >= 3:
                corner_sum_ms.append(model.last_kernel_ms)
        ecg_corner_sum = d_ecg.cpu().numpy()
        barrier()
        # a step is ~0.7 ms, the 256 MiB L2 flush before it ~0.07 ms: every step gets its own event pair, the flush stays
        # between the timed intervals
        sep_kernel_ms, sep_events = [], []
        for _ in range(args.steps):
            flush.fill_(1)
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            model.simulate_device(d_k.data_ptr(), d_leads.data_ptr(), B, L, d_ecg.data_ptr(), "3D4", 100.0, 1.0, float(T_FULL),
                                  mode=ek.MODE_SEPARABLE | ek.FLAG_TIME_KERNEL, stream=stream)
            f1.record()
            sep_events.append((f0, f1))
            sep_kernel_ms.append(model.last_kernel_ms)
        barrier()
        sms = sum(a.elapsed_time(b) for a, b in sep_events) / args.steps
        t = torch.tensor([sms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sms = float(t.item())
        sep_launches = int(model.last_launch_count)
        sep_kernel = model.last_kernel_name
        # the same with the caller's knowledge of the coefficients (ekg_simulate_device_hinted): no read-back inside the call
        sms_hinted = None
        try:
            k1_min_b, decay_max_b = ek.coefficient_hints(layer_k)
            hint_events = []
            for i in range(2 + args.steps):
                flush.fill_(1)
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record()
                model.simulate_device_hinted(d_k.data_ptr(), d_leads.data_ptr(), B, L, d_ecg.data_ptr(), k1_min_b, decay_max_b, "3D4", 100.0, 1.0,
                                             float(T_FULL), mode=ek.MODE_SEPARABLE, stream=stream)
                f1.record()
                if i >= 2:
                    hint_events.append((f0, f1))
            torch.cuda.synchronize()
            sms_hinted = sum(a.elapsed_time(b) for a, b in hint_events) / len(hint_events)
        except Exception:
            sms_hinted = None
        # through the host-buffer C-ABI call in its default mode (what EkgSim::run / runBatch issue)
        for _ in range(2):
            model.simulate(layer_k, leads, "3D4", 100.0, 1.0, float(T_FULL), mode=ek.MODE_DEFAULT)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            flush.fill_(1)
            model.simulate(layer_k, leads, "3D4", 100.0, 1.0, float(T_FULL), mode=ek.MODE_DEFAULT)
        torch.cuda.synchronize()
        sep_e2e = (time.perf_counter() - t0) / args.steps
        t = torch.tensor([sep_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sep_e2e = float(t.item())
        # moment kernel, per (voxel, vector) at 2 leads and the 3D4 stencil:
        #   boundary voxel (a corner missing): 9 lead-field evaluations -> ~100 packed fp32x2 operations = 200 fp32 lane-operations
        #                                      (an FFMA2 occupies the FMA pipe for two cycles), 16-18 rsqrt + 2 ex2
        #   interior voxel (all 8 corners):    the series of the corner sum -> 30 packed + 8 scalar = 68 lane-operations, 2 rsqrt + 2 ex2
        #   (the two kinds sit in segments of their own: ecg_moment_interior_kernel + ecg_moment_corners_kernel; moment_kernel_ms
        #   is the device time of both launches)
        occ = (m24["layers"] & 0x0FFF) > 0
        inner = occ.copy()
        pad = np.pad(occ, 1)
        Zs, Ys, Xs = occ.shape
        for dz in (0, 2):
            for dy in (0, 2):
                for dx in (0, 2):
                    inner &= pad[dz:dz + Zs, dy:dy + Ys, dx:dx + Xs]
        f_int = float(inner.sum()) / float(occ.sum())
        lane_ops = f_int * 68 + (1 - f_int) * 200
        mufu_ops = f_int * 4 + (1 - f_int) * 20
        mk = sum(sep_kernel_ms) / len(sep_kernel_ms)
        mk_sum = sum(corner_sum_ms) / len(corner_sum_ms)
        ecg_series = d_ecg.cpu().numpy()
        separable = {"kernel": sep_kernel, "ms_per_step": sms, "sims_per_s": world * B / (sms * 1e-3),
                     "ms_per_step_hinted": sms_hinted, "hinted_note": "ekg_simulate_device_hinted: the caller passes the batch's smallest k1 and largest "
                     "decay rate, the call does not read them back from the device (this rank's time)",
                     "equivalent_voxel_timesteps_per_s": vts_step_global / (sms * 1e-3),
                     "moment_kernel_ms": mk, "launches_per_step": sep_launches,
                     "interior_voxel_fraction": f_int,
                     "moment_kernel_mufu_gops": mufu_ops * B * N_VOX / (mk * 1e-3) / 1e9,
                     "moment_kernel_fp32_lane_gops": lane_ops * B * N_VOX / (mk * 1e-3) / 1e9,
                     "moment_kernel_ms_with_direct_corner_sum": mk_sum,
                     "series_vs_corner_sum_max_diff_of_peak": float((np.abs(ecg_series - ecg_corner_sum) / np.abs(ecg_series).max(axis=-1, keepdims=True)).max()),
                     "e2e_ms_per_step": 1e3 * sep_e2e, "e2e_sims_per_s": world * B / sep_e2e,
                     "e2e_api": "ekg_simulate (C ABI, host buffers, EKG_MODE_DEFAULT)",
                     "note": "valid because every sample of the run is later than the last activation time + 25/(k1 log2 e) "
                             "(the HOISTED kernel's own saturation test); earlier samples would go through the time loop"}

    # -- BASELINE configs[0]: ONE simulation (B = 1), device time of the C-ABI call with resident inputs
    single = {}
    for nm, md in (("direct", ek.MODE_DIRECT), ("hoisted", ek.MODE_HOISTED), ("separable", ek.MODE_SEPARABLE)):
        for _ in range(3):
            model.simulate_device(d_k.data_ptr(), d_leads.data_ptr(), 1, L, d_ecg.data_ptr(), "3D4", 100.0, 1.0, float(T_FULL), mode=md, stream=stream)
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(20):
            model.simulate_device(d_k.data_ptr(), d_leads.data_ptr(), 1, L, d_ecg.data_ptr(), "3D4", 100.0, 1.0, float(T_FULL), mode=md, stream=stream)
        s1.record()
        torch.cuda.synchronize()
        single[nm + "_ms"] = s0.elapsed_time(s1) / 20
    # the default mode once more with the caller's knowledge of the coefficients (ekg_simulate_device_hinted: no read-back of
    # the batch's rates in the middle of the call -- what ekg_simulate with host buffers does internally)
    try:
        k1_min, decay_max = ek.coefficient_hints(layer_k[:1])
        def hinted():
            model.simulate_device_hinted(d_k.data_ptr(), d_leads.data_ptr(), 1, L, d_ecg.data_ptr(), k1_min, decay_max, "3D4", 100.0, 1.0,
                                         float(T_FULL), mode=ek.MODE_SEPARABLE, stream=stream)
        for _ in range(3):
            hinted()
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(20):
            hinted()
        s1.record()
        torch.cuda.synchronize()
        single["separable_hinted_ms"] = s0.elapsed_time(s1) / 20
    except Exception as e:
        single["separable_hinted_error"] = repr(e)
    single["reference_cpu_s"] = 205.28  # SURVEY.md section 6, measured with the compiled reference on one core

    # -- whole pipeline, parameter vectors -> criteria: Evaluator::evalBatch = border APs on the host +
    #    ekg_evaluate (layer fit, simulation in the default mode, curve comparison on the device), B x 2 criteria back.
    #    Every rank evaluates its own 256 vectors on its own GPU (weak, like `value`); then rank 0 alone evaluates ONE
    #    batch of 256 through the product's C++ multi-device evaluator on all N GPUs of the box (config 3 as worded:
    #    "256 vectors sharded across 1/2/4/8 B200" -- strong scaling, 32 vectors per GPU at N = 8).
    pipeline = None
    if not args.no_pipeline:
        try:
            import tempfile
            import hostlib
            wd = tempfile.mkdtemp(prefix="ekg_pipe_")
            ekgio.materialise_testrun(wd)
            os.environ["EKGSIM_B200_DEVICE"] = str(local)
            ev = hostlib.Evaluator(wd, with_device=True)
            pipeline = {"api": "Evaluator::evalBatch (parameter vectors -> border APs on the host -> ekg_evaluate: layer fit, "
                               "simulation in the default mode, curve comparison -> B x 2 criteria back to the host)", "host_threads": host_cores()}
            for pb in sorted({B, 1024} if world == 1 else {B}):
                genes = np.tile(g["params"], ((pb + 255) // 256, 1))[:pb]
                ev.eval_batch(genes)
                best = 1e9
                for _ in range(3):
                    barrier()
                    t0 = time.perf_counter()
                    crit, viol = ev.eval_batch(genes)
                    dt = time.perf_counter() - t0
                    t = torch.tensor([dt], dtype=torch.float64, device=dev)
                    if world > 1:
                        dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    best = min(best, float(t.item()))
                pipeline["batch_%d" % pb] = {"seconds": best, "sims_per_s": world * pb / best, "per_gpu": pb}
            pipeline["sims_per_s"] = pipeline["batch_%d" % B]["sims_per_s"]
            t0 = time.perf_counter()
            for i in range(20):
                ev.eval(g["params"][i])
            pipeline["single_eval_ms"] = 1e3 * (time.perf_counter() - t0) / 20
            ev.close()
            # the layer fit alone (the dominant kernel of this pipeline), device time at B = 256 and B = 1
            nb = g["layer_k"].shape[1]
            border = np.ascontiguousarray(g["layer_k"][:, [0, 14, nb - 1]])
            d_b = torch.from_numpy(border).to(dev)
            d_lk = torch.empty((border.shape[0], nb, 9), dtype=torch.float64, device=dev)
            fit_ms = {}
            for fb in (1, 256):
                for _ in range(2):
                    model.fit_layers_device(d_b.data_ptr(), fb, 3, d_lk.data_ptr(), mid=14, stream=stream)
                torch.cuda.synchronize()
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record()
                for _ in range(5):
                    model.fit_layers_device(d_b.data_ptr(), fb, 3, d_lk.data_ptr(), mid=14, stream=stream)
                f1.record()
                torch.cuda.synchronize()
                fit_ms["B%d" % fb] = f0.elapsed_time(f1) / 5
            pipeline["fit_ms"] = fit_ms
            pipeline["fit_bit_identical_to_reference_glue"] = bool(d_lk.cpu().numpy().tobytes() == np.ascontiguousarray(g["layer_k"]).tobytes())
            idle_barrier()
            if rank == 0:
                # strong scaling of ONE 256-vector batch over the N GPUs of the box, inside one process (C++ host threads)
                evN = hostlib.Evaluator(wd, devices=str(world))
                genes = g["params"][:256]
                evN.eval_batch(genes)
                best = 1e9
                for _ in range(5):
                    t0 = time.perf_counter()
                    critN, violN = evN.eval_batch(genes)
                    best = min(best, time.perf_counter() - t0)
                want = np.load(os.path.join(ROOT, "tests", "golden", "golden_criteria256.npz"))
                pipeline["strong_256"] = {"api": "Evaluator::evalBatch on %d device(s) from one process (`ekgSim -batch -devices %d`)" % (evN.n_devices, world),
                                          "devices": evN.n_devices, "vectors_per_gpu": 256 // max(world, 1), "seconds": best, "sims_per_s": 256 / best,
                                          "criteria_max_abs_diff_vs_reference_pinned_oracle": float(np.abs(critN - want["criteria"]).max())}
                evN.close()
            idle_barrier()
        except Exception as e:
            pipeline = {"error": str(e)}

    # -- BASELINE configs[3]: the synthetic finer heart, z-slab sharded (all ranks)
    heart = None
    if not args.no_heart:
        try:
            heart = heart_block(args, rank, world, local)
        except Exception as e:
            heart = {"error": repr(e)}
        barrier()
        # the same configuration through the C++ host alone (rank 0 drives all N GPUs from one child process; the other ranks
        # wait on the CPU so that no barrier kernel of theirs shares a GPU with it)
        idle_barrier()
        if rank == 0 and not args.no_heart_cli and isinstance(heart, dict) and "error" not in heart:
            try:
                heart["cxx_host"] = heart_cli_block(args.heart_factor, world)
            except Exception as e:   # the CLI leg is extra evidence, never a reason to lose the line
                heart["cxx_host"] = {"error": "%s: %s" % (type(e).__name__, e)}
        idle_barrier()

    # -- BASELINE configs[4]: one AMS-DEMO generation (population 100) through the evaluation boundary, all N GPUs (rank 0)
    generation = None
    idle_barrier()
    if rank == 0 and not args.no_pipeline:
        try:
            generation = generation_block(world)
        except Exception as e:
            generation = {"error": repr(e)}
    idle_barrier()

    # -- parity spot check inside the bench (first vectors against the reference-pinned goldens)
    gf = np.load(os.path.join(ROOT, "tests", "golden", "golden_eval_full.npz"))
    chk = model.simulate(gf["layer_k"], gf["leads_zyx"], "3D4", 100.0, 1.0, float(T_FULL), mode=mode)
    parity = float((np.abs(chk - gf["ecg"]) / np.abs(gf["ecg"]).max(axis=2, keepdims=True)).max())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # -- roofline of the dominant kernel (ecg_kernel): MUFU pipe
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    k_ms = sum(kernel_ms) / len(kernel_ms)
    sm_count = torch.cuda.get_device_properties(local).multi_processor_count
    sm_mhz = clocks["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)
    vts_launch = B * N_VOX * T_FULL
    achieved_gops = SFU_OPS_PER_VTS * vts_launch / (k_ms * 1e-3) / 1e9
    peak_gops = sm_count * SFU_LANES_PER_SM * sm_mhz * 1e6 / 1e9
    peak_gops_max = sm_count * SFU_LANES_PER_SM * (clocks["sm_max_mhz"] or peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e9
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    alg_bytes = B * N_VOX * BYTES_PER_VOXEL + 2 * model_partial_bytes(B, L)
    traffic = None
    if B == 256 and mode == ek.MODE_DIRECT:
        pm = profile_metrics(PROFILE_DIRECT, "ecg_kernel")
        if pm and "dram__bytes_read.sum" in pm:
            traffic = pm["dram__bytes_read.sum"] + pm.get("dram__bytes_write.sum", 0.0)
    roofline = {
        "bound": "sfu", "kernel": direct_kernel_name,
        "achieved": achieved_gops, "peak": peak_gops, "unit": "Gop/s (MUFU)", "frac": achieved_gops / peak_gops,
        "peak_basis": "148 SM x 16 MUFU lanes/clk x SM clock sampled under load (%s MHz); at clocks.max.sm the peak is %.0f Gop/s "
                      "-> frac %.3f" % (sm_mhz, peak_gops_max, achieved_gops / peak_gops_max),
        "algorithmic_ops_per_voxel_timestep": SFU_OPS_PER_VTS,
        "executed_on_xu_pipe_per_voxel_timestep": 6 if mode == ek.MODE_DIRECT else 2,
        "note": "DIRECT evaluates all 7 transcendental operations per voxel-timestep; the reciprocal of the depolarisation "
                "sigmoid runs as a Newton iteration on the FMA pipe, the other 6 on the MUFU/XU pipe",
        "kernel_ms_per_launch": k_ms, "kernel_share_of_step": k_ms / ms_per_step,
        # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this launch at B = 256, parsed from
        # the committed summary (null when the batch / mode differs from the captured one or the file is absent)
        "traffic": traffic, "traffic_source": PROFILE_DIRECT if traffic is not None else None,
        "hbm": {"achieved_gbs": alg_bytes / (k_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback",
                "note": "algorithmic bytes: 16 B/voxel per individual + f64 partials; the kernel is MUFU-bound, not HBM-bound"},
    }

    # config 1 (ONE simulation, the north_star's Target): the same 7-op accounting on the B = 1 device time
    single["roofline"] = {"bound": "sfu", "kernel": "ecg_kernel<DIRECT> (+ params + reduce launches)",
                          "achieved": SFU_OPS_PER_VTS * N_VOX * T_FULL / (single["direct_ms"] * 1e-3) / 1e9, "peak": peak_gops_max,
                          "unit": "Gop/s (MUFU)", "frac": SFU_OPS_PER_VTS * N_VOX * T_FULL / (single["direct_ms"] * 1e-3) / 1e9 / peak_gops_max,
                          "peak_basis": "148 SM x 16 lanes x clocks.max.sm (a 0.4 ms call is too short for the clock sampler)"}
    if pipeline and "fit_ms" in pipeline:
        # dominant kernel of the genes -> criteria pipeline: fit_descent_kernel, f64 on the FP64 pipe (16 lanes per SM
        # sub-partition: one warp instruction per 2 cycles).  Instruction counts from the committed ncu summary.
        pf = profile_metrics(PROFILE_FIT, "fit_descent_kernel")
        fit_roof = {"bound": "fp64", "kernel": "fit_descent_kernel", "ms_B256": pipeline["fit_ms"]["B256"], "ms_B1": pipeline["fit_ms"]["B1"],
                    "share_of_pipeline_batch_256": pipeline["fit_ms"]["B256"] * 1e-3 / pipeline["batch_%d" % B]["seconds"] if B == 256 else None,
                    "source": PROFILE_FIT if pf else None}
        if pf and "smsp__inst_executed_pipe_fp64.sum" in pf:
            n64 = pf["smsp__inst_executed_pipe_fp64.sum"]
            fit_roof.update(achieved=n64 / (pipeline["fit_ms"]["B256"] * 1e-3) / 1e9, peak=sm_count * 4 * 0.5 * sm_mhz * 1e6 / 1e9,
                            unit="G warp-instructions/s (FP64 pipe)")
            fit_roof["frac"] = fit_roof["achieved"] / fit_roof["peak"]
        pipeline["roofline"] = fit_roof
    if separable:
        fp32_peak = sm_count * 128 * sm_mhz * 1e6 / 1e9
        separable["moment_kernel_roofline"] = {"bound": "fp32", "achieved": separable["moment_kernel_fp32_lane_gops"], "peak": fp32_peak,
                                               "unit": "G lane-op/s (FMA pipe)", "frac": separable["moment_kernel_fp32_lane_gops"] / fp32_peak,
                                               "mufu_frac": separable["moment_kernel_mufu_gops"] / peak_gops}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "dtype_note": "f32 AP evaluation and lead-field coefficients, compensated (f64-grade) accumulation, f64 partial sums and outputs",
        "data": "reference testRun model_24 (compact fixture) + 256 seeded vectors "
        "(layer coefficients from the reference glue); random-free, no checkpoint needed",
        "config": dict(workload_config(args, world), l2="flushed between iterations (256 MiB fill)", ecg_mode=args.mode),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * e2e_s / args.steps, "api": "ekg_simulate (C ABI, host buffers)"},
        "gpu_launches": direct_launches * args.steps,
        "launches_per_step": direct_launches,
        "sims_per_s": value / (N_VOX * T_FULL), "e2e_sims_per_s": e2e_value / (N_VOX * T_FULL),
        "roofline": roofline, "clocks": clocks,
        "automaton": {"ms": automaton_ms, "ms_first_call_of_the_process": automaton_first_ms, "ms_all": automaton_all,
                      "timing": "median of 5 calls of ekg_model_activation after the first one; device time (CUDA events) of the whole "
                                "call: grid and queue initialisation + the frontier kernel",
                      "kernel": "automaton_brick_kernel (4^3-brick frontier; FIFO work ring on model_24, time-bucket queue on large models)", "brick_visits": sweeps,
                      "bit_exact_vs_reference": bool(act_ok), "edges_per_s": 26 * N_VOX / (automaton_ms * 1e-3),
                      # 8 B neighbour time + 1 B layer read + 8 B atomicMin per relaxed edge, every edge relaxed once; the
                      # kernel is bound by the wave's dependency chain, not by bandwidth -- reported for honesty
                      "hbm": {"algorithmic_bytes": 17 * 26 * N_VOX, "achieved_gbs": 17 * 26 * N_VOX / (automaton_ms * 1e-3) / 1e9,
                              "peak_gbs": json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs") if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else None},
                      "reference_cpu_s": 2.0},
        "parity_max_err_of_peak": parity,
        "fast_path": fast, "separable_path": separable, "pipeline": pipeline, "single_sim": single,
        "heart4x": heart, "generation": generation,
    }

    if not args.no_cpu_baseline and world == 1:
        try:
            vec_lines = open(os.path.join(ROOT, "tests", "golden", "vectors256.txt")).read().strip().split("\n")
            cores = host_cores()
            v, secs, wall = reference_sample(args.ref_length, cores, vec_lines)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "reference",
                                    "sample": "%d concurrent processes of the unmodified reference CLI, one vector each, "
                                              "length %d of 400 samples (%.1f s wall)" % (cores, args.ref_length, wall)}
            if args.ref_full:   # one process, all 400 samples: the reference's ~5 s of fixed cost per run fully amortised
                vf, sf, wf = reference_sample(T_FULL, 1, vec_lines)
                line["cpu_baseline"]["full_length_one_process"] = {"seconds": sf[0], "voxel_timesteps_per_s_per_core": vf, "sims_per_s_per_core": 1.0 / sf[0]}
        except Exception as e:  # the reference binary is a prebuilt artefact; say so instead of failing the bench
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def model_partial_bytes(B, L):
    # partial sums written once and read once by the reduce kernel; segment count is chosen by the
    # library (about 2 per layer at B = 256), so this is an estimate used only for the HBM side note
    return 48 * B * L * T_FULL * 8


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_b200_arm(a)
