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
    automaton_ms = model.activation_ms
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
    single["reference_cpu_s"] = 205.28  # SURVEY.md section 6, measured with the compiled reference on one core

    # -- whole pipeline, parameter vectors -> criteria: host glue (C++, threads) + one GPU batch
    pipeline = None
    if world == 1 and not args.no_pipeline:
        try:
            import tempfile
            import hostlib
            wd = tempfile.mkdtemp(prefix="ekg_pipe_")
            ekgio.materialise_testrun(wd)
            ev = hostlib.Evaluator(wd, with_device=True)
            pipeline = {"api": "Evaluator::evalBatch (parameter vectors -> border APs on the host -> ekg_evaluate: layer fit, "
                               "simulation in the default mode, curve comparison -> B x 2 criteria back to the host)", "host_threads": host_cores()}
            for pb in sorted({B, 1024}):
                genes = np.tile(g["params"], ((pb + 255) // 256, 1))[:pb]
                ev.eval_batch(genes)
                best = 1e9
                for _ in range(3):
                    t0 = time.perf_counter()
                    crit, viol = ev.eval_batch(genes)
                    best = min(best, time.perf_counter() - t0)
                pipeline["batch_%d" % pb] = {"seconds": best, "sims_per_s": pb / best}
            pipeline["sims_per_s"] = pipeline["batch_%d" % B]["sims_per_s"]
            t0 = time.perf_counter()
            for i in range(20):
                ev.eval(g["params"][i])
            pipeline["single_eval_ms"] = 1e3 * (time.perf_counter() - t0) / 20
            ev.close()
        except Exception as e:
            pipeline = {"error": str(e)}

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
    roofline = {
        "bound": "sfu", "kernel": model.last_kernel_name,
        "achieved": achieved_gops, "peak": peak_gops, "unit": "Gop/s (MUFU)", "frac": achieved_gops / peak_gops,
        "peak_basis": "148 SM x 16 MUFU lanes/clk x SM clock sampled under load (%s MHz); at clocks.max.sm the peak is %.0f Gop/s "
                      "-> frac %.3f" % (sm_mhz, peak_gops_max, achieved_gops / peak_gops_max),
        "algorithmic_ops_per_voxel_timestep": SFU_OPS_PER_VTS,
        "executed_on_xu_pipe_per_voxel_timestep": 6 if mode == ek.MODE_DIRECT else 2,
        "note": "DIRECT evaluates all 7 transcendental operations per voxel-timestep; the reciprocal of the depolarisation "
                "sigmoid runs as a Newton iteration on the FMA pipe, the other 6 on the MUFU/XU pipe",
        "kernel_ms_per_launch": k_ms, "kernel_share_of_step": k_ms / ms_per_step,
        # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this launch at
        # B = 256 (profiles/r01_ecg_direct_v11_b256_ncu_full.txt): 14.1 MB read + 190.7 MB written (f64 partials)
        "traffic": 204.8e6 if (B == 256 and mode == ek.MODE_DIRECT) else None,
        "hbm": {"achieved_gbs": alg_bytes / (k_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback",
                "note": "algorithmic bytes: 16 B/voxel per individual + f64 partials; the kernel is MUFU-bound, not HBM-bound"},
    }

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
        "gpu_launches": int(model.last_launch_count) * args.steps,
        "launches_per_step": int(model.last_launch_count),
        "sims_per_s": value / (N_VOX * T_FULL), "e2e_sims_per_s": e2e_value / (N_VOX * T_FULL),
        "roofline": roofline, "clocks": clocks,
        "automaton": {"ms": automaton_ms, "kernel": "automaton_brick_kernel (4^3-brick frontier, work ring)", "brick_visits": sweeps,
                      "bit_exact_vs_reference": bool(act_ok), "edges_per_s": 26 * N_VOX / (automaton_ms * 1e-3),
                      "reference_cpu_s": 2.0},
        "parity_max_err_of_peak": parity,
        "fast_path": fast, "separable_path": separable, "pipeline": pipeline, "single_sim": single,
    }

    if not args.no_cpu_baseline and world == 1:
        try:
            vec_lines = open(os.path.join(ROOT, "tests", "golden", "vectors256.txt")).read().strip().split("\n")
            cores = host_cores()
            v, secs, wall = reference_sample(args.ref_length, cores, vec_lines)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "reference",
                                    "sample": "%d concurrent processes of the unmodified reference CLI, one vector each, "
                                              "length %d of 400 samples (%.1f s wall)" % (cores, args.ref_length, wall)}
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
