#!/usr/bin/env python
"""tools/bench_generation.py -- BASELINE config 5: one AMS-DEMO generation (population 100) on the B200 path.

 (a) the reference's own optimizer (unmodified main.cpp / AMS-DEMO, sequential -DNO_MPI build) with the B200
     simulator library underneath (oracle/_ref/ekgSim_refglue_b200, built by `make -C oracle ref`): population
     100, 1 generation -> 200 sequential evaluations (initial population + one generation of offspring);
 (b) the same 100 + 100 parameter vectors through the product's batched evaluator (`ekgSim -batch`): glue on all
     host threads, one GPU launch per 100 vectors.
The reference needs 205 s per evaluation on one core, i.e. 100 x 205 s / P per generation on P cores."""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ekgio  # noqa: E402

out = {"population": 100, "reference_cpu_core_seconds_per_generation": 100 * 205.28}
d = tempfile.mkdtemp(prefix="ekg_gen_")
ekgio.materialise_testrun(d, targets="target_ecg_v2_v6.column",
                          ini_edit=lambda s: s.replace("number of generations = 100", "number of generations = 1"))
exe = os.path.join(ROOT, "oracle", "_ref", "ekgSim_refglue_b200")
if os.path.exists(exe):
    t0 = time.time()
    r = subprocess.run([exe], cwd=d, capture_output=True, text=True)
    dt = time.time() - t0
    n_eval = len([ln for ln in open(os.path.join(d, "evaluations.txt")) if ln.strip() and not ln.startswith("#")])
    out["reference_optimizer_on_b200_simlib"] = {"evaluations": n_eval, "wall_s": dt, "s_per_generation_of_100": dt * 100 / max(n_eval, 1),
                                                 "ok": r.stdout.rstrip().endswith("All done")}
# (c) the reference's stand-alone optimizer (AMS-DEMO/main.cpp, sequential build) evaluating through its
#     ExternalEvaluation protocol: one `ekgSim -extern` process per individual, answered by a resident `ekgSim -serve`
demo = os.path.join(ROOT, "oracle", "_ref", "DEMO_ref")
cli = os.path.join(ROOT, "ekgsim_b200", "bin", "ekgSim")
if os.path.exists(demo):
    d2 = tempfile.mkdtemp(prefix="ekg_gen_ext_")
    ekgio.materialise_testrun(d2, targets="target_ecg_v2_v6.column")
    open(os.path.join(d2, "settings.ini"), "w").write("""[evaluation]
command line = %s -extern
input file name = input.txt
output file name = output.txt
chromosome vector length = 16
criteria vector length = 2
properties vector length = 0

[optimization]
random seed = 11
population size = 100
max number of generations = 2
DE schema = rand/1/bin
p crossover = 0.3
scaling factors = 0.5
queue length = 1

[initial population]
gene min = 0.0003, 0.01, 0.01, 200, 0.0003, 0.01, 0.01, 200, 0.0003, 0.01, 0.01, 200, -50, -50, -50, -50
gene max = 0.001, 0.1, 0.1, 400, 0.001, 0.1, 0.1, 400, 0.001, 0.1, 0.1, 400, 50, 50, 50, 50
""" % cli)
    sock = os.path.join(d2, "ekg.sock")
    srv = subprocess.Popen([cli, "-serve", sock], cwd=d2, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    for _ in range(600):
        if os.path.exists(sock) or srv.poll() is not None:
            break
        time.sleep(0.1)
    t0 = time.time()
    r = subprocess.run([demo], cwd=d2, capture_output=True, text=True, env=dict(os.environ, EKGSIM_B200_SERVER=sock))
    dt = time.time() - t0
    subprocess.run([cli, "-shutdown", sock], cwd=d2, capture_output=True)
    err = srv.communicate()[1]
    n_eval = len([ln for ln in open(os.path.join(d2, "evaluations.txt")) if ln.strip() and not ln.startswith("#")])
    out["reference_ams_demo_via_extern_server"] = {"evaluations": n_eval, "wall_s": dt, "s_per_generation_of_100": dt * 100 / max(n_eval, 1),
                                                   "ms_per_evaluation": 1e3 * dt / max(n_eval, 1), "ok": "front.txt" in r.stdout,
                                                   "server": err.strip().split("\n")[-1]}
vec = open(os.path.join(ROOT, "tests", "golden", "vectors256.txt")).read().strip().split("\n")[:100]
open(os.path.join(d, "pop.txt"), "w").write("\n".join(vec) + "\n")
t0 = time.time()
r = subprocess.run([cli, "-batch", "pop.txt", "-batchout", "crit.txt"], cwd=d, capture_output=True, text=True)
dt = time.time() - t0
m = re.search(r"batch of (\d+) simulations done in ([0-9.e+-]+) seconds \(GPU part ([0-9.e+-]+) s\)", r.stdout)
out["batched_cli"] = {"process_wall_s": dt, "evaluation_s": float(m.group(2)) if m else None, "gpu_s": float(m.group(3)) if m else None,
                      "host_threads": os.cpu_count()}
print(json.dumps(out))
