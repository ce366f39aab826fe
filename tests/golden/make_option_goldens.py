#!/usr/bin/env python
"""tests/golden/make_option_goldens.py -- goldens for the evaluation glue's NON-DEFAULT branches.

Build container only (needs oracle/_ref/ref_dump = the unmodified reference's sim.cpp + simlib compiled from
/root/reference by `make -C oracle ref`).  For every variant below a testRun-style directory is materialised
(tests/ekgio.py, the same files the GPU tests give the product), its simulator.ini edited, and `ref_dump eval` run on
a handful of parameter vectors.  Covered (reference sim.cpp:600-702 calculateFitness, :712-747 runApproxAndSim,
:751-821 simUsingBorderAps, SimSettings.h:158-356):

    comparison mode 1 / 3 / 4, mode = 9 (leads_sum: the reference throws), peak position is criterion,
    fast approximation is criterion, fast approximation limit < 2 (the gate that SKIPS Simulation::run),
    endo-epi minimization criterion epi delay, optimize measuring points = 0, interpolation = endo-epi,
    and one variant with every optional criterion switched on at once (ordering of the criteria vector).

The simulation length is cut to LENGTH samples in the ini (the reference needs 0.5 s per sample and vector); both
sides read the same ini, so the comparison is like for like.  Output: tests/golden/golden_options.npz +
golden_options.json (the ini edits, so the tests rebuild the same directories).
"""
import json
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import ekgio  # noqa: E402
from oracle import oracle  # noqa: E402

REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
SCRATCH = os.path.join(ROOT, ".scratch", "golden_opts")
LENGTH = 24

# name -> (ini edits {key: value}, which genes of the 16-gene vectors are used)
ALL16 = list(range(16))
VARIANTS = {
    "cmp_rms": ({"comparison mode": "1"}, ALL16),
    "cmp_devlin": ({"comparison mode": "3"}, ALL16),
    "cmp_vcorr": ({"comparison mode": "4"}, ALL16),
    "leads_sum": ({"mode": "9"}, ALL16),
    "peak_pos": ({"peak position is criterion": "1"}, ALL16),
    "approx_crit": ({"fast approximation is criterion": "1"}, ALL16),
    "approx_gate": ({"fast approximation limit": None}, ALL16),      # limit chosen below so that some vectors skip run()
    "approx_gate_rms": ({"fast approximation limit": None, "comparison mode": "1", "fast approximation is criterion": "1"}, ALL16),
    "endo_epi_delay": ({"endo-epi minimization criterion epi delay": "5"}, ALL16),
    "fixed_leads": ({"optimize measuring points": "0"}, list(range(12))),
    "interp_endo_epi": ({"interpolation": "endo-epi"}, [0, 1, 2, 3, 8, 9, 10, 11, 12, 13, 14, 15]),
    "everything": ({"peak position is criterion": "1", "fast approximation is criterion": "1",
                    "endo-epi minimization criterion epi delay": "2.5", "comparison mode": "4"}, ALL16),
}
N_VEC = 5


def edit_ini(ini, edits):
    ini = re.sub(r"(?m)^length = \d+$", "length = %d" % LENGTH, ini)
    for key, val in edits.items():
        pat = r"(?m)^" + re.escape(key) + r" = .*$"
        assert re.search(pat, ini), key
        ini = re.sub(pat, "%s = %s" % (key, val), ini)
    return ini


def vectors():
    lines = open(os.path.join(HERE, "vectors256.txt")).read().strip().split("\n")
    return np.array([[float(x) for x in ln.replace(",", " ").split()] for ln in lines])


def run_variant(name, edits, genes, vecs):
    d = os.path.join(SCRATCH, name)
    os.makedirs(d, exist_ok=True)
    ekgio.materialise_testrun(d, ini_edit=lambda s: edit_ini(s, edits))
    with open(os.path.join(d, "vec.txt"), "w") as f:
        for v in vecs:
            f.write(",".join("%.17g" % v[g] for g in genes) + "\n")
    r = subprocess.run([REF_DUMP, "eval", "vec.txt", "eval.bin"], cwd=d, capture_output=True, text=True)
    return name, r.returncode, r.stderr[-400:]


def main():
    vec = vectors()
    os.makedirs(SCRATCH, exist_ok=True)
    # pre-screen for the gate: approximate-ECG criterion of every vector (glue only: limit -1 skips every run())
    screen = {}
    for nm in ("approx_gate", "approx_gate_rms"):
        ed = dict(VARIANTS[nm][0])
        ed["fast approximation limit"] = "-1"
        ed["fast approximation is criterion"] = "1"
        n, rc, err = run_variant("screen_" + nm, ed, ALL16, vec[:48])
        assert rc == 0, err
        recs = oracle.read_eval_dump(os.path.join(SCRATCH, "screen_" + nm, "eval.bin"))
        ac = np.array([r["criteria"][2] for r in recs])       # ECG1, ECG2, [fast approx]
        order = np.argsort(ac)
        pick = sorted([int(order[i]) for i in (0, len(order) // 4, len(order) // 2, 3 * len(order) // 4, len(order) - 1)])
        limit = float(np.round(0.5 * (ac[order[len(order) // 2]] + ac[order[len(order) // 2 + 1]]), 6))
        screen[nm] = dict(limit=limit, pick=pick, approx=[float(ac[i]) for i in pick])
        print(nm, "approx criteria of the picked vectors", screen[nm], flush=True)

    jobs, meta = [], {}
    for name, (edits, genes) in VARIANTS.items():
        edits = dict(edits)
        idx = list(range(N_VEC))
        if name in screen:
            edits["fast approximation limit"] = repr(screen[name]["limit"])
            idx = screen[name]["pick"]
        meta[name] = dict(ini_edits=edits, genes=genes, vector_index=idx, length=LENGTH)
        jobs.append((name, edits, genes, vec[idx]))
    with ThreadPoolExecutor(max_workers=7) as ex:
        res = list(ex.map(lambda j: run_variant(*j), jobs))
    out = {}
    for name, rc, err in res:
        if name == "leads_sum":
            assert rc != 0, "mode = 9 is expected to throw in the reference"
            m = re.search(r"ref_dump: (.*)", err)
            meta[name]["error"] = m.group(1).strip() if m else err.strip()
            print(name, "->", meta[name]["error"])
            continue
        assert rc == 0, (name, err)
        recs = oracle.read_eval_dump(os.path.join(SCRATCH, name, "eval.bin"))
        assert len(recs) == len(meta[name]["vector_index"])
        out[name + "/params"] = np.array([r["params"] for r in recs])
        out[name + "/layer_k"] = np.array([r["layer_k"] for r in recs])
        out[name + "/leads_zyx"] = np.array([r["leads_zyx"] for r in recs])
        out[name + "/ecg"] = np.array([r["ecg"] for r in recs])
        out[name + "/criteria"] = np.array([r["criteria"] for r in recs])
        out[name + "/violation"] = np.array([r["violation"] for r in recs])
        out[name + "/simulation_done"] = np.array([r["simulation_done"] for r in recs])
        print(name, "criteria", out[name + "/criteria"].tolist(), "done", out[name + "/simulation_done"].tolist(), flush=True)
    np.savez_compressed(os.path.join(HERE, "golden_options.npz"), **out)
    json.dump(meta, open(os.path.join(HERE, "golden_options.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
