#!/usr/bin/env python
"""tests/golden/make_fixtures.py -- regenerates the committed golden fixtures.

Runs ONLY in the build container (needs /root/reference and oracle/_ref built by
`make -C oracle ref`).  Nothing in tests/, smoke() or bench.py reads /root/reference at run time;
they read the small files this script leaves in tests/golden/:

  testrun_model24.npz   compact copy of the reference's testRun inputs (layer map u8, start voxel,
                        conduction matrix, lead positions, the two target ECG columns, and the
                        key=value lines of simulator.ini without comments)
  golden_activation.json  fingerprint of the reference automaton output (sha256 of the raw f64
                        raster array, sum/min/max, class count)  -- `ref_dump activation`
  golden_eval_full.npz  full-length (T=400) reference evaluations: params, 24x9 layer coefficients,
                        displaced leads, ECG f64 [2][400], criteria, violation -- `ref_dump eval`
  golden_len16.npz      24 vectors of the config-3 batch, reference run with length = 16
  golden_glue256.npz    all 256 vectors of the batch, glue outputs only (layer coefficients,
                        leads, violation) -- `ref_dump eval --glue-only`

The reference runs themselves (about 4 core-minutes per full-length vector) are started by
tests/golden/run_reference.sh into .scratch/golden/.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import ekgio  # noqa: E402
from oracle import oracle  # noqa: E402

REF = "/root/reference/testRun"
SCRATCH = os.path.join(ROOT, ".scratch", "golden")


def inputs():
    layers, _ = ekgio.read_matrix(os.path.join(REF, "model_24.matrix"))
    transfer, _ = ekgio.read_matrix(os.path.join(REF, "conduction_24.matrix"), shape_load=False)
    leads = ekgio.read_points(os.path.join(REF, "model_24_measuring_pos.txt"))
    start = np.flatnonzero(layers.reshape(-1) & ekgio.START_FLAG)
    l8 = (layers & ~np.uint16(ekgio.START_FLAG)).astype(np.uint8)
    ini_lines = []
    for line in open(os.path.join(REF, "simulator.ini")):
        s = line.strip()
        if s and not s.startswith(";"):
            ini_lines.append(s)
    np.savez_compressed(
        os.path.join(HERE, "testrun_model24.npz"),
        layers_u8=l8, start_index=start.astype(np.int64), transfer=transfer, leads_zyx=leads,
        target_v5=ekgio.read_column(os.path.join(REF, "target_ecg_v2_v5.column")),
        target_v6=ekgio.read_column(os.path.join(REF, "target_ecg_v2_v6.column")),
        simulator_ini=np.array("\n".join(ini_lines) + "\n"),
    )
    print("inputs: layers", layers.shape, "occupied", int((l8 > 0).sum()), "start", start,
          "sha256(u8)", hashlib.sha256(l8.tobytes()).hexdigest())


def activation():
    layers, delay = oracle.read_activation_dump(os.path.join(SCRATCH, "act", "act.bin"))
    occ = layers > 0
    K, _ = oracle.ap_classes(layers, delay, int(layers.max()))
    fp = dict(
        source="oracle/_ref/ref_dump activation (unmodified reference, simulator.cpp:248-286)",
        shape_zyx=list(layers.shape),
        sha256_f64_raster=hashlib.sha256(delay.tobytes()).hexdigest(),
        sum=float(delay.sum()), min_occupied=float(delay[occ].min()), max=float(delay.max()),
        occupied=int(occ.sum()), classes=int(K),
    )
    json.dump(fp, open(os.path.join(HERE, "golden_activation.json"), "w"), indent=1)
    print("activation:", fp)


def _pack(recs):
    return dict(
        params=np.array([r["params"] for r in recs]),
        layer_k=np.array([r["layer_k"] for r in recs]),
        leads_zyx=np.array([r["leads_zyx"] for r in recs]),
        ecg=np.array([r["ecg"] for r in recs]),
        criteria=np.array([r["criteria"] for r in recs]),
        violation=np.array([r["violation"] for r in recs]),
        seconds=np.array([r["seconds"] for r in recs]),
        simulation_done=np.array([r["simulation_done"] for r in recs]),
    )


def evals():
    full, names = [], []
    for d in ("full1", "v6full", "full2", "full3"):
        p = os.path.join(SCRATCH, d, "eval.bin")
        if os.path.exists(p) and os.path.getsize(p) > 48:
            r = oracle.read_eval_dump(p)
            if r:
                full += r
                names += [d] * len(r)
    if full:
        pk = _pack(full)
        pk["name"] = np.array(names)
        pk["targets"] = np.array(["v6" if n == "v6full" else "v5" for n in names])
        np.savez_compressed(os.path.join(HERE, "golden_eval_full.npz"), **pk)
        for n, r in zip(names, full):
            print("full", n, "criteria", r["criteria"], "violation", r["violation"], "peaks", np.abs(r["ecg"]).max(axis=1), "%.1fs" % r["seconds"])
    p = os.path.join(SCRATCH, "len16", "eval.bin")
    if os.path.exists(p):
        r = oracle.read_eval_dump(p)
        pk = _pack(r)
        # peak amplitude of the full 400-sample trace (the tolerance's denominator); the reference
        # was only run for 16 samples here, so this comes from the oracle (pinned to 2e-13 of peak
        # against the reference on the full-length goldens)
        import ekgio as _io
        m = _io.load_model24()
        d = oracle.activation(m["layers"], m["transfer"])
        pk["peak_full"] = np.array([np.abs(oracle.run_factored(m["layers"], d, k, l, "3D4", 100.0, 1.0, 400.0)).max(axis=1)
                                    for k, l in zip(pk["layer_k"], pk["leads_zyx"])])
        np.savez_compressed(os.path.join(HERE, "golden_len16.npz"), **pk)
        print("len16:", len(r), "vectors")
    p = os.path.join(SCRATCH, "glue", "eval.bin")
    if os.path.exists(p):
        r = oracle.read_eval_dump(p)
        pk = _pack(r)
        del pk["ecg"]
        np.savez_compressed(os.path.join(HERE, "golden_glue256.npz"), **pk)
        print("glue:", len(r), "vectors")


if __name__ == "__main__":
    what = sys.argv[1:] or ["inputs", "activation", "evals"]
    for w in what:
        globals()[w]()
