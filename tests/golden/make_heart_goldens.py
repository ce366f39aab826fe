#!/usr/bin/env python
"""tests/golden/make_heart_goldens.py -- golden values for the synthetic finer-resolution hearts of BASELINE config 4
(SURVEY 8(d)): ekgio.scaled_heart(f), f = 2 and 4.  Runs the oracle (oracle/ekg_oracle.c; its automaton and ECG loops
are pinned to the compiled reference on model_24, tests/test_oracle_golden.py) once here in the build container
(f = 4: ~2 min automaton + the ECG samples) and leaves a small JSON per factor in tests/golden/:

  golden_heart{f}x.json   sha256 of the raw f64 raster activation map, its sum / max / occupied count, and the oracle's
                          ECG (class-factored loop, f64) for the first 16 samples of the testRun time axis
                          (start 100 ms, step 1 ms) with the layer coefficients and displaced leads of vector 0 of the
                          config-3 batch (golden_glue256.npz; leads x f), plus the peak |ECG| per lead over all 400
                          samples (the denominator of the 1e-5 tolerance).

bench.py (the heart4x block) and tests/test_gpu_heart.py compare the CUDA path with these.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import ekgio  # noqa: E402
from oracle import oracle  # noqa: E402


def make(f, full_peak=True):
    layers, transfer, leads = ekgio.scaled_heart(f)
    g = np.load(os.path.join(HERE, "golden_glue256.npz"))
    k = g["layer_k"][0]
    lead_v = g["leads_zyx"][0] * f          # displaced leads of vector 0, scaled like the model
    t0 = time.time()
    delay = oracle.activation(layers, transfer)
    t_auto = time.time() - t0
    occ = (layers & 0x0FFF) > 0
    t0 = time.time()
    ecg16 = oracle.run_factored(layers, delay, k, lead_v, "3D4", 100.0, 1.0, 16.0)
    t_ecg16 = time.time() - t0
    out = dict(
        source="oracle/ekg_oracle.c (ekg_oracle_activation, ekg_oracle_run_factored) on ekgio.scaled_heart(%d)" % f,
        factor=f, shape_zyx=list(layers.shape), occupied=int(occ.sum()),
        sha256_f64_raster=hashlib.sha256(delay.tobytes()).hexdigest(),
        sum=float(delay.sum()), max=float(delay.max()), min_occupied=float(delay[occ].min()),
        vector=0, t_start=100.0, t_step=1.0, leads_zyx=lead_v.tolist(),
        ecg16=ecg16.tolist(), oracle_automaton_s=t_auto, oracle_ecg16_s=t_ecg16,
    )
    if full_peak:
        t0 = time.time()
        full = oracle.run_factored(layers, delay, k, lead_v, "3D4", 100.0, 1.0, 400.0)
        out["peak_full"] = np.abs(full).max(axis=1).tolist()
        out["ecg_full_sha256"] = hashlib.sha256(full.tobytes()).hexdigest()
        out["ecg_at_t"] = {str(t): full[:, t].tolist() for t in (0, 100, 200, 399)}
        out["oracle_ecg400_s"] = time.time() - t0
    json.dump(out, open(os.path.join(HERE, "golden_heart%dx.json" % f), "w"), indent=1)
    print("heart%dx:" % f, {k_: v for k_, v in out.items() if k_ not in ("ecg16", "leads_zyx", "ecg_at_t")}, flush=True)


if __name__ == "__main__":
    for f in [int(a) for a in sys.argv[1:]] or [2, 4]:
        make(f)
