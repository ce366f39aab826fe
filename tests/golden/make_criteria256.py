#!/usr/bin/env python
"""tests/golden/make_criteria256.py -- criteria of all 256 vectors of the config-3 batch, derived with
the (reference-pinned) oracle: ECG = oracle.run_factored on the reference glue's layer coefficients
(golden_glue256.npz), criteria = 1 - Pearson against the normalised v5 targets exactly like
calculateFitness / statisticalCorrelationCoeff (sim.cpp:600-702, vectorMath.h:287-317).  The same numpy
restatement reproduces the four criteria pairs the compiled reference printed (checked below) before
anything is written.  ~10 CPU-minutes, parallel over processes.  Output: golden_criteria256.npz"""
import os
import sys
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import ekgio  # noqa: E402
from oracle import oracle  # noqa: E402

_m = _d = None


def _init():
    global _m, _d
    _m = ekgio.load_model24()
    _d = oracle.activation(_m["layers"], _m["transfer"])


def targets_v5():
    tv = ekgio.load_model24()["target_v5"]
    return np.stack([tv[:, c] / (tv[:, c].max() - tv[:, c].min()) for c in (1, 2)])   # sim.cpp:1016-1019; resample factor 1 -> copy


def pearson_criteria(ecg, targets):
    out = []
    for l in range(ecg.shape[0]):
        a, t = ecg[l], targets[l][: ecg.shape[1]]
        out.append(1.0 - np.mean((a - a.mean()) * (t - t.mean())) / (a.std() * t.std()))
    return np.array(out)


def _one(args):
    k, leads = args
    ecg = oracle.run_factored(_m["layers"], _d, k, leads, "3D4", 100.0, 1.0, 400.0)
    return pearson_criteria(ecg, targets_v5()), np.abs(ecg).max(axis=1)


if __name__ == "__main__":
    gf = np.load(os.path.join(HERE, "golden_eval_full.npz"))
    tg = targets_v5()
    for i, n in enumerate(gf["name"]):
        if n != "v6full":
            assert np.abs(pearson_criteria(gf["ecg"][i], tg) - gf["criteria"][i]).max() < 1e-12, n   # the restatement is pinned
    g = np.load(os.path.join(HERE, "golden_glue256.npz"))
    with Pool(int(sys.argv[1]) if len(sys.argv) > 1 else 8, initializer=_init) as p:
        res = p.map(_one, list(zip(g["layer_k"], g["leads_zyx"])), chunksize=4)
    np.savez_compressed(os.path.join(HERE, "golden_criteria256.npz"), criteria=np.array([r[0] for r in res]),
                        peak=np.array([r[1] for r in res]), violation=g["violation"])
    print("criteria range", np.array([r[0] for r in res]).min(), np.array([r[0] for r in res]).max())
