#!/usr/bin/env python
"""tools/bench_fit.py -- timing of the device-side layer fit and of the whole border-APs -> criteria pass
(ekg_fit_layers / ekg_evaluate) on model_24 with the 256 seeded vectors; one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ekgio  # noqa: E402
import ekgsim_b200  # noqa: E402


def main():
    m24 = ekgio.load_model24()
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_glue256.npz"))
    tv = m24["target_v5"]
    targets = np.stack([tv[:, c] / (tv[:, c].max() - tv[:, c].min()) for c in (1, 2)])
    m = ekgsim_b200.Model(m24["layers"], m24["transfer"])
    m.activation()
    border = g["layer_k"][:, [0, 14, 23]].copy()
    out = {}
    for B in (1, 16, 256, 1024, 4096):
        b = np.tile(border, (max(1, B // 256), 1, 1))[:B]
        leads = np.tile(g["leads_zyx"], (max(1, B // 256), 1, 1))[:B]
        m.fit_layers(b, mid=14)
        t = []
        for _ in range(5):
            t0 = time.perf_counter(); m.fit_layers(b, mid=14); t.append(time.perf_counter() - t0)
        out["fit_ms_B%d" % B] = round(1e3 * min(t), 3)
        for mode, name in ((3, "separable"), (2, "hoisted"), (1, "direct")):
            if B > 256 and mode != 3:
                continue
            m.evaluate(b, leads, targets, mid=14, mode=mode)
            t = []
            for _ in range(3):
                t0 = time.perf_counter(); m.evaluate(b, leads, targets, mid=14, mode=mode); t.append(time.perf_counter() - t0)
            out["evaluate_%s_ms_B%d" % (name, B)] = round(1e3 * min(t), 3)
            out["evaluate_%s_sims_per_s_B%d" % (name, B)] = round(B / min(t), 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
