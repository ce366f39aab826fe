#!/usr/bin/env python
"""tools/prof_case.py MODE [B] -- one model_24 batch through ekg_simulate in the given mode (1 direct, 2 hoisted,
3 separable), twice; meant to be run under ncu with a kernel-name filter (see profiles/)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ekgio  # noqa: E402
import ekgsim_b200  # noqa: E402

mode = int(sys.argv[1])
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
m24 = ekgio.load_model24()
g = np.load(os.path.join(ROOT, "tests", "golden", "golden_glue256.npz"))
m = ekgsim_b200.Model(m24["layers"], m24["transfer"])
m.activation()
k = np.tile(g["layer_k"], ((B + 255) // 256, 1, 1))[:B]
leads = np.tile(g["leads_zyx"], ((B + 255) // 256, 1, 1))[:B]
for _ in range(2):
    m.simulate(k, leads, "3D4", 100.0, 1.0, 400.0, mode=mode)
print(m.last_kernel_name, m.last_launch_count)
