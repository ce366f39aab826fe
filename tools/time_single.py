#!/usr/bin/env python
"""tools/time_single.py -- device time of ONE simulation (BASELINE config 1: model_24, T = 400, 2 leads) and of a 256-vector
batch in the three ECG modes, for kernel tuning (the segment / slice knobs are read from the environment once per process:
EKGSIM_B200_ECG_SUB, EKGSIM_B200_ECG_SLICES, EKGSIM_B200_MOMENT_FUSED, EKGSIM_B200_MOMENT_CTAS).  One JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ekgio  # noqa: E402
import ekgsim_b200 as ek  # noqa: E402

m24 = ekgio.load_model24()
model = ek.Model(m24["layers"], m24["transfer"], device=0)
model.activation(download=False)
g = np.load(os.path.join(ROOT, "tests", "golden", "golden_glue256.npz"))
gf = np.load(os.path.join(ROOT, "tests", "golden", "golden_eval_full.npz"))
dev = torch.device("cuda", 0)
d_k = torch.from_numpy(np.ascontiguousarray(g["layer_k"])).to(dev)
d_l = torch.from_numpy(np.ascontiguousarray(g["leads_zyx"])).to(dev)
d_e = torch.empty((256, 2, 400), dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream().cuda_stream
out = {"env": {k: v for k, v in os.environ.items() if k.startswith("EKGSIM_B200_")}}
batches = [int(b) for b in sys.argv[1:]] or [1, 256]
for B in batches:
    for nm, md in (("direct", ek.MODE_DIRECT), ("hoisted", ek.MODE_HOISTED), ("separable", ek.MODE_SEPARABLE)):
        run = lambda: model.simulate_device(d_k.data_ptr(), d_l.data_ptr(), B, 2, d_e.data_ptr(), "3D4", 100.0, 1.0, 400.0, mode=md, stream=stream)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        n = 20 if B == 1 else 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            run()
        e1.record()
        torch.cuda.synchronize()
        out["B%d_%s_ms" % (B, nm)] = e0.elapsed_time(e1) / n
        out["B%d_%s_launches" % (B, nm)] = model.last_launch_count
# parity of the single-simulation path against the reference goldens (4 full-length vectors, one at a time)
for nm, md in (("direct", ek.MODE_DIRECT), ("hoisted", ek.MODE_HOISTED), ("separable", ek.MODE_SEPARABLE)):
    worst = 0.0
    for i in range(gf["ecg"].shape[0]):
        e = model.simulate(gf["layer_k"][i], gf["leads_zyx"][i], "3D4", 100.0, 1.0, 400.0, mode=md)[0]
        err = float((np.abs(e - gf["ecg"][i]) / np.abs(gf["ecg"][i]).max(axis=1, keepdims=True)).max())
        worst = err if not (err <= worst) else worst   # a NaN must show
    out["B1_%s_err_of_peak" % nm] = worst
print(json.dumps(out))
