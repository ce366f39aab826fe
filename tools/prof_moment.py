"""tools/prof_moment.py -- device time of the SEPARABLE path's moment kernel on model_24 (B = 256 and B = 1) under the
tuning knobs the library reads from the environment (one subprocess per setting, the knobs are read once):
EKGSIM_B200_MOMENT_CTAS (CTAs per SM the segment table is sized for).
Usage: python tools/prof_moment.py [--sweep]   -> one JSON line per setting."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def one():
    import numpy as np
    import torch
    import ekgsim_b200 as ek
    import ekgio
    m24 = ekgio.load_model24()
    model = ek.Model(m24["layers"], m24["transfer"], device=0)
    model.activation(download=False)
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_glue256.npz"))
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    out = {"ctas_per_sm": os.environ.get("EKGSIM_B200_MOMENT_CTAS", "default")}
    for B in (256, 1):
        d_k = torch.from_numpy(np.ascontiguousarray(g["layer_k"][:B])).to(dev)
        d_l = torch.from_numpy(np.ascontiguousarray(g["leads_zyx"][:B])).to(dev)
        d_e = torch.empty((B, 2, 400), dtype=torch.float64, device=dev)
        for flag, name in ((0, "series"), (ek.FLAG_CORNER_SUM, "corner_sum")):
            ms = []
            for i in range(13):
                flush.fill_(1)
                model.simulate_device(d_k.data_ptr(), d_l.data_ptr(), B, 2, d_e.data_ptr(), "3D4", 100.0, 1.0, 400.0,
                                      mode=ek.MODE_SEPARABLE | ek.FLAG_TIME_KERNEL | flag, stream=stream)
                if i >= 3:
                    ms.append(model.last_kernel_ms)
            out["B%d_%s_ms" % (B, name)] = float(np.median(ms))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if "--sweep" in sys.argv:
        for ctas in ("24", "48", "96"):
            env = dict(os.environ, EKGSIM_B200_MOMENT_CTAS=ctas)
            subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, check=False)
    else:
        one()
