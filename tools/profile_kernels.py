#!/usr/bin/env python
"""tools/profile_kernels.py -- one pass of a named workload between cudaProfilerStart/Stop, for ncu:

    ncu --set full --clock-control none --import-source on --profile-from-start off \
        --metrics smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_alu.sum \
        -o gpurun_out/r02_<name> python tools/profile_kernels.py <name>

Workloads (model_24, T = 400, 2 leads, the 256 seeded vectors of the config-3 batch):
    direct256 hoisted256 separable256 cornersum256 direct1 hoisted1 separable1 fit256 fit1 evaluate256 automaton
    automatonx4 (the same on ekgio.scaled_heart(4): the time-bucket queue; EKGSIM_B200_AUTOMATON_QUEUE=fifo for the FIFO ring)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ekgio  # noqa: E402
import ekgsim_b200 as ek  # noqa: E402

name = sys.argv[1]
m24 = ekgio.load_model24()
if name == "automatonx4":
    big, transfer4, _ = ekgio.scaled_heart(4)
    model = ek.Model(big, transfer4, device=0)
    name = "automaton"
else:
    model = ek.Model(m24["layers"], m24["transfer"], device=0)
model.activation(download=False)
g = np.load(os.path.join(ROOT, "tests", "golden", "golden_glue256.npz"))
dev = torch.device("cuda", 0)
d_k = torch.from_numpy(np.ascontiguousarray(g["layer_k"])).to(dev)
d_l = torch.from_numpy(np.ascontiguousarray(g["leads_zyx"])).to(dev)
d_e = torch.empty((256, 2, 400), dtype=torch.float64, device=dev)
nl = g["layer_k"].shape[1]
d_b = torch.from_numpy(np.ascontiguousarray(g["layer_k"][:, [0, 14, nl - 1]])).to(dev)
d_lk = torch.empty((256, nl, 9), dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream().cuda_stream
B = 1 if name.endswith("1") else 256
modes = {"direct": ek.MODE_DIRECT, "hoisted": ek.MODE_HOISTED, "separable": ek.MODE_SEPARABLE, "cornersum": ek.MODE_SEPARABLE | ek.FLAG_CORNER_SUM}
kind = name.rstrip("0123456789")


def run():
    if kind in modes:
        model.simulate_device(d_k.data_ptr(), d_l.data_ptr(), B, 2, d_e.data_ptr(), "3D4", 100.0, 1.0, 400.0, mode=modes[kind], stream=stream)
    elif kind == "fit":
        model.fit_layers_device(d_b.data_ptr(), B, 3, d_lk.data_ptr(), mid=14, stream=stream)
    elif kind == "evaluate":
        tg = np.stack([m24["target_v5"][100:500, 1], m24["target_v5"][100:500, 2]])
        model.evaluate(g["layer_k"][:B, [0, 14, nl - 1]], g["leads_zyx"][:B], tg, mid=14)
    elif kind == "automaton":
        model.activation(download=False)
    else:
        raise SystemExit("unknown workload " + name)


for _ in range(2):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
