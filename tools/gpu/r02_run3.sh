set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/r2c_pytest.log 2>&1
tail -25 gpurun_out/r2c_pytest.log
python - <<'PY' > gpurun_out/r2c_create.json 2> gpurun_out/r2c_create.err
import json, sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
import ekgio, ekgsim_b200 as ek
out = {}
for f in (1, 4):
    layers, transfer, leads = ekgio.scaled_heart(f) if f > 1 else (ekgio.load_model24()["layers"], ekgio.load_model24()["transfer"], None)
    ek.Model(layers, transfer).close()
    t = []
    for _ in range(3):
        t0 = time.perf_counter(); m = ek.Model(layers, transfer); t.append(time.perf_counter() - t0)
        m.close()
    out["model_create_s_%dx" % f] = min(t)
print(json.dumps(out))
PY
cat gpurun_out/r2c_create.json; tail -3 gpurun_out/r2c_create.err
python - <<'PY'
import sys, json
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, ekgio, ekgsim_b200 as ek
m24 = ekgio.load_model24(); model = ek.Model(m24["layers"], m24["transfer"]); model.activation(download=False)
gf = np.load("tests/golden/golden_eval_full.npz")
e = model.simulate(gf["layer_k"][0], gf["leads_zyx"][0], "3D4", 100.0, 1.0, 400.0, mode=1)[0]
print("debug single:", e.shape, e[:, :3], gf["ecg"][0][:, :3], float(np.abs(e - gf["ecg"][0]).max()))
PY
