#!/usr/bin/env python
"""tools/ab_automaton.py -- A/B of the frontier automaton's work queue on one GPU: the FIFO ring against the time-bucket
queue (EKGSIM_B200_AUTOMATON_QUEUE=fifo|timed) and bucket widths (EKGSIM_B200_AUTOMATON_DELTA, ms), on model_24 and the
finer hearts.  Every variant must produce the same bits (checked against the committed sha256 goldens)."""
import argparse
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ekgio  # noqa: E402
import ekgsim_b200 as ek  # noqa: E402

GOLD1 = "4e51f1bbc04590c3a9a71b61f8575955ff4e778aa380c93292e5e4c406c64bda"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--factors", default="1,2,4")
    ap.add_argument("--deltas", default="")
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--ctas", default="", help="comma list of EKGSIM_B200_AUTOMATON_CTAS_PER_SM values to try for every variant")
    a = ap.parse_args()
    out = []
    for f in [int(x) for x in a.factors.split(",")]:
        layers, transfer, _ = ekgio.load_model24()["layers"], None, None
        if f == 1:
            m24 = ekgio.load_model24()
            layers, transfer = m24["layers"], m24["transfer"]
            gold = GOLD1
        else:
            layers, transfer, _ = ekgio.scaled_heart(f)
            g = os.path.join(ROOT, "tests", "golden", "golden_heart%dx.json" % f)
            gold = json.load(open(g))["sha256_f64_raster"] if os.path.exists(g) else None
        variants = [("fifo", None, c) for c in (a.ctas.split(",") if a.ctas else [None])] + [("timed", None, c) for c in (a.ctas.split(",") if a.ctas else [None])] \
            + [("timed", d, None) for d in a.deltas.split(",") if d]
        for queue, delta, ctas in variants:
            os.environ["EKGSIM_B200_AUTOMATON_QUEUE"] = queue
            if ctas is None:
                os.environ.pop("EKGSIM_B200_AUTOMATON_CTAS_PER_SM", None)
            else:
                os.environ["EKGSIM_B200_AUTOMATON_CTAS_PER_SM"] = ctas
            if delta is None:
                os.environ.pop("EKGSIM_B200_AUTOMATON_DELTA", None)
            else:
                os.environ["EKGSIM_B200_AUTOMATON_DELTA"] = delta
            m = ek.Model(layers, transfer, device=0)
            ms, visits = [], []
            for _ in range(a.reps):
                _, v = m.activation(download=False)
                ms.append(m.activation_ms)
                visits.append(v)
            sha = hashlib.sha256(m.get_activation().tobytes()).hexdigest()
            m.close()
            row = {"factor": f, "queue": queue, "delta_ms": delta or "default", "ctas_per_sm": ctas or "default", "ms_best": min(ms), "ms_all": [round(x, 3) for x in ms],
                   "brick_visits": visits[-1], "bit_exact": (sha == gold) if gold else None}
            print(json.dumps(row), flush=True)
            out.append(row)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ab_automaton.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
