#!/usr/bin/env python
"""tools/bench_heart4x_cli.py -- BASELINE config 4 through the product's C++ host alone (no Python on the path):
`ekgSim test -sim <README vector> -out result -slabs N` in a directory whose simulator.ini names the 4x finer heart as a
208 MB text .matrix.  Prints the JSON block bench.py carries as heart4x.cxx_host.

    python tools/bench_heart4x_cli.py --gpus 8 [--factor 4] [--mode direct]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--factor", type=int, default=4)
    ap.add_argument("--mode", default="", help="EKGSIM_B200_MODE for the runs: direct | hoisted | separable (default: the library's)")
    ap.add_argument("--automaton", default="", help="EKGSIM_B200_SLAB_AUTOMATON: replicated = every device computes the whole map")
    a = ap.parse_args()
    if a.mode:
        os.environ["EKGSIM_B200_MODE"] = a.mode
    if a.automaton:
        os.environ["EKGSIM_B200_SLAB_AUTOMATON"] = a.automaton
    print(json.dumps(dict(bench.heart_cli_block(a.factor, a.gpus), mode=a.mode or "default", automaton_env=a.automaton or "linked")))
