"""ekgsim_b200/build.py -- compiles libekgsim_b200.so in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["csrc/capi.cu", "csrc/automaton.cu", "csrc/ecg.cu", "csrc/fit.cu"]
HEADERS = ["csrc/ekg_internal.cuh", "../include/ekgsim_b200.h"]
OUT = os.path.join(HERE, "libekgsim_b200.so")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v",
              # the image's g++ links libstdc++ statically: keep that private copy's symbols local
              "-Xlinker", "--exclude-libs=ALL"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", OUT] + [os.path.join(HERE, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed")
    with open(os.path.join(HERE, "_build_log.txt"), "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
