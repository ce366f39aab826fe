"""ekgsim_b200 -- B200-native implementation of EkgSim's simulation hot path.

The product is native: `libekgsim_b200.so` (hand-written sm_100a CUDA kernels behind the C ABI of
include/ekgsim_b200.h) and the host-side C++ facade / CLI in ekgsim_b200/host/.  The Python layer
only binds the ABI for tests, bench.py and __graft_entry__.py.
"""
from .capi import (EkgError, FLAG_CORNER_SUM, FLAG_TIME_KERNEL, LIB_PATH, MODE_DEFAULT, MODE_DIRECT, MODE_HOISTED, MODE_SEPARABLE, FIT_D9, NBHD, START_FLAG, SYMBOLS, Model, coefficient_hints, lib,
                   n_steps)

__all__ = ["EkgError", "FLAG_CORNER_SUM", "FLAG_TIME_KERNEL", "LIB_PATH", "MODE_DEFAULT", "MODE_DIRECT", "MODE_HOISTED", "MODE_SEPARABLE", "FIT_D9", "NBHD", "START_FLAG", "SYMBOLS",
           "Model", "coefficient_hints", "lib", "n_steps"]
