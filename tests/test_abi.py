"""CPU tests of the drop-in boundary: the shared library builds (nvcc cross-compiles without a
GPU), loads, and exports exactly the symbols include/ekgsim_b200.h declares.  No compute calls."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ekgsim_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ekg_[a-z0-9_]+)\s*\(", src)))


def test_header_compiles_as_c():
    subprocess.check_call(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "ekgsim_b200.h")])


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert sorted(built.SYMBOLS) == names, "ctypes table and header disagree"
    assert built.lib().ekg_abi_version() == 3


def test_sass_is_sm100a_and_uses_mufu(built):
    out = subprocess.run(["cuobjdump", "-lelf", built.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", built.LIB_PATH], capture_output=True, text=True).stdout
    assert "MUFU.EX2" in sass and "MUFU.LG2" in sass and "MUFU.RCP" in sass


def test_no_cpu_fallback(built):
    """Without a CUDA device the compute entry points must fail loudly (EKG_E_CUDA)."""
    import numpy as np
    if built.lib().ekg_device_count() > 0:
        pytest.skip("a GPU is present")
    import synth
    layers, transfer, _ = synth.small_heart()
    with pytest.raises(built.EkgError) as e:
        built.Model(layers, transfer)
    assert e.value.code == -5 and "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under ekgsim_b200/ may reference it."""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "ekgsim_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"oracle[/.]|ekg_oracle|import oracle|from oracle", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
