"""ekgsim_b200/csrc/fit_math.cuh restates glibc's exp / log / pow (the reference's libm, FMA code path) for the device
fit.  Its HOST build must agree with the machine's libm bit for bit (tests/native/libm_port_check.cpp): that is what
makes the device fit take the reference glue's decisions.  Needs an x86-64 CPU with FMA3 and glibc >= 2.28 (both true for
the build container and the B200 boxes)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_fma():
    try:
        return " fma " in open("/proc/cpuinfo").read().replace("\n", " ")
    except OSError:
        return False


@pytest.mark.skipif(not _has_fma(), reason="host CPU without FMA3: glibc selects its non-FMA code path there")
def test_fit_math_matches_host_libm_bit_for_bit(tmp_path):
    exe = str(tmp_path / "libm_port_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-o", exe,
                           os.path.join(ROOT, "tests", "native", "libm_port_check.cpp"), "-lm"])
    r = subprocess.run([exe, "400000"], capture_output=True, text=True)
    print(r.stdout)
    assert r.returncode == 0, r.stdout
    lines = [ln.split() for ln in r.stdout.strip().split("\n")]
    assert len(lines) == 13 and all(ln[-1] == "0" for ln in lines)


def test_libm_tables_are_what_the_generator_extracts():
    """the committed table header is exactly what tools/gen_libm_tables.py reads out of this machine's libm"""
    path = os.path.join(ROOT, "ekgsim_b200", "csrc", "libm_tables.h")
    before = open(path).read()
    libm = "/lib/x86_64-linux-gnu/libm.so.6"
    if not os.path.exists(libm):
        pytest.skip("no glibc libm at the usual place")
    try:
        subprocess.check_call(["python", os.path.join(ROOT, "tools", "gen_libm_tables.py"), libm], stdout=subprocess.DEVNULL)
        assert open(path).read() == before
    finally:
        open(path, "w").write(before)
