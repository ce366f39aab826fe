"""The lock-free work queue of the time-bucket automaton (csrc/automaton.cu), restated with std::atomic and host threads
(tests/cpp/queue_emulation.cpp): every item pushed is processed exactly once, whatever the interleaving."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emulation(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("queue") / "queue_emulation")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", os.path.join(HERE, "cpp", "queue_emulation.cpp"), "-o", exe])
    return exe


@pytest.mark.parametrize("threads", [3, 16, 64])
def test_every_brick_is_taken_exactly_once(emulation, threads):
    r = subprocess.run([emulation, str(threads), "200000", str(1 << 18)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "duplicates 0" in r.stdout and "pending 0" in r.stdout, r.stdout
