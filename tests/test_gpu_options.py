"""GPU tests of the evaluation glue's NON-DEFAULT branches against the unmodified reference
(tests/golden/golden_options.{npz,json}, made by tests/golden/make_option_goldens.py: `ref_dump eval` = the
reference's sim.cpp + simlib under the same ini edits).  Reference code covered: sim.cpp:600-702 (calculateFitness:
comparison modes 1/3/4, peak-position criterion, fast-approximation criterion, endo-epi criterion, mode 9),
:712-747 (runApproxAndSim: the gate that skips Simulation::run), :751-821 (simUsingBorderAps, interpolation =
endo-epi), SimSettings.h:158-356.  Tolerances (north_star): comparison criteria within 1e-4; integer criteria (peak
positions) exact; criteria computed from the layer APs alone within 1e-9 relative; violations within 1e-9."""
import json
import os
import re

import numpy as np
import pytest

import ekgio
import hostlib

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
META = json.load(open(os.path.join(GOLDEN, "golden_options.json")))


def _edit(ini, edits, length):
    ini = re.sub(r"(?m)^length = \d+$", "length = %d" % length, ini)
    for key, val in edits.items():
        pat = r"(?m)^" + re.escape(key) + r" = .*$"
        assert re.search(pat, ini), key
        ini = re.sub(pat, "%s = %s" % (key, val), ini)
    return ini


def _workdir(tmp_path, name):
    meta = META[name]
    d = str(tmp_path / name)
    ekgio.materialise_testrun(d, ini_edit=lambda s: _edit(s, meta["ini_edits"], meta["length"]))
    return d, meta


def _check(name, crit, viol, g, simulated_only=None):
    want_c, want_v = g[name + "/criteria"], g[name + "/violation"]
    assert crit.shape == want_c.shape, (crit.shape, want_c.shape)
    assert np.abs(viol - want_v).max() < 1e-9
    big = np.abs(want_c) > 1e3          # the endo-epi criterion: a sum of 699 squared AP differences, host f64 from the layer APs
    integral = (want_c == np.round(want_c)) & (np.abs(want_c) >= 1) & ~big   # peak-position criteria
    rest = ~big & ~integral
    assert (crit[integral] == want_c[integral]).all(), (name, crit, want_c)
    if big.any():
        assert (np.abs(crit[big] - want_c[big]) / np.abs(want_c[big])).max() < 1e-9, (name, crit, want_c)
    assert np.abs(crit[rest] - want_c[rest]).max() < 1e-4, (name, crit, want_c)


@pytest.mark.parametrize("name", [n for n in META if "error" not in META[n]])
def test_option_criteria_match_reference(built, tmp_path, name):
    d, meta = _workdir(tmp_path, name)
    g = np.load(os.path.join(GOLDEN, "golden_options.npz"))
    params = g[name + "/params"]
    ev = hostlib.Evaluator(d, with_device=True)
    try:
        assert ev.n_crit == g[name + "/criteria"].shape[1]
        one = [ev.eval(p) for p in params]
        _check(name, np.array([c for c, _ in one]), np.array([v for _, v in one]), g)
        crit, viol = ev.eval_batch(params, threads=2)
        _check(name, crit, viol, g)
    finally:
        ev.close()
    if name.startswith("approx_gate"):   # the golden must exercise both sides of the gate
        done = g[name + "/simulation_done"]
        assert 0 < done.sum() < len(done)


def test_option_host_fit_and_host_criteria_agree(built, tmp_path, monkeypatch):
    """The same variants through the host restatement of the fit and the host-side comparison (the paths taken when the
    device-side shortcuts do not apply): every combined-criteria variant once more with EKGSIM_B200_FIT=host."""
    g = np.load(os.path.join(GOLDEN, "golden_options.npz"))
    monkeypatch.setenv("EKGSIM_B200_FIT", "host")
    monkeypatch.setenv("EKGSIM_B200_HOST_CRITERIA", "1")
    for name in ("everything", "cmp_devlin", "approx_gate_rms"):
        d, meta = _workdir(tmp_path, name)
        ev = hostlib.Evaluator(d, with_device=True)
        try:
            crit, viol = ev.eval_batch(g[name + "/params"], threads=3)
            _check(name, crit, viol, g)
        finally:
            ev.close()


def test_leads_sum_mode_throws_like_the_reference(built, tmp_path):
    """[optimization] mode = 9: the reference constructs (one criterion) and throws in calculateFitness (sim.cpp:612-614)."""
    d, meta = _workdir(tmp_path, "leads_sum")
    ev = hostlib.Evaluator(d, with_device=True)
    try:
        assert ev.n_crit == 1
        vec = np.loadtxt(os.path.join(GOLDEN, "vectors256.txt"), delimiter=",")[0]
        with pytest.raises(RuntimeError) as e:
            ev.eval(vec)
        assert meta["error"] in str(e.value)
        with pytest.raises(RuntimeError) as e:
            ev.eval_batch(vec[None], threads=1)
        assert meta["error"] in str(e.value)
    finally:
        ev.close()
