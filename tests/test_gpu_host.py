"""GPU tests of the drop-in surface: the `ekgSim` CLI and the evaluation glue driving the CUDA path,
against the criteria / ECGs the compiled reference produced for the same inputs.
Tolerances (north_star): ECG within 1e-5 of peak lead amplitude, criteria within 1e-4."""
import os
import re
import subprocess

import numpy as np
import pytest

import ekgio
import hostlib

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
README_VECTOR = "0.00035813,0.0890636,0.0632915,226.183,0.000369406,0.0965625,0.0523254,232.278,0.000710767,0.0720323,0.0187579,200.93,23,22,15,13"


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN, "golden_eval_full.npz"))


@pytest.fixture(scope="module")
def testrun(built, tmp_path_factory):
    d = str(tmp_path_factory.mktemp("testrun"))
    ekgio.materialise_testrun(d)
    return d


def test_cli_single_simulation_transcript_and_column(testrun, golden):
    r = subprocess.run([hostlib.CLI, "test", "-sim", README_VECTOR, "-out", "result"], cwd=testrun, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout
    # console contract: same lines, same formats
    assert " parameters for single simulator run: <0.00035813,0.0890636,0.0632915,226.183," in out
    assert "model: model_24.matrix (124, 124, 93)\n" in out
    assert "neighbourhood: 3D, 6 neighbours\n" in out
    assert "simulation start = 100\nsimulation time step = 1\nsimulation length = 400\ntotal steps = 400\n" in out
    assert "eval 1  " in out and out.rstrip().endswith("All done")  # "\reval 1  " (text mode turns \r into \n)
    m = re.search(r" criteria = <([0-9.e+-]+),([0-9.e+-]+)>, violation = ([0-9.e+-]+)\n", out)
    assert m, out
    i = list(golden["name"]).index("full1")
    assert abs(float(m.group(1)) - golden["criteria"][i][0]) < 1e-4
    assert abs(float(m.group(2)) - golden["criteria"][i][1]) < 1e-4
    assert m.group(3) == "240.609"
    secs = float(re.search(r" simulation done in ([0-9.e+-]+) seconds", out).group(1))
    assert secs < 60  # the reference needs ~205 s
    # result.column: header + 400 rows, values within 1e-5 of peak of the reference's f64 ECG
    lines = open(os.path.join(testrun, "result.column")).read().split("\n")
    assert lines[0].startswith("#Comment: input=<0.00035813,0.0890636,") and lines[0].endswith(">; ")
    assert lines[1] == "# "
    assert lines[2] == "#Time[ms]\t69.0145,128.133,-71.8312\t336.761,-112.667,183.886"
    assert len(lines) == 403 and lines[3].startswith(" 100.00000\t") and lines[-1].startswith(" 499.00000\t")
    vals = np.array([[float(x) for x in ln.split("\t")[1:]] for ln in lines[3:]]).T
    peak = np.abs(golden["ecg"][i]).max(axis=1, keepdims=True)
    assert (np.abs(vals - golden["ecg"][i]) / peak).max() < 1e-5 + 5e-6  # + the 6-digit print rounding
    assert re.fullmatch(r" 100\.00000\t[ -]?\d\.\d{5}e[+-]\d\d\t[ -]?\d\.\d{5}e[+-]\d\d", lines[3])
    assert os.path.exists(os.path.join(testrun, "target_chk.column"))


def test_eval_and_batch_criteria_match_reference(testrun, golden):
    ev = hostlib.Evaluator(testrun, with_device=True)
    names = list(golden["name"])
    idx = [names.index(n) for n in ("full1", "full2", "full3")]
    for i in idx:
        crit, viol = ev.eval(golden["params"][i])
        assert np.abs(crit - golden["criteria"][i]).max() < 1e-4, (names[i], crit, golden["criteria"][i])
        assert abs(viol - golden["violation"][i]) < 1e-9
    crit, viol = ev.eval_batch(golden["params"][idx], threads=3)
    assert np.abs(crit - golden["criteria"][idx]).max() < 1e-4
    assert np.abs(viol - golden["violation"][idx]).max() < 1e-9
    ev.close()


def test_v6_target_criteria(built, tmp_path, golden):
    d = str(tmp_path)
    ekgio.materialise_testrun(d, targets="target_ecg_v2_v6.column")
    i = list(golden["name"]).index("v6full")
    ev = hostlib.Evaluator(d, with_device=True)
    crit, viol = ev.eval(golden["params"][i])
    assert np.abs(crit - golden["criteria"][i]).max() < 1e-4 and viol == 0.0
    ev.close()


def test_batch_256_glue_plus_gpu(testrun):
    """Config 3 end to end: 256 parameter vectors -> criteria, glue on host threads, one GPU batch."""
    g = np.load(os.path.join(GOLDEN, "golden_glue256.npz"))
    ev = hostlib.Evaluator(testrun, with_device=True)
    crit, viol = ev.eval_batch(g["params"], threads=0)
    assert np.isfinite(crit).all(), crit[~np.isfinite(crit).all(axis=1)]
    assert (crit >= 0).all() and (crit <= 2.0 + 1e-9).all(), (crit.min(), crit.max())   # 1 - Pearson
    assert np.abs(viol - g["violation"]).max() < 1e-9
    one, v1 = ev.eval(g["params"][17])
    # same individual alone or inside the batch: only the fp32 grouping of partial sums differs
    assert np.abs(one - crit[17]).max() < 1e-6, (one, crit[17])
    ev.close()


def test_cli_extern_protocol(testrun, golden):
    """AMS-DEMO ExternalEvaluation: <cmd> <homeDir>, input.txt -> output.txt (ExternalEvaluation.h:95-151)."""
    home = os.path.join(testrun, "process1")
    os.makedirs(home, exist_ok=True)
    i = list(golden["name"]).index("full3")
    with open(os.path.join(home, "input.txt"), "w") as f:
        f.write("# file generated by ExternalEvaluation class\n" + "\n".join("%.17g" % v for v in golden["params"][i]) + "\n")
    r = subprocess.run([hostlib.CLI, "-extern", "process1"], cwd=testrun, capture_output=True, text=True)
    assert r.returncode == 0 and "caught" not in r.stdout, r.stdout
    txt = open(os.path.join(home, "output.txt")).read().split("\n")
    crit = [float(x) for x in txt[:2]]
    assert np.abs(np.array(crit) - golden["criteria"][i]).max() < 1e-4
    assert txt[2].startswith("# violation 0")


def test_reference_glue_on_b200_simlib(testrun, golden):
    """The reference's UNMODIFIED main.cpp + sim.cpp, compiled from /root/reference against the B200
    facade (oracle/Makefile target refglue, ekgsim_b200/host/compat/refglue_shim.h): only simlib is
    swapped, so this isolates the simulator library as a drop-in."""
    exe = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "ekgSim_refglue_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ekgSim_refglue_b200 not built (needs /root/reference at build time)")
    r = subprocess.run([exe, "test", "-sim", README_VECTOR, "-out", "result"], cwd=testrun, capture_output=True, text=True)
    assert r.returncode == 0 and "caught" not in r.stdout, r.stdout[-500:]
    m = re.search(r" criteria = <([0-9.e+-]+),([0-9.e+-]+)>, violation = ([0-9.e+-]+)\n", r.stdout)
    i = list(golden["name"]).index("full1")
    assert abs(float(m.group(1)) - golden["criteria"][i][0]) < 1e-4 and abs(float(m.group(2)) - golden["criteria"][i][1]) < 1e-4
    assert m.group(3) == "240.609"
    lines = open(os.path.join(testrun, "result.column")).read().split("\n")
    assert lines[2] == "#Time[ms]\t69.0145,128.133,-71.8312\t336.761,-112.667,183.886" and len(lines) == 403
