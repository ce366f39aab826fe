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


def test_batch_256_criteria_match_oracle(testrun):
    """All 256 vectors of the config-3 batch, parameter vector -> criteria through the product's own glue
    and GPU path, against criteria derived with the reference-pinned oracle
    (tests/golden/make_criteria256.py).  Tolerance 1e-4 (north_star)."""
    g = np.load(os.path.join(GOLDEN, "golden_glue256.npz"))
    want = np.load(os.path.join(GOLDEN, "golden_criteria256.npz"))
    ev = hostlib.Evaluator(testrun, with_device=True)
    crit, viol = ev.eval_batch(g["params"], threads=0)
    ev.close()
    err = np.abs(crit - want["criteria"])
    print("256-vector criteria: worst |diff| %.3g" % err.max())
    assert err.max() < 1e-4
    assert np.abs(viol - want["violation"]).max() < 1e-9


def test_cli_extern_protocol(testrun, golden):
    """AMS-DEMO ExternalEvaluation: <cmd> <homeDir>, input.txt -> output.txt (ExternalEvaluation.h:95-151)."""
    home = os.path.join(testrun, "process1")
    os.makedirs(home, exist_ok=True)
    i = list(golden["name"]).index("full3")
    with open(os.path.join(home, "input.txt"), "w") as f:
        f.write("# file generated by ExternalEvaluation class\n" + "\n".join("%.17g" % v for v in golden["params"][i]) + "\n")
    r = subprocess.run([hostlib.CLI, "-extern", "process1"], cwd=testrun, capture_output=True, text=True)
    assert r.returncode == 0 and "caught" not in r.stdout, r.stdout
    crit, viol = hostlib.read_extern_output(os.path.join(home, "output.txt"), 2)     # the reference's own reader, restated
    assert np.abs(np.array(crit) - golden["criteria"][i]).max() < 1e-4
    assert viol == golden["violation"][i]
    # an out-of-bounds vector: the violation must reach the optimizer through that reader
    j = list(golden["name"]).index("full1")
    with open(os.path.join(home, "input.txt"), "w") as f:
        f.write("# file generated by ExternalEvaluation class\n" + "\n".join("%.17g" % v for v in golden["params"][j]) + "\n")
    subprocess.run([hostlib.CLI, "-extern", "process1"], cwd=testrun, capture_output=True, text=True)
    crit, viol = hostlib.read_extern_output(os.path.join(home, "output.txt"), 2)
    assert viol > 0 and abs(viol - golden["violation"][j]) < 1e-9 and np.abs(np.array(crit) - golden["criteria"][j]).max() < 1e-4


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


def test_excitation_sequence_dump_and_reload(built, tmp_path, model24, model24_delay):
    """[output files] excitation sequence -> 3-decimal dump (simulator.cpp:71-106); [model] excitation
    sequence -> loadExcitationSequence path (simulator.cpp:288-367) instead of the automaton."""
    d = str(tmp_path)
    ekgio.materialise_testrun(d, ini_edit=lambda s: s.replace("results filename = result.column",
                                                              "results filename = result.column\nexcitation sequence = es_out.matrix"))
    r = subprocess.run([hostlib.CLI, "test", "-sim", README_VECTOR, "-out", "result"], cwd=d, capture_output=True, text=True)
    assert "caught" not in r.stdout, r.stdout[-400:]
    assert "calculating excitation sequence" in r.stderr
    lines = open(os.path.join(d, "es_out.matrix")).read().split("\n")
    assert lines[0].startswith("Excitation file") and lines[1] == "3D 93 x 124 x 124 tab"
    dump = np.array(" ".join(lines[2:]).split(), dtype=np.float64).reshape(124, 124, 93)
    assert np.abs(dump - model24_delay).max() <= 0.0005 + 1e-12 and dump[62, 62, 67] == 1.0
    first = np.array([[float(x) for x in ln.split("\t")[1:]] for ln in open(os.path.join(d, "result.column")).read().split("\n")[3:]]).T
    # second run: read the (rounded) sequence back instead of computing it
    ekgio.materialise_testrun(d, ini_edit=lambda s: s.replace("points = model_24_measuring_pos.txt",
                                                              "points = model_24_measuring_pos.txt\nexcitation sequence = es_out.matrix"))
    r = subprocess.run([hostlib.CLI, "test", "-sim", README_VECTOR, "-out", "result"], cwd=d, capture_output=True, text=True)
    assert "caught" not in r.stdout, r.stdout[-400:]
    assert "loading excitation sequence" in r.stderr and "calculating excitation sequence" not in r.stderr
    second = np.array([[float(x) for x in ln.split("\t")[1:]] for ln in open(os.path.join(d, "result.column")).read().split("\n")[3:]]).T
    # the dump came with its full-precision side-car (es_out.matrix.b200bin: raw f64, the input-pipeline cache of SURVEY
    # 8(f)), which is fresh -> the reloaded map is the computed one, bit for bit, and so is the ECG
    side = os.path.join(d, "es_out.matrix.b200bin")
    assert os.path.getsize(side) == 48 + 124 * 124 * 93 * 8
    assert np.fromfile(side, dtype=np.float64, offset=48).tobytes() == model24_delay.tobytes()
    assert np.abs(second - first).max() == 0.0
    # without the side-car (EKGSIM_B200_CACHE=0, or a dump written by the reference): the 3-decimal text, like the
    # reference's loadExcitationSequence -- delays rounded to 1e-3 ms move the ECG a little, but only a little
    r = subprocess.run([hostlib.CLI, "test", "-sim", README_VECTOR, "-out", "result"], cwd=d, capture_output=True, text=True,
                       env=dict(os.environ, EKGSIM_B200_CACHE="0"))
    assert "caught" not in r.stdout and "loading excitation sequence" in r.stderr
    third = np.array([[float(x) for x in ln.split("\t")[1:]] for ln in open(os.path.join(d, "result.column")).read().split("\n")[3:]]).T
    assert 0 < np.abs(third - first).max() < 5e-2 * np.abs(first).max()
    # a text file that changed after the side-car was written is trusted over it
    os.utime(os.path.join(d, "es_out.matrix"), None)
    st = os.stat(os.path.join(d, "es_out.matrix"))
    os.utime(os.path.join(d, "es_out.matrix"), (st.st_atime, st.st_mtime + 7))
    subprocess.run([hostlib.CLI, "test", "-sim", README_VECTOR, "-out", "result"], cwd=d, capture_output=True, text=True)
    fourth = np.array([[float(x) for x in ln.split("\t")[1:]] for ln in open(os.path.join(d, "result.column")).read().split("\n")[3:]]).T
    assert np.abs(fourth - third).max() == 0.0


def test_cli_layer_and_cell_ap_outputs(testrun, golden, model24, model24_delay):
    """-out layer_aps / -out "cell_aps i j": the per-class AP table in first-seen raster order
    (Simulation::setApIndices, simulator.cpp:561-621; sim.cpp:918-991)."""
    from oracle import oracle
    r = subprocess.run([hostlib.CLI, "test", "-sim", README_VECTOR, "-out", "layer_aps", "-out", "cell_aps 0 5 50647 99999"],
                       cwd=testrun, capture_output=True, text=True)
    assert "caught" not in r.stdout, r.stdout[-400:]
    i = list(golden["name"]).index("full1")
    lay = open(os.path.join(testrun, "layer_aps.column")).read().split("\n")
    assert lay[0] == "#Comment: " and lay[1].startswith("# ap_0 : k = [0, 2.5, 100, 0.9, 0.1, 0.00035813, 0.0890636, 0.0632915, 226.183]")
    hdr = [ln for ln in lay if ln.startswith("#Time[ms]")][0].split("\t")
    assert hdr[1:] == ["ap_%d" % n for n in range(24)]
    rows = [ln for ln in lay if not ln.startswith("#")]
    assert len(rows) == 700
    v = np.array([float(x) for x in rows[123].split("\t")])
    assert v[0] == 123.0
    want = np.array([oracle.wohlfart_plus(golden["layer_k"][i][l], 123.0) for l in range(24)])
    assert np.abs(v[1:] - want).max() < 1e-4 * np.abs(want).max()
    cell = open(os.path.join(testrun, "cell_aps.column")).read().split("\n")
    chdr = [ln for ln in cell if ln.startswith("#Time[ms]")][0].split("\t")
    assert chdr[1:] == ["ap_0", "ap_5", "ap_50647"]          # 99999 is beyond the 50648 classes -> skipped
    K, idx = oracle.ap_classes(model24["layers"], model24_delay, 24)
    crow = np.array([float(x) for x in [ln for ln in cell if not ln.startswith("#")][200].split("\t")])
    for col, cls in enumerate((0, 5, 50647)):
        vox = np.argwhere(idx == cls)[0]
        layer = int(model24["layers"][tuple(vox)] & 0x0FFF)
        ref = oracle.lib().ekg_oracle_ap(golden["layer_k"][i][layer - 1].ctypes.data, float(model24_delay[tuple(vox)]), 200.0)
        assert abs(crow[1 + col] - ref) < 1e-4 * max(1.0, abs(ref)), (cls, crow[1 + col], ref)


def test_reference_optimizer_drives_the_b200_simlib(built, tmp_path):
    """BASELINE config 5 in miniature: the reference's own AMS-DEMO optimizer (unmodified main.cpp, sequential
    -DNO_MPI path, main.cpp:283-361) evaluating its population through the B200 simulator library.
    Population 12, 2 generations; checks that the run completes and logs every evaluation."""
    exe = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "ekgSim_refglue_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ekgSim_refglue_b200 not built (needs /root/reference at build time)")
    d = str(tmp_path)
    ekgio.materialise_testrun(d, ini_edit=lambda s: s.replace("population size = 100", "population size = 12")
                              .replace("number of generations = 100", "number of generations = 2"))
    r = subprocess.run([exe], cwd=d, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "caught" not in r.stdout, r.stdout[-600:]
    assert r.stdout.rstrip().endswith("All done")
    ev = [ln for ln in open(os.path.join(d, "evaluations.txt")).read().split("\n") if ln and not ln.startswith("#")]
    assert len(ev) >= 24, len(ev)          # initial population + the offspring of the generations
    assert os.path.exists(os.path.join(d, "individuals.txt"))


def test_batch_device_fit_equals_host_fit(testrun, monkeypatch):
    """evalBatch with the layer fits on the device (default, ekg_evaluate) and on host threads
    (EKGSIM_B200_FIT=host, the reference's procedure restated on the CPU): same criteria."""
    import time
    g = np.load(os.path.join(GOLDEN, "golden_glue256.npz"))
    ev = hostlib.Evaluator(testrun, with_device=True)
    ev.eval_batch(g["params"])
    t0 = time.time(); crit_dev, viol_dev = ev.eval_batch(g["params"]); t_dev = time.time() - t0
    one, _ = ev.eval(g["params"][5])                      # single evaluation: ekg_fit_layers + ekg_simulate
    ev.close()
    monkeypatch.setenv("EKGSIM_B200_FIT", "host")
    ev = hostlib.Evaluator(testrun, with_device=True)
    ev.eval_batch(g["params"][:16])
    t0 = time.time(); crit_host, viol_host = ev.eval_batch(g["params"]); t_host = time.time() - t0
    ev.close()
    # device fit, but the curve comparison on the host from downloaded ECGs
    monkeypatch.delenv("EKGSIM_B200_FIT")
    monkeypatch.setenv("EKGSIM_B200_HOST_CRITERIA", "1")
    ev = hostlib.Evaluator(testrun, with_device=True)
    ev.eval_batch(g["params"])
    t0 = time.time(); crit_hc, _ = ev.eval_batch(g["params"]); t_hc = time.time() - t0
    ev.close()
    assert np.abs(crit_dev - crit_hc).max() < 1e-12
    print("evalBatch(256): device fit + device criteria %.2f ms, device fit + host criteria %.2f ms, host fit %.1f ms"
          % (t_dev * 1e3, t_hc * 1e3, t_host * 1e3))
    assert np.abs(crit_dev - crit_host).max() < 1e-7
    assert (viol_dev == viol_host).all()
    assert np.abs(one - crit_dev[5]).max() < 1e-6


def test_evaluation_server_batches_concurrent_extern_clients(testrun, golden):
    """`ekgSim -serve <socket>` keeps the model on the GPU; 12 concurrent `ekgSim -extern process<i>`
    clients (what AMS-DEMO's MPI worker ranks start, ExternalEvaluation.h:95-151) are answered from it, several per batch."""
    import time
    g = np.load(os.path.join(GOLDEN, "golden_glue256.npz"))
    want = np.load(os.path.join(GOLDEN, "golden_criteria256.npz"))
    sock = os.path.join(testrun, "ekg.sock")
    srv = subprocess.Popen([hostlib.CLI, "-serve", sock], cwd=testrun, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    try:
        for _ in range(600):
            if os.path.exists(sock) or srv.poll() is not None:
                break
            time.sleep(0.1)
        assert os.path.exists(sock), srv.stderr.read()
        n = 12
        for i in range(n):
            os.makedirs(os.path.join(testrun, "process%d" % i), exist_ok=True)
            with open(os.path.join(testrun, "process%d" % i, "input.txt"), "w") as f:
                f.write("# file generated by ExternalEvaluation class\n" + "\n".join("%.17g" % v for v in g["params"][i]) + "\n")
        env = dict(os.environ, EKGSIM_B200_SERVER=sock)
        t0 = time.time()
        procs = [subprocess.Popen([hostlib.CLI, "-extern", "process%d" % i], cwd=testrun, env=env, stdout=subprocess.PIPE,
                                  stderr=subprocess.DEVNULL, text=True) for i in range(n)]
        outs = [p.communicate(timeout=120)[0] for p in procs]
        dt = time.time() - t0
        for i in range(n):
            assert "All done" in outs[i] and "error" not in outs[i], outs[i]
            crit, viol = hostlib.read_extern_output(os.path.join(testrun, "process%d" % i, "output.txt"), 2)
            assert np.abs(np.array(crit) - want["criteria"][i]).max() < 1e-4, (i, crit)
            assert abs(viol - want["violation"][i]) < 1e-9
        # one more, alone, timed: connection + evaluation + files
        t1 = time.time()
        subprocess.run([hostlib.CLI, "-extern", "process0"], cwd=testrun, env=env, capture_output=True, text=True, timeout=60)
        one = time.time() - t1
        print("evaluation server: %d concurrent -extern clients in %.3f s, one client alone %.1f ms" % (n, dt, one * 1e3))
        # a chromosome of the wrong length is refused without disturbing the server
        with open(os.path.join(testrun, "process1", "input.txt"), "w") as f:
            f.write("1\n2\n3\n")
        r = subprocess.run([hostlib.CLI, "-extern", "process1"], cwd=testrun, env=env, capture_output=True, text=True, timeout=60)
        assert "runtime error caught: chromosome size does not agree" in r.stdout
        r = subprocess.run([hostlib.CLI, "-extern", "process2"], cwd=testrun, env=env, capture_output=True, text=True, timeout=60)
        assert "All done" in r.stdout and "error" not in r.stdout
        # a client that stalls in the middle of its request (a crashed or paused worker) is dropped after the server's
        # 2 s socket timeout instead of blocking everybody else; several devices behind one server ("0,0": two replicas)
        import socket
        import struct
        stalled = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        stalled.connect(sock)
        stalled.sendall(struct.pack("<I", 0x31474B45))          # half a header, then silence
        t1 = time.time()
        r = subprocess.run([hostlib.CLI, "-extern", "process3"], cwd=testrun, env=env, capture_output=True, text=True, timeout=60)
        assert "All done" in r.stdout and "error" not in r.stdout and time.time() - t1 < 15
        stalled.close()
    finally:
        subprocess.run([hostlib.CLI, "-shutdown", sock], cwd=testrun, capture_output=True, timeout=30)
        try:
            err = srv.communicate(timeout=30)[1]
        except subprocess.TimeoutExpired:
            srv.kill()
            err = srv.communicate()[1]
    assert "evaluations in" in err, err[-500:]
    assert not os.path.exists(sock)


AMS_DEMO_SETTINGS = """[evaluation]
command line = {cli} -extern
input file name = input.txt
output file name = output.txt
chromosome vector length = 16
criteria vector length = 2
properties vector length = 0

[optimization]
random seed = 11
population size = {pop}
max number of generations = {gens}
DE schema = rand/1/bin
p crossover = 0.3
scaling factors = 0.5
queue length = 1

[initial population]
gene min = 0.0003, 0.01, 0.01, 200, 0.0003, 0.01, 0.01, 200, 0.0003, 0.01, 0.01, 200, -50, -50, -50, -50
gene max = 0.001, 0.1, 0.1, 400, 0.001, 0.1, 0.1, 400, 0.001, 0.1, 0.1, 400, 50, 50, 50, 50
"""


def test_reference_ams_demo_drives_extern_and_server(built, tmp_path):
    """BASELINE config 5 as a drop-in: the reference's UNMODIFIED stand-alone optimizer (AMS-DEMO/main.cpp, sequential
    build, oracle/_ref/DEMO_ref) evaluates through its ExternalEvaluation protocol -- `ekgSim -extern <homeDir>` per
    individual, answered by the resident `ekgSim -serve`.  Every row of its evaluations.txt must carry the criteria and
    the violation this build computes for the genes the optimizer wrote to input.txt (6 significant digits,
    ExternalEvaluation.h:111-112)."""
    import time
    demo = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "DEMO_ref")
    if not os.path.exists(demo):
        pytest.skip("oracle/_ref/DEMO_ref not built (needs /root/reference at build time)")
    d = str(tmp_path)
    ekgio.materialise_testrun(d, targets="target_ecg_v2_v6.column")
    pop, gens = 12, 2
    with open(os.path.join(d, "settings.ini"), "w") as f:
        f.write(AMS_DEMO_SETTINGS.format(cli=hostlib.CLI, pop=pop, gens=gens))
    sock = os.path.join(d, "ekg.sock")
    srv = subprocess.Popen([hostlib.CLI, "-serve", sock], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    try:
        for _ in range(600):
            if os.path.exists(sock) or srv.poll() is not None:
                break
            time.sleep(0.1)
        assert os.path.exists(sock)
        t0 = time.time()
        r = subprocess.run([demo], cwd=d, capture_output=True, text=True, timeout=600, env=dict(os.environ, EKGSIM_B200_SERVER=sock))
        dt = time.time() - t0
    finally:
        subprocess.run([hostlib.CLI, "-shutdown", sock], cwd=d, capture_output=True, timeout=30)
        try:
            err = srv.communicate(timeout=30)[1]
        except subprocess.TimeoutExpired:
            srv.kill()
            err = srv.communicate()[1]
    assert "last front saved as front.txt" in r.stdout, r.stdout[-500:]
    rows = [ln.split("\t") for ln in open(os.path.join(d, "evaluations.txt")) if ln.strip() and not ln.startswith("#")]
    assert len(rows) >= pop * gens
    vec = lambda s: np.array([float(x) for x in s.strip().strip("<>").split(",") if x])
    genes = np.array([vec(rw[1]) for rw in rows])
    seen = np.array([[float("%g" % v) for v in gn] for gn in genes])        # what writeIn put into input.txt
    viol = np.array([float(rw[2]) for rw in rows])
    crit = np.array([vec(rw[4]) for rw in rows])
    ev = hostlib.Evaluator(d, with_device=True)
    want_c, want_v = ev.eval_batch(seen)
    ev.close()
    # (one vector per server batch there, one batch of all here: the fp32 partial sums group differently)
    assert np.abs(crit - want_c).max() < 1e-6 and np.abs(viol - want_v).max() < 1e-9
    assert (viol > 0).any() and (viol == 0).any()      # the gene box is wider than the k8 limits of the ini: both cases occur
    print("AMS-DEMO (reference optimizer) + ekgSim -extern + server: %d evaluations in %.2f s (%.1f ms each); server: %s"
          % (len(rows), dt, 1e3 * dt / len(rows), err.strip().split("\n")[-1]))


def test_reference_main_with_embedded_optimizer_on_the_b200_evaluator(built, tmp_path, golden):
    """The reference's UNMODIFIED main.cpp -- CLI and embedded AMS-DEMO optimizer -- linked against
    ekgsim_b200/host/compat/sim_b200.cpp (struct OptimizationFunction of sim.h, the in-process
    VirtualOptimizationFunction interface, over ekg::Evaluator) instead of sim.cpp + simlib:
    (1) `test -sim` prints the reference's criteria, (2) `ekgSim` without arguments runs the optimizer and every
    row of its evaluations.txt carries the criteria this build computes for that chromosome."""
    import re
    import time
    exe = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "ekgSim_refmain_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ekgSim_refmain_b200 not built (needs /root/reference at build time)")
    d = str(tmp_path)
    ekgio.materialise_testrun(d, ini_edit=lambda s: s.replace("population size = 100", "population size = 20")
                              .replace("number of generations = 100", "number of generations = 3"))
    i = list(golden["name"]).index("full1")
    r = subprocess.run([exe, "test", "-sim", ",".join("%.17g" % v for v in golden["params"][i]), "-out", "result"], cwd=d,
                       capture_output=True, text=True, timeout=300)
    m = re.search(r" criteria = <([0-9.e+-]+),([0-9.e+-]+)>, violation = ([0-9.e+-]+)\n", r.stdout)
    assert m, r.stdout[-600:]
    assert abs(float(m.group(1)) - golden["criteria"][i][0]) < 1e-4 and abs(float(m.group(2)) - golden["criteria"][i][1]) < 1e-4
    assert abs(float(m.group(3)) - golden["violation"][i]) < 1e-3
    t0 = time.time()
    r = subprocess.run([exe], cwd=d, capture_output=True, text=True, timeout=600)
    dt = time.time() - t0
    assert r.returncode == 0 and "caught" not in r.stdout and r.stdout.rstrip().endswith("All done"), r.stdout[-600:]
    rows = [ln.split("\t") for ln in open(os.path.join(d, "evaluations.txt")) if ln.strip() and not ln.startswith("#")]
    assert len(rows) >= 60
    vec = lambda s: np.array([float(x) for x in s.strip().strip("<>").split(",") if x])
    genes = np.array([vec(rw[1]) for rw in rows])
    ev = hostlib.Evaluator(d, with_device=True)
    want_c, want_v = ev.eval_batch(genes)
    ev.close()
    assert np.abs(np.array([vec(rw[4]) for rw in rows]) - want_c).max() < 1e-6
    assert np.abs(np.array([float(rw[2]) for rw in rows]) - want_v).max() < 1e-9
    print("reference main.cpp + embedded AMS-DEMO on the B200 evaluator: %d evaluations, %.2f s for the whole process" % (len(rows), dt))


def test_evaluator_splits_batches_over_devices(testrun):
    """Multi-GPU in the product's own C++ host (the in-process form of the reference's MPI task farm, main.cpp:283-361,
    ParallelFramework.h:388-429): one model replica per listed device, evalBatch splits the batch, one host thread per
    device.  "0,0" puts two replicas on GPU 0, so the splitting logic is exercised on a single-GPU box; "all" uses every
    GPU of the box.  Individuals are independent: the criteria must not depend on how the batch was cut."""
    g = np.load(os.path.join(GOLDEN, "golden_glue256.npz"))
    want = np.load(os.path.join(GOLDEN, "golden_criteria256.npz"))
    genes = g["params"][:67]   # an odd count: uneven shares
    ev1 = hostlib.Evaluator(testrun, with_device=True)
    c1, v1 = ev1.eval_batch(genes)
    ev1.close()
    for spec, n_dev in (("0,0", 2), ("0,0,0", 3), ("all", None)):
        ev = hostlib.Evaluator(testrun, devices=spec)
        if n_dev is not None:
            assert ev.n_devices == n_dev
        c, v = ev.eval_batch(genes)
        one, _ = ev.eval(genes[5])          # eval() runs on the first device
        ev.close()
        assert np.abs(c - want["criteria"][:67]).max() < 1e-4
        assert np.abs(c - c1).max() < 1e-6 and (v == v1).all()
        assert np.abs(one - c[5]).max() < 1e-6
    # fewer individuals than devices, and the CLI spelling
    ev = hostlib.Evaluator(testrun, devices="0,0,0")
    c, v = ev.eval_batch(genes[:2])
    ev.close()
    assert np.abs(c - c1[:2]).max() < 1e-6
    with open(os.path.join(testrun, "vec5.txt"), "w") as f:
        for p in genes[:5]:
            f.write(",".join("%.17g" % x for x in p) + "\n")
    r = subprocess.run([hostlib.CLI, "-batch", "vec5.txt", "-batchout", "crit5.txt", "-evalout", "evaluations5.txt", "-devices", "0,0"], cwd=testrun,
                       capture_output=True, text=True)
    assert r.returncode == 0 and "on 2 GPU(s)" in r.stdout, r.stdout + r.stderr
    got = np.loadtxt(os.path.join(testrun, "crit5.txt"))
    assert np.abs(got[:, :2] - c1[:5]).max() < 1e-6
    # -evalout: the batch in the format of the reference optimizer's evaluations.txt (AMS-DEMO/Individual.h:361-380)
    lines = open(os.path.join(testrun, "evaluations5.txt")).read().split("\n")
    assert lines[0] == "# format of this file:" and lines[1].startswith("# evaluation_number \t[function_input_vector] \t violation")
    rows = [ln.split("\t") for ln in lines[2:] if ln]
    assert len(rows) == 5 and all(len(rw) == 8 for rw in rows)
    for i, rw in enumerate(rows):
        assert int(rw[0]) == i and rw[3] == "<>" and rw[5] == "0"
        chrom = np.array([float(x) for x in rw[1].strip("<>").split(",")])
        assert (chrom == genes[i]).all()                                   # 26 significant digits: the genes survive exactly
        crit = np.array([float(x) for x in rw[4].strip("<>").split(",")])
        assert np.abs(crit - c1[i]).max() < 1e-6 and float(rw[2]) == v1[i]
    r = subprocess.run([hostlib.CLI, "-batch", "vec5.txt", "-devices", "7,99"], cwd=testrun, capture_output=True, text=True)
    assert "runtime error caught: no such CUDA device in device list" in r.stdout


def test_one_model_over_z_slabs_in_the_cxx_host(testrun, golden):
    """BASELINE config 4's sharding in the product's own C++ host: `EKGSIM_B200_SLABS` / `ekgSim -slabs` spreads ONE model
    over several GPUs -- every device sums the ECG over its z-slab, the facade adds the partial ECGs, the excitation
    sequence comes from the peer-linked automaton over the same slabs (one kernel per device, face planes through peer
    memory).  "0,0,0" puts three slabs on GPU 0, so the whole path runs on a single-GPU box; "all" uses every GPU.
    Same criteria as the unsharded evaluator, same console contract."""
    names = list(golden["name"])
    i = names.index("full1")
    g = np.load(os.path.join(GOLDEN, "golden_glue256.npz"))
    ev1 = hostlib.Evaluator(testrun, with_device=True)
    c1, v1 = ev1.eval(golden["params"][i])
    cb1, vb1 = ev1.eval_batch(g["params"][:9])
    ev1.close()
    for spec in ("0,0,0", "all", "0,0"):
        ev = hostlib.Evaluator(testrun, slabs=spec)
        c, v = ev.eval(golden["params"][i])
        cb, vb = ev.eval_batch(g["params"][:9])
        ev.close()
        assert np.abs(c - golden["criteria"][i]).max() < 1e-4 and v == v1, (spec, c)
        assert np.abs(c - c1).max() < 1e-6, (spec, c, c1)
        assert np.abs(cb - cb1).max() < 1e-6 and (vb == vb1).all(), spec
    # the CLI spelling; stderr says how the excitation sequence was computed
    r = subprocess.run([hostlib.CLI, "test", "-sim", README_VECTOR, "-out", "result", "-slabs", "0,0,0,0"], cwd=testrun, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "one model on 4 z-slabs (peer-linked excitation sequence)" in r.stderr, r.stderr
    m = re.search(r" criteria = <([0-9.e+-]+),([0-9.e+-]+)>, violation = ([0-9.e+-]+)\n", r.stdout)
    assert m, r.stdout
    assert abs(float(m.group(1)) - golden["criteria"][i][0]) < 1e-4 and abs(float(m.group(2)) - golden["criteria"][i][1]) < 1e-4
    assert m.group(3) == "240.609"
    lines = open(os.path.join(testrun, "result.column")).read().split("\n")
    vals = np.array([[float(x) for x in ln.split("\t")[1:]] for ln in lines[3:]]).T
    peak = np.abs(golden["ecg"][i]).max(axis=1, keepdims=True)
    assert (np.abs(vals - golden["ecg"][i]) / peak).max() < 1e-5 + 5e-6
    # the replicated automaton as the fallback (devices without native peer atomics take this path by themselves)
    env = dict(os.environ, EKGSIM_B200_SLAB_AUTOMATON="replicated")
    r = subprocess.run([hostlib.CLI, "test", "-sim", README_VECTOR, "-slabs", "2"], cwd=testrun, capture_output=True, text=True, env=env)
    if "no such CUDA device" not in r.stdout:      # a one-GPU box has no device 1
        assert "one model on 2 z-slabs (replicated excitation sequence)" in r.stderr, r.stderr
        assert m.group(0) in r.stdout
    r = subprocess.run([hostlib.CLI, "test", "-sim", README_VECTOR, "-slabs", "0,99"], cwd=testrun, capture_output=True, text=True)
    assert "runtime error caught: no such CUDA device in slab list" in r.stdout
