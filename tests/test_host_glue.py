"""Host-side C++ (ekgsim_b200/host): ini/format readers, the evaluation glue and the CLI.
CPU tests compare the newly written glue with the reference's own glue outputs
(tests/golden/golden_glue256.npz, dumped from the compiled reference by oracle/ref_dump)."""
import os
import subprocess

import numpy as np
import pytest

import ekgio
import hostlib

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
README_VECTOR = "0.00035813,0.0890636,0.0632915,226.183,0.000369406,0.0965625,0.0523254,232.278,0.000710767,0.0720323,0.0187579,200.93,23,22,15,13"


@pytest.fixture(scope="module")
def testrun(built, tmp_path_factory):
    d = str(tmp_path_factory.mktemp("testrun"))
    ekgio.materialise_testrun(d)
    return d


def test_layer_coefficients_bit_identical_to_reference_glue(testrun):
    """sim.cpp:825-916 + nonlinearFit.h:92-168 restated: 24x9 coefficients, displaced leads, violation."""
    ev = hostlib.Evaluator(testrun, with_device=False)
    g = np.load(os.path.join(GOLDEN, "golden_glue256.npz"))
    for i in range(0, 256, 2):
        k, leads, viol = ev.layer_coefficients(g["params"][i])
        assert k.tobytes() == g["layer_k"][i].tobytes(), i
        assert np.abs(leads - g["leads_zyx"][i]).max() < 1e-12
        assert abs(viol - g["violation"][i]) < 1e-12
    ev.close()


def test_ap_formula_matches_oracle():
    from oracle import oracle
    rng = np.random.default_rng(0)
    for _ in range(200):
        k = np.array([rng.uniform(-90, 0), rng.uniform(1.5, 3.5), 100.0, rng.uniform(0.85, 0.95), rng.uniform(0.05, 0.2),
                      rng.uniform(3e-4, 1e-3), rng.uniform(0.01, 0.1), rng.uniform(0.01, 0.1), rng.uniform(200, 450)])
        t = rng.uniform(-20, 700)
        assert hostlib.lib().ekg_host_wohlfart_plus(k.ctypes.data, float(t)) == oracle.wohlfart_plus(k, t)


def test_cli_without_gpu_fails_loudly_like_the_reference(testrun, built):
    """Errors surface as 'runtime error caught: ...' on stdout and exit code 0 (main.cpp:405-413)."""
    if built.lib().ekg_device_count() > 0:
        pytest.skip("a GPU is present")
    r = subprocess.run([hostlib.CLI, "test", "-sim", README_VECTOR, "-out", "result"], cwd=testrun, capture_output=True, text=True)
    assert r.returncode == 0
    assert " parameters for single simulator run: <0.00035813,0.0890636" in r.stdout
    assert " u0 = -0,0.619547,0.78496" in r.stdout          # SURVEY 9 console transcript
    assert "runtime error caught: " in r.stdout and "no CPU fallback" in r.stdout
    assert "criteria" not in r.stdout


def test_ini_semantics(tmp_path, built):
    """Last duplicate wins, names are space/case sensitive, unknown neighbourhood throws."""
    d = str(tmp_path)
    ekgio.materialise_testrun(d, ini_edit=lambda s: s.replace("neighbourhood type = 3D4", "neighbourhood type = 3D4\nneighbourhood type = 5D9"))
    with pytest.raises(RuntimeError) as e:
        hostlib.Evaluator(d, with_device=False)
    assert "unknown neighbourhood" in str(e.value)
    ekgio.materialise_testrun(d, ini_edit=lambda s: s.replace("interpolation = endo-mid-epi", "interpolation = sideways"))
    with pytest.raises(RuntimeError) as e:
        hostlib.Evaluator(d, with_device=False)
    assert "unknown interpolation type [sideways]" in str(e.value)


def test_missing_files(tmp_path, built):
    d = str(tmp_path)
    ekgio.materialise_testrun(d)
    os.remove(os.path.join(d, "conduction_24.matrix"))
    with pytest.raises(RuntimeError) as e:
        hostlib.Evaluator(d, with_device=False)
    assert "could not open file conduction_24.matrix" in str(e.value)


def test_char_matrix_and_2d_model(tmp_path, built):
    """One-character-per-voxel .matrix (matrix.h:206-218): digits, letters = 10.., X = start in layer 1."""
    d = str(tmp_path)
    ekgio.materialise_testrun(d)
    rows = ["0000000000", "0112233440", "01X2233440", "0112233AB0", "0000000000"]
    with open(os.path.join(d, "model_24.matrix"), "w") as f:
        f.write("tiny\n2d 10 x 5\n" + "\n".join(rows) + "\n")
    ev = hostlib.Evaluator(d, with_device=False, n_layers=11)   # highest layer 'B' = 11
    ev.close()


def test_binary_shape_cache(tmp_path, built, monkeypatch):
    """EKGSIM_B200_CACHE=1: <shape>.b200bin is written on the first parse; a fresh side-car is used by default (also
    without the variable), ignored when the text file changes or with EKGSIM_B200_CACHE=0."""
    d = str(tmp_path)
    ekgio.materialise_testrun(d)
    monkeypatch.setenv("EKGSIM_B200_CACHE", "1")
    hostlib.Evaluator(d, with_device=False).close()
    cache = os.path.join(d, "model_24.matrix.b200bin")
    assert os.path.getsize(cache) == 48 + 124 * 124 * 93 * 2
    raw = np.fromfile(cache, dtype=np.uint16, offset=48).reshape(124, 124, 93)
    assert (raw == ekgio.load_model24()["layers"]).all()
    # corrupt the text but keep size and mtime -> the cache is trusted
    st = os.stat(os.path.join(d, "model_24.matrix"))
    txt = open(os.path.join(d, "model_24.matrix")).read()
    at = txt.index(" 1 ", 100) + 1          # a layer-1 voxel somewhere in the body -> layer 2, same file size
    with open(os.path.join(d, "model_24.matrix"), "r+") as f:
        f.seek(at)
        f.write("2")
    os.utime(os.path.join(d, "model_24.matrix"), (st.st_atime, st.st_mtime))
    ev = hostlib.Evaluator(d, with_device=False)
    g = np.load(os.path.join(GOLDEN, "golden_glue256.npz"))
    k, _, _ = ev.layer_coefficients(g["params"][0])
    assert k.tobytes() == g["layer_k"][0].tobytes()
    ev.close()
    # reading a fresh side-car needs no opt-in; nothing is written without it
    monkeypatch.delenv("EKGSIM_B200_CACHE")
    before = os.stat(cache).st_mtime_ns
    hostlib.Evaluator(d, with_device=False).close()
    assert os.stat(cache).st_mtime_ns == before
    monkeypatch.setenv("EKGSIM_B200_CACHE", "1")
    # a newer text file invalidates it
    os.utime(os.path.join(d, "model_24.matrix"), (st.st_atime, st.st_mtime + 5))
    hostlib.Evaluator(d, with_device=False).close()
    raw2 = np.fromfile(cache, dtype=np.uint16, offset=48)
    assert (raw2 != raw.reshape(-1)).sum() == 1


def test_extern_client_speaks_the_server_protocol(tmp_path, built):
    """`ekgSim -extern <homeDir> -server <socket>` against a stand-in server written in Python: request =
    'EKG1' | u32 n | f64 genes[n], reply = i32 0 | u32 n | f64 violation | f64 criteria[n]  (ekg_server.h).
    No GPU and no simulator.ini needed on the client side."""
    import socket
    import struct
    import threading
    d = str(tmp_path)
    os.makedirs(os.path.join(d, "process3"))
    genes = [0.00035813, 0.0890636, 226.183, -13.5]
    with open(os.path.join(d, "process3", "input.txt"), "w") as f:
        f.write("# file generated by ExternalEvaluation class\n" + "\n".join("%.17g" % g for g in genes) + "\n")
    path = os.path.join(d, "s.sock")
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(path)
    srv.listen(4)
    seen = {}

    def serve_two():
        for k in range(2):
            c, _ = srv.accept()
            magic, n = struct.unpack("<II", c.recv(8, socket.MSG_WAITALL))
            seen[k] = (magic, struct.unpack("<%dd" % n, c.recv(8 * n, socket.MSG_WAITALL)))
            if k == 0:
                c.sendall(struct.pack("<iId2d", 0, 2, 240.609, 1.48975, 1.22272))
            else:
                msg = b"chromosome size does not agree"
                c.sendall(struct.pack("<iI", -1, len(msg)) + msg)
            c.close()

    th = threading.Thread(target=serve_two, daemon=True)
    th.start()
    r = subprocess.run([hostlib.CLI, "-extern", "process3", "-server", path], cwd=d, capture_output=True, text=True, timeout=30)
    assert "All done" in r.stdout, r.stdout
    crit, viol = hostlib.read_extern_output(os.path.join(d, "process3", "output.txt"), 2)   # ExternalEvaluation::readOut restated
    assert crit == [1.48975, 1.22272] and viol == 240.609
    assert seen[0] == (0x31474B45, tuple(genes))
    # an error reply surfaces like every other error of the CLI (main.cpp:405-413)
    r = subprocess.run([hostlib.CLI, "-extern", "process3"], cwd=d, capture_output=True, text=True, timeout=30,
                       env=dict(os.environ, EKGSIM_B200_SERVER=path))
    th.join(timeout=10)
    assert "runtime error caught: chromosome size does not agree" in r.stdout
    srv.close()


def test_builtin_test_shape_matches_reference(tmp_path):
    """`[model] shape =` (empty) makes the reference generate a 2-D test ring and export it as autoGeneratedShape.matrix
    (InputLoader::generateTestShape / loadShape, simulator.h:600-651).  The facade does the same: the exported file equals
    the reference's byte for byte (golden written by `ref_dump testshape`, tests/golden/run_reference.sh), and reading it
    back through the one-character-per-voxel branch of the .matrix parser (matrix.h:206-218) returns the same layers."""
    out = str(tmp_path / "autoGeneratedShape.matrix")
    layers = hostlib.generate_test_shape(out)
    assert layers.shape == (1, 160, 120)
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "autoGeneratedShape_ref.matrix")
    assert open(out, "rb").read() == open(golden, "rb").read()
    starts = np.argwhere(layers & 0x1000)
    assert starts.tolist() == [[0, 132, 30]] and int(layers[0, 132, 30]) == 0x1001      # first ring voxel above the centre line
    assert int((layers & 0x0FFF).max()) == 12 and int(((layers & 0x0FFF) > 0).sum()) == 6346   # the ring is clipped by the grid
    back = hostlib.load_shape(out)
    assert back.shape == layers.shape and (back == layers).all()      # 'X' reads back as start flag + layer 1, which is what it was


def test_run_approximation_matches_oracle(built):
    """EkgSim::runApproximation (the "string model", simulator.cpp:552-559) of the facade against the oracle's restatement:
    same expression per sample, host f64 on both sides -> identical bits.  Sizes: the testRun axis, a fractional step
    (result.size() = (size_t)(length / step)), a non-zero start."""
    from oracle import oracle
    g = np.load(os.path.join(GOLDEN, "golden_glue256.npz"))
    for vec, (start, length, step, delay) in zip((0, 7, 100), ((100, 400, 1.0, 10.0), (0, 50, 0.5, 0.0), (37, 100, 1.0 / 3.0, 2.5))):
        k = g["layer_k"][vec]
        got = hostlib.run_approximation(k, start, length, step, delay)
        want = oracle.run_approximation(k, float(start), step, float(length), delay)
        assert got.shape == want.shape == (int(length / step),)
        assert got.tobytes() == want.tobytes()
        # and against the formula itself for one sample (endo delayed minus epi, layer APs carry at = 0)
        i = len(got) // 2
        t = start + i * step
        assert got[i] == hostlib.lib().ekg_host_wohlfart_plus(k[0].ctypes.data, t + delay) - hostlib.lib().ekg_host_wohlfart_plus(k[-1].ctypes.data, t)


def test_cxx_host_slab_cuts_equal_the_python_driver(built, model24):
    """`ekgSim -slabs N` (EkgSim::balancedSlabs, sim_lib.h) cuts a model into z-slabs exactly like ekgsim_b200/dist.py::slab_ranges
    does for the torchrun layout: the two hosts shard one model the same way, also when slabs come out empty."""
    import ctypes as C
    from ekgsim_b200 import dist as ekdist
    rng = np.random.default_rng(5)
    tiny = np.zeros((3, 4, 5), dtype=np.uint16)
    tiny[1, 2, 3] = 1
    sparse = (rng.random((17, 9, 11)) < 0.2).astype(np.uint16) * 3
    sparse[5:9] = 0
    for layers in (model24["layers"], tiny, sparse):
        l = np.ascontiguousarray(layers, dtype=np.uint16)
        occ = ((l & 0x0FFF) > 0).sum(axis=(1, 2))
        for n in (1, 2, 3, 4, 8, 16):
            out = np.zeros((n, 2), dtype=np.int64)
            rc = hostlib.lib().ekg_host_balanced_slabs(l.ctypes.data_as(C.c_void_p), C.c_int64(l.shape[0]), C.c_int64(l.shape[1]), C.c_int64(l.shape[2]),
                                                       C.c_int(n), out.ctypes.data_as(C.c_void_p))
            assert rc == 0
            assert [tuple(int(v) for v in row) for row in out] == ekdist.slab_ranges(occ, n), (l.shape, n)
