"""GPU tests on a model LARGER than model_24 (BASELINE config 4): the synthetic finer-resolution heart
ekgio.scaled_heart(f).  f = 2 (4.4 M occupied voxels, 248 x 248 x 186 grid) runs in the default GPU suite; its goldens
come from the oracle (tests/golden/make_heart_goldens.py; the oracle's loops are pinned to the compiled reference on
model_24).  Bar: activation map bit-exact (sha256 of the raw f64 raster), ECG within 1e-5 of the peak lead amplitude."""
import hashlib
import json
import os

import numpy as np
import pytest

import ekgio

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def heart2x(built):
    layers, transfer, leads = ekgio.scaled_heart(2)
    gold = json.load(open(os.path.join(GOLDEN, "golden_heart2x.json")))
    m = built.Model(layers, transfer, device=0)
    delay, visits = m.activation()
    g = np.load(os.path.join(GOLDEN, "golden_glue256.npz"))
    return dict(model=m, layers=layers, transfer=transfer, delay=delay, gold=gold, k=g["layer_k"][0],
                leads=np.array(gold["leads_zyx"]), visits=visits)


def test_heart2x_activation_bit_exact(heart2x):
    d, gold = heart2x["delay"], heart2x["gold"]
    assert list(d.shape) == gold["shape_zyx"]
    assert int(((heart2x["layers"] & 0x0FFF) > 0).sum()) == gold["occupied"] == heart2x["model"].num_voxels
    assert hashlib.sha256(d.tobytes()).hexdigest() == gold["sha256_f64_raster"]
    assert d.sum() == gold["sum"] and d.max() == gold["max"]


def test_heart2x_sweep_automaton_agrees(built, heart2x, monkeypatch):
    """the plain label-correcting sweep (the independent cross-check of the brick frontier) reaches the same bits"""
    monkeypatch.setenv("EKGSIM_B200_AUTOMATON", "sweep")
    m = built.Model(heart2x["layers"], heart2x["transfer"], device=0)
    d, _ = m.activation()
    assert d.tobytes() == heart2x["delay"].tobytes()
    m.close()


@pytest.mark.parametrize("mode", ["direct", "hoisted", "separable"])
def test_heart2x_ecg_vs_oracle(built, heart2x, mode):
    md = {"direct": built.MODE_DIRECT, "hoisted": built.MODE_HOISTED, "separable": built.MODE_SEPARABLE}[mode]
    gold = heart2x["gold"]
    peak = np.array(gold["peak_full"])[:, None]
    ecg = heart2x["model"].simulate(heart2x["k"], heart2x["leads"], "3D4", 100.0, 1.0, 400.0, mode=md)[0]
    e16 = np.abs(ecg[:, :16] - np.array(gold["ecg16"])) / peak
    assert e16.max() < 1e-5, (mode, e16.max())
    for t, want in gold["ecg_at_t"].items():
        assert (np.abs(ecg[:, int(t)] - np.array(want)) / peak[:, 0]).max() < 1e-5, (mode, t)
    assert (np.abs(np.abs(ecg).max(axis=1) - peak[:, 0]) / peak[:, 0]).max() < 1e-5
    # 16 samples alone (what SURVEY 8(d) config 4 prescribes for the CPU side) give the same values
    short = heart2x["model"].simulate(heart2x["k"], heart2x["leads"], "3D4", 100.0, 1.0, 16.0, mode=md)[0]
    assert (np.abs(short - np.array(gold["ecg16"])) / peak).max() < 1e-5


def test_heart2x_slabs_add_up(built, heart2x):
    """z-slab decomposition (config 4's sharding) on one device: partial ECGs of 3 slabs sum to the whole"""
    from ekgsim_b200 import dist as ekdist
    m, gold = heart2x["model"], heart2x["gold"]
    occ_z = ((heart2x["layers"] & 0x0FFF) > 0).sum(axis=(1, 2))
    slabs = ekdist.slab_ranges(occ_z, 3)
    peak = np.array(gold["peak_full"])[:, None]
    try:
        for md in (built.MODE_DIRECT, built.MODE_SEPARABLE):
            total, n = 0.0, 0
            for z0, z1 in slabs:
                m.set_slab(z0, z1)
                n += m.num_voxels
                total = total + m.simulate(heart2x["k"], heart2x["leads"], "3D4", 100.0, 1.0, 16.0, mode=md)[0]
            assert n == gold["occupied"]
            assert (np.abs(total - np.array(gold["ecg16"])) / peak).max() < 1e-5
    finally:
        m.set_slab(0, heart2x["layers"].shape[0])
