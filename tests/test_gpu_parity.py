"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden
vectors produced by the compiled reference.  Tolerances are BASELINE.json's: activation times
bit-exact, ECG traces within 1e-5 of the peak lead amplitude."""
import hashlib
import json
import os

import numpy as np
import pytest

import synth
from oracle import oracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ECG_TOL = 1e-5  # relative to peak lead amplitude (north_star)


def rel_err(ecg, ref):
    peak = np.abs(ref).max(axis=-1, keepdims=True)
    return float((np.abs(ecg - ref) / peak).max())


@pytest.fixture(scope="module")
def gpu_model24(built, model24):
    m = built.Model(model24["layers"], model24["transfer"], device=0)
    m._layers = model24["layers"]
    yield m
    m.close()


def gpu_model24_layers(m):
    return m._layers


def test_activation_model24_bit_exact(gpu_model24, model24_delay):
    delay, sweeps = gpu_model24.activation()
    fp = json.load(open(os.path.join(GOLDEN, "golden_activation.json")))
    assert hashlib.sha256(delay.tobytes()).hexdigest() == fp["sha256_f64_raster"]
    assert delay.tobytes() == model24_delay.tobytes()
    assert sweeps > 1
    K, idx = gpu_model24.ap_classes()
    assert K == fp["classes"]
    Ko, idxo = oracle.ap_classes(gpu_model24_layers(gpu_model24), model24_delay, 24)
    assert Ko == K and (idxo == idx).all()  # same first-seen raster numbering as setApIndices
    print("automaton: %d sweeps, %.3f ms on device" % (sweeps, gpu_model24.activation_ms))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_activation_small_bit_exact(built, seed):
    layers, transfer, _ = synth.small_heart(seed=seed, shape=(17 + seed, 21, 19 + 2 * seed), n_layers=4 + seed)
    m = built.Model(layers, transfer)
    delay, _ = m.activation()
    assert delay.tobytes() == oracle.activation(layers, transfer).tobytes()
    m.close()


@pytest.mark.parametrize("n_layers", [30, 40, 60])
def test_activation_many_layers(built, n_layers):
    """More layers than the reference's 24: up to 36 the (layers + 1)^2 x 3 weight table sits in shared memory, 40 layers read
    it from global memory (the kernel variant without the shared table), 60 are beyond the 16-bit row offsets of the
    frontier kernel and take the plain sweeps by themselves.  Same bits as the oracle in every case."""
    layers, transfer, _ = synth.small_heart(seed=n_layers, shape=(40, 44, 42), n_layers=n_layers, hole=True)
    assert int((layers & 0x0FFF).max()) == n_layers
    ref = oracle.activation(layers, transfer)
    m = built.Model(layers, transfer)
    delay, _ = m.activation()
    assert delay.tobytes() == ref.tobytes()
    m.close()


@pytest.mark.parametrize("seed", [11, 12, 13, 14])
def test_activation_random_conduction_and_many_starts(built, seed):
    """Asymmetric random conduction matrices (T[exciting][excited] != T[excited][exciting], three orders of
    magnitude apart), several start voxels, odd grid extents that do not fill whole 4^3 bricks, both automaton
    kernels: the least fixed point of the relaxation must come out with the reference's bits whatever the visiting order."""
    rng = np.random.default_rng(seed)
    shape = (int(rng.integers(9, 30)), int(rng.integers(9, 34)), int(rng.integers(9, 31)))
    nl = int(rng.integers(2, 9))
    layers, transfer, _ = synth.small_heart(seed=seed, shape=shape, n_layers=nl)
    n = transfer.shape[0]
    transfer[1:, 1:] = 10.0 ** rng.uniform(-2, 1, size=(n - 1, n - 1))
    occ = np.argwhere((layers & 0x0FFF) > 0)
    for idx in rng.choice(len(occ), size=int(rng.integers(1, 6)), replace=False):
        layers[tuple(occ[idx])] |= 0x1000
    ref = oracle.activation(layers, transfer)
    m = built.Model(layers, transfer)
    delay, _ = m.activation()
    assert delay.tobytes() == ref.tobytes()
    m.close()
    os.environ["EKGSIM_B200_AUTOMATON"] = "sweep"
    try:
        m = built.Model(layers, transfer)
        delay, _ = m.activation()
        assert delay.tobytes() == ref.tobytes()
        m.close()
    finally:
        del os.environ["EKGSIM_B200_AUTOMATON"]


def test_activation_2d_and_unreachable(built):
    layers, transfer, _ = synth.small_heart(seed=4, shape=(1, 40, 36), n_layers=5, hole=False)
    layers[0, :, 18] = 0  # cut the ring in two places -> still connected around; then isolate an island
    layers[0, 2:4, 2:4] = 3
    m = built.Model(layers, transfer)
    delay, _ = m.activation()
    ref = oracle.activation(layers, transfer)
    assert delay.tobytes() == ref.tobytes()
    assert (ref[0, 2:4, 2:4] == 0).all()  # island never reached -> delay stays 0 (simulator.cpp:219)
    m.close()


def test_activation_map_stays_on_device(built):
    """The raster host copy of the activation map is optional (ekg_model_activation with delay_out = NULL, what the
    facade passes) and made lazily: the ECG entry points work without it, ekg_model_get_activation /
    ekg_model_ap_classes fetch it on demand -- here through more than one 32 MB download chunk."""
    small, transfer, leads = synth.small_heart(seed=5, shape=(23, 27, 25))
    layers = np.zeros((172, 160, 161), dtype=np.uint16)          # 4.43 M voxels = 35 MB of doubles: two chunks
    layers[140:163, 100:127, 120:145] = small
    leads = leads + np.array([140.0, 100.0, 120.0])
    m = built.Model(layers, transfer)
    none, visits = m.activation(download=False)
    assert none is None and visits > 0
    ref = oracle.activation(layers, transfer)
    layer_k = synth.layer_params(int((layers & 0xFFF).max()), seed=2)
    want = oracle.run_direct(layers, ref, layer_k, leads, "3D4", 0.0, 1.0, 64.0)
    for mode in (1, 2, 3):
        assert rel_err(m.simulate(layer_k, leads, "3D4", 0.0, 1.0, 64.0, mode=mode)[0], want) < ECG_TOL
    got = m.get_activation()
    assert got.tobytes() == ref.tobytes()
    K, idx = m.ap_classes()
    Ko, idxo = oracle.ap_classes(layers, ref, int((layers & 0xFFF).max()))
    assert K == Ko and (idx == idxo).all()
    delay, _ = m.activation()                                    # and the eager form still returns the same map
    assert delay.tobytes() == ref.tobytes()
    m.close()


def test_set_activation_range_and_copy(built):
    """ekg_model_set_activation (loadExcitationSequence, simulator.cpp:288-367): the range of the loaded times -- also
    negative ones -- is found on the device (centre of the hoisted exponentials, saturation time of the SEPARABLE
    path), and ekg_model_get_activation returns the caller's array as given."""
    layers, transfer, leads = synth.small_heart(seed=6)
    m = built.Model(layers, transfer)
    base = oracle.activation(layers, transfer)
    shifted = np.where((layers & 0xFFF) > 0, base - 7.5, -3.0)    # negative times; junk in the empty voxels
    m.set_activation(shifted)
    assert m.get_activation().tobytes() == shifted.tobytes()
    layer_k = synth.layer_params(int((layers & 0xFFF).max()), seed=9)
    want = oracle.run_direct(layers, shifted, layer_k, leads, "3D4", -5.0, 1.0, 80.0)
    for mode in (1, 2, 3):
        assert rel_err(m.simulate(layer_k, leads, "3D4", -5.0, 1.0, 80.0, mode=mode)[0], want) < ECG_TOL
    m.close()


def test_activation_errors(built):
    layers, transfer, _ = synth.small_heart(seed=0)
    nostart = layers & 0x0FFF
    m = built.Model(nostart, transfer)
    with pytest.raises(built.EkgError) as e:
        m.activation()
    assert "starting point" in str(e.value)
    m.close()
    with pytest.raises(built.EkgError) as e:
        built.Model(layers, transfer[:3, :3])
    assert "transfer matrix too small" in str(e.value)


@pytest.mark.parametrize("mode", [1, 2, 3])
@pytest.mark.parametrize("nbhd", ["3D4", "3D8"])
def test_ecg_small_vs_oracle(built, mode, nbhd):
    layers, transfer, leads = synth.small_heart(seed=7)
    nl = int((layers & 0xFFF).max())
    k = synth.layer_params(nl, seed=11, batch=3)
    delay = oracle.activation(layers, transfer)
    m = built.Model(layers, transfer)
    m.set_activation(delay)
    for (t0, dt, tot) in [(0.0, 1.0, 128.0), (100.0, 1.0, 400.0), (3.0, 0.5, 50.0)]:
        ecg = m.simulate(k, leads, nbhd, t0, dt, tot, mode=mode)
        for b in range(3):
            ref = oracle.run_direct(layers, delay, k[b], leads, nbhd, t0, dt, tot)
            assert rel_err(ecg[b], ref) < ECG_TOL, (mode, nbhd, t0, b, rel_err(ecg[b], ref))
    m.close()


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_ecg_model24_golden_full(gpu_model24, model24_delay, mode):
    """Full-length model_24 simulations against ECGs dumped from the compiled reference."""
    g = np.load(os.path.join(GOLDEN, "golden_eval_full.npz"))
    gpu_model24.set_activation(model24_delay)
    ecg = gpu_model24.simulate(g["layer_k"], g["leads_zyx"], "3D4", 100.0, 1.0, 400.0, mode=mode)
    for i in range(ecg.shape[0]):
        e = rel_err(ecg[i], g["ecg"][i])
        print("golden %s mode %d: max err %.3g of peak" % (g["name"][i], mode, e))
        assert e < ECG_TOL


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_ecg_model24_golden_single(gpu_model24, model24_delay, mode):
    """BASELINE config 1 as the facade issues it: ONE simulation per call (B = 1: the time-loop kernels then split the
    segment's voxels into 16 slices so that 16 x 400 (slice, sample) pairs fill 25 whole tiles), each of the four
    full-length reference ECGs on its own."""
    g = np.load(os.path.join(GOLDEN, "golden_eval_full.npz"))
    gpu_model24.set_activation(model24_delay)
    for i in range(g["ecg"].shape[0]):
        ecg = gpu_model24.simulate(g["layer_k"][i], g["leads_zyx"][i], "3D4", 100.0, 1.0, 400.0, mode=mode)
        assert ecg.shape == (1, 2, 400) and np.isfinite(ecg).all()
        e = rel_err(ecg[0], g["ecg"][i])
        print("golden %s alone, mode %d: max err %.3g of peak" % (g["name"][i], mode, e))
        assert e < ECG_TOL


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_ecg_model24_len16_batch(gpu_model24, model24_delay, mode):
    g = np.load(os.path.join(GOLDEN, "golden_len16.npz"))
    gpu_model24.set_activation(model24_delay)
    ecg = gpu_model24.simulate(g["layer_k"], g["leads_zyx"], "3D4", 100.0, 1.0, 16.0, mode=mode)
    # tolerance is relative to the peak amplitude of the lead (north_star), i.e. of the full
    # 400-sample trace, not of this 16-sample window (peak_full: make_fixtures / oracle)
    errs = [float((np.abs(ecg[i] - g["ecg"][i]) / g["peak_full"][i][:, None]).max()) for i in range(ecg.shape[0])]
    print("len16 batch of %d, mode %d: worst %.3g of peak" % (len(errs), mode, max(errs)))
    assert max(errs) < ECG_TOL


def test_ecg_slabs_sum_to_whole(built):
    layers, transfer, leads = synth.small_heart(seed=9)
    nl = int((layers & 0xFFF).max())
    k = synth.layer_params(nl, seed=2)
    delay = oracle.activation(layers, transfer)
    m = built.Model(layers, transfer)
    m.set_activation(delay)
    whole = m.simulate(k, leads, "3D4", 0.0, 1.0, 64.0, mode=1)
    Z = layers.shape[0]
    parts = np.zeros_like(whole)
    nvox = 0
    for z0, z1 in [(0, Z // 3), (Z // 3, Z // 2), (Z // 2, Z)]:
        m.set_slab(z0, z1)
        nvox += m.num_voxels
        parts += m.simulate(k, leads, "3D4", 0.0, 1.0, 64.0, mode=1)
    assert nvox == int(((layers & 0xFFF) > 0).sum())
    # same sum, different fp32 grouping of the per-segment partials -> equal to fp32 rounding
    assert rel_err(parts, whole) < 2e-6
    m.close()


def test_determinism(gpu_model24, model24_delay):
    g = np.load(os.path.join(GOLDEN, "golden_len16.npz"))
    gpu_model24.set_activation(model24_delay)
    a = gpu_model24.simulate(g["layer_k"][:4], g["leads_zyx"][:4], "3D4", 100.0, 1.0, 16.0, mode=1)
    b = gpu_model24.simulate(g["layer_k"][:4], g["leads_zyx"][:4], "3D4", 100.0, 1.0, 16.0, mode=1)
    assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("n_leads", [1, 3, 5])
def test_ecg_lead_counts(built, n_leads):
    """Lead passes: 1..4 leads per launch, more than 4 in several passes."""
    layers, transfer, leads2 = synth.small_heart(seed=12)
    rng = np.random.default_rng(n_leads)
    leads = np.concatenate([leads2, rng.uniform(-60, 90, size=(4, 3))])[:n_leads]
    nl = int((layers & 0xFFF).max())
    k = synth.layer_params(nl, seed=1, batch=2)
    delay = oracle.activation(layers, transfer)
    m = built.Model(layers, transfer)
    m.set_activation(delay)
    for mode in (1, 2, 3):
        ecg = m.simulate(k, leads, "3D4", 0.0, 1.0, 80.0, mode=mode)
        assert ecg.shape == (2, n_leads, 80)
        for b in range(2):
            ref = oracle.run_direct(layers, delay, k[b], leads, "3D4", 0.0, 1.0, 80.0)
            assert rel_err(ecg[b], ref) < ECG_TOL, (mode, b, rel_err(ecg[b], ref))
    m.close()


@pytest.mark.parametrize("nbhd", ["2D4", "2D8"])
def test_ecg_2d_model(built, nbhd):
    layers, transfer, _ = synth.small_heart(seed=4, shape=(1, 40, 36), n_layers=5, hole=False)
    leads = np.array([[30.0, -20.0, 50.0], [-25.0, 70.0, 10.0]])
    k = synth.layer_params(5, seed=8)
    delay = oracle.activation(layers, transfer)
    m = built.Model(layers, transfer)
    m.set_activation(delay)
    for mode in (1, 2, 3):
        ecg = m.simulate(k, leads, nbhd, 0.0, 1.0, 60.0, mode=mode)[0]
        ref = oracle.run_direct(layers, delay, k, leads, nbhd, 0.0, 1.0, 60.0)
        assert rel_err(ecg, ref) < ECG_TOL, (mode, nbhd, rel_err(ecg, ref))
    m.close()


def test_builtin_test_shape_on_the_device(built):
    """The reference's built-in 2-D test ring (InputLoader::generateTestShape, what loadShape falls back to without a shape
    file; the facade's copy is pinned byte for byte in tests/test_host_glue.py) through the CUDA path: activation times
    bit-exact, ECG in both 2-D stencils within tolerance."""
    import hostlib
    layers = hostlib.generate_test_shape()
    n = int((layers & 0x0FFF).max()) + 2
    transfer = np.full((n, n), -1.0)
    for i in range(1, n):
        for j in range(1, n):
            transfer[i, j] = 0.166667 if i == j else float(abs(i - j))
    m = built.Model(layers, transfer)
    delay, _ = m.activation()
    ref_delay = oracle.activation(layers, transfer)
    assert delay.tobytes() == ref_delay.tobytes()
    leads = np.array([[60.0, -40.0, 30.0], [-35.0, 200.0, 160.0]])
    k = synth.layer_params(n - 2, seed=12)
    for nbhd in ("2D4", "2D8"):
        ref = oracle.run_direct(layers, ref_delay, k, leads, nbhd, 0.0, 1.0, 80.0)
        for mode in (1, 2, 3):
            ecg = m.simulate(k, leads, nbhd, 0.0, 1.0, 80.0, mode=mode)[0]
            assert rel_err(ecg, ref) < ECG_TOL, (mode, nbhd, rel_err(ecg, ref))
    m.close()


@pytest.mark.parametrize("B,t0,dt,total", [(10, 0.0, 1.0, 7.0), (37, 5.0, 0.25, 3.0), (3, 0.0, 2.0, 1500.0), (1, 20.0, 1.0, 1.0)])
def test_ecg_ragged_pair_tiles(built, B, t0, dt, total):
    """Few samples per vector (one 256-pair tile spans many vectors and is cut at 4), more samples
    than one tile, a single sample, fractional (exactly representable) time steps."""
    layers, transfer, leads = synth.small_heart(seed=21, shape=(14, 15, 16), n_layers=4)
    nl = int((layers & 0xFFF).max())
    k = synth.layer_params(nl, seed=B, batch=B)
    if B == 1:
        k = k.reshape(1, nl, 9)
    delay = oracle.activation(layers, transfer)
    m = built.Model(layers, transfer)
    m.set_activation(delay)
    lb = np.stack([leads + i for i in range(B)])  # every vector has its own lead positions
    full = [oracle.run_direct(layers, delay, k[b], lb[b], "3D4", 0.0, 1.0, 300.0) for b in range(min(B, 4))]
    for mode in (1, 2, 3):
        ecg = m.simulate(k, lb, "3D4", t0, dt, total, mode=mode)
        for b in range(min(B, 4)):
            ref = oracle.run_direct(layers, delay, k[b], lb[b], "3D4", t0, dt, total)
            peak = np.abs(full[b]).max(axis=1, keepdims=True)   # peak of the whole trace, not of a short window
            assert (np.abs(ecg[b] - ref) / peak).max() < ECG_TOL, (mode, b)
    m.close()


def test_simulate_argument_errors(built):
    layers, transfer, leads = synth.small_heart(seed=1)
    nl = int((layers & 0xFFF).max())
    k = synth.layer_params(nl, seed=1)
    m = built.Model(layers, transfer)
    with pytest.raises(built.EkgError) as e:   # no activation map yet
        m.simulate(k, leads, "3D4", 0.0, 1.0, 10.0)
    assert e.value.code == -4
    m.activation()
    with pytest.raises(built.EkgError):
        m.simulate(k, leads, 9, 0.0, 1.0, 10.0)      # unknown neighbourhood id
    with pytest.raises(built.EkgError):
        m.simulate(k, leads, "3D4", 0.0, -1.0, 10.0)  # bad step
    with pytest.raises(built.EkgError):
        m.set_slab(5, 2)
    m.close()


def test_two_gpus_slabs_and_individuals(built, model24, model24_delay):
    """Needs 2 GPUs (skipped otherwise): z-slab shards of model_24 on two devices add up to the
    single-device ECG, and individuals split over two devices reproduce the single-device batch."""
    if built.lib().ekg_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from ekgsim_b200 import dist as ekdist
    g = np.load(os.path.join(GOLDEN, "golden_len16.npz"))
    k, leads = g["layer_k"][:6], g["leads_zyx"][:6]
    m0 = built.Model(model24["layers"], model24["transfer"], device=0)
    m1 = built.Model(model24["layers"], model24["transfer"], device=1)
    d0, _ = m0.activation()
    d1, _ = m1.activation()
    assert d0.tobytes() == d1.tobytes() == model24_delay.tobytes()
    whole = m0.simulate(k, leads, "3D4", 100.0, 1.0, 16.0, mode=1)
    # individuals: 3 + 3
    split = np.concatenate([m0.simulate(k[:3], leads[:3], "3D4", 100.0, 1.0, 16.0, mode=1),
                            m1.simulate(k[3:], leads[3:], "3D4", 100.0, 1.0, 16.0, mode=1)])
    assert (np.abs(split - whole) / g["peak_full"][:6, :, None]).max() < 2e-6
    # z-slabs balanced by occupied voxels
    occ = ((model24["layers"] & 0x0FFF) > 0).sum(axis=(1, 2))
    (a0, a1), (b0, b1) = ekdist.slab_ranges(occ, 2)
    m0.set_slab(a0, a1)
    m1.set_slab(b0, b1)
    assert m0.num_voxels + m1.num_voxels == 555868 and abs(m0.num_voxels - m1.num_voxels) < 0.05 * 555868
    parts = m0.simulate(k, leads, "3D4", 100.0, 1.0, 16.0, mode=1) + m1.simulate(k, leads, "3D4", 100.0, 1.0, 16.0, mode=1)
    assert (np.abs(parts - whole) / g["peak_full"][:6, :, None]).max() < 2e-6
    assert (np.abs(parts - g["ecg"][:6]) / g["peak_full"][:6, :, None]).max() < ECG_TOL
    # the automaton over the same two slabs, peer-linked ACROSS the devices (peer access + native atomics over NVLink; with
    # both queue disciplines): one kernel per device, no host round, the reference's bits on both
    for queue in ("fifo", "timed"):
        os.environ["EKGSIM_B200_AUTOMATON_QUEUE"] = queue
        try:
            infos = [m0.activation_link_info(), m1.activation_link_info()]
            slabs = [(a0, a1), (b0, b1)]
            m0.activation_link(0, infos, slabs)
            m1.activation_link(1, infos, slabs)
            for m in (m0, m1):
                m.activation_begin()
            for m in (m0, m1):
                m.activation_linked_launch()
            res = [m.activation_linked_wait() for m in (m0, m1)]
            for m in (m0, m1):
                m.activation_linked_gather()
            for m in (m0, m1):
                assert m.activation_end().tobytes() == model24_delay.tobytes(), queue
            assert res[0][1][0] == res[1][1][1] > 0 and res[1][1][0] == res[0][1][1] > 0, res
            for m in (m0, m1):
                m.activation_unlink()
        finally:
            os.environ.pop("EKGSIM_B200_AUTOMATON_QUEUE", None)
    m0.close()
    m1.close()


def test_ecg_large_batch_is_cut_into_sub_batches(built):
    """B * L * T large enough that the library splits the batch (partial-sum scratch cap): vectors of
    the later sub-batches must come out exactly like the same vectors evaluated alone."""
    layers, transfer, leads = synth.small_heart(seed=33, shape=(10, 11, 12), n_layers=3)
    nl = int((layers & 0xFFF).max())
    B = 260
    k = synth.layer_params(nl, seed=77, batch=B)
    delay = oracle.activation(layers, transfer)
    m = built.Model(layers, transfer)
    m.set_activation(delay)
    ecg = m.simulate(k, leads, "3D4", 0.0, 0.25, 800.0, mode=1)      # T = 3200 -> sub-batches of ~131 vectors
    assert ecg.shape == (B, 2, 3200)
    for b in (0, 130, 131, 259):
        ref = oracle.run_factored(layers, delay, k[b], leads, "3D4", 0.0, 0.25, 800.0)
        assert rel_err(ecg[b], ref) < ECG_TOL, b
    m.close()


def test_ecg_model24_batch_sample_vs_oracle(gpu_model24, model24, model24_delay):
    """A spread of the 256-vector batch at full length against the (reference-pinned) oracle."""
    g = np.load(os.path.join(GOLDEN, "golden_glue256.npz"))
    pick = [5, 40, 77, 101, 150, 199, 230, 255]
    gpu_model24.set_activation(model24_delay)
    for mode in (1, 2, 3):
        ecg = gpu_model24.simulate(g["layer_k"], g["leads_zyx"], "3D4", 100.0, 1.0, 400.0, mode=mode)
        worst = 0.0
        for b in pick:
            ref = oracle.run_factored(model24["layers"], model24_delay, g["layer_k"][b], g["leads_zyx"][b], "3D4", 100.0, 1.0, 400.0)
            worst = max(worst, rel_err(ecg[b], ref))
        print("batch sample mode %d: worst %.3g of peak" % (mode, worst))
        assert worst < ECG_TOL


def test_criteria_on_device(gpu_model24, model24, model24_delay):
    """ekg_simulate_criteria: the four curve comparisons of calculateFitness (vectorMath.h) computed on
    the device agree with a numpy restatement applied to the ECGs of the same call, and comparison 2
    reproduces the criteria the reference printed (1e-4)."""
    g = np.load(os.path.join(GOLDEN, "golden_eval_full.npz"))
    sel = [i for i, n in enumerate(g["name"]) if n != "v6full"]
    tv = model24["target_v5"]
    targets = np.stack([tv[:, c] / (tv[:, c].max() - tv[:, c].min()) for c in (1, 2)])      # loadTargets, sim.cpp:1016-1019
    offsets = np.array([1.0 - tv[:, c].min() / (tv[:, c].max() - tv[:, c].min()) for c in (1, 2)])
    gpu_model24.set_activation(model24_delay)
    k, leads = g["layer_k"][sel], g["leads_zyx"][sel]
    for cmp_mode in (1, 2, 3, 4):
        crit, ecg = gpu_model24.simulate_criteria(k, leads, targets, comparison=cmp_mode, target_offsets=offsets, mode=1, want_ecg=True)
        for b in range(len(sel)):
            for l in range(2):
                a, t = ecg[b, l], targets[l][:400]
                if cmp_mode == 1:
                    want = np.sqrt(np.mean((a - t) ** 2))
                elif cmp_mode == 2:
                    want = 1.0 - np.mean((a - a.mean()) * (t - t.mean())) / (a.std() * t.std())
                elif cmp_mode == 4:
                    want = 1.0 - a.dot(t) / (np.sqrt(a.dot(a)) * np.sqrt(t.dot(t)))
                else:
                    d = (a / (a.max() - a.min()) + offsets[l]) / (t + offsets[l])
                    want = np.sqrt(d.var())
                assert abs(crit[b, l] - want) < 1e-11 * max(1.0, abs(want)), (cmp_mode, b, l, crit[b, l], want)
        if cmp_mode == 2:
            assert np.abs(crit - g["criteria"][sel]).max() < 1e-4
    only = gpu_model24.simulate_criteria(k, leads, targets, comparison=2, mode=2)     # no ECG download at all
    assert np.abs(only - g["criteria"][sel]).max() < 1e-4


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_ecg_wide_parameter_ranges(built, seed):
    """Coefficients far outside the testRun ranges (slow depolarisation sigmoid that never saturates,
    large k4, non-zero k0, negative start time so that many samples lie before the activation time):
    both kernels must still match the f64 oracle."""
    rng = np.random.default_rng(100 + seed)
    layers, transfer, leads = synth.small_heart(seed=40 + seed, shape=(13, 14, 15), n_layers=5)
    nl = int((layers & 0xFFF).max())
    B = 6
    k = np.zeros((B, nl, 9))
    for b in range(B):
        for l in range(nl):
            k[b, l] = [rng.uniform(-90, 10), rng.uniform(0.05, 3.5), rng.uniform(50, 150), rng.uniform(0.5, 0.99), rng.uniform(0.01, 0.5),
                       rng.uniform(1e-4, 5e-3), rng.uniform(0.005, 0.3), rng.uniform(0.005, 0.3), rng.uniform(40, 500)]
    delay = oracle.activation(layers, transfer)
    m = built.Model(layers, transfer)
    m.set_activation(delay)
    for (t0, dt, tot) in [(-10.0, 1.0, 200.0), (0.0, 2.5, 600.0)]:
        refs = [oracle.run_direct(layers, delay, k[b], leads, "3D4", t0, dt, tot) for b in range(B)]
        for mode in (1, 2, 3):
            ecg = m.simulate(k, leads, "3D4", t0, dt, tot, mode=mode)
            assert np.isfinite(ecg).all()
            worst = max(rel_err(ecg[b], refs[b]) for b in range(B))
            assert worst < ECG_TOL, (seed, mode, t0, worst)
    m.close()


def test_full_size_properties(gpu_model24, model24, model24_delay):
    """Size-independent properties at the BASELINE size (model_24, T = 400), no oracle needed:
    (1) the ECG is linear in the AP amplitude k2 and blind to the offset k0 (sum of the lead-field
    coefficients over the model is zero); (2) z-slabs add up to the whole model; (3) DIRECT and HOISTED
    agree; (4) permuting the batch permutes the result."""
    g = np.load(os.path.join(GOLDEN, "golden_glue256.npz"))
    k, leads = g["layer_k"][:8].copy(), g["leads_zyx"][:8].copy()
    gpu_model24.set_activation(model24_delay)
    base = gpu_model24.simulate(k, leads, "3D4", 100.0, 1.0, 400.0, mode=1)
    peak = np.abs(base).max(axis=2, keepdims=True)
    k2 = k.copy(); k2[:, :, 2] *= 2.0; k2[:, :, 0] = -37.5
    scaled = gpu_model24.simulate(k2, leads, "3D4", 100.0, 1.0, 400.0, mode=1)
    assert (np.abs(scaled - 2.0 * base) / peak).max() < 2e-5          # fp32 evaluation on both sides
    hoisted = gpu_model24.simulate(k, leads, "3D4", 100.0, 1.0, 400.0, mode=2)
    assert (np.abs(hoisted - base) / peak).max() < ECG_TOL
    separable = gpu_model24.simulate(k, leads, "3D4", 100.0, 1.0, 400.0, mode=3)
    assert (np.abs(separable - base) / peak).max() < ECG_TOL
    sep_parts = np.zeros_like(base)
    for z0, z1 in [(0, 62), (62, 124)]:
        gpu_model24.set_slab(z0, z1)
        sep_parts += gpu_model24.simulate(k, leads, "3D4", 100.0, 1.0, 400.0, mode=3)
    gpu_model24.set_slab(0, 124)
    assert (np.abs(sep_parts - separable) / peak).max() < 2e-6
    perm = np.array([3, 0, 7, 1, 6, 2, 5, 4])
    shuffled = gpu_model24.simulate(k[perm], leads[perm], "3D4", 100.0, 1.0, 400.0, mode=1)
    assert (np.abs(shuffled - base[perm]) / peak[perm]).max() < 2e-6
    parts = np.zeros_like(base)
    for z0, z1 in [(0, 40), (40, 41), (41, 90), (90, 124)]:
        gpu_model24.set_slab(z0, z1)
        parts += gpu_model24.simulate(k, leads, "3D4", 100.0, 1.0, 400.0, mode=1)
    gpu_model24.set_slab(0, 124)
    assert (np.abs(parts - base) / peak).max() < 2e-6


def test_separable_split_and_single_vector(gpu_model24, model24, model24_delay):
    """EKG_MODE_SEPARABLE: a run that starts inside the QRS complex is split into time-loop samples (before
    the last voxel's depolarisation sigmoid saturates) and moment samples; the seam must not show.  Also B = 1
    (all threads of a CTA stride over voxels) and odd batch sizes (partly filled vector groups)."""
    g = np.load(os.path.join(GOLDEN, "golden_glue256.npz"))
    gpu_model24.set_activation(model24_delay)
    k, leads = g["layer_k"][:5], g["leads_zyx"][:5]
    # activation ends at 39.02 ms, k1 = 2.5 -> saturation from 45.95 ms on: samples 0..21 loop, 22.. moments
    ecg = gpu_model24.simulate(k, leads, "3D4", 25.0, 1.0, 60.0, mode=3)
    assert gpu_model24.last_kernel_name == "ecg_kernel<HOISTED> + ecg_moment_kernel"
    for b in (0, 4):
        ref = oracle.run_factored(model24["layers"], model24_delay, k[b], leads[b], "3D4", 25.0, 1.0, 60.0)
        assert rel_err(ecg[b], ref) < ECG_TOL, (b, rel_err(ecg[b], ref))
    hoisted = gpu_model24.simulate(k, leads, "3D4", 25.0, 1.0, 60.0, mode=2)
    peak = np.abs(hoisted).max(axis=2, keepdims=True)
    assert (np.abs(ecg - hoisted) / peak).max() < 2e-6
    # entirely before saturation: the whole run goes through the time loop
    early = gpu_model24.simulate(k[:2], leads[:2], "3D4", 0.0, 1.0, 40.0, mode=3)
    assert gpu_model24.last_kernel_name == "ecg_kernel<HOISTED>"
    assert np.abs(early - gpu_model24.simulate(k[:2], leads[:2], "3D4", 0.0, 1.0, 40.0, mode=2)).max() == 0.0
    # entirely after: moments only; one vector alone equals the same vector inside a batch of 5 or 37
    late1 = gpu_model24.simulate(k[3], leads[3], "3D4", 100.0, 1.0, 400.0, mode=3)
    assert gpu_model24.last_kernel_name == "ecg_moment_kernel"
    late5 = gpu_model24.simulate(k, leads, "3D4", 100.0, 1.0, 400.0, mode=3)
    late37 = gpu_model24.simulate(g["layer_k"][:37], g["leads_zyx"][:37], "3D4", 100.0, 1.0, 400.0, mode=3)
    pk = np.abs(late1).max(axis=2, keepdims=True)
    assert (np.abs(late1[0] - late5[3]) / pk[0]).max() < 1e-6      # different voxel-lane grouping of the fp32 sums
    assert (np.abs(late37[:5] - late5) / np.abs(late5).max(axis=2, keepdims=True)).max() < 1e-6
    # a slow depolarisation (k1 = 0.2) pushes the seam far out: 39.02 + 25 / (0.2 log2 e) = 125.7 ms
    slow = k.copy(); slow[:, :, 1] = 0.2
    e3 = gpu_model24.simulate(slow, leads, "3D4", 100.0, 1.0, 100.0, mode=3)
    assert gpu_model24.last_kernel_name == "ecg_kernel<HOISTED> + ecg_moment_kernel"
    ref = oracle.run_factored(model24["layers"], model24_delay, slow[1], leads[1], "3D4", 100.0, 1.0, 100.0)
    assert rel_err(e3[1], ref) < ECG_TOL


def test_separable_series_vs_corner_sum(built, gpu_model24, model24, model24_delay):
    """The moment kernel evaluates the stencil sum of interior voxels (all 8 corners occupied, 86 % of model_24) by its
    harmonic series; EKG_FLAG_CORNER_SUM adds the 8 corner terms instead.  Both must agree with the reference's ECG, the
    series at least as closely as the sum (it has no cancellation), for a batch and for one vector alone."""
    f = np.load(os.path.join(GOLDEN, "golden_eval_full.npz"))      # four full-length runs of the compiled reference
    gpu_model24.set_activation(model24_delay)
    k, leads, ref = f["layer_k"], f["leads_zyx"], f["ecg"]
    ser = gpu_model24.simulate(k, leads, "3D4", 100.0, 1.0, 400.0, mode=3)
    assert gpu_model24.last_kernel_name == "ecg_moment_kernel"
    csum = gpu_model24.simulate(k, leads, "3D4", 100.0, 1.0, 400.0, mode=3 | built.FLAG_CORNER_SUM)
    e_ser, e_sum = rel_err(ser, ref), rel_err(csum, ref)
    print("series %.3g, corner sum %.3g of peak" % (e_ser, e_sum))
    assert e_ser < ECG_TOL and e_sum < ECG_TOL
    assert e_ser < 2e-6
    assert rel_err(ser, csum) < 3e-6
    one = gpu_model24.simulate(k[1], leads[1], "3D4", 100.0, 1.0, 400.0, mode=3)      # VB = 1: lanes stride over voxels
    assert rel_err(one[0], ref[1]) < 2e-6


@pytest.mark.parametrize("scale", [0.6, 1.0, 3.0, 12.0])
@pytest.mark.parametrize("n_leads", [2, 3])
def test_separable_series_lead_distance(built, scale, n_leads):
    """Leads next to the heart (closer than 32 voxels: the series is not used), at a few heart diameters (series for some
    voxels, direct sum for others, mixed inside one packed lead pair) and far away (series everywhere; the remainder of
    the cancelling corner terms is all that is left of an interior voxel) against the f64 oracle."""
    layers, transfer, leads = synth.small_heart(seed=7, shape=(26, 30, 28), n_layers=5, hole=False)
    centre = np.array(layers.shape, dtype=np.float64) / 2
    leads = np.vstack([leads, [[40.0, 45.0, -30.0]]])[:n_leads]
    leads = centre + (leads - centre) * scale
    m = built.Model(layers, transfer)
    delay, _ = m.activation()
    layer_k = synth.layer_params(int((layers & 0xFFF).max()), seed=1, batch=3)
    leads_b = np.stack([leads, leads + 1.25, leads - 0.5])
    # the whole trace: the QRS samples run through the time loop, everything after the last depolarisation through the
    # moments.  The tolerance is relative to the peak lead amplitude of the trace (north_star); on the plateau alone the
    # ECG all but vanishes (a uniform V gives exactly 0), so a window without the QRS complex would measure nothing.
    got = m.simulate(layer_k, leads_b, "3D4", 0.0, 2.0, 400.0, mode=3)
    assert m.last_kernel_name == "ecg_kernel<HOISTED> + ecg_moment_kernel"
    csum = m.simulate(layer_k, leads_b, "3D4", 0.0, 2.0, 400.0, mode=3 | built.FLAG_CORNER_SUM)
    first_moment_sample = int((float(delay.max()) + 25.0 / (2.5 * 1.4426950408889634)) / 2.0) + 2
    for b in range(3):
        want = oracle.run_direct(layers, delay, layer_k[b], leads_b[b], "3D4", 0.0, 2.0, 400.0)
        e_ser, e_sum = rel_err(got[b], want), rel_err(csum[b], want)
        peak = np.abs(want).max(axis=-1, keepdims=True)
        late = float((np.abs(got[b] - want)[:, first_moment_sample:] / peak).max())
        print("scale %g vector %d: series %.3g, corner sum %.3g of peak; moment samples %.3g" % (scale, b, e_ser, e_sum, late))
        assert e_ser < ECG_TOL and e_sum < ECG_TOL, (b, e_ser, e_sum)
        assert late < 2e-6, (b, late)
    m.close()


def run_sharded(ranks, slabs, budget=0):
    """The exchange loop of ekgsim_b200.dist.sharded_activation with the planes handed over directly."""
    world = len(ranks)
    for p in ranks:
        p.begin()
    rounds, visits = 0, 0
    while True:
        rounds += 1
        improved = 0
        for p in ranks:
            v, left = p.relax(budget)
            visits += v
            improved += left
        for r in range(world - 1):
            up = ranks[r].export(slabs[r][1] - 1, slabs[r][1])            # last plane of r -> halo of r + 1
            dn = ranks[r + 1].export(slabs[r + 1][0], slabs[r + 1][0] + 1)  # first plane of r + 1 -> halo of r
            improved += ranks[r + 1].merge(slabs[r][1] - 1, up)
            improved += ranks[r].merge(slabs[r + 1][0], dn)
        if improved == 0:
            break
        assert rounds < 5000
    for s in range(world):
        buf = ranks[s].export(*slabs[s])
        for r in range(world):
            if r != s:
                ranks[r].merge(slabs[s][0], buf)
    return rounds, visits, [p.end() for p in ranks]


def run_linked(models, slabs, gather=True):
    """The peer-linked sharded automaton with all ranks' handles on ONE device (raw pointers instead of CUDA IPC; every
    rank's kernel gets an equal share of the SMs so that all of them are resident): link, begin everywhere, launch
    everywhere, wait everywhere, pull the other slabs, end."""
    import torch
    infos = [m.activation_link_info() for m in models]
    for r, m in enumerate(models):
        m.activation_link(r, infos, slabs)
    for m in models:
        m.activation_begin()
    cap = max(1, torch.cuda.get_device_properties(0).multi_processor_count * 2 // len(models))
    for m in models:
        m.activation_linked_launch(max_ctas=cap)
    res = [m.activation_linked_wait() for m in models]
    if gather:
        for m in models:
            m.activation_linked_gather()
    return res, [m.activation_end() for m in models]


def test_sharded_automaton_many_crossings(built):
    """A serpentine path crosses the slab faces once per turn: the wave has to be handed back and forth ~10 times, through
    faces that cut bricks in the middle as well as faces on brick boundaries."""
    import torch
    from ekgsim_b200 import dist as ekdist
    layers, transfer = synth.serpentine(n_turns=10, height=12)
    ref = oracle.activation(layers, transfer)
    assert (ref[(layers & 0x0FFF) > 0] > 0).all()
    for slabs in ([(0, 8), (8, 16)], [(0, 6), (6, 7), (7, 16)], [(0, 4), (4, 12), (12, 16)]):
        ranks = []
        for z0, z1 in slabs:
            m = built.Model(layers, transfer, device=0)
            m.set_slab(z0, z1)
            ranks.append(ekdist.ModelPlanes(m, torch.device("cuda", 0)))
        rounds, visits, delays = run_sharded(ranks, slabs)
        for d in delays:
            assert d.tobytes() == ref.tobytes(), slabs
        assert rounds >= 10, (slabs, rounds)
        print("serpentine, slabs %s: %d rounds, %d brick visits" % (slabs, rounds, visits))
        # bounded relaxation: at most ~3 brick visits per rank and round, the leftovers are carried over
        rounds_b, visits_b, delays_b = run_sharded(ranks, slabs, budget=3)
        for d in delays_b:
            assert d.tobytes() == ref.tobytes(), slabs
        assert rounds_b > rounds
        # peer-linked: no rounds at all, the slabs hand the wave back and forth from inside their kernels
        res, delays_l = run_linked([p.model for p in ranks], slabs)
        for d in delays_l:
            assert d.tobytes() == ref.tobytes(), slabs
        sent = sum(r[1][0] for r in res)
        assert sent == sum(r[1][1] for r in res) and sent >= 10, res       # every brick queued elsewhere arrived; >= one per crossing
        print("  linked: brick visits %s, bricks queued across the faces %d, cells written across %d" % ([r[0] for r in res], sent, sum(r[1][2] for r in res)))
        for p in ranks:
            p.model.close()


def test_linked_automaton_empty_slab_and_relink(built):
    """A rank with an empty slab takes part idle; linking again with other slabs works on the same handles; ranks whose
    slab is one plane thick send that plane both ways."""
    layers, transfer = synth.serpentine(n_turns=6, height=12)
    ref = oracle.activation(layers, transfer)
    models = [built.Model(layers, transfer, device=0) for _ in range(4)]
    for slabs in ([(0, 8), (8, 8), (8, 9), (9, 16)], [(0, 3), (3, 4), (4, 5), (5, 16)], [(0, 0), (0, 16), (16, 16), (16, 16)]):
        for m, (z0, z1) in zip(models, slabs):
            m.set_slab(z0, z1)
        res, delays = run_linked(models, slabs)
        for d in delays:
            assert d.tobytes() == ref.tobytes(), slabs
        assert sum(r[1][0] for r in res) == sum(r[1][1] for r in res)
    # without the gather a rank is only exact on its own slab
    slabs = [(0, 5), (5, 9), (9, 12), (12, 16)]
    for m, (z0, z1) in zip(models, slabs):
        m.set_slab(z0, z1)
    res, delays = run_linked(models, slabs, gather=False)
    for d, (z0, z1) in zip(delays, slabs):
        assert d[z0:z1].tobytes() == ref[z0:z1].tobytes(), (z0, z1)
    # errors: launch without link / begin, slabs that do not partition
    lone = built.Model(layers, transfer, device=0)
    with pytest.raises(built.EkgError):
        lone.activation_linked_launch()
    with pytest.raises(built.EkgError):
        lone.activation_link(0, [lone.activation_link_info()], [(0, 15)])
    lone.activation_link(0, [lone.activation_link_info()], [(0, 16)])
    with pytest.raises(built.EkgError):
        lone.activation_linked_launch()                                    # no begin
    lone.activation_begin()
    lone.activation_linked_launch()
    lone.activation_linked_wait()
    assert lone.activation_end().tobytes() == ref.tobytes()               # one rank alone: the detector sees itself idle
    lone.activation_unlink()
    lone.close()
    for m in models:
        m.close()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_automaton_entry_points(built, model24, model24_delay, world):
    """ekg_model_activation_begin / _relax / _export / _merge / _end: `world` handles on ONE device play the ranks of a
    z-slab sharded run (planes handed over directly instead of through NCCL; the collective driver itself is covered
    by the gloo test and tools/bench_heart4x.py --sharded-automaton).  Every handle must end with the reference's bits."""
    import torch
    from ekgsim_b200 import dist as ekdist
    occ = ((model24["layers"] & 0x0FFF) > 0).sum(axis=(1, 2))
    slabs = ekdist.slab_ranges(occ, world)
    # the start voxel (z = 62) lies in one slab only; the other ranks have nothing to do until a halo arrives
    slabs = [(0, 41), (41, 124)] if world == 2 else [(0, 30), (30, 61), (61, 124)]
    dev = torch.device("cuda", 0)
    ranks = []
    for z0, z1 in slabs:
        m = built.Model(model24["layers"], model24["transfer"], device=0)
        m.set_slab(z0, z1)
        ranks.append(ekdist.ModelPlanes(m, dev))
    rounds, visits, delays = run_sharded(ranks, slabs)
    assert rounds >= 2
    for delay in delays:
        assert delay.tobytes() == model24_delay.tobytes()
    print("sharded automaton, %d slabs on model_24: %d rounds, %d brick visits in total" % (world, rounds, visits))
    # bounded rounds (what the NCCL driver uses so that the ranks work side by side): same bits, more rounds
    rounds_b, visits_b, delays_b = run_sharded(ranks, slabs, budget=4000)
    assert rounds_b > rounds
    for delay in delays_b:
        assert delay.tobytes() == model24_delay.tobytes()
    print("  bounded to 4000 visits per round: %d rounds, %d brick visits" % (rounds_b, visits_b))
    # peer-linked: one kernel per rank, face planes and brick pushes go through the neighbours' memory
    res, delays_l = run_linked([p.model for p in ranks], slabs)
    for delay in delays_l:
        assert delay.tobytes() == model24_delay.tobytes()
    assert sum(r[1][0] for r in res) == sum(r[1][1] for r in res) > 0
    print("  linked: brick visits %s, remote (queued there, queued here, cells written) %s" % ([r[0] for r in res], [r[1] for r in res]))
    # the handles are usable afterwards like after ekg_model_activation: slab ECGs add up
    g = np.load(os.path.join(GOLDEN, "golden_glue256.npz"))
    parts = sum(p.model.simulate(g["layer_k"][:2], g["leads_zyx"][:2], "3D4", 100.0, 1.0, 50.0, mode=3) for p in ranks)
    whole = ranks[0].model
    whole.set_slab(0, 124)
    ref = whole.simulate(g["layer_k"][:2], g["leads_zyx"][:2], "3D4", 100.0, 1.0, 50.0, mode=3)
    assert rel_err(parts, ref) < 2e-6
    with pytest.raises(built.EkgError):
        whole.activation_relax()                                           # not between begin and end
    with pytest.raises(built.EkgError):
        whole.activation_relax_bounded(100)
    for p in ranks:
        p.model.close()


def test_fast_decay_falls_back_to_direct(gpu_model24, model24, model24_delay):
    """The HOISTED / SEPARABLE factorisation exp(-k (t - at)) = exp(-k (t - t0)) exp(k (at - t0)) is clamped at 2^60 per
    factor; a plateau decay k4 = 4 /ms over the 39 ms activation range would need 2^112.  The library notices (it knows
    the coefficients of a host-buffer call) and runs DIRECT instead of returning clamped numbers."""
    g = np.load(os.path.join(GOLDEN, "golden_glue256.npz"))
    gpu_model24.set_activation(model24_delay)
    k = g["layer_k"][:3].copy()
    k[:, :, 4] = 4.0
    leads = g["leads_zyx"][:3]
    ref = oracle.run_factored(model24["layers"], model24_delay, k[1], leads[1], "3D4", 100.0, 1.0, 80.0)
    for mode in (2, 3, 0):
        ecg = gpu_model24.simulate(k, leads, "3D4", 100.0, 1.0, 80.0, mode=mode)
        assert gpu_model24.last_kernel_name == "ecg_kernel<DIRECT>", mode
        assert rel_err(ecg[1], ref) < ECG_TOL, (mode, rel_err(ecg[1], ref))
    # the ordinary range keeps the fast paths
    gpu_model24.simulate(g["layer_k"][:3], leads, "3D4", 100.0, 1.0, 80.0, mode=2)
    assert gpu_model24.last_kernel_name == "ecg_kernel<HOISTED>"
    gpu_model24.simulate(g["layer_k"][:3], leads, "3D4", 100.0, 1.0, 80.0, mode=0)
    assert gpu_model24.last_kernel_name == "ecg_moment_kernel"


def test_simulate_device_hinted_matches_unhinted(built, gpu_model24, model24_delay):
    """ekg_simulate_device_hinted: the caller passes the batch's smallest k1 and largest decay rate, so the default mode does
    not read them back from the device in the middle of the call.  Same kernels, same results as the unhinted call (the
    device computes the same two numbers in fp32, so the split between time loop and moments may move by one sample:
    tolerance instead of bit equality) and as the reference goldens; bad hints are rejected."""
    import torch
    g = np.load(os.path.join(GOLDEN, "golden_glue256.npz"))
    gf = np.load(os.path.join(GOLDEN, "golden_eval_full.npz"))
    gpu_model24.set_activation(model24_delay)
    dev = torch.device("cuda", 0)
    for k, leads in ((g["layer_k"][:5], g["leads_zyx"][:5]), (gf["layer_k"], gf["leads_zyx"])):
        B = k.shape[0]
        d_k = torch.from_numpy(np.ascontiguousarray(k)).to(dev)
        d_l = torch.from_numpy(np.ascontiguousarray(leads)).to(dev)
        e0 = torch.empty((B, 2, 400), dtype=torch.float64, device=dev)
        e1 = torch.empty_like(e0)
        stream = torch.cuda.current_stream().cuda_stream
        k1_min, decay_max = built.coefficient_hints(k)
        for t_start in (100.0, 0.0):                     # all samples after the QRS complex / the QRS complex inside the run
            gpu_model24.simulate_device(d_k.data_ptr(), d_l.data_ptr(), B, 2, e0.data_ptr(), "3D4", t_start, 1.0, 400.0, mode=0, stream=stream)
            gpu_model24.simulate_device_hinted(d_k.data_ptr(), d_l.data_ptr(), B, 2, e1.data_ptr(), k1_min, decay_max, "3D4", t_start, 1.0, 400.0,
                                               mode=0, stream=stream)
            torch.cuda.synchronize()
            a, b = e0.cpu().numpy(), e1.cpu().numpy()
            assert rel_err(b, a) < 2e-6, (t_start, rel_err(b, a))
    assert rel_err(b, b) == 0.0
    # the last batch at t_start = 100 is the reference's own configuration
    gpu_model24.simulate_device_hinted(d_k.data_ptr(), d_l.data_ptr(), B, 2, e1.data_ptr(), k1_min, decay_max, "3D4", 100.0, 1.0, 400.0, mode=0, stream=stream)
    torch.cuda.synchronize()
    assert rel_err(e1.cpu().numpy(), gf["ecg"]) < ECG_TOL
    for bad in ((0.0, 1.0), (float("nan"), 1.0), (1.0, -1.0), (1.0, float("inf"))):
        with pytest.raises(built.EkgError):
            gpu_model24.simulate_device_hinted(d_k.data_ptr(), d_l.data_ptr(), B, 2, e1.data_ptr(), bad[0], bad[1], "3D4", 100.0, 1.0, 400.0, mode=0, stream=stream)
