"""GPU: the device-side layer-AP construction (csrc/fit.cu, ekg_fit_layers / ekg_evaluate) against the
coefficients of the compiled reference's glue (golden_glue256.npz), against the oracle's restatement for
settings the goldens do not cover, and end to end (border APs -> criteria) against the reference-pinned
criteria of all 256 vectors.  f64 on both sides; the only difference is CUDA's exp/log/pow vs glibc's
(<= 2 ulp), which can flip an accept/reject decision of the descent late in a fit: coefficients are
required to agree to 1e-9 relative and are bit-identical for almost every vector."""
import os

import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIT_TOL = 1e-9


@pytest.fixture(scope="module")
def gpu_model24(built, model24, model24_delay):
    m = built.Model(model24["layers"], model24["transfer"], device=0)
    m.set_activation(model24_delay)
    yield m
    m.close()


def rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def test_fit_256_vectors_against_reference_glue(gpu_model24):
    g = np.load(os.path.join(GOLDEN, "golden_glue256.npz"))
    border = g["layer_k"][:, [0, 14, 23]]
    k = gpu_model24.fit_layers(border, mid=14)
    assert gpu_model24.last_launch_count == 2
    same = np.array([k[i].tobytes() == g["layer_k"][i].tobytes() for i in range(256)])
    worst = rel(k, g["layer_k"]).max()
    print("device fit: %d of 256 vectors bit-identical to the reference glue, worst relative difference %.3g" % (same.sum(), worst))
    assert worst < FIT_TOL
    assert same.sum() >= 240
    # a vector alone gives the same bits as inside the batch
    one = gpu_model24.fit_layers(border[17], mid=14)
    assert one[0].tobytes() == k[17].tobytes()


@pytest.mark.parametrize("n_layers,mid", [(6, -1), (9, 3), (5, 1), (3, 1), (2, -1)])
def test_fit_small_models_against_oracle(built, n_layers, mid):
    """endo-epi (2 border APs) and endo-mid-epi on models with other layer counts; wide coefficient draws,
    among them APs that never repolarise (apd90 = -1 -> 700-sample cap) and equal border APs."""
    import synth
    layers, transfer, _ = synth.small_heart(n_layers=n_layers, shape=(12, 14, 13))
    m = built.Model(layers, transfer)
    rng = np.random.default_rng(n_layers * 10 + mid)
    nb = 2 if mid < 0 else 3
    B = 24
    border = np.zeros((B, nb, 9))
    for b in range(B):
        for j in range(nb):
            border[b, j] = [rng.uniform(-90, 0) if b % 3 == 0 else 0.0, rng.uniform(1.5, 3.5), 100.0, rng.uniform(0.85, 0.95),
                            rng.uniform(0.05, 0.2), rng.uniform(3e-4, 1e-3), rng.uniform(0.01, 0.1), rng.uniform(0.01, 0.1),
                            rng.uniform(200, 450)]
    border[1, :, 5] = 0.0; border[1, :, 8] = 5000.0      # never repolarises
    border[2, 1:] = border[2, 0]                         # identical border APs: i1 == i2 everywhere
    got = m.fit_layers(border, mid=mid)
    m.close()
    same = 0
    for b in range(B):
        want = oracle.fit_layers(border[b], n_layers, mid)
        assert rel(got[b], want).max() < FIT_TOL, (b, rel(got[b], want).max())
        same += got[b].tobytes() == want.tobytes()
    assert same >= B - 3, same


def test_fit_other_descent_settings(gpu_model24):
    g = np.load(os.path.join(GOLDEN, "golden_glue256.npz"))
    border = g["layer_k"][:6, [0, 14, 23]]
    # also fit k4 and the plateau height k2 (not k1: beyond 10 ms exp(-k1 t) < 1e-10, its finite-difference
    # gradient is rounding noise whose sign no two libms agree on)
    d9 = (0.0, 0.0, 0.05, 0.001, 0.002, 0.00005, 0.0005, 0.01, 0.2)
    for step, eps, it in [(0.5, 1e-3, 7), (0.25, 1e-1, 100), (1.0, 1e-6, 40)]:
        got = gpu_model24.fit_layers(border, mid=14, d9=d9, step=step, eps=eps, iterations=it)
        for b in range(6):
            want = oracle.fit_layers(border[b], 24, 14, d9=d9, step=step, eps=eps, iterations=it)
            assert rel(got[b], want).max() < FIT_TOL, (step, eps, it, b)
    got = gpu_model24.fit_layers(border, mid=14, iterations=0)       # no descent: the linear blend
    want = oracle.fit_layers(border[0], 24, 14, iterations=0)
    assert got[0].tobytes() == want.tobytes()


def test_fit_argument_errors(gpu_model24, built):
    g = np.load(os.path.join(GOLDEN, "golden_glue256.npz"))
    border = g["layer_k"][:2, [0, 14, 23]]
    for mid in (0, 23, 24, -1):
        with pytest.raises(built.EkgError):
            gpu_model24.fit_layers(border, mid=mid)
    with pytest.raises(built.EkgError):
        gpu_model24.fit_layers(np.zeros((2, 4, 9)), mid=3)


def test_evaluate_256_border_aps_to_criteria(gpu_model24, model24):
    """ekg_evaluate: 27 coefficients + 2 leads per vector in, 2 criteria out, nothing else crosses PCIe."""
    g = np.load(os.path.join(GOLDEN, "golden_glue256.npz"))
    want = np.load(os.path.join(GOLDEN, "golden_criteria256.npz"))
    tv = model24["target_v5"]
    targets = np.stack([tv[:, c] / (tv[:, c].max() - tv[:, c].min()) for c in (1, 2)])   # sim.cpp:1016-1019
    border = g["layer_k"][:, [0, 14, 23]]
    for mode in (1, 2):
        crit, lk, ecg = gpu_model24.evaluate(border, g["leads_zyx"], targets, mid=14, comparison=2, mode=mode, want_layer_k=True,
                                             want_ecg=(mode == 2))
        err = np.abs(crit - want["criteria"]).max()
        print("ekg_evaluate mode %d: worst criteria difference %.3g, launches %d" % (mode, err, gpu_model24.last_launch_count))
        assert err < 1e-4                                    # north_star tolerance
        assert rel(lk, g["layer_k"]).max() < FIT_TOL
    # the ECGs it returns are the ones ekg_simulate gives for the same coefficients
    ref = gpu_model24.simulate(lk, g["leads_zyx"], "3D4", 100.0, 1.0, 400.0, mode=2)
    assert np.abs(ref - ecg).max() == 0.0


def test_device_pointer_entry_points_on_a_side_stream(gpu_model24):
    """ekg_fit_layers_device -> ekg_simulate_device chained on a caller's (non-default) stream with device-resident
    buffers (torch only provides the memory and the stream): same coefficients and ECGs as the host-buffer calls.
    In SEPARABLE mode the library does not know k1 on the host here and reads its minimum back from the device."""
    import torch
    g = np.load(os.path.join(GOLDEN, "golden_glue256.npz"))
    B = 40
    border = np.ascontiguousarray(g["layer_k"][:B, [0, 14, 23]])
    leads = np.ascontiguousarray(g["leads_zyx"][:B])
    dev = torch.device("cuda", 0)
    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        d_border = torch.from_numpy(border).to(dev, non_blocking=False)
        d_leads = torch.from_numpy(leads).to(dev)
        d_k = torch.empty((B, 24, 9), dtype=torch.float64, device=dev)
        d_ecg = torch.empty((3, B, 2, 400), dtype=torch.float64, device=dev)
        gpu_model24.fit_layers_device(d_border.data_ptr(), B, 3, d_k.data_ptr(), mid=14, stream=side.cuda_stream)
        for i, mode in enumerate((1, 2, 3)):
            gpu_model24.simulate_device(d_k.data_ptr(), d_leads.data_ptr(), B, 2, d_ecg[i].data_ptr(), "3D4", 100.0, 1.0, 400.0,
                                        mode=mode, stream=side.cuda_stream)
        side.synchronize()
        k = d_k.cpu().numpy()
        ecg = d_ecg.cpu().numpy()
    assert k.tobytes() == gpu_model24.fit_layers(border, mid=14).tobytes()
    for i, mode in enumerate((1, 2, 3)):
        host = gpu_model24.simulate(k, leads, "3D4", 100.0, 1.0, 400.0, mode=mode)
        assert np.abs(host - ecg[i]).max() == 0.0, mode
    # a run that starts inside the QRS complex: the device-side k1 minimum must put the seam where the host-side one does
    d_e2 = torch.empty((B, 2, 60), dtype=torch.float64, device=dev)
    gpu_model24.simulate_device(d_k.data_ptr(), d_leads.data_ptr(), B, 2, d_e2.data_ptr(), "3D4", 25.0, 1.0, 60.0, mode=3, stream=0)
    torch.cuda.synchronize()
    assert gpu_model24.last_kernel_name == "ecg_kernel<HOISTED> + ecg_moment_kernel"
    assert np.abs(gpu_model24.simulate(k, leads, "3D4", 25.0, 1.0, 60.0, mode=3) - d_e2.cpu().numpy()).max() == 0.0
