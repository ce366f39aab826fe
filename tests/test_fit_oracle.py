"""CPU: the oracle's restatement of the layer-AP construction (oracle/ekg_oracle.c: connectors of
WohlfartInterpolationEvaluator sim.cpp:91-313, steepestDescend nonlinearFit.h:92-168, the layer loop
sim.cpp:751-916) against the coefficients the compiled reference's own glue produced for the 256
seeded vectors (tests/golden/golden_glue256.npz, dumped by oracle/ref_dump)."""
import os

import numpy as np

from oracle import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_fit_bit_identical_to_reference_glue():
    g = np.load(os.path.join(GOLDEN, "golden_glue256.npz"))
    for i in range(0, 256, 3):
        k = oracle.fit_layers(g["layer_k"][i][[0, 14, 23]], 24, mid=14)
        assert k.tobytes() == g["layer_k"][i].tobytes(), i


def test_oracle_apd90_and_border_layers():
    k = np.array([0.0, 2.5, 100.0, 0.9, 0.1, 0.001, 0.1, 0.1, 250.0])   # simulator.ini base AP
    apd = oracle.lib().ekg_oracle_apd90(k.ctypes.data)
    v = lambda t: oracle.wohlfart_plus(k, float(t))
    assert v(int(np.floor(apd))) > 10.0 >= v(int(np.floor(apd)) + 1)    # k0 + 0.1 k2 crossed inside that millisecond
    never = k.copy(); never[5] = 0.0; never[8] = 5000.0                 # does not repolarise within 1000 ms
    assert oracle.lib().ekg_oracle_apd90(never.ctypes.data) == -1.0
    b = np.stack([k, k * [1, 1, 1, 1, 1, 0.7, 0.8, 0.9, 1.1]])
    out = oracle.fit_layers(b, 6)                                        # endo-epi
    assert out[0].tobytes() == b[0].tobytes() and out[5].tobytes() == b[1].tobytes()
    blend = np.array([b[0] * (1 - i / 5.0) + b[1] * (i / 5.0) for i in range(6)])
    assert np.abs(out[:, [0, 1, 2, 4]] - blend[:, [0, 1, 2, 4]]).max() == 0.0     # coefficients with d = 0 keep the blend
    assert np.abs(out[1:5] - blend[1:5]).max() > 0.0                              # the others were moved by the descent
