"""tests/synth.py -- small synthetic voxel hearts and parameter sets for tests and smoke()."""
from __future__ import annotations

import numpy as np

START_FLAG = 0x1000


def small_heart(shape=(20, 24, 22), n_layers=6, seed=0, hole=True):
    """Thick ellipsoidal shell with `n_layers` concentric layers, one start voxel on the inner
    surface, a conduction matrix shaped like testRun/conduction_24.matrix (0.166667 within a
    layer, |i-j| across layers) and two leads outside the grid."""
    rng = np.random.default_rng(seed)
    Z, Y, X = shape
    z, y, x = np.meshgrid(np.arange(Z), np.arange(Y), np.arange(X), indexing="ij")
    c = np.array([(Z - 1) / 2, (Y - 1) / 2, (X - 1) / 2])
    r = np.sqrt(((z - c[0]) / (Z / 2)) ** 2 + ((y - c[1]) / (Y / 2)) ** 2 + ((x - c[2]) / (X / 2)) ** 2)
    lo, hi = 0.35, 0.95
    layers = np.zeros(shape, dtype=np.uint16)
    inside = (r >= lo) & (r < hi)
    layers[inside] = (1 + np.floor((r[inside] - lo) / (hi - lo) * n_layers)).astype(np.uint16)
    layers = np.minimum(layers, n_layers).astype(np.uint16)
    if hole:  # open the shell at the top like a ventricle base, and knock out a few random voxels
        layers[: Z // 5] = 0
        kill = rng.random(shape) < 0.01
        layers[kill] = 0
    cand = np.argwhere(layers == 1)
    s = cand[len(cand) // 2]
    layers[tuple(s)] |= START_FLAG
    n = n_layers + 2
    transfer = np.full((n, n), -1.0)
    for i in range(1, n):
        for j in range(1, n):
            transfer[i, j] = 0.166667 if i == j else float(abs(i - j))
    leads = np.array([[-0.9 * Z, 1.3 * Y, 0.8 * X], [1.6 * Z, -0.7 * Y, 2.1 * X]], dtype=np.float64)
    return layers, transfer, leads


def layer_params(n_layers, seed=0, batch=None):
    """WohlfartPlus coefficient sets in the range the reference's ini allows (simulator.ini
    [wohlfart ap] k min / k max), smoothly varying over layers.  Shape [n_layers, 9] or
    [batch, n_layers, 9]."""
    rng = np.random.default_rng(seed)
    nb = batch or 1
    out = np.zeros((nb, n_layers, 9))
    for b in range(nb):
        k5 = rng.uniform(3e-4, 1e-3, 2)
        k6 = rng.uniform(0.01, 0.1, 2)
        k7 = rng.uniform(0.01, 0.1, 2)
        k8 = rng.uniform(200, 400, 2)
        for l in range(n_layers):
            f = l / max(n_layers - 1, 1)
            out[b, l] = [0.0, 2.5, 100.0, 0.9, 0.1,
                         k5[0] * (1 - f) + k5[1] * f, k6[0] * (1 - f) + k6[1] * f,
                         k7[0] * (1 - f) + k7[1] * f, k8[0] * (1 - f) + k8[1] * f]
    return out if batch else out[0]


def serpentine(n_turns=10, height=12):
    """A one-voxel-thick path that runs up and down in z while advancing in x (square wave): the excitation has to
    cross any z = const cut once per turn, so a z-slab sharded automaton needs about n_turns exchange rounds.
    Two layers alternate along the path.  Returns (layers, transfer)."""
    Z, Y, X = height + 4, 3, 4 * n_turns + 2
    layers = np.zeros((Z, Y, X), dtype=np.uint16)
    lo, hi = 2, height + 1
    for i in range(n_turns):
        z_run = lo if i % 2 == 0 else hi
        layers[z_run, 1, 4 * i:4 * i + 4] = 1 + i % 2
        layers[lo:hi + 1, 1, 4 * i + 3] = 1 + i % 2          # the connector to the other level
    layers[lo, 1, 0] |= START_FLAG
    transfer = np.full((4, 4), -1.0)
    transfer[1:3, 1:3] = [[0.166667, 1.0], [2.0, 0.25]]
    return layers, transfer
