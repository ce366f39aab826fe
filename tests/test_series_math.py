"""The harmonic series that replaces the 8-corner stencil sum of interior voxels in the SEPARABLE path (csrc/ecg.cu,
corner_series2; derivation: tools/derive_corner_series.py) against the sum itself in f64, and the
constants the kernel carries against the ones this test pins.  CPU only: the GPU side of the same statement is
tests/test_gpu_parity.py::test_separable_series_vs_corner_sum / test_separable_series_lead_distance."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

A4 = (112.0, -560.0)                         # 112 - 560 e2
A6 = (-288.0, 3024.0, -33264.0)              # -288 + 3024 e2 - 33264 e3
A8 = (-792.0, 14256.0, -51480.0, 41184.0)    # -792 + 14256 e2 - 51480 e2^2 + 41184 e3


def corner_sum(q):
    """sum over d in {+-1}^3 of d.(q + d) / |q + d|^3, the stencil coefficient of an interior voxel (simulator.cpp:505-527
    re-associated per voxel), q = lead - voxel"""
    g = np.zeros(len(q))
    for dz in (-1.0, 1.0):
        for dy in (-1.0, 1.0):
            for dx in (-1.0, 1.0):
                d = np.array([dz, dy, dx])
                p = q + d
                g += (p @ d) / np.sum(p * p, axis=1) ** 1.5
    return g


def series(q, dtype=np.float64):
    q = q.astype(dtype)
    sq = q * q
    r2 = (sq[:, 0] + sq[:, 1]) + sq[:, 2]
    y = (1.0 / np.sqrt(r2.astype(np.float64))).astype(dtype)
    u = y * y
    xz, xy, xx = sq[:, 0] * u, sq[:, 1] * u, sq[:, 2] * u
    zy = xz * xy
    e2 = (xz + xy) * xx + zy
    e3 = zy * xx
    t = dtype
    a4 = t(A4[1]) * e2 + t(A4[0])
    a6 = t(A6[1]) * e2 + (t(A6[2]) * e3 + t(A6[0]))
    a8 = (t(A8[2]) * e2 + t(A8[1])) * e2 + (t(A8[3]) * e3 + t(A8[0]))
    return ((u * u) * y) * ((a8 * u + a6) * u + a4)


def directions(n, seed):
    v = np.random.default_rng(seed).normal(size=(n, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def test_series_equals_corner_sum_beyond_32_voxels():
    """truncation error of the three-term series < 8e-8 of the local scale at the kernel's validity radius (|q| >= 32 voxels,
    kSeriesMinR2 = 1024) and falling like |q|^-6; the f64 corner sum itself is only good to ~1e-16 |q|^3, hence the floor"""
    n = directions(4000, 1)
    for radius, tol in ((32.0, 8e-8), (64.0, 2e-9), (150.0, 1e-9)):
        q = n * radius
        want, got = corner_sum(q), series(q)
        scale = np.abs(want).max()
        assert np.abs(got - want).max() / scale < tol, (radius, np.abs(got - want).max() / scale)


def test_series_in_fp32_has_no_cancellation():
    """evaluated in fp32 like the kernel does, the series stays within ~1e-6 of the f64 value; the direct sum in fp32 is
    off by 1e-4 .. 1e-2 at these distances (8 terms of size |q|^-2 cancelling down to |q|^-5)"""
    q = directions(4000, 2) * np.random.default_rng(3).uniform(60.0, 300.0, size=(4000, 1))
    want = corner_sum(q)
    got = series(q, np.float32).astype(np.float64)
    assert np.abs(got - want).max() / np.abs(want).max() < 5e-6
    q32 = q.astype(np.float32)
    g32 = np.zeros(len(q), np.float32)
    for dz in (-1.0, 1.0):
        for dy in (-1.0, 1.0):
            for dx in (-1.0, 1.0):
                d = np.array([dz, dy, dx], np.float32)
                p = q32 + d
                g32 += (p @ d) / (np.sum(p * p, axis=1, dtype=np.float32) ** np.float32(1.5))
    assert np.abs(g32 - want).max() / np.abs(want).max() > 1e-4     # what the series replaces


def test_kernel_carries_the_pinned_constants():
    src = open(os.path.join(ROOT, "ekgsim_b200", "csrc", "ecg.cu")).read()
    body = src[src.index("corner_series2(f2 qz"):]
    body = body[:body.index("\n}\n")]
    found = sorted(float(m) for m in re.findall(r"c2\((-?[0-9.]+)f\)", body))
    assert found == sorted(A4 + A6 + A8), found
    assert "kSeriesMinR2 = 1024.f" in src
