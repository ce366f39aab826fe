"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes binding of the plain-C restatement (oracle/ekg_oracle.c) plus helpers that read the
binary dumps written by oracle/ref_dump.cpp (the compiled, unmodified reference).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module;
the product package (ekgsim_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libekg_oracle.so")

NBHD = {"2D4": 0, "2D8": 1, "3D4": 2, "3D8": 3, "cube": 3}
START_FLAG = 0x1000


def build(force: bool = False) -> str:
    """Compile the C restatement (and, when /root/reference exists, oracle/_ref)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "ekg_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        d, i64, p = C.c_double, C.c_int64, C.c_void_p
        L.ekg_oracle_wohlfart_plus.restype = d
        L.ekg_oracle_wohlfart_plus.argtypes = [p, d]
        L.ekg_oracle_ap.restype = d
        L.ekg_oracle_ap.argtypes = [p, d, d]
        L.ekg_oracle_neighbourhood.restype = C.c_int
        L.ekg_oracle_neighbourhood.argtypes = [C.c_int, p]
        L.ekg_oracle_activation.restype = C.c_int
        L.ekg_oracle_activation.argtypes = [p, i64, i64, i64, p, i64, i64, p]
        for f in (L.ekg_oracle_run_direct, L.ekg_oracle_run_factored):
            f.restype = i64
            f.argtypes = [p, p, i64, i64, i64, p, i64, p, i64, C.c_int, d, d, d, p]
        L.ekg_oracle_run_factored_slab.restype = i64
        L.ekg_oracle_run_factored_slab.argtypes = [p, p, i64, i64, i64, p, i64, p, i64, C.c_int, d, d, d, i64, i64, p]
        L.ekg_oracle_ap_classes.restype = i64
        L.ekg_oracle_ap_classes.argtypes = [p, p, i64, i64, p]
        L.ekg_oracle_run_approximation.restype = i64
        L.ekg_oracle_run_approximation.argtypes = [p, i64, d, d, d, d, p]
        L.ekg_oracle_apd90.restype = d
        L.ekg_oracle_apd90.argtypes = [p]
        L.ekg_oracle_fit_layers.restype = C.c_int
        L.ekg_oracle_fit_layers.argtypes = [p, i64, i64, i64, p, d, d, i64, p]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def wohlfart_plus(k, t: float) -> float:
    k = np.ascontiguousarray(k, dtype=np.float64)
    return lib().ekg_oracle_wohlfart_plus(_ptr(k), float(t))


def neighbourhood(nbhd) -> np.ndarray:
    dif = np.zeros((26, 3), dtype=np.int32)
    n = lib().ekg_oracle_neighbourhood(NBHD[nbhd] if isinstance(nbhd, str) else int(nbhd), _ptr(dif))
    if n < 0:
        raise ValueError("unknown neighbourhood")
    return dif[:n].copy()


def activation(layers: np.ndarray, transfer: np.ndarray) -> np.ndarray:
    """layers: uint16 [Z,Y,X] with START_FLAG on start voxels; transfer: f64 [rows, cols]."""
    layers = np.ascontiguousarray(layers, dtype=np.uint16)
    transfer = np.ascontiguousarray(transfer, dtype=np.float64)
    Z, Y, X = layers.shape
    out = np.empty(layers.shape, dtype=np.float64)
    rc = lib().ekg_oracle_activation(_ptr(layers), Z, Y, X, _ptr(transfer), transfer.shape[0], transfer.shape[1], _ptr(out))
    if rc == -1:
        raise RuntimeError("Could not find starting point for excitation sequence")
    if rc == -2:
        raise RuntimeError("transfer (conduction) matrix does not define layer")
    if rc:
        raise RuntimeError("oracle activation failed: %d" % rc)
    return out


def _run(fn, layers, delay, layer_k, leads_zyx, nbhd, t_start, t_step, total_time):
    layers = np.ascontiguousarray(layers, dtype=np.uint16)
    delay = np.ascontiguousarray(delay, dtype=np.float64)
    layer_k = np.ascontiguousarray(layer_k, dtype=np.float64).reshape(-1, 9)
    leads = np.ascontiguousarray(leads_zyx, dtype=np.float64).reshape(-1, 3)
    Z, Y, X = layers.shape
    steps = int(np.ceil(total_time / t_step))
    out = np.zeros((leads.shape[0], steps), dtype=np.float64)
    n = fn(_ptr(layers), _ptr(delay), Z, Y, X, _ptr(layer_k), layer_k.shape[0], _ptr(leads), leads.shape[0],
           NBHD[nbhd] if isinstance(nbhd, str) else int(nbhd), float(t_start), float(t_step), float(total_time), _ptr(out))
    if n != steps:
        raise RuntimeError("oracle run failed: %d" % n)
    return out


def run_direct(layers, delay, layer_k, leads_zyx, nbhd="3D4", t_start=100.0, t_step=1.0, total_time=400.0):
    return _run(lib().ekg_oracle_run_direct, layers, delay, layer_k, leads_zyx, nbhd, t_start, t_step, total_time)


def run_factored(layers, delay, layer_k, leads_zyx, nbhd="3D4", t_start=100.0, t_step=1.0, total_time=400.0):
    return _run(lib().ekg_oracle_run_factored, layers, delay, layer_k, leads_zyx, nbhd, t_start, t_step, total_time)


def run_factored_slab(layers, delay, layer_k, leads_zyx, z0, z1, nbhd="3D4", t_start=100.0, t_step=1.0, total_time=400.0):
    fn = lambda *a: lib().ekg_oracle_run_factored_slab(*a[:-1], int(z0), int(z1), a[-1])
    return _run(fn, layers, delay, layer_k, leads_zyx, nbhd, t_start, t_step, total_time)


def ap_classes(layers, delay, n_layers):
    layers = np.ascontiguousarray(layers, dtype=np.uint16)
    delay = np.ascontiguousarray(delay, dtype=np.float64)
    idx = np.empty(layers.size, dtype=np.int64)
    K = lib().ekg_oracle_ap_classes(_ptr(layers), _ptr(delay), layers.size, int(n_layers), _ptr(idx))
    return int(K), idx.reshape(layers.shape)


def run_approximation(layer_k, t_start, t_step, total_time, delay):
    layer_k = np.ascontiguousarray(layer_k, dtype=np.float64).reshape(-1, 9)
    n = int(total_time / t_step)
    out = np.zeros(n, dtype=np.float64)
    lib().ekg_oracle_run_approximation(_ptr(layer_k), layer_k.shape[0], float(t_start), float(t_step), float(total_time), float(delay), _ptr(out))
    return out


FIT_D9 = (0.0, 0.0, 0.0, 0.001, 0.0, 0.00005, 0.0005, 0.01, 0.2)   # sim.cpp:877 `kd`


def fit_layers(border_k, n_layers, mid=-1, d9=FIT_D9, step=0.5, eps=1e-3, iterations=100):
    """Layer coefficients [n_layers, 9] of one parameter vector from its 2 or 3 border APs."""
    border_k = np.ascontiguousarray(border_k, dtype=np.float64).reshape(-1, 9)
    d9 = np.ascontiguousarray(d9, dtype=np.float64)
    out = np.zeros((int(n_layers), 9), dtype=np.float64)
    rc = lib().ekg_oracle_fit_layers(_ptr(border_k), border_k.shape[0], int(n_layers), int(mid), _ptr(d9), float(step), float(eps),
                                     int(iterations), _ptr(out))
    if rc:
        raise ValueError("ekg_oracle_fit_layers: bad arguments")
    return out


# ---- readers for oracle/ref_dump.cpp output --------------------------------------------------

def read_activation_dump(path):
    with open(path, "rb") as f:
        if f.read(8) != b"EKGACT1\0":
            raise ValueError("bad magic")
        Z, Y, X = struct.unpack("<3Q", f.read(24))
        layers = np.frombuffer(f.read(Z * Y * X * 2), dtype=np.uint16).reshape(Z, Y, X).copy()
        delay = np.frombuffer(f.read(Z * Y * X * 8), dtype=np.float64).reshape(Z, Y, X).copy()
    return layers, delay


def read_eval_dump(path):
    out = []
    with open(path, "rb") as f:
        if f.read(8) != b"EKGEVL1\0":
            raise ValueError("bad magic")
        nvec, nl, L, T, nc = struct.unpack("<5Q", f.read(40))
        for _ in range(nvec):
            hdr = f.read(8)
            if len(hdr) < 8:
                break
            (npar,) = struct.unpack("<Q", hdr)
            rd = lambda n: np.frombuffer(f.read(8 * n), dtype=np.float64).copy()
            rec = dict(params=rd(npar), layer_k=rd(nl * 9).reshape(nl, 9), leads_zyx=rd(L * 3).reshape(L, 3),
                       ecg=rd(L * T).reshape(L, T), criteria=rd(nc), violation=float(rd(1)[0]), seconds=float(rd(1)[0]))
            (rec["simulation_done"],) = struct.unpack("<Q", f.read(8))
            out.append(rec)
    return out
