"""ekgsim_b200/capi.py -- ctypes binding of the C ABI (include/ekgsim_b200.h).

The product is the shared library `libekgsim_b200.so` (hand-written sm_100a CUDA behind an
extern "C" boundary) plus the host-side C++ facade in ekgsim_b200/host/.  This module is the
thin Python view of the same ABI used by tests/, bench.py and __graft_entry__.py; it adds no
compute of its own and has no CPU fallback: if the library is missing or no CUDA device is
present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libekgsim_b200.so")

NBHD = {"2D4": 0, "2D8": 1, "3D4": 2, "3D8": 3, "cube": 3}
MODE_DEFAULT, MODE_DIRECT, MODE_HOISTED, MODE_SEPARABLE = 0, 1, 2, 3
FLAG_TIME_KERNEL = 0x100
FLAG_CORNER_SUM = 0x200   # SEPARABLE: interior voxels by the direct corner sum instead of the series (cross-check)
FIT_D9 = (0.0, 0.0, 0.0, 0.001, 0.0, 0.00005, 0.0005, 0.01, 0.2)   # sim.cpp:877 `kd`
START_FLAG = 0x1000
LINK_INFO_BYTES = 256   # EKG_LINK_INFO_BYTES

# every symbol include/ekgsim_b200.h declares: name -> (restype, argtypes)
_p, _i64, _d, _int = C.c_void_p, C.c_int64, C.c_double, C.c_int
SYMBOLS = {
    "ekg_abi_version": (_int, []),
    "ekg_last_error": (C.c_char_p, []),
    "ekg_device_count": (_int, []),
    "ekg_model_create": (_int, [_p, _i64, _i64, _i64, _p, _i64, _i64, _int, C.POINTER(_p)]),
    "ekg_model_destroy": (None, [_p]),
    "ekg_model_set_slab": (_int, [_p, _i64, _i64]),
    "ekg_model_num_voxels": (_i64, [_p]),
    "ekg_model_num_layers": (_i64, [_p]),
    "ekg_model_activation": (_int, [_p, _p, C.POINTER(_i64)]),
    "ekg_model_activation_ms": (_d, [_p]),
    "ekg_model_activation_brick_visits": (_i64, [_p]),
    "ekg_model_set_activation": (_int, [_p, _p]),
    "ekg_model_get_activation": (_int, [_p, _p]),
    "ekg_model_activation_begin": (_int, [_p]),
    "ekg_model_activation_relax": (_int, [_p, C.POINTER(_i64)]),
    "ekg_model_activation_relax_bounded": (_int, [_p, _i64, C.POINTER(_i64), C.POINTER(_i64)]),
    "ekg_model_plane_elems": (_i64, [_p]),
    "ekg_model_activation_export": (_int, [_p, _i64, _i64, _p, _p]),
    "ekg_model_activation_merge": (_int, [_p, _i64, _i64, _p, C.POINTER(_i64), _p]),
    "ekg_model_activation_merge_async": (_int, [_p, _i64, _i64, _p, _p, _p]),
    "ekg_model_activation_end": (_int, [_p, _p]),
    "ekg_model_activation_link_info": (_int, [_p, _p]),
    "ekg_model_activation_link": (_int, [_p, _int, _int, _p, _p]),
    "ekg_model_activation_linked_launch": (_int, [_p, _int]),
    "ekg_model_activation_linked_wait": (_int, [_p, C.POINTER(_i64), C.POINTER(_i64)]),
    "ekg_model_activation_linked_gather": (_int, [_p]),
    "ekg_model_activation_unlink": (_int, [_p]),
    "ekg_model_ap_classes": (_int, [_p, _p, C.POINTER(_i64)]),
    "ekg_simulate": (_int, [_p, _p, _p, _i64, _i64, _int, _d, _d, _d, _int, _p]),
    "ekg_simulate_criteria": (_int, [_p, _p, _p, _i64, _i64, _int, _d, _d, _d, _int, _p, _i64, _p, _int, _p, _p]),
    "ekg_simulate_device": (_int, [_p, _p, _p, _i64, _i64, _int, _d, _d, _d, _int, _p, _p]),
    "ekg_simulate_device_hinted": (_int, [_p, _p, _p, _i64, _i64, _int, _d, _d, _d, _int, _d, _d, _p, _p]),
    "ekg_fit_layers": (_int, [_p, _p, _i64, _i64, _i64, _p, _d, _d, _i64, _p]),
    "ekg_fit_layers_device": (_int, [_p, _p, _i64, _i64, _i64, _p, _d, _d, _i64, _p, _p]),
    "ekg_evaluate": (_int, [_p, _p, _i64, _i64, _p, _d, _d, _i64, _p, _i64, _i64, _int, _d, _d, _d, _int, _p, _i64, _p, _int, _p, _p, _p]),
    "ekg_last_kernel_ms": (_d, [_p]),
    "ekg_last_launch_count": (_i64, [_p]),
    "ekg_last_kernel_name": (C.c_char_p, [_p]),
}

_lib = None


class EkgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (code %d)" % (msg, code))
        self.code = code


def lib():
    """Loads libekgsim_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libekgsim_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise EkgError(rc, lib().ekg_last_error().decode("utf-8", "replace"))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def coefficient_hints(layer_k):
    """(k1_min, decay_max) of layer coefficients [..., 9] for simulate_device_hinted: the smallest k1, the largest of
    |k4 + k5| and |k5| (what ekg_simulate derives from its host buffers, capi.cu simulate_host)"""
    k = np.asarray(layer_k, dtype=np.float64).reshape(-1, 9)
    return float(k[:, 1].min()), float(np.maximum(np.abs(k[:, 4] + k[:, 5]), np.abs(k[:, 5])).max())


def n_steps(total_time, t_step):
    return int(np.ceil(total_time / t_step))  # simulator.cpp:471


class Model:
    """Device-resident voxel model (layers + conduction matrix + activation map)."""

    def __init__(self, layers, transfer, device=0):
        layers = np.ascontiguousarray(layers, dtype=np.uint16)
        if layers.ndim == 2:
            layers = layers[None]
        transfer = np.ascontiguousarray(transfer, dtype=np.float64)
        self.shape = layers.shape
        self.device = device
        self._h = C.c_void_p()
        Z, Y, X = layers.shape
        _check(lib().ekg_model_create(_ptr(layers), Z, Y, X, _ptr(transfer), transfer.shape[0], transfer.shape[1],
                                      device, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().ekg_model_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    @property
    def num_voxels(self):
        return int(lib().ekg_model_num_voxels(self._h))

    @property
    def num_layers(self):
        return int(lib().ekg_model_num_layers(self._h))

    def set_slab(self, z0, z1):
        _check(lib().ekg_model_set_slab(self._h, int(z0), int(z1)))

    def activation(self, download=True):
        """Runs the automaton on the GPU; returns (delay[Z,Y,X] f64, sweeps).  download=False leaves the map on the
        device (delay is None; `get_activation` fetches it later if it is wanted after all)."""
        out = np.empty(self.shape, dtype=np.float64) if download else None
        sweeps = C.c_int64(0)
        _check(lib().ekg_model_activation(self._h, _ptr(out) if download else None, C.byref(sweeps)))
        return out, int(sweeps.value)

    @property
    def activation_ms(self):
        return float(lib().ekg_model_activation_ms(self._h))

    @property
    def activation_brick_visits(self):
        return int(lib().ekg_model_activation_brick_visits(self._h))

    # -- z-slab sharded automaton (driver: ekgsim_b200/dist.py::sharded_activation) --
    @property
    def plane_elems(self):
        return int(lib().ekg_model_plane_elems(self._h))

    def activation_begin(self):
        _check(lib().ekg_model_activation_begin(self._h))

    def activation_relax(self):
        v = C.c_int64(0)
        _check(lib().ekg_model_activation_relax(self._h, C.byref(v)))
        return int(v.value)

    def activation_relax_bounded(self, max_visits):
        """-> (brick visits, bricks still queued); max_visits = 0: until the slab's fixed point"""
        v, left = C.c_int64(0), C.c_int64(0)
        _check(lib().ekg_model_activation_relax_bounded(self._h, int(max_visits), C.byref(v), C.byref(left)))
        return int(v.value), int(left.value)

    def activation_export(self, z_begin, z_end, d_planes, stream=0):
        _check(lib().ekg_model_activation_export(self._h, int(z_begin), int(z_end), C.c_void_p(d_planes), C.c_void_p(stream)))

    def activation_merge(self, z_begin, z_end, d_planes, stream=0):
        n = C.c_int64(0)
        _check(lib().ekg_model_activation_merge(self._h, int(z_begin), int(z_end), C.c_void_p(d_planes), C.byref(n), C.c_void_p(stream)))
        return int(n.value)

    def activation_merge_async(self, z_begin, z_end, d_planes, d_counter, stream=0):
        """like activation_merge, but the number of improved cells is added to the u64 device counter at `d_counter`"""
        _check(lib().ekg_model_activation_merge_async(self._h, int(z_begin), int(z_end), C.c_void_p(d_planes), C.c_void_p(d_counter), C.c_void_p(stream)))

    def activation_end(self, download=True):
        out = np.empty(self.shape, dtype=np.float64) if download else None
        _check(lib().ekg_model_activation_end(self._h, _ptr(out) if download else None))
        return out

    # peer-linked sharded automaton (include/ekgsim_b200.h, "peer-linked")
    def activation_link_info(self):
        buf = C.create_string_buffer(LINK_INFO_BYTES)
        _check(lib().ekg_model_activation_link_info(self._h, buf))
        return buf.raw

    def activation_link(self, rank, infos, slabs):
        """infos: every rank's activation_link_info() in rank order; slabs: every rank's (z_begin, z_end)"""
        blob = b"".join(infos)
        assert len(blob) == LINK_INFO_BYTES * len(slabs)
        sl = np.ascontiguousarray(np.asarray(slabs, dtype=np.int64).reshape(-1, 2))
        _check(lib().ekg_model_activation_link(self._h, int(rank), len(slabs), C.c_char_p(blob), _ptr(sl)))

    def activation_linked_launch(self, max_ctas=0):
        _check(lib().ekg_model_activation_linked_launch(self._h, int(max_ctas)))

    def activation_linked_wait(self):
        """-> (brick visits, (bricks queued at other ranks, bricks queued here by others, cells written to other ranks))"""
        v = _i64(0)
        rem = (_i64 * 3)()
        _check(lib().ekg_model_activation_linked_wait(self._h, C.byref(v), rem))
        return int(v.value), tuple(int(x) for x in rem)

    def activation_linked_gather(self):
        _check(lib().ekg_model_activation_linked_gather(self._h))

    def activation_unlink(self):
        _check(lib().ekg_model_activation_unlink(self._h))

    def set_activation(self, delay):
        delay = np.ascontiguousarray(delay, dtype=np.float64)
        assert delay.size == int(np.prod(self.shape))
        _check(lib().ekg_model_set_activation(self._h, _ptr(delay)))

    def get_activation(self):
        out = np.empty(self.shape, dtype=np.float64)
        _check(lib().ekg_model_get_activation(self._h, _ptr(out)))
        return out

    def ap_classes(self):
        idx = np.empty(self.shape, dtype=np.int64)
        K = C.c_int64(0)
        _check(lib().ekg_model_ap_classes(self._h, _ptr(idx), C.byref(K)))
        return int(K.value), idx

    def simulate(self, layer_k, leads_zyx, nbhd="3D4", t_start=100.0, t_step=1.0, total_time=400.0, mode=MODE_DEFAULT):
        """Host buffers in, host ECG [B, L, T] out (the call EkgSim::run maps to)."""
        layer_k = np.ascontiguousarray(layer_k, dtype=np.float64)
        if layer_k.ndim == 2:
            layer_k = layer_k[None]
        B = layer_k.shape[0]
        assert layer_k.shape[1:] == (self.num_layers, 9), layer_k.shape
        leads = np.ascontiguousarray(leads_zyx, dtype=np.float64)
        if leads.ndim == 2:
            leads = np.broadcast_to(leads[None], (B,) + leads.shape).copy()
        L = leads.shape[1]
        T = n_steps(total_time, t_step) if t_step > 0 and total_time > 0 else 0
        out = np.empty((B, L, max(T, 0)), dtype=np.float64)
        _check(lib().ekg_simulate(self._h, _ptr(layer_k), _ptr(leads), B, L, NBHD[nbhd] if isinstance(nbhd, str) else int(nbhd),
                                  float(t_start), float(t_step), float(total_time), int(mode), _ptr(out)))
        return out

    def simulate_criteria(self, layer_k, leads_zyx, targets, comparison=2, target_offsets=None, nbhd="3D4", t_start=100.0,
                          t_step=1.0, total_time=400.0, mode=MODE_DEFAULT, want_ecg=False):
        """Simulation + the reference's curve comparison on the device; returns criteria [B, L] (and the ECGs)."""
        layer_k = np.ascontiguousarray(layer_k, dtype=np.float64)
        if layer_k.ndim == 2:
            layer_k = layer_k[None]
        B = layer_k.shape[0]
        leads = np.ascontiguousarray(leads_zyx, dtype=np.float64)
        if leads.ndim == 2:
            leads = np.broadcast_to(leads[None], (B,) + leads.shape).copy()
        L = leads.shape[1]
        targets = np.ascontiguousarray(targets, dtype=np.float64)
        assert targets.shape[0] == L
        off = None if target_offsets is None else np.ascontiguousarray(target_offsets, dtype=np.float64)
        T = n_steps(total_time, t_step)
        crit = np.empty((B, L), dtype=np.float64)
        ecg = np.empty((B, L, T), dtype=np.float64) if want_ecg else None
        _check(lib().ekg_simulate_criteria(self._h, _ptr(layer_k), _ptr(leads), B, L, NBHD[nbhd] if isinstance(nbhd, str) else int(nbhd),
                                           float(t_start), float(t_step), float(total_time), int(mode), _ptr(targets), targets.shape[1],
                                           _ptr(off) if off is not None else None, int(comparison), _ptr(crit),
                                           _ptr(ecg) if ecg is not None else None))
        return (crit, ecg) if want_ecg else crit

    def simulate_device(self, d_layer_k, d_leads, B, L, d_ecg, nbhd="3D4", t_start=100.0, t_step=1.0, total_time=400.0,
                        mode=MODE_DEFAULT, stream=0):
        """Raw device pointers (ints) on this model's device; asynchronous on `stream`."""
        _check(lib().ekg_simulate_device(self._h, C.c_void_p(d_layer_k), C.c_void_p(d_leads), int(B), int(L),
                                         NBHD[nbhd] if isinstance(nbhd, str) else int(nbhd), float(t_start), float(t_step),
                                         float(total_time), int(mode), C.c_void_p(d_ecg), C.c_void_p(stream)))

    def simulate_device_hinted(self, d_layer_k, d_leads, B, L, d_ecg, k1_min, decay_max, nbhd="3D4", t_start=100.0, t_step=1.0,
                               total_time=400.0, mode=MODE_DEFAULT, stream=0):
        """simulate_device for a caller who knows the coefficients on the host (`coefficient_hints`): no read-back inside"""
        _check(lib().ekg_simulate_device_hinted(self._h, C.c_void_p(d_layer_k), C.c_void_p(d_leads), int(B), int(L),
                                                NBHD[nbhd] if isinstance(nbhd, str) else int(nbhd), float(t_start), float(t_step),
                                                float(total_time), int(mode), float(k1_min), float(decay_max), C.c_void_p(d_ecg), C.c_void_p(stream)))

    def fit_layers(self, border_k, mid=-1, d9=FIT_D9, step=0.5, eps=1e-3, iterations=100):
        """Border APs [B, 2|3, 9] -> layer coefficients [B, n_layers, 9] (sim.cpp:751-916 on the device)."""
        border_k = np.ascontiguousarray(border_k, dtype=np.float64)
        if border_k.ndim == 2:
            border_k = border_k[None]
        B, nb = border_k.shape[:2]
        d9 = np.ascontiguousarray(d9, dtype=np.float64)
        out = np.empty((B, self.num_layers, 9), dtype=np.float64)
        _check(lib().ekg_fit_layers(self._h, _ptr(border_k), B, nb, int(mid), _ptr(d9), float(step), float(eps), int(iterations), _ptr(out)))
        return out

    def fit_layers_device(self, d_border_k, B, n_border, d_layer_k, mid=-1, d9=FIT_D9, step=0.5, eps=1e-3, iterations=100, stream=0):
        """Raw device pointers (ints) on this model's device; asynchronous on `stream`."""
        d9 = np.ascontiguousarray(d9, dtype=np.float64)
        _check(lib().ekg_fit_layers_device(self._h, C.c_void_p(d_border_k), int(B), int(n_border), int(mid), _ptr(d9), float(step),
                                           float(eps), int(iterations), C.c_void_p(d_layer_k), C.c_void_p(stream)))

    def evaluate(self, border_k, leads_zyx, targets, mid=-1, comparison=2, target_offsets=None, d9=FIT_D9, step=0.5, eps=1e-3,
                 iterations=100, nbhd="3D4", t_start=100.0, t_step=1.0, total_time=400.0, mode=MODE_DEFAULT, want_layer_k=False,
                 want_ecg=False):
        """Border APs + leads in, criteria [B, L] out; fit, simulation and comparison stay on the device."""
        border_k = np.ascontiguousarray(border_k, dtype=np.float64)
        if border_k.ndim == 2:
            border_k = border_k[None]
        B, nb = border_k.shape[:2]
        leads = np.ascontiguousarray(leads_zyx, dtype=np.float64)
        if leads.ndim == 2:
            leads = np.broadcast_to(leads[None], (B,) + leads.shape).copy()
        L = leads.shape[1]
        targets = np.ascontiguousarray(targets, dtype=np.float64)
        assert targets.shape[0] == L
        off = None if target_offsets is None else np.ascontiguousarray(target_offsets, dtype=np.float64)
        d9 = np.ascontiguousarray(d9, dtype=np.float64)
        T = n_steps(total_time, t_step)
        crit = np.empty((B, L), dtype=np.float64)
        lk = np.empty((B, self.num_layers, 9), dtype=np.float64) if want_layer_k else None
        ecg = np.empty((B, L, T), dtype=np.float64) if want_ecg else None
        _check(lib().ekg_evaluate(self._h, _ptr(border_k), nb, int(mid), _ptr(d9), float(step), float(eps), int(iterations),
                                  _ptr(leads), B, L, NBHD[nbhd] if isinstance(nbhd, str) else int(nbhd),
                                  float(t_start), float(t_step), float(total_time), int(mode), _ptr(targets), targets.shape[1],
                                  _ptr(off) if off is not None else None, int(comparison), _ptr(crit),
                                  _ptr(lk) if lk is not None else None, _ptr(ecg) if ecg is not None else None))
        return crit, lk, ecg

    @property
    def last_kernel_ms(self):
        return float(lib().ekg_last_kernel_ms(self._h))

    @property
    def last_launch_count(self):
        return int(lib().ekg_last_launch_count(self._h))

    @property
    def last_kernel_name(self):
        return lib().ekg_last_kernel_name(self._h).decode()
