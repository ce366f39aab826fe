"""tests/hostlib.py -- ctypes view of libekgsim_host.so (the C shim over the host-side C++ glue)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "ekgsim_b200", "libekgsim_host.so")
CLI = os.path.join(ROOT, "ekgsim_b200", "bin", "ekgSim")
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(PATH)
        L.ekg_host_last_error.restype = C.c_char_p
        L.ekg_host_wohlfart_plus.restype = C.c_double
        L.ekg_host_wohlfart_plus.argtypes = [C.c_void_p, C.c_double]
        L.ekg_host_apd90.restype = C.c_double
        L.ekg_host_apd90.argtypes = [C.c_void_p]
        L.ekg_host_evaluator_create.restype = C.c_void_p
        L.ekg_host_evaluator_create.argtypes = [C.c_char_p, C.c_int]
        L.ekg_host_evaluator_create_on.restype = C.c_void_p
        L.ekg_host_evaluator_create_on.argtypes = [C.c_char_p, C.c_char_p]
        L.ekg_host_num_devices.argtypes = [C.c_void_p]
        L.ekg_host_evaluator_destroy.argtypes = [C.c_void_p]
        L.ekg_host_num_criteria.argtypes = [C.c_void_p]
        L.ekg_host_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ekg_host_eval_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ekg_host_layer_coefficients.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ekg_host_generate_test_shape.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
        L.ekg_host_load_shape.restype = C.c_int64
        L.ekg_host_load_shape.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.ekg_host_run_approximation.restype = C.c_int64
        L.ekg_host_run_approximation.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p]
        _lib = L
    return _lib


def run_approximation(layer_k, start, length, step, delay):
    """EkgSim::runApproximation of the facade (host only): endo AP delayed minus epi AP."""
    k = np.ascontiguousarray(layer_k, dtype=np.float64).reshape(-1, 9)
    out = np.zeros(int(length / step) + 8)
    n = lib().ekg_host_run_approximation(k.ctypes.data, k.shape[0], int(start), int(length), float(step), float(delay), out.ctypes.data)
    if n < 0:
        raise RuntimeError(lib().ekg_host_last_error().decode())
    return out[:n].copy()


def generate_test_shape(export_as=""):
    """The facade's built-in test shape (InputLoader::generateTestShape): layers u16 [1, 160, 120] with the start flag."""
    layers = np.zeros(160 * 120, dtype=np.uint16)
    dims = np.zeros(3, dtype=np.int64)
    if lib().ekg_host_generate_test_shape(layers.ctypes.data, dims.ctypes.data, export_as.encode()) != 0:
        raise RuntimeError(lib().ekg_host_last_error().decode())
    return layers.reshape(tuple(int(d) for d in dims))


def load_shape(fname, capacity=1 << 22):
    """A .matrix shape file through the facade's parser: layers u16 [Z, Y, X] with the start flag."""
    layers = np.zeros(capacity, dtype=np.uint16)
    dims = np.zeros(3, dtype=np.int64)
    n = lib().ekg_host_load_shape(fname.encode(), layers.ctypes.data, capacity, dims.ctypes.data)
    if n < 0:
        raise RuntimeError(lib().ekg_host_last_error().decode())
    return layers[:n].reshape(tuple(int(d) for d in dims)).copy()


class Evaluator:
    """Runs in `workdir` (must hold simulator.ini + inputs, like the reference CLI)."""

    def __init__(self, workdir, with_device=True, n_layers=24, n_leads=2, devices=None, slabs=None):
        """devices: None (one GPU) | "all" | "<count>" | "0,1,..." -- batches are split over them (Evaluator::evalBatch);
        slabs: the same spellings -- ONE model as z-slabs over those GPUs (EKGSIM_B200_SLABS, EkgSim::setSlabDevices)"""
        self.workdir, self.n_layers, self.n_leads = workdir, n_layers, n_leads
        cwd = os.getcwd()
        os.chdir(workdir)
        old = os.environ.pop("EKGSIM_B200_SLABS", None)
        if slabs is not None:
            os.environ["EKGSIM_B200_SLABS"] = str(slabs)
        try:
            if devices is not None:
                self.h = lib().ekg_host_evaluator_create_on(b"simulator.ini", str(devices).encode())
            else:
                self.h = lib().ekg_host_evaluator_create(b"simulator.ini", 1 if with_device else 0)
        finally:
            os.chdir(cwd)
            os.environ.pop("EKGSIM_B200_SLABS", None)
            if old is not None:
                os.environ["EKGSIM_B200_SLABS"] = old
        if not self.h:
            raise RuntimeError(lib().ekg_host_last_error().decode())
        self.n_crit = lib().ekg_host_num_criteria(self.h)
        self.n_devices = lib().ekg_host_num_devices(self.h) if with_device else 0

    def close(self):
        if self.h:
            lib().ekg_host_evaluator_destroy(self.h)
            self.h = None

    def _in(self, fn):
        cwd = os.getcwd()
        os.chdir(self.workdir)
        try:
            return fn()
        finally:
            os.chdir(cwd)

    def layer_coefficients(self, genes):
        g = np.ascontiguousarray(genes, dtype=np.float64)
        k = np.zeros((self.n_layers, 9))
        leads = np.zeros((self.n_leads, 3))
        viol = C.c_double(0)
        rc = lib().ekg_host_layer_coefficients(self.h, g.ctypes.data, len(g), k.ctypes.data, leads.ctypes.data, C.byref(viol))
        if rc:
            raise RuntimeError(lib().ekg_host_last_error().decode())
        return k, leads, viol.value

    def eval(self, genes):
        g = np.ascontiguousarray(genes, dtype=np.float64)
        crit = np.zeros(self.n_crit)
        viol = C.c_double(0)
        rc = self._in(lambda: lib().ekg_host_eval(self.h, g.ctypes.data, len(g), crit.ctypes.data, C.byref(viol)))
        if rc:
            raise RuntimeError(lib().ekg_host_last_error().decode())
        return crit, viol.value

    def eval_batch(self, genes, threads=0):
        g = np.ascontiguousarray(genes, dtype=np.float64)
        B, n = g.shape
        crit = np.zeros((B, self.n_crit))
        viol = np.zeros(B)
        rc = self._in(lambda: lib().ekg_host_eval_batch(self.h, g.ctypes.data, n, B, threads, crit.ctypes.data, viol.ctypes.data))
        if rc:
            raise RuntimeError(lib().ekg_host_last_error().decode())
        return crit, viol


def read_extern_output(path, n_criteria, n_properties=0):
    """ExternalEvaluation::readOut restated (ExternalEvaluation.h:118-151): at every position where a value is
    expected a '#' starts a comment line ("violation <v>" sets the violation); reading stops after the last value."""
    data = open(path).read()
    pos, values, violation = 0, [], 0.0
    while len(values) < n_criteria + n_properties:
        if pos < len(data) and data[pos] == "#":          # file.peek() == '#'
            end = data.find("\n", pos)
            end = len(data) if end < 0 else end
            words = data[pos + 1:end].split()
            if words and words[0] == "violation":
                violation = float(words[1])
            pos = end + 1                                   # getline consumes the newline
        else:
            while pos < len(data) and data[pos].isspace():  # operator>> skips leading whitespace ...
                pos += 1
            end = pos
            while end < len(data) and not data[end].isspace():
                end += 1
            try:
                values.append(float(data[pos:end]))
            except ValueError:                              # ... and a failed read leaves infinity behind
                values.append(float("inf"))
                violation = float("inf")
            pos = end                                       # the newline after a value is NOT consumed
    return values[:n_criteria], violation
