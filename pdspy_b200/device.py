"""Process-local device state: dataset handles keyed on the caller's arrays.

The reference passes the same u, v (and the same data arrays) to every likelihood call;
uploading them once and keeping a handle is what makes the GPU path a drop-in.  Handles live
here, never inside Visibilities objects (those get pickled by the samplers)."""
import ctypes
from collections import OrderedDict

import numpy as np

from . import _lib
from ._lib import HOST, DEVICE, check, ptr, f64

_CACHE = OrderedDict()
_CACHE_MAX = 8


def _fingerprint(a):
    """Cheap content fingerprint (<= 4096 strided samples + the ends)."""
    if a.size == 0:
        return (0,)
    step = max(1, a.size // 4096)
    flat = a.reshape(-1)
    s = flat[::step]
    return (a.size, float(s.sum()), float(flat[0]), float(flat[-1]), float(flat[a.size // 2]))


class Dataset:
    """Owns a pdsb_dataset handle: uv points (+ optionally the observed data)."""

    def __init__(self, u, v, kind=HOST):
        L = _lib.lib()
        self._h = ctypes.c_void_p()
        if kind == HOST:
            u, v = f64(u), f64(v)
            if u.shape != v.shape or u.ndim != 1:
                raise ValueError("u and v must be 1-D arrays of the same length")
            n = u.size
        else:
            u, v, n = u
        check(L.pdsb_dataset_create(ptr(u), ptr(v), n, kind, ctypes.byref(self._h)))
        nuv, nuvh, nf, herm = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int(), ctypes.c_int()
        check(L.pdsb_dataset_info(self._h, ctypes.byref(nuv), ctypes.byref(nuvh), ctypes.byref(nf),
                                  ctypes.byref(herm)))
        self.nuv, self.nuv_unique, self.hermitian = nuv.value, nuvh.value, bool(herm.value)
        self.nf = 0
        self.data_key = None

    @property
    def handle(self):
        if not self._h:
            raise _lib.PdsbError("dataset was destroyed")
        return self._h

    def set_data(self, real, imag, weights, kind=HOST):
        L = _lib.lib()
        if kind == HOST:
            real, imag, weights = f64(real), f64(imag), f64(weights)
            if real.ndim != 2 or real.shape[0] != self.nuv or imag.shape != real.shape or weights.shape != real.shape:
                raise ValueError("real, imag, weights must be [nuv, nf]")
            nf = real.shape[1]
        else:
            real, imag, weights, nf = real
        check(L.pdsb_dataset_set_data(self.handle, ptr(real), ptr(imag), ptr(weights), nf, kind))
        self.nf = nf

    def destroy(self):
        if self._h and _lib._lib is not None:
            _lib._lib.pdsb_dataset_destroy(self._h)
        self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def dataset_for(u, v, data=None):
    """Cached Dataset for the caller's u, v (and optionally data = (real, imag, weights))."""
    key = (u.ctypes.data, v.ctypes.data, u.size) if isinstance(u, np.ndarray) and isinstance(v, np.ndarray) \
        else None
    fp = (_fingerprint(np.asarray(u)), _fingerprint(np.asarray(v)))
    ent = _CACHE.get(key) if key is not None else None
    if ent is None or ent[0] != fp:
        if ent is not None:
            ent[1].destroy()
        ds = Dataset(np.asarray(u), np.asarray(v))
        if key is not None:
            _CACHE[key] = (fp, ds)
            while len(_CACHE) > _CACHE_MAX:
                _, (_, old) = _CACHE.popitem(last=False)
                old.destroy()
    else:
        ds = ent[1]
        _CACHE.move_to_end(key)
    if data is not None:
        real, imag, weights = data
        dkey = (real.ctypes.data, imag.ctypes.data, weights.ctypes.data, real.shape,
                _fingerprint(real), _fingerprint(imag), _fingerprint(weights))
        if ds.data_key != dkey:
            ds.set_data(real, imag, weights)
            ds.data_key = dkey
    return ds


def clear_cache():
    """Drop every cached device handle (call after mutating u, v or data arrays in place)."""
    while _CACHE:
        _, (_, ds) = _CACHE.popitem()
        ds.destroy()


DFT_KERNELS = {"fp32": 0, "tcgen05": 200, "fp64": 300}


def set_dft_kernel(kind="fp32"):
    """Select the DFT kernel every later interpolate_model / likelihood call runs.

    "fp32" (default): the FP32-pipe kernel of BASELINE.json's north star.  "tcgen05": the tensor-core
    kernel (tcgen05.mma / TMEM, lattice-split fp16 operands, ~10x faster, same parity bounds).
    "fp64": the all-fp64 reference kernel (1e-13 of max|V| from the CPU oracle, ~8x slower than "fp32").
    An int is passed to pdsb_set_dft_variant unchanged.  The environment variable PDSPY_B200_DFT
    sets the initial choice."""
    variant = DFT_KERNELS[kind] if isinstance(kind, str) else int(kind)
    _lib.check(_lib.lib().pdsb_set_dft_variant(variant))
    return variant
