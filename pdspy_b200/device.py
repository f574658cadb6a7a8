"""Process-local device state: dataset handles keyed on the caller's arrays.

The reference passes the same u, v (and the same data arrays) to every likelihood call;
uploading them once and keeping a handle is what makes the GPU path a drop-in.  Handles live
here, never inside Visibilities objects (those get pickled by the samplers)."""
import ctypes
from collections import OrderedDict

import numpy as np
import numpy

from . import _lib
from ._lib import HOST, DEVICE, check, ptr, f64

_CACHE = OrderedDict()
_CACHE_MAX = 8

# How dataset_for decides that the device copy of caller-owned arrays is still valid.
#   "hash"   (default) every call hashes the arrays in full on the host (pdsb_hash64: multi-threaded,
#            memory-bound - about 30 GB/s, i.e. 1 ms for a uv list, 40 ms for 1M x 64 data) and compares with
#            the hash taken at upload.  Any in-place edit is seen (the reference edits arrays in place:
#            invert.py:15-47).
#   "freeze" the arrays are made read-only (flags.writeable = False) when they are uploaded, and an array that
#            is still read-only at the same address is trusted without reading it: an in-place edit raises
#            numpy's "assignment destination is read-only" instead of going unnoticed.  clear_cache() /
#            release(array) make them writeable again.  Zero cost per call; edits through an older writeable
#            view of the same memory are NOT seen.
_POLICY = "hash"


def set_cache_policy(policy):
    """"hash" (default) or "freeze": see above."""
    global _POLICY
    if policy not in ("hash", "freeze"):
        raise ValueError("cache policy must be 'hash' or 'freeze'")
    _POLICY = policy


class _Stamp:
    """What is remembered about one uploaded host array."""

    def __init__(self, a):
        self.addr, self.shape, self.hash = a.ctypes.data, a.shape, _lib.host_hash(a)
        self.frozen = None
        if _POLICY == "freeze" and a.flags.writeable:
            try:
                a.flags.writeable = False
                self.frozen = a
            except ValueError:
                pass

    def valid_for(self, a):
        if a.ctypes.data != self.addr or a.shape != self.shape:
            return False
        if _POLICY == "freeze" and self.frozen is a and not a.flags.writeable:
            return True
        return _lib.host_hash(a) == self.hash

    def thaw(self):
        if self.frozen is not None:
            try:
                self.frozen.flags.writeable = True
            except ValueError:
                pass
            self.frozen = None


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
    """Cached Dataset for the caller's u, v (and optionally data = (real, imag, weights)): uploaded once, reused
    while the arrays are unchanged (see the cache policy above)."""
    u, v = f64(u), f64(v)
    key = (u.ctypes.data, v.ctypes.data, u.size)
    ent = _CACHE.get(key)
    if ent is not None and not (ent["u"].valid_for(u) and ent["v"].valid_for(v)):
        _drop(key)
        ent = None
    if ent is None:
        ent = {"ds": Dataset(u, v), "u": _Stamp(u), "v": _Stamp(v), "data": None}
        _CACHE[key] = ent
        while len(_CACHE) > _CACHE_MAX:
            _drop(next(iter(_CACHE)))
    else:
        _CACHE.move_to_end(key)
    ds = ent["ds"]
    if data is not None:
        arrays = [f64(a) for a in data]
        st = ent["data"]
        if st is None or not all(s.valid_for(a) for s, a in zip(st, arrays)):
            if st is not None:
                for s in st:
                    s.thaw()
            ds.set_data(*arrays)
            ent["data"] = [_Stamp(a) for a in arrays]
    return ds


def _drop(key):
    ent = _CACHE.pop(key)
    for st in [ent["u"], ent["v"]] + list(ent["data"] or []):
        st.thaw()
    ent["ds"].destroy()


def clear_cache():
    """Drop every cached device handle (and un-freeze the arrays the "freeze" policy made read-only)."""
    while _CACHE:
        _drop(next(iter(_CACHE)))
    _POOL.clear()


def release(array):
    """Forget the device copies made from `array` (a u, v, real, imag or weights array handed to a likelihood
    call) and make it writeable again."""
    addr = array.ctypes.data
    for key in [k for k, e in _CACHE.items()
                if any(st.addr == addr for st in [e["u"], e["v"]] + list(e["data"] or []))]:
        _drop(key)


# ---------------------------------------------------------------------------------------------------
# Device-resident model visibilities (interpolate_model's result before anybody reads it on the host).
# Visibilities objects carry only an integer token; the buffers live here, so nothing holding a CUDA
# allocation is reachable from an object the samplers pickle.
_MODELS = {}
_POOL = {}            # nbytes -> [free DeviceBuffer, ...]  (cudaMalloc of 2 x 512 MB per likelihood call is not free)
_next_token = [1]


def _take(nbytes):
    free = _POOL.get(nbytes)
    if free:
        return free.pop()
    return _lib.DeviceBuffer(nbytes)


def register_model(shape):
    """Two device buffers (real, imag) of `shape` doubles; returns (token, real, imag)."""
    nbytes = int(numpy.prod(shape)) * 8
    re, im = _take(max(nbytes, 8)), _take(max(nbytes, 8))
    token = _next_token[0]
    _next_token[0] += 1
    _MODELS[token] = (re, im, tuple(shape))
    return token, re, im


def model_buffers(token):
    """(real DeviceBuffer, imag DeviceBuffer, shape) or None when the token is unknown (other process, released)."""
    return _MODELS.get(token)


def release_model(token):
    ent = _MODELS.pop(token, None)
    if ent is not None:
        for b in ent[:2]:
            free = _POOL.setdefault(b.nbytes, [])
            if len(free) < 4:
                free.append(b)
            else:
                b.free()


DFT_KERNELS = {"fp32": 0, "tcgen05": 200, "fp64": 300, "nufft": 400}


def set_dft_kernel(kind="fp32"):
    """Select the DFT kernel every later interpolate_model / likelihood call runs.

    "fp32" (default): the FP32-pipe kernel of BASELINE.json's north star.  "tcgen05": the tensor-core
    kernel (tcgen05.mma / TMEM, lattice-split fp16 operands, ~10x faster, same parity bounds).
    "fp64": the all-fp64 reference kernel (1e-13 of max|V| from the CPU oracle, ~8x slower than "fp32").
    "nufft": not a direct sum at all - the same exact transform through a type-2 non-uniform FFT with an 8-point kernel
    (4e-8 of max|V|, ~100x faster than "fp32" on cubes; even image sides up to 2048, other shapes take "fp32";
    DESIGN.md 4.2f).
    An int is passed to pdsb_set_dft_variant unchanged.  The environment variable PDSPY_B200_DFT
    sets the initial choice."""
    variant = DFT_KERNELS[kind] if isinstance(kind, str) else int(kind)
    _lib.check(_lib.lib().pdsb_set_dft_variant(variant))
    return variant
