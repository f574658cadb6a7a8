"""ctypes binding of libpdsb.so (include/pdsb.h).  No torch, no CPU fallback: if the
library or a CUDA device is missing, every compute call raises PdsbError."""
import ctypes
import os

import numpy as np

HOST, DEVICE = 0, 1
CONV = {"pillbox": 0, "expsinc": 1}
WEIGHTING = {"natural": 0, "uniform": 1, "superuniform": 2, "robust": 3}
MODE = {"continuum": 0, "spectralline": 1}

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PDSPY_B200_LIB") or os.path.join(_HERE, "libpdsb.so")     # override: tuning builds


class PdsbError(RuntimeError):
    pass


_c_int, _c_i64, _c_dbl = ctypes.c_int, ctypes.c_int64, ctypes.c_double
_P = ctypes.c_void_p
_PP = ctypes.POINTER(ctypes.c_void_p)

# name -> argtypes ; every function returns int except pdsb_last_error / pdsb_version
SIGNATURES = {
    "pdsb_version": [],
    "pdsb_device_count": [ctypes.POINTER(_c_int)],
    "pdsb_init": [_c_int],
    "pdsb_shutdown": [],
    "pdsb_last_error": [],
    "pdsb_get_stream": [ctypes.POINTER(ctypes.c_uint64)],
    "pdsb_set_stream": [ctypes.c_uint64],
    "pdsb_reset_stream": [],
    "pdsb_synchronize": [],
    "pdsb_device_info": [ctypes.POINTER(_c_int), ctypes.POINTER(_c_int), ctypes.POINTER(_c_i64),
                         ctypes.POINTER(_c_int), ctypes.POINTER(_c_int)],
    "pdsb_device_alloc": [_PP, _c_i64],
    "pdsb_device_free": [_P],
    "pdsb_host_alloc_pinned": [_PP, _c_i64],
    "pdsb_host_free_pinned": [_P],
    "pdsb_memcpy": [_P, _c_int, _P, _c_int, _c_i64],
    "pdsb_memset": [_P, _c_int, _c_i64],
    "pdsb_timer_start": [],
    "pdsb_timer_stop": [ctypes.POINTER(_c_dbl)],
    "pdsb_profile_enable": [_c_int],
    "pdsb_profile_reset": [],
    "pdsb_profile_get": [ctypes.c_char_p, ctypes.POINTER(_c_dbl), ctypes.POINTER(_c_i64)],
    "pdsb_launch_count": [ctypes.POINTER(_c_i64)],
    "pdsb_dataset_create": [_P, _P, _c_i64, _c_int, _PP],
    "pdsb_dataset_set_data": [_P, _P, _P, _P, _c_int, _c_int],
    "pdsb_dataset_info": [_P, ctypes.POINTER(_c_i64), ctypes.POINTER(_c_i64), ctypes.POINTER(_c_int),
                          ctypes.POINTER(_c_int)],
    "pdsb_dataset_destroy": [_P],
    "pdsb_sample_image": [_P, _P, _c_int, _c_int, _c_int, _c_int, _c_dbl, _c_dbl, _c_dbl, _P, _P, _c_int],
    "pdsb_loglike": [_P, _P, _c_int, _c_int, _c_int, _c_int, _c_dbl, _c_dbl, _c_dbl, _P, _P],
    "pdsb_sample_image_ex": [_P, _P, _c_int, _c_int, _c_int, _c_int, _c_dbl, _c_dbl, _c_dbl, _P, _c_dbl, _c_dbl, _c_dbl,
                             _P, _P, _c_int],
    "pdsb_loglike_ex": [_P, _P, _c_int, _c_int, _c_int, _c_int, _c_dbl, _c_dbl, _c_dbl, _P, _c_dbl, _c_dbl, _c_dbl,
                        _P, _P],
    "pdsb_loglike_device": [_P, _P, _c_int, _c_int, _c_int, _c_int, _c_dbl, _c_dbl, _c_dbl, _P],
    "pdsb_dataset_logsum": [_P, ctypes.POINTER(_c_dbl)],
    "pdsb_loglike_batch": [_P, _P, _c_int, _c_int, _c_int, _c_int, _c_int, _c_dbl, _P, _P, _P],
    "pdsb_sample_triangles": [_P, _P, _c_int, _P, _c_i64, _c_int, _c_int, _c_dbl, _c_dbl, _P, _P, _c_int],
    "pdsb_triangle_record_bytes": [],
    "pdsb_chi2": [_P, _P, _P, _P, _P, _c_i64, _c_int, _P],
    "pdsb_chi2_dataset": [_P, _P, _P, _c_int, _P],
    "pdsb_hash64": [_P, _c_i64, ctypes.POINTER(ctypes.c_uint64)],
    "pdsb_chisq": [_P, _P, _P, _P, _P, _c_i64, _c_int, _c_int, ctypes.POINTER(ctypes.c_float)],
    "pdsb_grid": [_P, _P, _P, _P, _P, _P, _c_i64, _c_int, _c_int, _c_int, _c_dbl, _P, _P, _c_int, _c_int,
                  _c_dbl, _c_int, _c_int, _c_int, _c_int, _P, _P, _P, _P, _P, _P, _c_int,
                  ctypes.POINTER(_c_i64)],
    "pdsb_grid_normalise": [_P, _P, _P, _c_int, _c_int, _c_int],
    "pdsb_freqcorrect": [_P, _P, _P, _c_i64, _c_int, _c_dbl, _c_int, _P, _P],
    "pdsb_average": [_P, _P, _P, _P, _P, _P, _P, _P, _c_i64, _c_int, _c_int, _c_dbl, _c_int, _c_dbl, _c_int, _c_dbl, _c_dbl,
                     _P, _c_int, _P, _P, _P, _P, _P, ctypes.POINTER(_c_i64), ctypes.POINTER(_c_i64)],
    "pdsb_center": [_P, _P, _P, _P, _P, _c_i64, _c_int, _c_dbl, _c_dbl, _c_dbl, _c_int, _P, _P],
    "pdsb_set_grid_band": [_c_int, _c_int],
    "pdsb_grid_weights_map": [_P, _P, _P, _P, _P, _P, _c_i64, _c_int, _c_int, _c_int, _c_dbl, _P, _P, _c_int, _c_int, _c_int,
                              _c_int, _c_int, _P, _P, ctypes.POINTER(_c_i64)],
    "pdsb_set_grid_reweight": [_P, _c_i64, _P, _c_int],
    "pdsb_sample_image_fft": [_P, _P, _c_int, _c_int, _c_int, _c_dbl, _c_dbl, _c_dbl, _P, _P, _c_int],
    "pdsb_loglike_fft": [_P, _P, _c_int, _c_int, _c_int, _c_dbl, _c_dbl, _c_dbl, _P],
    "pdsb_sample_image_nufft": [_P, _P, _c_int, _c_int, _c_int, _c_int, _c_dbl, _c_dbl, _c_dbl, _P, _P, _c_int],
    "pdsb_loglike_nufft": [_P, _P, _c_int, _c_int, _c_int, _c_int, _c_dbl, _c_dbl, _c_dbl, _P],
    "pdsb_regrid_linear": [_P, _c_i64, _P, _P, _c_i64, _c_int, _c_dbl, _c_int, _P],
    "pdsb_channel_postprocess": [_P, _c_i64, _c_int, _c_int, _c_int, _c_int, _c_int, _P],
    "pdsb_channel_postprocess_scaled": [_P, _c_i64, _c_int, _c_int, _c_int, _c_int, _P, _c_int, _P],
    "pdsb_channel_slice": [_P, _c_int, _c_i64, _c_int, _c_int, _c_int, _P],
    "pdsb_invert_image": [_P, _P, _P, _c_int, _c_int, _c_int, _P],
    "pdsb_mad_std": [_P, _c_i64, _c_int, ctypes.POINTER(_c_dbl)],
    "pdsb_clean_loop": [_P, _P, _c_int, _c_int, _c_int, _c_dbl, _c_dbl, _c_int, _c_dbl, _c_int, _P, _P,
                        ctypes.POINTER(_c_int), ctypes.POINTER(_c_dbl)],
    "pdsb_clean_restore": [_P, _P, _P, _c_int, _c_int, _c_int, _c_int, _P],
    "pdsb_set_dft_variant": [_c_int],
    "pdsb_set_dft_split": [_c_int],
    "pdsb_bench_fma": [_c_int, _c_int, ctypes.POINTER(_c_dbl), ctypes.POINTER(_c_dbl)],
    "pdsb_tc5_accum_probe": [_P, _P, _P, _P, ctypes.c_float, _c_int, _P],
}

_lib = None
_inited = False


def load():
    """dlopen libpdsb.so and declare every prototype of include/pdsb.h.  Does not touch CUDA."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PdsbError(
            "libpdsb.so is not built (%s missing). Build it with `python pdspy_b200/csrc/build.py`; "
            "there is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_char_p if name == "pdsb_last_error" else _c_int
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise PdsbError("libpdsb error %d: %s" % (rc, load().pdsb_last_error().decode(errors="replace")))


def lib():
    """The library with a device initialised (LOCAL_RANK / PDSB_DEVICE select the GPU)."""
    global _inited
    L = load()
    if not _inited:
        dev = int(os.environ.get("PDSB_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        check(L.pdsb_init(dev))
        _inited = True
        kernel = os.environ.get("PDSPY_B200_DFT", "")          # initial DFT kernel: fp32 (default) | tcgen05 | fp64 | <int>
        if kernel:
            variants = {"fp32": 0, "tcgen05": 200, "fp64": 300}
            check(L.pdsb_set_dft_variant(variants[kernel] if kernel in variants else int(kernel)))
    return L


def ptr(a):
    """void* of a numpy array (must be C-contiguous), an int device address, or None."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return ctypes.c_void_p(int(a))
    if isinstance(a, DeviceBuffer):
        return ctypes.c_void_p(a.ptr)
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def host_hash(a):
    """64-bit content hash of a C-contiguous numpy array (pdsb_hash64; no device needed)."""
    a = np.ascontiguousarray(a)
    out = ctypes.c_uint64()
    check(load().pdsb_hash64(a.ctypes.data_as(ctypes.c_void_p), a.nbytes, ctypes.byref(out)))
    return out.value


class DeviceBuffer:
    """A raw device allocation owned by Python (plain bytes; no torch needed)."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        p = ctypes.c_void_p()
        check(lib().pdsb_device_alloc(ctypes.byref(p), self.nbytes))
        self.ptr = p.value or 0

    @classmethod
    def from_numpy(cls, a):
        a = np.ascontiguousarray(a)
        b = cls(a.nbytes)
        b.upload(a)
        return b

    def upload(self, a):
        a = np.ascontiguousarray(a)
        assert a.nbytes <= self.nbytes
        check(lib().pdsb_memcpy(ctypes.c_void_p(self.ptr), DEVICE, ptr(a), HOST, a.nbytes))
        check(lib().pdsb_synchronize())

    def download(self, shape, dtype=np.float64):
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        check(lib().pdsb_memcpy(ptr(out), HOST, ctypes.c_void_p(self.ptr), DEVICE, out.nbytes))
        check(lib().pdsb_synchronize())
        return out

    def free(self):
        if self.ptr and _lib is not None:
            _lib.pdsb_device_free(ctypes.c_void_p(self.ptr))
        self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedArray:
    """numpy view over page-locked host memory (for the end-to-end timed copies)."""

    def __init__(self, shape, dtype=np.float64):
        self.shape = tuple(shape)
        dt = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * dt.itemsize
        p = ctypes.c_void_p()
        check(lib().pdsb_host_alloc_pinned(ctypes.byref(p), max(self.nbytes, 1)))
        self._ptr = p.value
        buf = (ctypes.c_char * self.nbytes).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=dt).reshape(self.shape)

    def free(self):
        if self._ptr and _lib is not None:
            self.array = None
            _lib.pdsb_host_free_pinned(ctypes.c_void_p(self._ptr))
        self._ptr = None
