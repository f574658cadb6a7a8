"""Runtime pieces of the C-ABI on a real device: buffers, timers, per-kernel profile, launch
counter, error reporting."""
import ctypes

import numpy as np
import pytest

import synth
from pdspy_b200 import _lib, DeviceBuffer, PinnedArray, Dataset

pytestmark = pytest.mark.gpu


def test_device_info_is_a_b200(gpu):
    sm, khz, mem, maj, mnr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int64(), ctypes.c_int(), ctypes.c_int()
    _lib.check(gpu.pdsb_device_info(ctypes.byref(sm), ctypes.byref(khz), ctypes.byref(mem), ctypes.byref(maj),
                                    ctypes.byref(mnr)))
    assert maj.value == 10 and sm.value >= 100


def test_buffers_round_trip_and_pinned(gpu):
    a = np.random.default_rng(0).normal(size=1000)
    b = DeviceBuffer.from_numpy(a)
    np.testing.assert_array_equal(b.download(a.shape), a)
    p = PinnedArray((1000,))
    p.array[:] = a
    b2 = DeviceBuffer(a.nbytes)
    _lib.check(gpu.pdsb_memcpy(_lib.ptr(b2), _lib.DEVICE, _lib.ptr(p.array), _lib.HOST, a.nbytes))
    _lib.check(gpu.pdsb_synchronize())
    np.testing.assert_array_equal(b2.download(a.shape), a)
    p.free()


def test_profile_and_launch_count(gpu):
    n0 = ctypes.c_int64()
    gpu.pdsb_launch_count(ctypes.byref(n0))
    _lib.check(gpu.pdsb_profile_reset())
    _lib.check(gpu.pdsb_profile_enable(1))
    c = synth.make_config("C1", nuv=4096)
    from pdspy_b200.interferometry import interpolate_model
    _lib.check(gpu.pdsb_set_dft_variant(1))        # the FP32-pipe kernel: the launch list below is its
    try:
        interpolate_model(c["u"], c["v"], c["freq"], c["model"])
    finally:
        _lib.check(gpu.pdsb_set_dft_variant(0))
    ms, cnt = ctypes.c_double(), ctypes.c_int64()
    _lib.check(gpu.pdsb_profile_get(b"dft_", ctypes.byref(ms), ctypes.byref(cnt)))
    assert cnt.value == 1 and ms.value > 0
    _lib.check(gpu.pdsb_profile_get(b"", ctypes.byref(ms), ctypes.byref(cnt)))
    assert cnt.value == 3                       # fold, dft, finish
    _lib.check(gpu.pdsb_profile_enable(0))
    n1 = ctypes.c_int64()
    gpu.pdsb_launch_count(ctypes.byref(n1))
    assert n1.value - n0.value == 3


def test_timer(gpu):
    _lib.check(gpu.pdsb_timer_start())
    tf, ms = ctypes.c_double(), ctypes.c_double()
    _lib.check(gpu.pdsb_bench_fma(1, 2000, ctypes.byref(tf), ctypes.byref(ms)))
    el = ctypes.c_double()
    _lib.check(gpu.pdsb_timer_stop(ctypes.byref(el)))
    assert el.value >= ms.value > 0 and tf.value > 10


def test_errors_surface_as_exceptions(gpu):
    ds = Dataset(np.zeros(4), np.ones(4))
    img = np.zeros((8, 8, 2))
    out = np.zeros(2)
    with pytest.raises(_lib.PdsbError, match="no data"):
        _lib.check(gpu.pdsb_loglike(ds.handle, _lib.ptr(img), 8, 8, 2, _lib.HOST, 1e-7, 0.0, 0.0, _lib.ptr(out),
                                    _lib.ptr(out)))
    ds.set_data(np.zeros((4, 3)), np.zeros((4, 3)), np.ones((4, 3)))
    with pytest.raises(_lib.PdsbError, match="channel count"):
        _lib.check(gpu.pdsb_loglike(ds.handle, _lib.ptr(img), 8, 8, 2, _lib.HOST, 1e-7, 0.0, 0.0, _lib.ptr(out),
                                    _lib.ptr(out)))
    with pytest.raises(_lib.PdsbError):
        _lib.check(gpu.pdsb_sample_image(ds.handle, _lib.ptr(img), 8, 8, 2, _lib.HOST, -1.0, 0.0, 0.0, _lib.ptr(out),
                                         _lib.ptr(out), _lib.HOST))


def test_staged_copies_of_large_pageable_arrays(gpu):
    """Host arrays of 16 MB and more travel through a ring of pinned buffers filled by several host threads
    (runtime.cu: copy_h2d / copy_d2h); the bytes must arrive unchanged, for sizes that are not a multiple of the
    4 MB chunk and for pinned memory (which goes straight through)."""
    rng = np.random.default_rng(11)
    for n in (2 * 1024 * 1024 + 7, 9 * 1024 * 1024 + 12345, 40 * 1024 * 1024 + 1):     # doubles: 16 MB+, 72 MB+, 320 MB+
        a = rng.random(n)
        d = DeviceBuffer.from_numpy(a)
        back = np.empty_like(a)
        _lib.check(gpu.pdsb_memcpy(_lib.ptr(back), _lib.HOST, _lib.ptr(d), _lib.DEVICE, a.nbytes))
        _lib.check(gpu.pdsb_synchronize())
        assert np.array_equal(a, back), n
        d.free()
    p = PinnedArray((3 * 1024 * 1024,))                  # 24 MB of pinned memory: not staged
    p.array[:] = rng.random(p.array.size)
    d = DeviceBuffer(p.array.nbytes)
    _lib.check(gpu.pdsb_memcpy(_lib.ptr(d), _lib.DEVICE, _lib.ptr(p.array), _lib.HOST, p.array.nbytes))
    _lib.check(gpu.pdsb_synchronize())
    np.testing.assert_array_equal(d.download(p.array.shape), p.array)
    p.free()
    d.free()
