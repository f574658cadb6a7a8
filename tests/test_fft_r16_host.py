"""The register-resident Stockham passes behind the half-spectrum transforms (pdspy_b200/csrc/fft_r16.cuh) on the CPU:
the header's index arithmetic, twiddles and radix-2/4/8/16 butterflies are __host__ __device__, so g++ builds the
very functions the kernels call (tests/fft_r16_host.cpp emulates the threads of one transform) and numpy.fft judges
them.  No GPU needed."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host_fft():
    out = os.path.join(tempfile.mkdtemp(prefix="r16_"), "libr16host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-o", out, os.path.join(HERE, "fft_r16_host.cpp")])
    lib = ctypes.CDLL(out)
    lib.r16_host_fft.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    return lib


@pytest.mark.parametrize("logn", [8, 9, 10, 11, 12])
def test_stockham_passes_match_numpy(host_fft, logn):
    n = 1 << logn
    rng = np.random.default_rng(logn)
    x = rng.normal(size=n) + 1j * rng.normal(size=n)
    got = np.empty(n, dtype=np.complex128)
    assert host_fft.r16_host_fft(x.ctypes.data, logn, got.ctypes.data) == 0
    ref = np.fft.ifft(x) * n                               # kernel e^{+2 pi i jk/n}
    assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max() * logn
    # a delta at c: a pure exponential, which pins the sign and the output order
    d = np.zeros(n, dtype=np.complex128)
    d[3] = 1.0
    assert host_fft.r16_host_fft(d.ctypes.data, logn, got.ctypes.data) == 0
    np.testing.assert_allclose(got, np.exp(2j * np.pi * 3 * np.arange(n) / n), atol=1e-14)


def test_sizes_outside_the_register_path_are_refused(host_fft):
    x = np.zeros(128, dtype=np.complex128)
    assert host_fft.r16_host_fft(x.ctypes.data, 7, x.ctypes.data) == 1
