"""invert() on the GPU against images produced by the reference's own
invert.py (scipy.fftpack + compiled grid), incl. the reference's smoke test tests/test.py:9.
Tolerance 1e-10 of the image peak: fp64 FFT of a different factorisation + expsinc last bits."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden import INVERT_CASES, INVERT_ANYSIZE_CASES            # noqa: E402

from pdspy_b200.interferometry import invert, Visibilities        # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "invert_golden.npz")
GOLD_ANY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "invert_anysize_golden.npz")
ALL_CASES = dict(INVERT_CASES, **INVERT_ANYSIZE_CASES)


@pytest.mark.parametrize("name", sorted(ALL_CASES))
@pytest.mark.parametrize("deterministic", [True, False])
def test_invert_vs_reference_images(gpu, fixture720, name, deterministic):
    """Power-of-two sizes (radix-2 rows) and the sizes the reference's scipy ifft2 also takes - even and odd,
    not powers of two (Bluestein rows): invert.py:65-84."""
    f = fixture720
    data = Visibilities(f["u"], f["v"], f["freq"], f["real"].copy(), f["imag"].copy(), f["weights"].copy())
    g = np.load(GOLD if name in INVERT_CASES else GOLD_ANY)
    with contextlib.redirect_stdout(io.StringIO()):
        r = invert(data, deterministic=deterministic, **ALL_CASES[name])
    im = r.image[:, :, 0, 0]
    sample = im[::4, ::4] if im.shape[0] > 128 else im
    peak = np.abs(g[name + "/sample"]).max()
    assert np.abs(sample - g[name + "/sample"]).max() <= 1e-10 * peak
    stats = np.array([im.sum(), im.max(), im.min(), np.abs(im).sum()])
    assert np.all(np.abs(stats - g[name + "/stats"]) <= 1e-9 * np.abs(g[name + "/stats"]).max())
    np.testing.assert_allclose(r.x, g[name + "/x"], rtol=1e-15)
    # beam=True / uvtaper mutate the data in place and must restore it (invert.py:40-47)
    np.testing.assert_array_equal(data.real, f["real"])
    np.testing.assert_array_equal(data.weights, f["weights"])


def test_invert_size_limit(gpu, fixture720):
    f = fixture720
    data = Visibilities(f["u"], f["v"], f["freq"], f["real"], f["imag"], f["weights"])
    with pytest.raises(NotImplementedError):
        invert(data, imsize=5000)


def test_bluestein_rows_against_numpy(gpu):
    """The chirp-z row transform against numpy's ifft2 on random complex maps of awkward sizes (prime, odd,
    2 x prime, 1000); a centred delta as the kernel map makes the gridding correction 1."""
    import ctypes
    from pdspy_b200 import _lib
    rng = np.random.default_rng(4)
    for n in (7, 97, 251, 502, 1000):
        re, im = rng.normal(size=(n * n, 1)), rng.normal(size=(n * n, 1))
        conv = np.zeros((n, n))
        conv[n // 2, n // 2] = n * n                                  # ifft2 of a centred delta: 1 everywhere
        out = np.empty((n, n, 1, 1))
        _lib.check(gpu.pdsb_invert_image(_lib.ptr(re), _lib.ptr(im), _lib.ptr(conv), n, 1, _lib.HOST, _lib.ptr(out)))
        comp = (re + 1j * im).reshape(n, n)
        ref = (np.fft.fftshift(np.fft.ifft2(np.fft.ifftshift(comp))).real * n * n)[:, ::-1]
        assert np.abs(out[:, :, 0, 0] - ref).max() <= 1e-11 * np.abs(ref).max(), n
