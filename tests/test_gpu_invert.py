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
from make_golden import INVERT_CASES            # noqa: E402

from pdspy_b200.interferometry import invert, Visibilities        # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "invert_golden.npz")


@pytest.mark.parametrize("name", sorted(INVERT_CASES))
@pytest.mark.parametrize("deterministic", [True, False])
def test_invert_vs_reference_images(gpu, fixture720, name, deterministic):
    f = fixture720
    data = Visibilities(f["u"], f["v"], f["freq"], f["real"].copy(), f["imag"].copy(), f["weights"].copy())
    g = np.load(GOLD)
    with contextlib.redirect_stdout(io.StringIO()):
        r = invert(data, deterministic=deterministic, **INVERT_CASES[name])
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


def test_invert_rejects_non_power_of_two(gpu, fixture720):
    f = fixture720
    data = Visibilities(f["u"], f["v"], f["freq"], f["real"], f["imag"], f["weights"])
    with pytest.raises(NotImplementedError):
        invert(data, imsize=300)
