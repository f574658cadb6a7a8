"""GPU parity: clean() (pdspy/interferometry/clean.py) through the C-ABI, against the oracle and against
outputs of the reference's own clean.py (tests/golden/clean_golden.npz)."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import CLEAN_CASES, two_source_set                     # noqa: E402
from oracle import clean as ocl                                         # noqa: E402
from pdspy_b200.interferometry import Visibilities                      # noqa: E402
from pdspy_b200.interferometry.clean import clean, clean_arrays, fit_clean_beam, mad_std   # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(HERE, "golden", "clean_golden.npz"))


@pytest.mark.parametrize("n", [1, 2, 7, 1000, 4097, 65536, 300001])
def test_mad_std_is_exact(gpu, n):
    rng = np.random.default_rng(n)
    x = rng.normal(size=n)
    x[rng.random(n) < 0.3] = 0.0                     # ties and +-0, as in a dirty image with blanked pixels
    x[rng.random(n) < 0.05] *= -0.0
    x[: n // 10] = x[n // 2]                          # more duplicates
    assert mad_std(x) == ocl.mad_std(x)
    assert mad_std(np.full(n, 3.25)) == 0.0


@pytest.mark.parametrize("name", list(CLEAN_CASES))
def test_loop_on_reference_inputs(gpu, name):
    """Same dirty image and beam as the reference's run: mask identical, components at the same pixels,
    everything else to 1e-12 of the peak (the reference subtracts through an FFT convolution)."""
    kw = CLEAN_CASES[name]
    dirty, beam = GOLD[name + "/dirty"], GOLD[name + "/dirty_beam"]
    cb = fit_clean_beam(beam)
    ci, res, model, mask, n, thr = clean_arrays(dirty, beam, cb, gain=kw.get("gain", 0.1), maxiter=kw["maxiter"],
                                                nsigma=kw.get("nsigma", 5.))
    o = ocl.loop(dirty, beam, cb, gain=kw.get("gain", 0.1), maxiter=kw["maxiter"], nsigma=kw.get("nsigma", 5.))
    assert n == o[4] and abs(thr - o[5]) <= 1e-15 * abs(o[5])
    assert np.array_equal(mask, GOLD[name + "/mask"])
    assert np.array_equal(model != 0, GOLD[name + "/model"] != 0)
    peak = np.abs(GOLD[name + "/clean_image"]).max()
    for got, key in ((ci, "clean_image"), (res, "residuals"), (model, "model")):
        assert np.abs(got - GOLD[name + "/" + key]).max() < 1e-12 * peak, key


def test_maxiter_zero_and_blank_image(gpu):
    dirty, beam = GOLD["fx_expsinc_64/dirty"], GOLD["fx_expsinc_64/dirty_beam"]
    cb = fit_clean_beam(beam)
    ci, res, model, mask, n, thr = clean_arrays(dirty, beam, cb, maxiter=0)
    assert n == 0 and not model.any() and np.array_equal(res, dirty) and np.array_equal(ci, dirty)
    z = np.zeros_like(dirty)
    ci, res, model, mask, n, thr = clean_arrays(z, beam, cb, maxiter=10)
    assert n == 0 and not ci.any() and not mask.any()


@pytest.mark.parametrize("name", ["fx_expsinc_64", "two_sources_line_32"])
def test_clean_end_to_end_vs_reference(gpu, fixture720, name):
    """The whole call (GPU invert of image and beam, leastsq beam fit, GPU loop and restore) against the
    reference's clean() on the same visibilities."""
    kw = CLEAN_CASES[name]
    if name.startswith("two"):
        u, v, freq, re, im, w = two_source_set()
    else:
        f = fixture720
        u, v, freq, re, im, w = (f[k].copy() for k in ("u", "v", "freq", "real", "imag", "weights"))
    data = Visibilities(u, v, freq, re, im, w)
    with contextlib.redirect_stdout(io.StringIO()):
        ci, res, beam, model, mask = clean(data, **kw)
    assert ci.image.shape == GOLD[name + "/clean_image"].shape + (1,)
    peak = np.abs(GOLD[name + "/clean_image"]).max()
    assert np.array_equal(mask.image[:, :, :, 0], GOLD[name + "/mask"])
    assert np.abs(beam.image[:, :, :, 0] - GOLD[name + "/dirty_beam"]).max() < 1e-9
    for got, key in ((ci, "clean_image"), (res, "residuals"), (model, "model")):
        assert np.abs(got.image[:, :, :, 0] - GOLD[name + "/" + key]).max() < 1e-8 * peak, key
