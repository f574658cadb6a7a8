"""average() and center() on the GPU against the live reference's
golden vectors and the oracle.  average(): every output bit-exact (ordered sums in the reference's
loop order).  center(): 1e-14 relative (the reference's phase factor comes from libm's complex exp,
the device's from CUDA sincos: last-ulp differences)."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden import AVERAGE_CASES, multi_channel_set           # noqa: E402

from oracle import average as oa                                   # noqa: E402
import synth                                       # noqa: E402
from pdspy_b200.interferometry import average, center, Visibilities   # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "average_golden.npz")


def _quiet(fn, *a, **kw):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        r = fn(*a, **kw)
    return r, buf.getvalue()


def _data():
    u, v, freq, re, im, w = multi_channel_set()
    u[5] = 0.0
    v[5] = 0.0
    return Visibilities(u, v, freq, re, im, w)


@pytest.mark.parametrize("name", sorted(AVERAGE_CASES))
def test_average_bit_exact_vs_live_reference_golden(gpu, name):
    g = np.load(GOLD)
    a, out = _quiet(average, _data(), **AVERAGE_CASES[name])
    for nm in ("u", "v", "freq", "real", "imag", "weights"):
        np.testing.assert_array_equal(getattr(a, nm), g["%s/%s" % (name, nm)], err_msg=nm)


def test_average_large_radial_profile_vs_oracle(gpu):
    """load_data.py:87-96 usage: 1-D log-radial profile of a big continuum data set."""
    u, v = synth.synth_uv(400_000, 0.01 * synth.ARCSEC)
    re, im, w = synth.synth_data(400_000, 1)
    d = Visibilities(u, v, synth.synth_freq(1), re, im, w)
    kw = dict(gridsize=40, radial=True, log=True, logmin=d.uvdist.min() * 0.95, logmax=d.uvdist.max() * 1.05)
    a, _ = _quiet(average, d, **kw)
    o, _ = _quiet(oa.average, u, v, d.freq, re, im, w, **kw)
    for got, exp in zip((a.u, a.v, a.freq, a.real, a.imag, a.weights), o):
        np.testing.assert_array_equal(got, exp)


def test_average_warning_and_empty(gpu):
    d = _data()
    a, out = _quiet(average, d, gridsize=8, binsize=8000.0)
    assert out.startswith("WARNING")
    e = Visibilities(np.zeros(0), np.zeros(0), np.array([230e9]), np.zeros((0, 1)), np.zeros((0, 1)), np.zeros((0, 1)))
    a, out = _quiet(average, e, gridsize=8, binsize=8000.0)
    assert a.real.shape == (0, 1)


def test_center_vs_reference_python_golden(gpu):
    g = np.load(GOLD)
    c = center(_data(), [0.31, -0.17, 1.0])
    for got, exp in ((c.real, g["center/real"]), (c.imag, g["center/imag"])):
        assert np.abs(got - exp).max() <= 1e-14 * np.abs(exp).max()
    d = _data()
    np.testing.assert_array_equal(c.u, d.u)
    np.testing.assert_array_equal(c.weights, d.weights)


def test_center_commutes_with_interpolate_model_shift(gpu):
    """center(data, [x0, y0]) undoes interpolate_model(dRA=x0, dDec=y0) (SURVEY.md section 8c pin 2),
    up to the reference's truncated pi (3.14159) in point_model: |delta phase| <= 2.7e-6 * 2 pi u x0."""
    from pdspy_b200.interferometry import interpolate_model
    c1 = synth.make_config("C1", nuv=4000)
    x0, y0 = 0.05, -0.03
    shifted = interpolate_model(c1["u"], c1["v"], c1["freq"], c1["model"], dRA=x0, dDec=y0)
    plain = interpolate_model(c1["u"], c1["v"], c1["freq"], c1["model"])
    back = center(shifted, [x0, y0, 1.0])
    err = np.abs((back.real + 1j * back.imag) - (plain.real + 1j * plain.imag)).max() / np.abs(plain.real).max()
    assert err < 2e-5
