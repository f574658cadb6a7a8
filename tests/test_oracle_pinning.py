"""Pin the CPU oracle: against the reference's own compiled module (oracle/_ref, when present)
and against the golden vectors generated from it (tests/golden/, always present)."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden import GRID_CASES_FIXTURE, GRID_CASES_MULTI, multi_channel_set   # noqa: E402

from oracle import grid as og, likelihood as ol                                   # noqa: E402


def _quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


def _check_against_golden(name, out, golden, exact):
    real, imag, weights = out[3], out[4], out[5]
    assert tuple(golden[name + "/shape"]) == real.shape
    np.testing.assert_array_equal(out[2], golden[name + "/freq"])
    np.testing.assert_array_equal(out[0][[0, 1, -1]], golden[name + "/u_ends"])
    for nm, arr in (("real", real), ("imag", imag), ("weights", weights)):
        idx, val = golden["%s/%s_idx" % (name, nm)], golden["%s/%s_val" % (name, nm)]
        assert np.count_nonzero(arr) == int(golden["%s/%s_nnz" % (name, nm)])
        got = arr.reshape(-1)[idx]
        if exact:
            np.testing.assert_array_equal(got, val)
        else:
            scale = np.abs(val).max()
            assert np.abs(got - val).max() <= 1e-10 * scale
        s = float(golden["%s/%s_sum" % (name, nm)])
        assert abs(arr.sum() - s) <= 1e-9 * max(abs(s), np.abs(arr).sum() * 1e-3, 1e-300)


@pytest.mark.parametrize("name", sorted(GRID_CASES_FIXTURE))
def test_grid_oracle_vs_golden_fixture(name, fixture720, grid_golden):
    f = fixture720
    kw = GRID_CASES_FIXTURE[name]
    out = _quiet(og.grid, f["u"], f["v"], f["freq"], f["real"], f["imag"], f["weights"], **kw)
    # pillbox kernel values are 0/1: the restatement is bit-exact; expsinc carries the reference's
    # -ffast-math codegen in its last bits (setup.py:11)
    exact = kw.get("convolution", "pillbox") == "pillbox"
    _check_against_golden(name, out, grid_golden, exact)


@pytest.mark.parametrize("name", sorted(GRID_CASES_MULTI))
def test_grid_oracle_vs_golden_multichannel(name, grid_golden):
    u, v, freq, re, im, w = multi_channel_set()
    kw = GRID_CASES_MULTI[name]
    out = _quiet(og.grid, u, v, freq, re, im, w, **kw)
    exact = kw.get("convolution", "pillbox") == "pillbox"
    _check_against_golden(name, out, grid_golden, exact)


def test_grid_oracle_vs_live_reference(live_ref):
    if live_ref is None:
        pytest.skip("live reference module not available")
    u, v, freq, re, im, w = multi_channel_set()
    data = live_ref.Visibilities(u, v, freq, re, im, w)
    for kw in (dict(gridsize=96, binsize=9000.0, convolution="pillbox", mode="spectralline"),
               dict(gridsize=97, binsize=9000.0, convolution="pillbox", weighting="uniform", npixels=2),
               dict(gridsize=96, binsize=9000.0, convolution="expsinc", weighting="robust", robust=-0.5,
                    mode="spectralline", imaging=True)):
        r = _quiet(live_ref.grid, data, **kw)
        o = _quiet(og.grid, u, v, freq, re, im, w, **kw)
        for a, nm in zip(o[:6], ("u", "v", "freq", "real", "imag", "weights")):
            b = getattr(r, nm)
            if kw["convolution"] == "pillbox" or nm in ("u", "v", "freq"):
                np.testing.assert_array_equal(a, b, err_msg=nm)
            else:
                assert np.abs(a - b).max() <= 1e-10 * np.abs(b).max(), nm


def test_index_maps_match_live_reference_cast_semantics(live_ref):
    """uint32 index maps incl. negative / huge coordinates (numpy's float64->uint32 cast)."""
    u = np.array([-3.7e5, -1.0, 0.0, 0.49999, 1e5, 7e9, -7e15, 3.3e5])
    v = u[::-1].copy()
    i, j = og.index_maps(u, v, np.array([230e9]), 64, 8000.0)
    assert i.dtype == np.uint32
    # trunc toward zero through int64 then wrap: -3.7e5/8000+32 = -14.25 -> -14 -> 2**32-14
    assert i[0, 0] == np.uint32(2 ** 32 - 14)
    assert i[2, 0] == 32 and i[3, 0] == 32


def test_freqcorrect_oracle_vs_live(live_ref):
    if live_ref is None:
        pytest.skip("live reference module not available")
    u, v, freq, re, im, w = multi_channel_set()
    r = live_ref.freqcorrect(live_ref.Visibilities(u, v, freq, re, im, w))
    o = og.freqcorrect(u, v, freq, re, im, w)
    for a, nm in zip(o, ("u", "v", "freq", "real", "imag", "weights")):
        np.testing.assert_array_equal(a, getattr(r, nm), err_msg=nm)


def test_chisq_oracle_vs_golden(fixture720):
    f = fixture720
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chisq_golden.npz"))
    got = ol.chisq_c(f["real"], f["imag"], f["weights"], g["m_real"], g["m_imag"])
    assert got == float(g["chisq"])            # same double accumulation, same float rounding


def test_lnlike_c_vs_verbatim_numpy():
    rng = np.random.default_rng(3)
    n, nf = 3000, 5
    d_re, d_im, m_re, m_im = (rng.normal(size=(n, nf)) for _ in range(4))
    w = rng.uniform(0.5, 2, (n, nf))
    w[rng.random((n, nf)) < 0.02] = 0.0
    w[rng.random((n, nf)) < 0.01] *= -1
    a = ol.lnlike_vis_numpy(d_re, d_im, w, m_re, m_im)
    b = ol.lnlike_vis_c(d_re, d_im, w, m_re, m_im)
    assert abs(a - b) <= 1e-12 * abs(a)


# ---- average() / center() ------------------------------------------
from make_golden import AVERAGE_CASES                                                # noqa: E402
from oracle import average as oa                                                     # noqa: E402


def _avg_inputs():
    u, v, freq, re, im, w = multi_channel_set()
    u[5] = 0.0
    v[5] = 0.0
    return u, v, freq, re, im, w


@pytest.fixture(scope="module")
def average_golden():
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "average_golden.npz"))
    return {k: d[k] for k in d.files}


@pytest.mark.parametrize("name", sorted(AVERAGE_CASES))
def test_average_oracle_vs_live_reference_golden(name, average_golden):
    u, v, freq, re, im, w = _avg_inputs()
    o = _quiet(oa.average, u, v, freq, re, im, w, **AVERAGE_CASES[name])
    for a, nm in zip(o, ("u", "v", "freq", "real", "imag", "weights")):
        np.testing.assert_array_equal(a, average_golden["%s/%s" % (name, nm)], err_msg=nm)


def test_center_oracle_vs_reference_python_golden(average_golden):
    u, v, freq, re, im, w = _avg_inputs()
    cr, ci = oa.center(u, v, freq, re, im, 0.31, -0.17)
    np.testing.assert_array_equal(cr, average_golden["center/real"])
    np.testing.assert_array_equal(ci, average_golden["center/imag"])


@pytest.mark.parametrize("subsample,averaging,hanning", [(1, 1, True), (3, 1, False), (2, 2, True), (1, 4, False), (5, 1, True)])
def test_channel_postprocess_oracle_vs_reference_expressions(subsample, averaging, hanning):
    """oracle/cube.py against the reference's own expression sequence (run_flared_model.py:311-339, with
    scipy.signal.fftconvolve); 1e-13 of the cube maximum (the FFT convolution is not exact)."""
    from oracle import cube as oc
    nfd = 6
    rng = np.random.default_rng(subsample * 10 + averaging)
    img = rng.random((9, 9, nfd * averaging * subsample, 1))
    lit = oc.literal(img, nfd, subsample, averaging, hanning)
    got = oc.post(img[:, :, :, 0], subsample, averaging, hanning)
    assert got.shape == (9, 9, nfd)
    assert np.abs(got - lit[:, :, :, 0]).max() < 1e-13 * img.max()


@pytest.mark.parametrize("name", ["fx_expsinc_64", "fx_pillbox_robust_64", "two_sources_line_32"])
def test_clean_oracle_vs_reference_clean_golden(name):
    """oracle/clean.py:loop against outputs of the reference's own clean.py (scipy fftconvolve, leastsq;
    tests/golden/make_golden.py) from the same dirty image and beam."""
    from make_golden import CLEAN_CASES
    from oracle import clean as ocl
    from pdspy_b200.interferometry.clean import fit_clean_beam
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "clean_golden.npz"))
    kw = CLEAN_CASES[name]
    cb = fit_clean_beam(g[name + "/dirty_beam"])
    ci, res, model, mask, n, thr = ocl.loop(g[name + "/dirty"], g[name + "/dirty_beam"], cb, gain=kw.get("gain", 0.1),
                                            maxiter=kw["maxiter"], nsigma=kw.get("nsigma", 5.))
    assert n > 5 and np.count_nonzero(model) >= 2
    assert np.array_equal(mask, g[name + "/mask"])
    peak = np.abs(g[name + "/clean_image"]).max()
    for got, key in ((ci, "clean_image"), (res, "residuals"), (model, "model")):
        assert np.abs(got - g[name + "/" + key]).max() < 1e-14 * peak, key
