"""GPU parity: channel post-processing of a model cube (run_flared_model.py:308-366) through the C-ABI."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cube as oc, dft as od                                # noqa: E402
import synth                                            # noqa: E402
from pdspy_b200.interferometry import postprocess_channels, model_visibilities, loglike_image, Visibilities  # noqa: E402

pytestmark = pytest.mark.gpu
A = synth.ARCSEC
TOL = 1e-14        # relative to the cube maximum: same sums in the same order, up to the last bit of the taps


@pytest.mark.parametrize("subsample,averaging,hanning", [(1, 1, True), (3, 1, False), (2, 2, True), (1, 4, False),
                                                          (5, 1, True), (1, 1, False), (4, 3, True)])
def test_regular_cube(gpu, subsample, averaging, hanning):
    nfd = 7
    rng = np.random.default_rng(subsample * 100 + averaging)
    img = rng.random((33, 31, nfd * averaging * subsample, 1)) * 3.0
    got = postprocess_channels(img, subsample, averaging, hanning)
    ref = oc.post(img[:, :, :, 0], subsample, averaging, hanning)
    assert got.shape == (33, 31, nfd, 1)
    assert np.abs(got[:, :, :, 0] - ref).max() <= TOL * img.max()
    if not hanning:                                  # plain block means: bit-exact
        assert np.array_equal(got[:, :, :, 0], ref)


def test_unstructured_image_and_edges(gpu):
    rng = np.random.default_rng(5)
    img = rng.random((1000, 12))
    got = postprocess_channels(img, 2, 3, True)
    assert got.shape == (1000, 2)
    assert np.abs(got - oc.post(img, 2, 3, True)).max() <= TOL
    one = postprocess_channels(rng.random((4, 4, 1, 1)), 1, 1, True)        # a single channel: only the centre tap
    assert one.shape == (4, 4, 1, 1)
    assert postprocess_channels(np.empty((0, 6)), 2, 1, False).shape == (0, 3)
    with pytest.raises(ValueError):
        postprocess_channels(img, 5, 1, False)                               # 12 is not a multiple of 5
    with pytest.raises(ValueError):
        postprocess_channels(np.zeros((3, 3, 3)), 1, 1, False)


def test_matches_reference_expression_sequence(gpu):
    """Against the reference's own numpy/scipy expressions (fftconvolve is only ~1e-16 accurate)."""
    rng = np.random.default_rng(11)
    img = rng.random((16, 16, 24, 1))
    lit = oc.literal(img, 6, 2, 2, True)
    got = postprocess_channels(img, 2, 2, True)
    assert np.abs(got - lit).max() < 1e-13


def test_chained_into_the_transform_and_the_likelihood(gpu):
    """model_visibilities / loglike_image with subsample / averaging / hanning == post-process on the host,
    then the plain call (the transform is linear in the cube)."""
    nfd, sub, avg = 3, 2, 2
    rng = np.random.default_rng(3)
    img = rng.random((48, 48, nfd * sub * avg, 1))
    m = synth.SynthImage(img, 0.05, synth.synth_freq(nfd * sub * avg))
    u, v = synth.synth_uv(600, 0.05 * A)
    freq = synth.synth_freq(nfd)
    vis = model_visibilities(u, v, freq, m, dRA=0.03, dDec=-0.02, subsample=sub, averaging=avg, hanning=True)
    post = oc.post(img[:, :, :, 0], sub, avg, True)[:, :, :, None]
    ref = od.exact_dft(u, v, post, 0.05 * A, 0.03 * A, -0.02 * A)
    assert vis.real.shape == (600, nfd)
    assert np.abs(vis.real + 1j * vis.imag - ref).max() / np.abs(ref).max() < 1e-5
    w = rng.uniform(0.5, 2.0, (600, nfd))
    data = Visibilities(u, v, freq, ref.real + rng.normal(0, 0.1, ref.shape), ref.imag + rng.normal(0, 0.1, ref.shape), w)
    ll, chi2 = loglike_image(data, m, dRA=0.03, dDec=-0.02, subsample=sub, averaging=avg, hanning=True)
    mp = synth.SynthImage(np.ascontiguousarray(post), 0.05, freq)
    ll0, chi20 = loglike_image(data, mp, dRA=0.03, dDec=-0.02)
    assert abs(ll - ll0) <= 1e-12 * abs(ll0) and np.allclose(chi2, chi20, rtol=1e-12)


@pytest.mark.parametrize("sub,avg,hanning", [(1, 1, True), (2, 2, True), (3, 1, False)])
def test_extinction_goes_before_the_post_processing(gpu, sub, avg, hanning):
    """The reference multiplies the rendered channels by extinction[i] (run_flared_model.py:286-299) and by
    flux_unc (:303) BEFORE the sub-sample mean / Hanning smoothing / binning (:308-366).  The two do not commute
    (with hanning and subsample = averaging = 1 the channel counts are even equal), so the chained calls must
    follow that order: compare with the literal sequence on the host."""
    nfd = 4
    nf_in = nfd * sub * avg
    rng = np.random.default_rng(sub * 10 + avg)
    img = rng.random((40, 40, nf_in, 1))
    ext = np.exp(-rng.uniform(0.0, 2.0, nf_in))                   # exp(-tau) per rendered channel
    flux_unc = 1.07
    m = synth.SynthImage(img, 0.05, synth.synth_freq(nf_in))
    u, v = synth.synth_uv(500, 0.05 * A)
    freq = synth.synth_freq(nfd)
    lit = img * ext[None, None, :, None]                          # :286-299
    lit = lit * flux_unc                                          # :303
    post = oc.post(lit[:, :, :, 0], sub, avg, hanning)[:, :, :, None]
    ref = od.exact_dft(u, v, post, 0.05 * A, 0.01 * A, 0.02 * A)
    vis = model_visibilities(u, v, freq, m, dRA=0.01, dDec=0.02, flux_unc=flux_unc, extinction=ext, subsample=sub,
                             averaging=avg, hanning=hanning)
    assert np.abs(vis.real + 1j * vis.imag - ref).max() / np.abs(ref).max() < 1e-5
    wrong = od.exact_dft(u, v, oc.post(img[:, :, :, 0], sub, avg, hanning)[:, :, :, None] * flux_unc, 0.05 * A, 0.01 * A,
                         0.02 * A)
    if nf_in == nfd:                                              # what round 1 computed: visibly different
        wrong = wrong * ext[None, :]
        assert np.abs(wrong - ref).max() / np.abs(ref).max() > 1e-3
    w = rng.uniform(0.5, 2.0, (500, nfd))
    data = Visibilities(u, v, freq, ref.real + rng.normal(0, 0.1, ref.shape), ref.imag + rng.normal(0, 0.1, ref.shape), w)
    ll, _ = loglike_image(data, m, dRA=0.01, dDec=0.02, flux_unc=flux_unc, extinction=ext, subsample=sub, averaging=avg,
                          hanning=hanning)
    ll0, _ = loglike_image(data, synth.SynthImage(np.ascontiguousarray(post), 0.05, freq), dRA=0.01, dDec=0.02)
    # (flux_unc multiplies the fp32-folded result on one side and the fp64 cube on the other: 1e-8, bound 1e-7)
    assert abs(ll - ll0) <= 1e-7 * abs(ll0)
    got = postprocess_channels(img, sub, avg, hanning, in_scale=ext)
    assert np.abs(got[:, :, :, 0] - oc.post((img * ext[None, None, :, None])[:, :, :, 0], sub, avg, hanning)).max() <= TOL
    if sub * avg > 1:
        with pytest.raises(ValueError):                           # a factor per OUTPUT channel is not what the reference has
            model_visibilities(u, v, freq, m, extinction=ext[:nfd], subsample=sub, averaging=avg, hanning=hanning)
