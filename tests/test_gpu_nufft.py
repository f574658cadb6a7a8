"""interpolate_model(code="nufft") / loglike_image_nufft: the exact transform through a type-2 non-uniform FFT
(vis.cu: 8-point exponential-of-semicircle kernel, image zero-padded to 2n).  It must meet the tolerances of the direct
transform - visibilities 1e-5 of max|V|, chi^2 1e-7 (BASELINE.json north_star) - against the SAME exact fp64 oracle;
what it actually delivers (4e-8 / 2e-10) is asserted too, so that a wrong kernel width or deapodisation shows."""
import numpy as np
import pytest

from oracle import dft as od, likelihood as ol
import synth
from pdspy_b200 import _lib
from pdspy_b200.interferometry import interpolate_model, loglike_image_nufft, loglike_image, Visibilities

pytestmark = pytest.mark.gpu
A = synth.ARCSEC


def relerr(vis, ref):
    got = vis.real + 1j * vis.imag
    scale = np.abs(ref).max(axis=0, keepdims=True)
    return (np.abs(got - ref) / scale).max()


@pytest.mark.parametrize("n,nf,nuv,herm,kind", [
    (64, 2, 600, True, "random"),        # Hermitian-doubled list: one evaluation per pair
    (64, 3, 601, False, "random"),       # odd count, odd channel count (a channel pair without a partner)
    (4, 1, 50, False, "random"),         # the smallest side: every tap wraps around the 8-point grid
    (8, 5, 80, True, "random"),
    (16, 40, 100, False, "random"),      # more channels than a lane group
    (128, 1, 2000, True, "disk"),
    (256, 7, 1500, False, "disk"),
    (1024, 1, 400, True, "disk"),        # transform length 2048
    (2048, 1, 200, False, "disk"),       # transform length 4096: the largest
    (2, 1, 30, False, "random"),         # grid of 4 cells: every tap index occurs twice
    (2, 32, 64, True, "random"),         # the same through the shared-memory-staged sampler (>= 16 channels)
    (4, 17, 70, False, "random"),
    (300, 2, 700, True, "disk"),         # sides that are not powers of two: embedded about the centre pixel
    (90, 3, 300, False, "random"),       # n / 2 odd: the row pair straddles the origin of the transform
    ((96, 250), 2, 400, False, "random"),    # rectangular, both ways
    ((250, 96), 1, 400, True, "random"),
])
def test_nufft_vs_exact_oracle(gpu, n, nf, nuv, herm, kind):
    px = 0.05
    if isinstance(n, tuple):
        img = np.random.default_rng(n[0]).random((n[0], n[1], nf, 1))
        n = max(n)
    else:
        img = synth.synth_image(n, nf, px, kind=kind)
    m = synth.SynthImage(img, px, synth.synth_freq(nf))
    if herm:
        u, v = synth.synth_uv(nuv, px * A)
    else:                                   # the whole Nyquist square, both signs of u: mirrored and wrapped taps
        rng = np.random.default_rng(n + nf)
        lim = 0.5 / (px * A)
        u, v = rng.uniform(-lim, lim, nuv), rng.uniform(-lim, lim, nuv)
        u[:4] = [0.0, lim * (1 - 1e-12), -lim * (1 - 1e-12), 1e-3 * lim]
        v[:4] = [0.0, -lim * (1 - 1e-12), lim * (1 - 1e-12), -1e-3 * lim]
        if nuv > 8:                         # beyond the image's Nyquist frequency: the sum is periodic in u dxy, so is the torus
            u[4:8] = [1.6 * lim, -2.3 * lim, 0.2 * lim, 7.9 * lim]
            v[4:8] = [0.3 * lim, 1.1 * lim, -3.7 * lim, -7.2 * lim]
    ref = od.exact_dft(u, v, m.image, px * A, 0.04 * A, -0.03 * A)
    vis = interpolate_model(u, v, m.freq, m, dRA=0.04, dDec=-0.03, code="nufft")
    assert vis.real.shape == (nuv, nf) and np.all(vis.weights == 1)
    err = relerr(vis, ref)
    assert err < 1e-5, err                  # the north star's bound
    assert err < 3e-7, err                  # what an 8-point kernel at oversampling 2 delivers


def test_nufft_likelihood_vs_oracle_chain_and_direct_transform(gpu):
    n, px, nf = 128, 0.05, 3
    img = synth.synth_image(n, nf, px, kind="disk")
    m = synth.SynthImage(img, px, synth.synth_freq(nf))
    u, v = synth.synth_uv(6000, px * A)
    ref = od.exact_dft(u, v, m.image, px * A, 0.01 * A, 0.02 * A)
    for model in ((ref.real, ref.imag), None):           # data around the model, and pure noise
        re, im, w = synth.synth_data(6000, nf, model=model)
        data = Visibilities(u, v, m.freq, re, im, w)
        ll, c_re, c_im = loglike_image_nufft(data, m, dRA=0.01, dDec=0.02)
        ll_ref = ol.lnlike_vis_numpy(re, im, w, ref.real, ref.imag)
        assert abs(ll - ll_ref) <= 1e-7 * abs(ll_ref)
        assert abs(ll - ll_ref) <= 2e-9 * abs(ll_ref)
        c_ref = ((re - ref.real) ** 2 * w).sum()
        assert abs(c_re - c_ref) <= 1e-7 * c_ref
        ll_direct, _ = loglike_image(data, m, dRA=0.01, dDec=0.02)
        assert abs(ll - ll_direct) <= 1e-7 * abs(ll_direct)
    # non-Hermitian list: no twin rows
    rng = np.random.default_rng(3)
    lim = 0.45 / (px * A)
    u2, v2 = rng.uniform(-lim, lim, 777), rng.uniform(-lim, lim, 777)
    re, im, w = (rng.normal(size=(777, nf)) for _ in range(3))
    w = np.abs(w)
    data = Visibilities(u2, v2, m.freq, re, im, w)
    ref2 = od.exact_dft(u2, v2, m.image, px * A, 0.0, 0.0)
    ll, _, _ = loglike_image_nufft(data, m)
    ll_ref = ol.lnlike_vis_numpy(re, im, w, ref2.real, ref2.imag)
    assert abs(ll - ll_ref) <= 2e-9 * abs(ll_ref)


def test_nufft_rejects_what_it_cannot_do(gpu):
    u, v = synth.synth_uv(10, 0.05 * A)
    odd = synth.SynthImage(np.random.default_rng(0).random((33, 64, 1, 1)), 0.05, synth.synth_freq(1))
    with pytest.raises(ValueError):                                  # odd side: the direct transform's job
        interpolate_model(u, v, odd.freq, odd, code="nufft")
    m = synth.SynthImage(synth.synth_image(48, 1, 0.05, kind="random"), 0.05, synth.synth_freq(1))
    out = np.empty((10, 1))
    with pytest.raises(_lib.PdsbError):                              # the C-ABI refuses it as well
        from pdspy_b200.device import dataset_for
        _lib.check(gpu.pdsb_sample_image_nufft(dataset_for(u, v).handle, _lib.ptr(np.zeros((33, 64, 1))), 33, 64, 1,
                                               _lib.HOST, 1e-7, 0.0, 0.0, _lib.ptr(out), _lib.ptr(out.copy()), _lib.HOST))
    e = interpolate_model(np.zeros(0), np.zeros(0), m.freq, synth.SynthImage(synth.synth_image(64, 2, 0.05), 0.05,
                                                                         synth.synth_freq(2)), code="nufft")
    assert e.real.shape == (0, 2)


@pytest.mark.parametrize("code", ["nufft", "galario-fft"])
def test_empty_channels_are_exactly_zero(gpu, code):
    """Two channels share one complex transform in the FFT-based paths; an all-zero channel must still come out as exact
    zeros (as it does from the direct-sum kernels), whether it is the first or the second of its pair, and its partner
    must be what it is next to a non-empty channel."""
    n, px, nf = 64, 0.05, 6
    img = synth.synth_image(n, nf, px, kind="random")
    img[n // 2, n // 2, :, 0] += 1e5                       # a bright pixel: the partner's rounding residue would be ~1e-11
    full = synth.SynthImage(img.copy(), px, synth.synth_freq(nf))
    img[:, :, 1, 0] = 0.0                                  # second of pair (0, 1)
    img[:, :, 4, 0] = 0.0                                  # first of pair (4, 5)
    m = synth.SynthImage(img, px, synth.synth_freq(nf))
    u, v = synth.synth_uv(500, px * A)
    a = interpolate_model(u, v, m.freq, m, dRA=0.02, dDec=-0.01, code=code)
    b = interpolate_model(u, v, full.freq, full, dRA=0.02, dDec=-0.01, code=code)
    va, vb = a.real + 1j * a.imag, b.real + 1j * b.imag
    assert np.all(va[:, 1] == 0.0) and np.all(va[:, 4] == 0.0)
    for i in (0, 2, 3, 5):
        assert np.abs(va[:, i] - vb[:, i]).max() <= 1e-13 * np.abs(vb[:, i]).max(), i
