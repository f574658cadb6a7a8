"""Parity of the CUDA image->visibility path (through the C-ABI) against the exact-DFT oracle.

Tolerance (BASELINE.json north_star): visibilities within 1e-5 relative - fp32 products with
fp64 phase seeding and accumulation - measured against max|V| of the channel, because individual
|V| pass through zero at nulls (hard part 2)."""
import ctypes
import os

import numpy as np
import pytest

from oracle import dft as od
import synth
from pdspy_b200 import _lib, Dataset, DeviceBuffer
from pdspy_b200.interferometry import interpolate_model, Visibilities

pytestmark = pytest.mark.gpu
A = synth.ARCSEC
TOL = 1e-5


def relerr(vis, ref):
    got = vis.real + 1j * vis.imag
    scale = np.abs(ref).max(axis=0, keepdims=True)
    return (np.abs(got - ref) / scale).max()


def _model(ny, nx, nf, px, seed=0, kind="random"):
    rng = np.random.default_rng(seed)
    img = rng.random((ny, nx, nf, 1)) if kind == "random" else synth.synth_image(nx, nf, px)
    m = synth.SynthImage(img, px, synth.synth_freq(nf))
    return m


@pytest.fixture(autouse=True)
def _reset_tuning(gpu):
    yield
    gpu.pdsb_set_dft_variant(0)
    gpu.pdsb_set_dft_split(0)


@pytest.mark.parametrize("ny,nx,nf,nuv,herm", [
    (64, 64, 1, 400, True),       # square even, Hermitian-doubled list (the reference's case)
    (64, 64, 3, 401, False),      # odd count: cannot be Hermitian-doubled
    (63, 65, 2, 333, False),      # odd sizes: self-paired middle row / column
    (40, 72, 33, 130, True),      # rectangular, more than one 32-channel epilogue tile
    (8, 8, 1, 5, False),          # smaller than one tile
    (2, 2, 1, 3, False),
    (130, 34, 65, 64, True),      # several row chunks, 65 channels
])
def test_parity_small_shapes(gpu, ny, nx, nf, nuv, herm):
    px = 0.1
    m = _model(ny, nx, nf, px, seed=ny * 131 + nx)
    if herm:
        u, v = synth.synth_uv(nuv, px * A)
    else:
        rng = np.random.default_rng(nuv)
        u, v = rng.normal(0, 3e5, nuv), rng.normal(0, 3e5, nuv)
    ref = od.exact_dft(u, v, m.image, px * A, 0.05 * A, -0.03 * A)
    vis = interpolate_model(u, v, m.freq, m, dRA=0.05, dDec=-0.03)
    assert vis.real.shape == (nuv, nf) and np.all(vis.weights == 1)
    assert relerr(vis, ref) < TOL


@pytest.mark.parametrize("variant", [1, 2, 200, 201, 300, 400])
@pytest.mark.parametrize("split", [1, 3])
def test_every_kernel_variant_and_split(gpu, variant, split):
    gpu.pdsb_set_dft_variant(variant)
    gpu.pdsb_set_dft_split(split)
    m = _model(96, 96, 2, 0.05, seed=variant)
    u, v = synth.synth_uv(1500, 0.05 * A)
    ref = od.exact_dft(u, v, m.image, 0.05 * A, -0.11 * A, 0.07 * A)
    vis = interpolate_model(u, v, m.freq, m, dRA=-0.11, dDec=0.07)
    assert relerr(vis, ref) < TOL


@pytest.mark.parametrize("shape", [(100, 400, 2, 300), (257, 129, 9, 131), (64, 64, 1, 1), (40, 520, 1, 700)])
@pytest.mark.parametrize("split", [1, 0])
def test_tcgen05_variant_shapes(gpu, shape, split):
    """The tcgen05/TMEM kernel (variant 200) on shapes that exercise its pipeline: four and more K tiles
    per CTA (A-operand double buffer reuse), plane groups (nf > 8), odd sizes (self-mirrored row / column),
    uv counts that do not fill a 128-row tile, and a K-split grid."""
    ny, nx, nf, nuv = shape
    gpu.pdsb_set_dft_variant(200)
    gpu.pdsb_set_dft_split(split)
    rng = np.random.default_rng(ny * 7 + nx)
    img = rng.random((ny, nx, nf, 1))
    m = synth.SynthImage(img, 0.05, synth.synth_freq(nf))
    u, v = rng.normal(0, 4e5, nuv), rng.normal(0, 4e5, nuv)
    ref = od.exact_dft(u, v, img, 0.05 * A, 0.02 * A, -0.04 * A)
    vis = interpolate_model(u, v, m.freq, m, dRA=0.02, dDec=-0.04)
    assert relerr(vis, ref) < TOL


def test_unknown_variant_is_rejected(gpu):
    assert gpu.pdsb_set_dft_variant(3) != 0
    assert gpu.pdsb_set_dft_variant(103) != 0
    assert gpu.pdsb_set_dft_variant(202) != 0
    assert gpu.pdsb_set_dft_variant(0) == 0


def test_against_literal_oracle_on_the_reference_fixture(gpu, fixture720):
    """The reference's fixture uv points (720, Hermitian-doubled) and the literal per-pixel
    oracle (one complex exponential per pixel-visibility pair), incl. the committed golden."""
    f = fixture720
    img = synth.synth_image(64, 2, 0.5, kind="disk")
    m = synth.SynthImage(img, 0.5, synth.synth_freq(2))
    vis = interpolate_model(f["u"], f["v"], m.freq, m, dRA=0.05, dDec=-0.03)
    ref = od.exact_dft_literal(f["u"], f["v"], img, 0.5 * A, 0.05 * A, -0.03 * A)
    assert relerr(vis, ref) < TOL
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dft_golden.npz"))
    assert relerr(vis, g["real"] + 1j * g["imag"]) < TOL


@pytest.mark.parametrize("kind", ["disk", "random"])
def test_config1_full(gpu, kind):
    """BASELINE.json configs[0]: 256x256 single channel onto 50k uv points, with and without
    the dRA/dDec offset."""
    c = synth.make_config("C1", kind=kind)
    for dra, ddec in ((0.0, 0.0), (c["dRA"], c["dDec"])):
        ref = od.exact_dft(c["u"], c["v"], c["model"].image, c["pixelsize"] * A, dra * A, ddec * A)
        vis = interpolate_model(c["u"], c["v"], c["freq"], c["model"], dRA=dra, dDec=ddec)
        assert relerr(vis, ref) < TOL


def test_gaussian_known_answer(gpu):
    """The reference's analytic gaussian_model (model.py:137-147) from a rendered image."""
    n, px, x0, y0, sig, flux = 256, 0.05, 0.35, -0.2, 0.3, 2.0
    x = (np.arange(n) - n / 2) * px
    X, Y = np.meshgrid(x, x)
    g = np.exp(-0.5 * ((X + x0) ** 2 + (Y - (y0 - px)) ** 2) / sig ** 2)
    img = (g * flux / g.sum())[:, :, None, None]
    m = synth.SynthImage(np.ascontiguousarray(img), px, synth.synth_freq(1))
    u, v = synth.synth_uv(2000, px * A)
    u, v = u * 0.2, v * 0.2
    vis = interpolate_model(u, v, m.freq, m, dRA=0.1, dDec=0.2)
    ref = flux * np.exp(-2 * np.pi ** 2 * (sig * A) ** 2 * (u ** 2 + v ** 2)) * \
        np.exp(-2j * np.pi * (u * (x0 + 0.1) * A + v * (y0 + 0.2) * A))
    assert np.abs(vis.real[:, 0] + 1j * vis.imag[:, 0] - ref).max() < TOL * flux


def test_empty_uv_list(gpu):
    m = _model(16, 16, 2, 0.1)
    vis = interpolate_model(np.zeros(0), np.zeros(0), m.freq, m)
    assert vis.real.shape == (0, 2)


def test_hermitian_detection_is_exact_not_assumed(gpu):
    u, v = synth.synth_uv(200, 0.1 * A)
    d = Dataset(u, v)
    assert d.hermitian and d.nuv_unique == 100
    v2 = v.copy()
    v2[150] = np.nextafter(v2[150], 1.0)          # one ulp off: must not be folded
    d2 = Dataset(u, v2)
    assert not d2.hermitian and d2.nuv_unique == 200
    m = _model(32, 32, 1, 0.1)
    ref = od.exact_dft(u, v2, m.image, 0.1 * A)
    vis = interpolate_model(u, v2, m.freq, m)
    assert relerr(vis, ref) < TOL


def _device_sample(gpu, ds, dimg, n, nf, px, dra, ddec, nuv):
    dre, dim_ = DeviceBuffer(nuv * nf * 8), DeviceBuffer(nuv * nf * 8)
    _lib.check(gpu.pdsb_sample_image(ds.handle, _lib.ptr(dimg), n, n, nf, _lib.DEVICE, px * A, dra * A, ddec * A,
                                     _lib.ptr(dre), _lib.ptr(dim_), _lib.DEVICE))
    return dre.download((nuv, nf)) + 1j * dim_.download((nuv, nf))


def test_config2_full_size_properties_and_subset_parity(gpu):
    """BASELINE.json configs[1] at full size (1024^2 x 1M uv, dRA/dDec offset), device-resident
    buffers: parity on a random 2048-point subset against the oracle, plus size-independent
    properties on all points: linearity and Hermitian symmetry."""
    c = synth.make_config("C2")
    n, nuv, px = c["npix"], c["u"].size, c["pixelsize"]
    img1 = np.ascontiguousarray(c["model"].image[:, :, :, 0])
    img2 = np.random.default_rng(5).random(img1.shape) * img1.max()
    ds = Dataset(c["u"], c["v"])
    d1, d2 = DeviceBuffer.from_numpy(img1), DeviceBuffer.from_numpy(img2)
    d3 = DeviceBuffer.from_numpy(2.0 * img1 - 0.5 * img2)
    V1 = _device_sample(gpu, ds, d1, n, 1, px, c["dRA"], c["dDec"], nuv)
    V2 = _device_sample(gpu, ds, d2, n, 1, px, c["dRA"], c["dDec"], nuv)
    V3 = _device_sample(gpu, ds, d3, n, 1, px, c["dRA"], c["dDec"], nuv)
    sub = np.random.default_rng(1).choice(nuv, 2048, replace=False)
    ref = od.exact_dft(c["u"][sub], c["v"][sub], c["model"].image, px * A, c["dRA"] * A, c["dDec"] * A)
    assert np.abs(V1[sub] - ref).max() / np.abs(V1).max() < TOL
    scale = max(np.abs(V1).max(), np.abs(V2).max())
    assert np.abs(V3 - (2.0 * V1 - 0.5 * V2)).max() / scale < 3 * TOL
    h = nuv // 2
    assert np.array_equal(V1[h:], np.conj(V1[:h]))


def test_config3_shape_channels_share_uv(gpu):
    """BASELINE.json configs[2] shape at reduced uv count (512^2 x 64 channels): every channel
    is an independent image on the same uv points."""
    c = synth.make_config("C3", nuv=4096)
    ref = od.exact_dft(c["u"], c["v"], c["model"].image, c["pixelsize"] * A, c["dRA"] * A, c["dDec"] * A)
    vis = interpolate_model(c["u"], c["v"], c["freq"], c["model"], dRA=c["dRA"], dDec=c["dDec"])
    assert vis.real.shape == (4096, 64)
    assert relerr(vis, ref) < TOL


def test_dataset_cache_invalidation_on_in_place_change(gpu):
    m = _model(32, 32, 1, 0.1)
    u, v = synth.synth_uv(64, 0.1 * A)
    a = interpolate_model(u, v, m.freq, m)
    u[:] = u * 0.5
    b = interpolate_model(u, v, m.freq, m)
    ref = od.exact_dft(u, v, m.image, 0.1 * A)
    assert relerr(b, ref) < TOL and relerr(a, ref) > 1e-3


@pytest.mark.parametrize("n,nf,nuv,herm", [(64, 2, 500, True), (128, 3, 1000, False), (256, 1, 3000, True),
                                           # the half-spectrum transform's corners: smallest sides (one radix-2 stage,
                                           # odd / even stage counts), odd channel counts (a pair without a partner,
                                           # ragged last block), more channels than a lane group, a side that needs
                                           # the strided per-thread loops and 128 KB of shared memory
                                           (2, 1, 40, False), (4, 2, 60, True), (8, 5, 80, False), (16, 7, 100, True),
                                           (32, 40, 200, False), (1024, 2, 300, True), (2048, 1, 200, False),
                                           (4096, 1, 100, True)])
def test_galario_fft_path_vs_restated_galario_algorithm(gpu, n, nf, nuv, herm):
    """code="galario-fft": the reference's own algorithm (FFT + bilinear interpolation, oracle/dft.py:galario_like)
    on the GPU, fp64 throughout: 1e-12 of max|V|; and the method gap to the exact transform is galario's."""
    px = 0.05
    img = synth.synth_image(n, nf, px, kind="disk")
    m = synth.SynthImage(img, px, synth.synth_freq(nf))
    if herm:
        u, v = synth.synth_uv(nuv, px * A)
    else:
        rng = np.random.default_rng(n)
        lim = 0.4 / (px * A)
        u, v = rng.uniform(-lim, lim, nuv), rng.uniform(-lim, lim, nuv)
    ref = od.galario_like(u, v, m.image, px * A, 0.04 * A, -0.03 * A)
    vis = interpolate_model(u, v, m.freq, m, dRA=0.04, dDec=-0.03, code="galario-fft")
    assert vis.real.shape == (nuv, nf)
    assert relerr(vis, ref) < 1e-12
    exact = interpolate_model(u, v, m.freq, m, dRA=0.04, dDec=-0.03)
    gap = np.abs((vis.real - exact.real) + 1j * (vis.imag - exact.imag)).max() / np.abs(ref).max()
    if n >= 64:
        assert gap < 0.1                               # the two methods describe the same sky


def test_galario_fft_likelihood_and_grid_points(gpu):
    """On FFT grid points (u, v multiples of du) interpolation is exact: the FFT path equals the exact transform
    to rounding; and the fused likelihood equals the numpy expression on the FFT-path visibilities."""
    from oracle import likelihood as ol
    from pdspy_b200.interferometry import loglike_image_fft, Visibilities
    n, px, nf = 64, 0.05, 2
    img = synth.synth_image(n, nf, px, kind="disk")
    m = synth.SynthImage(img, px, synth.synth_freq(nf))
    du = 1.0 / (n * px * A)
    rng = np.random.default_rng(8)
    u = rng.integers(-n // 2 + 1, n // 2, 400) * du
    v = rng.integers(-n // 2 + 1, n // 2, 400) * du
    a = interpolate_model(u, v, m.freq, m, code="galario-fft")
    b = interpolate_model(u, v, m.freq, m)
    scale = np.abs(b.real + 1j * b.imag).max()
    assert np.abs((a.real - b.real) + 1j * (a.imag - b.imag)).max() / scale < 2e-6     # fp32 products in b
    for nf2 in (nf, 1, 3, 37):                         # lane groups of 2, 1, 2 and 32 channels
        img2 = synth.synth_image(n, nf2, px, kind="disk")
        m2 = synth.SynthImage(img2, px, synth.synth_freq(nf2))
        u2, v2 = synth.synth_uv(800, px * A)
        re, im, w = synth.synth_data(800, nf2)
        data = Visibilities(u2, v2, m2.freq, re, im, w)
        ll, c_re, c_im = loglike_image_fft(data, m2, dRA=0.01, dDec=0.02)
        vis = interpolate_model(u2, v2, m2.freq, m2, dRA=0.01, dDec=0.02, code="galario-fft")
        ll_ref = ol.lnlike_vis_numpy(re, im, w, vis.real, vis.imag)
        assert abs(ll - ll_ref) <= 1e-10 * abs(ll_ref), nf2


@pytest.mark.parametrize("ny,nx,nf,nuv,herm", [(64, 64, 2, 400, True), (63, 65, 3, 333, False), (130, 34, 9, 64, True),
                                               (256, 256, 1, 2000, True), (17, 500, 1, 129, False)])
def test_fp64_reference_kernel_vs_cpu_oracle(gpu, ny, nx, nf, nuv, herm):
    """Variant 300 (all fp64, no fold): 1e-11 of max|V| from the exact CPU oracle (both sum ~1e4-1e5 fp64 terms) - the on-device reference the
    full-size tests hold the other kernels to."""
    gpu.pdsb_set_dft_variant(300)
    rng = np.random.default_rng(ny * 31 + nx)
    img = rng.random((ny, nx, nf, 1))
    m = synth.SynthImage(img, 0.05, synth.synth_freq(nf))
    if herm:
        u, v = synth.synth_uv(nuv, 0.05 * A)
    else:
        u, v = rng.normal(0, 4e5, nuv), rng.normal(0, 4e5, nuv)
    ref = od.exact_dft(u, v, img, 0.05 * A, 0.021 * A, -0.034 * A)
    vis = interpolate_model(u, v, m.freq, m, dRA=0.021, dDec=-0.034)
    assert relerr(vis, ref) < 1e-11
