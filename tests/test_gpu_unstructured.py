"""GPU parity: unstructured (scattered-point) images, interpolate_model(code="galario-unstructured")."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dft as od, unstructured as ou                        # noqa: E402
import pdspy_b200 as pb                                                  # noqa: E402
import synth                                             # noqa: E402
from pdspy_b200.interferometry import interpolate_model                  # noqa: E402
from pdspy_b200.interferometry.unstructured import regrid                # noqa: E402

pytestmark = pytest.mark.gpu
A = synth.ARCSEC


def _circular_image(nr=60, nphi=48, nf=3, rmax=1.5, seed=0):
    """Points laid out like RADMC-3D's circular images (Model.py:536-552): one centre point, then rings."""
    r = np.concatenate([[0.0], np.geomspace(0.01, rmax, nr)])
    phi = (np.arange(nphi) + 0.5) * 2 * np.pi / nphi
    rr, pp = np.meshgrid(r, phi)
    x, y = -rr * np.cos(pp), rr * np.sin(pp)
    x = np.concatenate(([x[0, 0]], x[:, 1:].ravel()))
    y = np.concatenate(([y[0, 0]], y[:, 1:].ravel()))
    chan = 1.0 + 0.3 * np.arange(nf)
    img = (np.exp(-0.5 * ((x - 0.2) ** 2 / 0.3 ** 2 + (y + 0.1) ** 2 / 0.5 ** 2))[:, None] * chan[None, :]) * 1e9   # Jy/sr
    return pb.UnstructuredImage(np.ascontiguousarray(img), x=x, y=y, freq=synth.synth_freq(nf))


def test_regrid_vs_scipy_linear_interpolator(gpu):
    m = _circular_image()
    got = regrid(m, 96, 0.04)
    ref = ou.regrid(m.x, m.y, m.image, 96, 0.04)
    assert got.shape == (96, 96, 3, 1) and np.count_nonzero(ref) > 1000
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()


def test_interpolate_model_unstructured_vs_oracle_chain(gpu):
    m = _circular_image()
    u, v = synth.synth_uv(3000, 0.04 * A)
    vis = interpolate_model(u, v, m.freq, m, dRA=0.05, dDec=-0.02, code="galario-unstructured", nxy=128, dxy=0.03)
    cube = ou.regrid(m.x, m.y, m.image, 128, 0.03)
    ref = od.exact_dft(u, v, cube, 0.03 * A, 0.05 * A, -0.02 * A)
    assert vis.real.shape == (3000, 3) and np.all(vis.weights == 1)
    assert np.abs(vis.real + 1j * vis.imag - ref).max() / np.abs(ref).max() < 1e-5


def test_grid_nodes_as_scattered_points_reproduce_the_structured_path(gpu):
    """Fixes the frame: a regular image whose pixels are handed over as scattered points (at the positions
    the reference's flip puts them: x_c, y_j + one pixel) must give the code="galario" visibilities."""
    n, px, nf = 48, 0.05, 2
    rng = np.random.default_rng(4)
    img = rng.random((n, n, nf, 1))
    sm = synth.SynthImage(img, px, synth.synth_freq(nf))
    u, v = synth.synth_uv(1500, px * A)
    ref = interpolate_model(u, v, sm.freq, sm, dRA=0.03, dDec=0.04)
    xs = (np.arange(n) - n / 2) * px
    X, Y = np.meshgrid(xs, xs + px)                                     # y_eff = y_j + px (the [::-1] of an even grid)
    um = pb.UnstructuredImage(np.ascontiguousarray(img[:, :, :, 0].reshape(n * n, nf) / (px * A) ** 2),
                              x=X.ravel(), y=Y.ravel(), freq=sm.freq)
    vis = interpolate_model(u, v, sm.freq, um, dRA=0.03, dDec=0.04, code="galario-unstructured", nxy=n, dxy=px)
    scale = np.abs(ref.real + 1j * ref.imag).max()
    assert np.abs((vis.real - ref.real) + 1j * (vis.imag - ref.imag)).max() / scale < 1e-9


def test_trift_vs_oracle(gpu):
    """code="trift" (interpolate_model.py:49-55): the exact transform of the triangulated image against the
    extended-precision oracle (explicit divided differences), same triangulation; baselines long enough that the
    outer triangles get sub-divided, a Hermitian-doubled and a plain uv list, u = v = 0 (total flux)."""
    from oracle import trift as ot
    m = _circular_image(nr=14, nphi=12, nf=3, rmax=1.5)
    rng = np.random.default_rng(8)
    for herm in (True, False):
        if herm:
            u, v = synth.synth_uv(64, 0.02 * A)
        else:
            u, v = np.concatenate([[0.0], rng.normal(0, 4e5, 40)]), np.concatenate([[0.0], rng.normal(0, 4e5, 40)])
        vis = interpolate_model(u, v, m.freq, m, dRA=0.05, dDec=-0.02, code="trift")
        ref = ot.trift(m.x, m.y, m.image, u, v, 0.05 * A, -0.02 * A)
        assert vis.real.shape == (u.size, 3) and np.all(vis.weights == 1)
        assert np.abs(vis.real + 1j * vis.imag - ref).max() / np.abs(ref).max() < 1e-11


def test_trift_is_the_small_pixel_limit_of_the_regridded_path(gpu):
    """The same sky through the reference's three codes must give the same visibilities: code="trift" is what
    code="galario-unstructured" converges to as its pixels shrink; fixes trift's frame and sign."""
    m = _circular_image(nr=40, nphi=36, nf=2, rmax=1.0)
    u, v = synth.synth_uv(400, 0.03 * A)
    t = interpolate_model(u, v, m.freq, m, dRA=0.03, dDec=0.01, code="trift")
    tv = t.real + 1j * t.imag
    errs = []
    for nxy, dxy in ((256, 0.01), (512, 0.005)):
        g = interpolate_model(u, v, m.freq, m, dRA=0.03, dDec=0.01, code="galario-unstructured", nxy=nxy, dxy=dxy)
        errs.append(np.abs(g.real + 1j * g.imag - tv).max() / np.abs(tv).max())
    assert errs[1] < 2e-4 and errs[1] < 0.75 * errs[0]           # measured 1.2e-4 -> 7.3e-5 (hull edge: O(dxy))


def test_trift_subdivision_is_exact(gpu):
    """Sub-dividing triangles (trift.py) must not change the result: the interpolant is linear on the parent."""
    from pdspy_b200.interferometry import trift as tr
    m = _circular_image(nr=10, nphi=8, nf=2, rmax=2.0)
    u, v = synth.synth_uv(200, 0.02 * A)
    a = interpolate_model(u, v, m.freq, m, code="trift")
    old = tr.PHASE_SPAN
    try:
        tr.PHASE_SPAN = 2.0                                     # 4x the triangles
        tr._CACHE.clear()
        b = interpolate_model(u, v, m.freq, m, code="trift")
    finally:
        tr.PHASE_SPAN = old
        tr._CACHE.clear()
    # (measured against the extended-precision oracle: 4e-14 with 55k triangles, 3e-13 with 220k: summation rounding)
    assert np.abs((a.real - b.real) + 1j * (a.imag - b.imag)).max() <= 1e-11 * np.abs(a.real + 1j * a.imag).max()


def test_unstructured_shapes_checked(gpu):
    m = _circular_image()
    u, v = synth.synth_uv(10, 0.04 * A)
    bad = pb.UnstructuredImage(m.image, x=m.x[:-1], y=m.y, freq=m.freq)
    with pytest.raises(ValueError):
        interpolate_model(u, v, m.freq, bad, code="trift")
    with pytest.raises(ValueError):
        interpolate_model(u, v, m.freq, bad, code="galario-unstructured", nxy=32, dxy=0.1)
