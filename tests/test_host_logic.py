"""Host-side mirrors of the reference interface: containers, argument checks, synthetic data."""
import pickle

import numpy as np
import pytest

import synth
from pdspy_b200 import Image
from pdspy_b200.interferometry import Visibilities, VisibilitiesObject
from pdspy_b200.interferometry.interpolate_model import interpolate_model
from pdspy_b200 import device


def _vis(n=6, nf=2):
    rng = np.random.default_rng(0)
    return Visibilities(rng.normal(size=n), rng.normal(size=n), np.arange(nf) + 1.0,
                        rng.normal(size=(n, nf)), rng.normal(size=(n, nf)), rng.uniform(1, 2, (n, nf)))


def test_visibilities_derived_arrays_match_reference_definitions(live_ref):
    d = _vis()
    np.testing.assert_array_equal(d.uvdist, np.sqrt(d.u ** 2 + d.v ** 2))
    np.testing.assert_array_equal(d.amp, np.sqrt(d.real ** 2 + d.imag ** 2))
    np.testing.assert_array_equal(d.phase, np.arctan2(d.imag, d.real))
    if live_ref is not None:
        r = live_ref.Visibilities(d.u, d.v, d.freq, d.real, d.imag, d.weights)
        for nm in ("uvdist", "amp", "phase", "weights"):
            np.testing.assert_array_equal(getattr(d, nm), getattr(r, nm))
        assert r.array_name == d.array_name == "CARMA"


def test_visibilities_default_weights_and_type_checks():
    d = Visibilities(np.zeros(3), np.zeros(3), np.ones(1), np.zeros((3, 1)), np.zeros((3, 1)))
    assert d.weights.shape == (3, 1) and np.all(d.weights == 1)
    with pytest.raises(ValueError):
        Visibilities(np.zeros(3, dtype=np.float32), np.zeros(3), np.ones(1), np.zeros((3, 1)), np.zeros((3, 1)))
    with pytest.raises(ValueError):
        Visibilities(np.zeros(3), np.zeros(3), np.ones(1), np.zeros(3), np.zeros((3, 1)))
    with pytest.raises(TypeError):
        Visibilities([0.0, 1.0], np.zeros(2))
    e = Visibilities()
    assert e.u is None and e.amp is None


def test_pickle_round_trip_returns_base_class_like_the_reference():
    d = _vis()
    d.baseline = np.arange(6)
    r = pickle.loads(pickle.dumps(d))
    assert type(r) is VisibilitiesObject           # libinterferometry.pyx:46-52 rebuild()
    np.testing.assert_array_equal(r.real, d.real)
    np.testing.assert_array_equal(r.amp, d.amp)
    np.testing.assert_array_equal(r.baseline, d.baseline)


def test_get_baselines():
    d = _vis()
    d.baseline = np.array([1, 2, 1, 2, 1, 3])
    s = d.get_baselines(1)
    assert s.u.size == 3 and s.real.shape == (3, 2)


def test_image_container_derives_freq_and_wave():
    img = Image(np.zeros((4, 4, 2, 1)), x=np.arange(4.0), y=np.arange(4.0), wave=np.array([0.1, 0.2]))
    np.testing.assert_allclose(img.freq * img.wave, 2.99792458e10)
    with pytest.raises(ValueError):
        Image(np.zeros((4, 4, 2), dtype=np.float64))


def test_interpolate_model_argument_errors_do_not_need_a_gpu():
    c = synth.make_config("C1", nuv=16)
    with pytest.raises(ValueError):
        interpolate_model(c["u"], c["v"], c["freq"], c["model"], code="nope")


def test_synthetic_uv_is_hermitian_doubled_and_nyquist_limited():
    u, v = synth.synth_uv(1000, 0.01 * synth.ARCSEC)
    h = 500
    np.testing.assert_array_equal(u[h:], -u[:h])
    np.testing.assert_array_equal(v[h:], -v[:h])
    assert np.hypot(u, v).max() <= 0.45 / (0.01 * synth.ARCSEC) * (1 + 1e-12)
    re, im, w = synth.synth_data(1000, 3)
    np.testing.assert_array_equal(im[h:], -im[:h])
    assert (w == 0).any() and (w < 0).any()


def test_host_hash_sees_any_single_element_edit():
    """The handle cache's content hash (pdsb_hash64; no GPU needed): one changed element anywhere, a different
    length and a different shape-preserving permutation all change it; equal content hashes equal."""
    from pdspy_b200 import _lib
    rng = np.random.default_rng(0)
    a = rng.random((30011, 7))
    h0 = _lib.host_hash(a)
    assert _lib.host_hash(a.copy()) == h0
    for idx in ((0, 0), (12345, 3), (30010, 6), (4097, 1)):         # incl. positions round 1's sampled fingerprint missed
        b = a.copy()
        b[idx] = np.nextafter(b[idx], 2.0)
        assert _lib.host_hash(b) != h0, idx
    assert _lib.host_hash(a[:-1]) != h0
    b = a.copy()
    b[[5, 6]] = b[[6, 5]]
    assert _lib.host_hash(b) != h0
    assert _lib.host_hash(np.zeros(0)) == _lib.host_hash(np.zeros(0))
    assert _lib.host_hash(np.zeros(3)) != _lib.host_hash(np.zeros(4))


def test_lazy_visibilities_without_a_device_copy_behave_like_plain_ones():
    from pdspy_b200.interferometry import Visibilities
    v = Visibilities(np.zeros(2), np.ones(2), np.ones(1), np.full((2, 1), 3.0), np.full((2, 1), 4.0))
    assert v._device_token is None and np.all(v.weights == 1) and np.all(v.amp == 5)
    v.real = np.zeros((2, 1))
    assert np.all(v.real == 0)
    w = pickle.loads(pickle.dumps(v))
    assert type(w).__name__ == "VisibilitiesObject" and np.array_equal(w.imag, v.imag)


def test_unstructured_triangulation_helper_vs_scipy_interpolator():
    """pdspy_b200.interferometry.unstructured.triangulate (host side of code="galario-unstructured"): the
    per-pixel triangles and barycentric weights reproduce scipy's LinearNDInterpolator (oracle/unstructured.py)."""
    import numpy as np
    from oracle import unstructured as ou
    from pdspy_b200.interferometry.unstructured import triangulate
    A = 4.84813681e-6
    rng = np.random.default_rng(1)
    r = np.concatenate([[0], np.sqrt(rng.random(2000)) * 1.2])
    ph = rng.random(2001) * 2 * np.pi
    x, y = -r * np.cos(ph), r * np.sin(ph)
    vals = np.exp(-0.5 * (x ** 2 + y ** 2) / 0.3 ** 2)[:, None] * np.array([1.0, 0.5])
    nxy, dxy = 48, 0.05
    tri, bary = triangulate(x * A, -y * A, nxy, dxy * A)
    got = np.where(tri[:, :1] >= 0, (bary[:, :, None] * vals[np.maximum(tri, 0)]).sum(axis=1), 0.0) * (dxy * A) ** 2
    ref = ou.regrid(x, y, vals, nxy, dxy)[::-1, :, :, 0].reshape(nxy * nxy, 2)
    assert (tri[:, 0] >= 0).mean() > 0.5
    assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max()


def test_mesh_filler_matches_meshgrid_and_hands_over_exceptions():
    """grid() / sharded_grid() fill the result's coordinate arrays (numpy.meshgrid, libinterferometry.pyx:383) on a
    second host thread while the device works; a failure there must surface in the caller, not vanish."""
    from pdspy_b200.interferometry.grid import MeshFiller
    uu, vv = np.linspace(-3, 2, 6), np.linspace(-1, 1, 5)
    for threaded in (True, False):
        a, b = MeshFiller(uu, vv, threaded=threaded).result()
        ra, rb = np.meshgrid(uu, vv)
        assert np.array_equal(a, ra) and np.array_equal(b, rb)
    huge = np.broadcast_to(np.float64(0.0), (1 << 40,))      # a view: the 2-D copies cannot be allocated
    bad = MeshFiller(huge, huge[:4], threaded=True)
    with pytest.raises(MemoryError):
        bad.result()
