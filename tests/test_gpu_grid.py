"""Convolutional gridding on the GPU against the live reference's golden vectors and the oracle.

Bars (north_star): index maps bit-exact; gridded weight (and real/imag) maps bit-exact for the
pillbox kernel in deterministic mode; 1e-10 relative for expsinc (the reference's own last bits
depend on its -ffast-math build, setup.py:11) and for the atomic mode."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden import GRID_CASES_FIXTURE, GRID_CASES_MULTI, multi_channel_set      # noqa: E402

from oracle import grid as og                                                        # noqa: E402
import synth
from pdspy_b200 import _lib                                                   # noqa: E402
from pdspy_b200.interferometry import grid, freqcorrect, Visibilities                # noqa: E402

pytestmark = pytest.mark.gpu


def _quiet(fn, *a, **kw):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        r = fn(*a, **kw)
    return r, buf.getvalue()


def _vs_golden(name, g, golden, exact, tol=1e-10):
    assert tuple(golden[name + "/shape"]) == g.real.shape
    np.testing.assert_array_equal(g.freq, golden[name + "/freq"])
    np.testing.assert_array_equal(g.u[[0, 1, -1]], golden[name + "/u_ends"])
    for nm in ("real", "imag", "weights"):
        arr = getattr(g, nm)
        idx, val = golden["%s/%s_idx" % (name, nm)], golden["%s/%s_val" % (name, nm)]
        got = arr.reshape(-1)[idx]
        nnz = int(golden["%s/%s_nnz" % (name, nm)])
        if exact:
            assert np.count_nonzero(arr) == nnz
            np.testing.assert_array_equal(got, val, err_msg="%s %s" % (name, nm))
        else:
            # cells where Hermitian-conjugate contributions cancel hold a rounding residue
            # (~1e-17) in the reference and may be exactly 0 here, or vice versa
            assert abs(np.count_nonzero(arr) - nnz) <= max(2, nnz // 1000)
            assert np.abs(got - val).max() <= tol * np.abs(val).max(), (name, nm)
        s = float(golden["%s/%s_sum" % (name, nm)])
        assert abs(arr.sum() - s) <= 1e-9 * max(np.abs(arr).sum(), 1e-300)


@pytest.mark.parametrize("name", sorted(GRID_CASES_FIXTURE))
@pytest.mark.parametrize("deterministic", [True, False])
def test_fixture_cases_vs_live_reference_golden(gpu, fixture720, grid_golden, name, deterministic):
    f = fixture720
    data = Visibilities(f["u"], f["v"], f["freq"], f["real"], f["imag"], f["weights"])
    kw = GRID_CASES_FIXTURE[name]
    g, _ = _quiet(grid, data, deterministic=deterministic, **kw)
    exact = deterministic and kw.get("convolution", "pillbox") == "pillbox" and \
        kw.get("weighting", "natural") != "robust" and not kw.get("imaging", False)
    _vs_golden(name, g, grid_golden, exact)


@pytest.mark.parametrize("name", sorted(GRID_CASES_MULTI))
@pytest.mark.parametrize("deterministic", [True, False])
def test_multichannel_cases_vs_live_reference_golden(gpu, grid_golden, name, deterministic):
    u, v, freq, re, im, w = multi_channel_set()
    kw = GRID_CASES_MULTI[name]
    g, out = _quiet(grid, Visibilities(u, v, freq, re, im, w), deterministic=deterministic, **kw)
    assert out.startswith("WARNING: uv.grid was supplied with a gridsize and binsize")   # :424-425
    exact = deterministic and kw.get("convolution", "pillbox") == "pillbox" and \
        kw.get("weighting", "natural") != "robust" and not kw.get("imaging", False)
    _vs_golden(name, g, grid_golden, exact)


@pytest.mark.parametrize("kw", [
    dict(gridsize=128, binsize=8000., convolution="pillbox"),
    dict(gridsize=128, binsize=8000., convolution="pillbox", mode="spectralline"),
    dict(gridsize=127, binsize=8000., convolution="pillbox", mode="spectralline", weighting="uniform", npixels=1),
    dict(gridsize=128, binsize=8000., convolution="pillbox", weighting="superuniform"),
    dict(gridsize=128, binsize=8000., convolution="pillbox", mfs=True),
    dict(gridsize=128, binsize=8000., convolution="pillbox", channel=1),
])
def test_pillbox_deterministic_is_bit_exact_incl_index_maps(gpu, kw):
    rng = np.random.default_rng(7)
    n, nf = 20000, 3
    u, v = rng.normal(0, 2e5, n), rng.normal(0, 2e5, n)
    freq = 230e9 + 1e8 * np.arange(nf)
    re, im = rng.normal(size=(n, nf)), rng.normal(size=(n, nf))
    w = rng.uniform(0.5, 2, (n, nf))
    w[rng.random((n, nf)) < 0.01] = 0
    w[rng.random((n, nf)) < 0.001] *= -1
    o, _ = _quiet(og.grid, u, v, freq, re, im, w, return_maps=True, **kw)
    (g, gi, gj, wm), _ = _quiet(grid, Visibilities(u, v, freq, re, im, w), return_maps=True, **kw)
    np.testing.assert_array_equal(gi, o[6])
    np.testing.assert_array_equal(gj, o[7])
    np.testing.assert_array_equal(wm, o[9])
    np.testing.assert_array_equal(g.weights, o[5])
    np.testing.assert_array_equal(g.real, o[3])
    np.testing.assert_array_equal(g.imag, o[4])
    np.testing.assert_array_equal(g.u, o[0])
    np.testing.assert_array_equal(g.v, o[1])


def test_expsinc_and_atomic_modes_within_tolerance(gpu):
    rng = np.random.default_rng(8)
    n, nf = 30000, 2
    u, v = rng.normal(0, 2e5, n), rng.normal(0, 2e5, n)
    freq = 230e9 + 1e8 * np.arange(nf)
    re, im = rng.normal(size=(n, nf)), rng.normal(size=(n, nf))
    w = rng.uniform(0.5, 2, (n, nf))
    d = Visibilities(u, v, freq, re, im, w)
    for kw in (dict(gridsize=128, binsize=9000., convolution="expsinc", mode="spectralline"),
               dict(gridsize=128, binsize=9000., convolution="expsinc", weighting="robust", robust=0.5),
               dict(gridsize=129, binsize=9000., convolution="expsinc", imaging=True, weighting="uniform")):
        o, _ = _quiet(og.grid, u, v, freq, re, im, w, **kw)
        for det in (True, False):
            g, _ = _quiet(grid, d, deterministic=det, **kw)
            for nm, ob in zip(("real", "imag", "weights"), o[3:6]):
                assert np.abs(getattr(g, nm) - ob).max() <= 1e-10 * np.abs(ob).max(), (kw, det, nm)


def test_edge_cases_empty_and_all_outside(gpu):
    e = Visibilities(np.zeros(0), np.zeros(0), np.array([230e9]), np.zeros((0, 1)), np.zeros((0, 1)), np.zeros((0, 1)))
    g, out = _quiet(grid, e, gridsize=16, binsize=1000.)
    assert g.real.shape == (256, 1) and not g.weights.any() and out == ""
    far = Visibilities(np.array([1e9, -1e9]), np.array([0., 0.]), np.array([230e9]), np.ones((2, 1)), np.ones((2, 1)),
                       np.ones((2, 1)))
    g, out = _quiet(grid, far, gridsize=16, binsize=1000., convolution="expsinc")
    assert not g.weights.any() and out.startswith("WARNING")
    # negative coordinate wraps through the uint32 cast exactly like numpy (index map parity)
    (g, gi, gj, _), _ = _quiet(grid, far, gridsize=16, binsize=1000., return_maps=True)
    i_ref, j_ref = og.index_maps(far.u, far.v, far.freq, 16, 1000.)
    np.testing.assert_array_equal(gi, i_ref)
    np.testing.assert_array_equal(gj, j_ref)


def test_point_on_cell_edge_is_dropped_by_pillbox(gpu):
    """ones() uses strict |u| < 0.5 (libinterferometry.pyx:576-585): a point exactly between two
    cell centres contributes to neither."""
    d = Visibilities(np.array([500.0, 1000.0]), np.array([0.0, 0.0]), np.array([230e9]), np.ones((2, 1)),
                     np.ones((2, 1)), np.ones((2, 1)))
    g, _ = _quiet(grid, d, gridsize=8, binsize=1000.)
    o, _ = _quiet(og.grid, d.u, d.v, d.freq, d.real, d.imag, d.weights, gridsize=8, binsize=1000.)
    np.testing.assert_array_equal(g.weights, o[5])
    assert g.weights.sum() == 1.0


def test_freqcorrect_bit_exact(gpu):
    u, v, freq, re, im, w = multi_channel_set()
    d = Visibilities(u, v, freq, re, im, w)
    r = freqcorrect(d)
    o = og.freqcorrect(u, v, freq, re, im, w)
    for a, nm in zip(o, ("u", "v", "freq", "real", "imag", "weights")):
        np.testing.assert_array_equal(getattr(r, nm), a, err_msg=nm)
    r2 = freqcorrect(d, freq=231e9)
    o2 = og.freqcorrect(u, v, freq, re, im, w, new_freq=231e9)
    np.testing.assert_array_equal(r2.u, o2[0])


def test_config4_scale_properties(gpu):
    """BASELINE.json configs[3] shape at 1M visibilities onto 2048^2 (the full 10M runs in
    bench.py): size-independent properties.  pillbox/natural/imaging=False conserves the total
    weight of the points that land on the grid; deterministic == atomic to rounding; index maps
    equal numpy's."""
    px = 0.01
    u, v = synth.synth_uv(1_000_000, px * synth.ARCSEC)
    re, im, w = synth.synth_data(1_000_000, 1)
    d = Visibilities(u, v, synth.synth_freq(1), re, im, w)
    G = 2048
    binsize = 2.2 * np.hypot(u, v).max() / G
    (g, gi, gj, wm), _ = _quiet(grid, d, gridsize=G, binsize=binsize, return_maps=True, deterministic=True)
    i_ref, j_ref = og.index_maps(u, v, d.freq, G, binsize)
    np.testing.assert_array_equal(gi, i_ref)
    np.testing.assert_array_equal(gj, j_ref)
    ga, _ = _quiet(grid, d, gridsize=G, binsize=binsize, deterministic=False)
    assert np.abs(ga.weights - g.weights).max() <= 1e-12 * g.weights.max()
    assert abs(g.weights.sum() - wm.sum()) <= 1e-9 * wm.sum()
    ge, _ = _quiet(grid, d, gridsize=G, binsize=binsize, convolution="expsinc", deterministic=False, imaging=True)
    assert abs(ge.weights.sum() - 1.0) < 1e-12


def test_sharded_grid_single_rank_equals_grid(gpu):
    """pdspy_b200.dist.sharded_grid (multi-GPU throughput mode) with one rank and no process group:
    raw-sum maps -> device normalisation must equal the ordinary grid()."""
    from pdspy_b200 import dist as pdist
    u, v, freq, re, im, w = multi_channel_set()
    d = Visibilities(u, v, freq, re, im, w)
    for kw in (dict(gridsize=128, binsize=8000., convolution="expsinc", mode="spectralline", imaging=True),
               dict(gridsize=128, binsize=8000., convolution="pillbox")):
        a, out = _quiet(pdist.sharded_grid, pdist.shard_visibilities(d, 0, 1), **kw)
        b, _ = _quiet(grid, d, **kw)
        assert out.startswith("WARNING")
        for nm in ("u", "v", "freq"):
            np.testing.assert_array_equal(getattr(a, nm), getattr(b, nm))
        for nm in ("real", "imag", "weights"):
            assert np.abs(getattr(a, nm) - getattr(b, nm)).max() <= 1e-12 * np.abs(getattr(b, nm)).max()


@pytest.mark.parametrize("kw", [dict(gridsize=128, binsize=8000., convolution="expsinc", mode="spectralline", imaging=True),
                                dict(gridsize=128, binsize=8000., convolution="pillbox"),
                                dict(gridsize=65, binsize=16000., convolution="expsinc")])
@pytest.mark.parametrize("nbands", [1, 3, 8])
def test_banded_grid_is_bit_identical_to_the_ordered_grid(gpu, kw, nbands):
    """pdspy_b200.dist.banded_grid (multi-GPU bit-exact mode): the row bands of `nbands`
    ranks, computed here one after the other on one GPU, summed and normalised, equal the single-GPU
    ordered grid() bit for bit (the N-process version is scripts/gpu_dist_grid.py under torchrun)."""
    import torch
    from pdspy_b200 import dist as pdist, _lib
    u, v, freq, re, im, w = multi_channel_set()
    d = Visibilities(u, v, freq, re, im, w)
    ref, _ = _quiet(grid, d, deterministic=True, **kw)
    try:
        if nbands == 1:
            got, out = _quiet(pdist.banded_grid, d, **kw)
            assert out.startswith("WARNING")
        else:
            total, nout = None, 0
            for b in range(nbands):
                maps, n = pdist.banded_grid(d, bands=(b, nbands), **kw)
                others = maps.clone()
                lo, hi = pdist.shard_bounds(kw["gridsize"], b, nbands)
                others.view(3, kw["gridsize"], kw["gridsize"], -1)[:, lo:hi] = 0
                assert not others.any()                                 # nothing outside the rank's own band
                total = maps if total is None else total + maps
                nout += n
            nch = total.shape[2]
            _lib.check(gpu.pdsb_grid_normalise(total[0].data_ptr(), total[1].data_ptr(), total[2].data_ptr(),
                                               kw["gridsize"], nch, 1 if kw.get("imaging") else 0))
            torch.cuda.synchronize()
            host = total.cpu().numpy()
            got = Visibilities(ref.u, ref.v, ref.freq, host[0], host[1], host[2])
            assert nout > 0
    finally:
        _lib.check(gpu.pdsb_reset_stream())
    for nm in ("real", "imag", "weights"):
        np.testing.assert_array_equal(getattr(got, nm), getattr(ref, nm))


@pytest.mark.parametrize("kw", [dict(gridsize=128, binsize=8000., convolution="expsinc", weighting="uniform", npixels=1),
                                dict(gridsize=128, binsize=8000., convolution="pillbox", weighting="robust", robust=0.5, npixels=2,
                                     mode="spectralline", imaging=True),
                                dict(gridsize=65, binsize=16000., convolution="expsinc", weighting="superuniform"),
                                dict(gridsize=128, binsize=8000., convolution="pillbox", weighting="robust", robust=-1.0, npixels=1)])
@pytest.mark.parametrize("nbands", [1, 3])
def test_banded_grid_with_reweighting_is_bit_identical(gpu, kw, nbands):
    """Re-weighting (libinterferometry.pyx:429-485) in the bit-exact multi-GPU mode: the binned-weight map is built
    band by band (ones + ordered box sums), summed, and handed to every band's main scatter; robust's global sums
    are taken over the full map and the full weight array.  Bands computed one after the other on one GPU."""
    import torch
    from pdspy_b200 import dist as pdist, _lib
    u, v, freq, re, im, w = multi_channel_set()
    d = Visibilities(u, v, freq, re, im, w)
    ref, _ = _quiet(grid, d, deterministic=True, **kw)
    G = kw["gridsize"]
    try:
        if nbands == 1:
            got, _ = _quiet(pdist.banded_grid, d, **kw)
        else:
            wm, sumw = None, None
            for b in range(nbands):
                m, sw = pdist.banded_weight_map(d, G, kw["binsize"], kw["weighting"], kw.get("npixels", 0),
                                                kw.get("mode", "continuum"), (b, nbands))
                wm = m if wm is None else wm + m
                sumw = sw
            total = None
            for b in range(nbands):
                maps, _n = pdist.banded_grid(d, bands=(b, nbands), weight_map=(wm, sumw), **kw)
                total = maps if total is None else total + maps
            nch = total.shape[2]
            _lib.check(gpu.pdsb_grid_normalise(total[0].data_ptr(), total[1].data_ptr(), total[2].data_ptr(), G, nch,
                                               1 if kw.get("imaging") else 0))
            torch.cuda.synchronize()
            host = total.cpu().numpy()
            got = Visibilities(ref.u, ref.v, ref.freq, host[0], host[1], host[2])
    finally:
        _lib.check(gpu.pdsb_reset_stream())
    for nm in ("real", "imag", "weights"):
        np.testing.assert_array_equal(getattr(got, nm), getattr(ref, nm))


@pytest.mark.parametrize("kw", [dict(weighting="uniform", npixels=1, convolution="expsinc"),
                                dict(weighting="robust", robust=0.5, npixels=2, convolution="pillbox", mode="spectralline")])
def test_sharded_grid_with_reweighting_single_rank(gpu, kw):
    """pdspy_b200.dist.sharded_grid on one rank (weights phase -> reduced map -> main scatter) against grid()."""
    from pdspy_b200 import dist as pdist, _lib
    u, v, freq, re, im, w = multi_channel_set()
    d = Visibilities(u, v, freq, re, im, w)
    ref, _ = _quiet(grid, d, gridsize=128, binsize=8000., deterministic=True, **kw)
    try:
        got, _ = _quiet(pdist.sharded_grid, d, gridsize=128, binsize=8000., **kw)
    finally:
        _lib.check(gpu.pdsb_reset_stream())
    for nm in ("real", "imag", "weights"):
        assert np.abs(getattr(got, nm) - getattr(ref, nm)).max() <= 1e-11 * np.abs(getattr(ref, nm)).max(), nm


def test_band_needs_ordered_raw_natural(gpu):
    from pdspy_b200 import _lib
    u, v, freq, re, im, w = multi_channel_set()
    d = Visibilities(u, v, freq, re, im, w)
    _lib.check(gpu.pdsb_set_grid_band(0, 16))
    try:
        with pytest.raises(_lib.PdsbError):
            _quiet(grid, d, gridsize=64, binsize=8000.)                 # normalised output + a band: refused
    finally:
        _lib.check(gpu.pdsb_set_grid_band(0, 0))
    assert gpu.pdsb_set_grid_band(5, 5) != 0


@pytest.mark.parametrize("G", [5, 8, 24, 300])
@pytest.mark.parametrize("convolution", ["pillbox", "expsinc"])
def test_fast_mode_hot_cells_and_tiny_grids(gpu, G, convolution):
    """The sorted-tile mode on what stresses its sort: 90 % of 200 k visibilities inside ONE cell (a key run cut
    into hundreds of work items, every rank issued by the shared-memory window), grids smaller than the window
    and smaller than one tile, sides that are not a multiple of the tile; against the ordered mode."""
    rng = np.random.default_rng(G)
    n, nf = 200_000, 1
    binsize = 5000.0
    hot = rng.random(n) < 0.9
    u = np.where(hot, rng.uniform(0.1, 0.9, n) * binsize, rng.normal(0, 0.3 * G * binsize, n))
    v = np.where(hot, rng.uniform(0.1, 0.9, n) * binsize, rng.normal(0, 0.3 * G * binsize, n))
    freq = np.array([230e9])
    re, im = rng.normal(size=(n, nf)), rng.normal(size=(n, nf))
    w = rng.uniform(0.5, 2, (n, nf))
    d = Visibilities(u, v, freq, re, im, w)
    kw = dict(gridsize=G, binsize=binsize, convolution=convolution, imaging=True)
    a, _ = _quiet(grid, d, deterministic=True, **kw)
    b, _ = _quiet(grid, d, deterministic=False, **kw)
    for nm in ("real", "imag", "weights"):
        x, y = getattr(a, nm), getattr(b, nm)
        assert np.abs(x - y).max() <= 1e-11 * np.abs(x).max(), (G, convolution, nm)
