"""BASELINE.json configurations at FULL size on one B200, through size-independent properties
(the oracle cannot finish these sizes): configs[2] (512^2 x 64 ch x 1M uv likelihood),
configs[3] (10M visibilities onto 2048^2), and a slice of configs[4] (walker batch)."""
import contextlib
import ctypes
import io

import numpy as np
import pytest

from oracle import dft as od, likelihood as ol, grid as og
import synth
from pdspy_b200 import _lib, Dataset, DeviceBuffer
from pdspy_b200.interferometry import Visibilities, grid, loglike_image

pytestmark = pytest.mark.gpu
A = synth.ARCSEC


def test_config3_full_likelihood(gpu):
    """Full C3: the fused likelihood equals (a) chi^2 recomputed with numpy from the GPU's own
    visibilities (1e-12), (b) the oracle chain on a random 1024-point subset of the uv list, and is
    (c) linear: chi^2 of data == model is ~0 relative to chi^2 against zero data."""
    c = synth.make_config("C3")
    n, nf, nuv, px = c["npix"], c["nf"], c["u"].size, c["pixelsize"]
    cube = np.ascontiguousarray(c["model"].image[:, :, :, 0])
    ds = Dataset(c["u"], c["v"])
    dcube = DeviceBuffer.from_numpy(cube)
    dre, dim_ = DeviceBuffer(nuv * nf * 8), DeviceBuffer(nuv * nf * 8)
    _lib.check(gpu.pdsb_sample_image(ds.handle, _lib.ptr(dcube), n, n, nf, _lib.DEVICE, px * A, c["dRA"] * A,
                                     c["dDec"] * A, _lib.ptr(dre), _lib.ptr(dim_), _lib.DEVICE))
    vre, vim = dre.download((nuv, nf)), dim_.download((nuv, nf))
    # (b) subset parity against the exact oracle, relative to the channel maxima of the full set
    sub = np.random.default_rng(3).choice(nuv, 1024, replace=False)
    ref = od.exact_dft(c["u"][sub], c["v"][sub], c["model"].image, px * A, c["dRA"] * A, c["dDec"] * A)
    scale = np.abs(vre + 1j * vim).max(axis=0)
    assert (np.abs((vre + 1j * vim)[sub] - ref) / scale).max() < 1e-5
    # Hermitian halves
    h = nuv // 2
    assert np.array_equal(vre[h:], vre[:h]) and np.array_equal(vim[h:], -vim[:h])
    # (a) fused likelihood vs numpy on the GPU's own visibilities
    re, im, w = synth.synth_data(nuv, nf)
    ds.set_data(re, im, w)
    chi2 = np.empty(nf)
    ll = ctypes.c_double()
    _lib.check(gpu.pdsb_loglike(ds.handle, _lib.ptr(dcube), n, n, nf, _lib.DEVICE, px * A, c["dRA"] * A, c["dDec"] * A,
                                _lib.ptr(chi2), ctypes.cast(ctypes.byref(ll), ctypes.c_void_p)))
    exp = ol.lnlike_vis_numpy(re, im, w, vre, vim)
    assert abs(ll.value - exp) <= 1e-11 * abs(exp)
    exp_chi2 = ol.chi2_per_channel_numpy(re, im, w, vre, vim)
    assert np.all(np.abs(chi2 - exp_chi2) <= 1e-11 * np.abs(exp_chi2))
    # (c) data == model: chi^2 collapses to rounding level
    ds.set_data(vre, vim, np.abs(w) + 0.5)
    _lib.check(gpu.pdsb_loglike(ds.handle, _lib.ptr(dcube), n, n, nf, _lib.DEVICE, px * A, c["dRA"] * A, c["dDec"] * A,
                                _lib.ptr(chi2), ctypes.cast(ctypes.byref(ll), ctypes.c_void_p)))
    assert chi2.max() == 0.0          # same kernels, same order: bit-identical model both times


@pytest.mark.parametrize("workload", ["C2", "C3"])
def test_every_uv_point_against_the_fp64_kernel(gpu, workload):
    """BASELINE.json configs[1] (1024^2 onto 1M uv points) and configs[2] (512^2 x 64 channels onto 1M uv points)
    IN FULL: the FP32-pipe default and the tcgen05 tensor-core kernel agree with the fp64 reference kernel (variant
    300, itself 1e-11 from the CPU oracle, re-checked here on a random subset) on EVERY point and channel
    within the 1e-5 bound."""
    from pdspy_b200.interferometry import interpolate_model
    c = synth.make_config(workload)

    def run(var):
        _lib.check(gpu.pdsb_set_dft_variant(var))
        try:
            v = interpolate_model(c["u"], c["v"], c["freq"], c["model"], dRA=c["dRA"], dDec=c["dDec"])
        finally:
            _lib.check(gpu.pdsb_set_dft_variant(0))
        return v.real, v.imag

    rr, ri = run(300)
    # the independent check of what the three kernels share (Hermitian fold, epilogue): 2048 random points out of both
    # halves of the uv list against the CPU oracle (about 3 s of host time for C3)
    sub = np.random.default_rng(5).choice(c["u"].size, 2048, replace=False)
    exact = od.exact_dft(c["u"][sub], c["v"][sub], c["model"].image, c["pixelsize"] * A, c["dRA"] * A, c["dDec"] * A)
    scale = np.sqrt((rr * rr + ri * ri).max(axis=0))
    assert (np.abs((rr[sub] + 1j * ri[sub]) - exact) / scale).max() < 1e-10
    for name, var in (("fp32", 0), ("tcgen05", 200)):
        vr, vi = run(var)
        vr -= rr
        vi -= ri
        err = (np.sqrt((vr * vr + vi * vi).max(axis=0)) / scale).max()
        del vr, vi
        assert err < 1e-5, (name, err)
        assert err < 3e-6, (name, err)                 # what these kernels actually deliver
    # the NUFFT path (an FFT-cost route to the same exact transform) on every point as well
    v = interpolate_model(c["u"], c["v"], c["freq"], c["model"], dRA=c["dRA"], dDec=c["dDec"], code="nufft")
    vr, vi = v.real - rr, v.imag - ri
    err = (np.sqrt((vr * vr + vi * vi).max(axis=0)) / scale).max()
    assert err < 3e-7, ("nufft", err)


def _star_cube(c):
    """The normal pdspy cube: a star pixel 1e5 x the disk peak in every channel (Model.py:522-527 puts the star
    on pixel (n/2, n/2)), plus one all-zero channel."""
    img = c["model"].image
    n = img.shape[0]
    img[n // 2, n // 2, :, 0] = 1e5 * img.max()
    if img.shape[2] > 3:
        img[:, :, 3, 0] = 0.0


@pytest.mark.parametrize("workload,star", [("C2", False), ("C3", False), ("C3", True)])
def test_lnlike_against_the_fp64_kernel(gpu, workload, star):
    """Full configs[1] / configs[2]: the fused log-likelihood and the per-channel chi^2 of the FP32-pipe default and
    of the tcgen05 kernel against the same call on the fp64 kernel, north-star bound 1e-7 PER CHANNEL for both
    (measured: FP32 1e-11 / 3e-9, tcgen05 6e-9 / 4e-8 - its lattice operand split leaves one truncation per
    accumulator round).  star=True: the high-dynamic-range case (a star pixel 1e5 x the disk peak,
    one empty channel); the visibilities are then one pixel's fp32 rounding (6e-8, the same for every uv point)
    from the fp64 result, so the bound there is what fp32 storage of the image allows: 4e-7."""
    import pdspy_b200 as pb
    c = synth.make_config(workload)
    if star:
        _star_cube(c)
    re, im, w = synth.synth_data(c["u"].size, c["nf"])
    d = Visibilities(c["u"], c["v"], c["freq"], re, im, w)
    out = {}
    try:
        for k in ("fp64", "fp32", "tcgen05", "nufft"):
            pb.set_dft_kernel(k)
            out[k] = loglike_image(d, c["model"], dRA=c["dRA"], dDec=c["dDec"])
    finally:
        pb.set_dft_kernel("fp32")
    ll0, chi0 = out["fp64"]
    assert np.all(np.isfinite(chi0))
    bound = 4e-7 if star else 1e-7
    for k in ("fp32", "tcgen05", "nufft"):
        ll, chi = out[k]
        assert abs(ll - ll0) <= bound * abs(ll0), (k, abs(ll - ll0) / abs(ll0))
        assert np.all(np.abs(chi - chi0) <= bound * np.abs(chi0)), (k, (np.abs(chi - chi0) / np.abs(chi0)).max())
    ll, chi = out["nufft"]                             # fp64 throughout; kernel accuracy 4e-8 of max|V|
    assert abs(ll - ll0) <= 1e-7 * abs(ll0) and np.all(np.abs(chi - chi0) <= 1e-7 * np.abs(chi0))
    if not star:
        ll, chi = out["fp32"]
        assert abs(ll - ll0) <= 1e-9 * abs(ll0)
        assert np.all(np.abs(chi - chi0) <= 2e-8 * np.abs(chi0))


def test_star_cube_visibilities(gpu):
    """High dynamic range at full C3 size: star pixel 1e5 x the disk peak and an all-zero channel; every kernel
    within 1e-5 of max|V| of the fp64 kernel on every point, the empty channel exactly zero."""
    from pdspy_b200.interferometry import interpolate_model
    c = synth.make_config("C3", nuv=200_000)
    _star_cube(c)
    out = {}
    for name, var in (("fp64", 300), ("fp32", 0), ("tcgen05", 200), ("nufft", 400)):
        _lib.check(gpu.pdsb_set_dft_variant(var))
        try:
            v = interpolate_model(c["u"], c["v"], c["freq"], c["model"], dRA=c["dRA"], dDec=c["dDec"])
        finally:
            _lib.check(gpu.pdsb_set_dft_variant(0))
        out[name] = v.real + 1j * v.imag
    scale = np.abs(out["fp64"]).max(axis=0)
    assert scale[3] == 0.0
    scale[3] = 1.0
    for name in ("fp32", "tcgen05", "nufft"):
        assert np.all(out[name][:, 3] == 0.0), name
        assert (np.abs(out[name] - out["fp64"]) / scale).max() < 1e-5, name


def test_config4_full_gridding(gpu):
    """Full C4: 10M visibilities onto 2048^2.  Index maps equal numpy's; pillbox ordered == fast mode to
    rounding; total weight conserved; expsinc imaging map sums to 1; a 1%-subset re-gridded by the oracle
    equals gridding the same subset on the GPU bit for bit (pillbox)."""
    px = 0.01
    nvis, G = 10_000_000, 2048
    u, v = synth.synth_uv(nvis, px * A)
    re, im, w = synth.synth_data(nvis, 1)
    freq = synth.synth_freq(1)
    d = Visibilities(u, v, freq, re, im, w)
    binsize = 2.2 * np.hypot(u, v).max() / G
    with contextlib.redirect_stdout(io.StringIO()):
        g, gi, gj, wm = grid(d, gridsize=G, binsize=binsize, return_maps=True, deterministic=True)
        gf = grid(d, gridsize=G, binsize=binsize, deterministic=False)
        ge = grid(d, gridsize=G, binsize=binsize, convolution="expsinc", imaging=True, deterministic=False)
    i_ref, j_ref = og.index_maps(u, v, freq, G, binsize)
    assert np.array_equal(gi, i_ref) and np.array_equal(gj, j_ref)
    assert np.abs(gf.weights - g.weights).max() <= 1e-12 * g.weights.max()
    assert np.abs(gf.real - g.real).max() <= 1e-9 * np.abs(g.real).max()
    assert abs(g.weights.sum() - wm.sum()) <= 1e-9 * wm.sum()
    assert abs(ge.weights.sum() - 1.0) < 1e-12
    sel = np.arange(0, nvis, 100)
    ds = Visibilities(u[sel].copy(), v[sel].copy(), freq, re[sel].copy(), im[sel].copy(), w[sel].copy())
    with contextlib.redirect_stdout(io.StringIO()):
        gs = grid(ds, gridsize=G, binsize=binsize)
        o = og.grid(ds.u, ds.v, freq, ds.real, ds.imag, ds.weights, gridsize=G, binsize=binsize)
    assert np.array_equal(gs.weights, o[5]) and np.array_equal(gs.real, o[3]) and np.array_equal(gs.imag, o[4])


def test_config5_walker_slice(gpu):
    """configs[4] shape, 3 walkers of the 128: the batch entry point equals three single calls."""
    c = synth.make_config("C3", nuv=100_000)
    n, nf, px = c["npix"], c["nf"], c["pixelsize"]
    re, im, w = synth.synth_data(c["u"].size, nf)
    ds = Dataset(c["u"], c["v"])
    ds.set_data(re, im, w)
    base = np.ascontiguousarray(c["model"].image[:, :, :, 0])
    cubes = np.stack([base * (1 + 0.01 * k) for k in range(3)])
    dra = np.ascontiguousarray(np.array([0.0, 0.01, -0.02]) * A)
    ddec = np.ascontiguousarray(np.array([0.0, -0.01, 0.03]) * A)
    out = np.empty(3)
    _lib.check(gpu.pdsb_loglike_batch(ds.handle, _lib.ptr(cubes), 3, n, n, nf, _lib.HOST, px * A, _lib.ptr(dra),
                                      _lib.ptr(ddec), _lib.ptr(out)))
    for k in range(3):
        one = ctypes.c_double()
        _lib.check(gpu.pdsb_loglike(ds.handle, _lib.ptr(np.ascontiguousarray(cubes[k])), n, n, nf, _lib.HOST, px * A,
                                    float(dra[k]), float(ddec[k]), None, ctypes.cast(ctypes.byref(one), ctypes.c_void_p)))
        assert one.value == out[k]
