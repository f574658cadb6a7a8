"""chi^2 / log-likelihood on the GPU against the reference expressions.
Tolerance (north_star): chi-squared within 1e-7 relative (fp64 reduction)."""
import ctypes
import os

import numpy as np
import pytest

from oracle import dft as od, likelihood as ol
import synth
from pdspy_b200 import _lib, utils
from pdspy_b200.interferometry import (interpolate_model, Visibilities, chisq, loglike_image, loglike_images)

pytestmark = pytest.mark.gpu
A = synth.ARCSEC


def _arrays(n, nf, seed=0):
    rng = np.random.default_rng(seed)
    d_re, d_im, m_re, m_im = (rng.normal(size=(n, nf)) for _ in range(4))
    w = rng.uniform(0.5, 2, (n, nf))
    w[rng.random((n, nf)) < 0.01] = 0.0
    w[rng.random((n, nf)) < 0.001] *= -1
    return d_re, d_im, w, m_re, m_im


@pytest.mark.parametrize("n,nf", [(1, 1), (720, 1), (5000, 7), (100000, 64)])
def test_visibility_lnlike_vs_verbatim_numpy(gpu, n, nf):
    d_re, d_im, w, m_re, m_im = _arrays(n, nf, seed=n)
    u = np.zeros(n)
    f = np.ones(nf)
    got = utils.visibility_lnlike(Visibilities(u, u, f, d_re, d_im, w), Visibilities(u, u, f, m_re, m_im, w))
    ref = ol.lnlike_vis_numpy(d_re, d_im, w, m_re, m_im)
    assert abs(got - ref) <= 1e-7 * abs(ref)
    assert abs(got - ref) <= 1e-11 * abs(ref)       # in practice: fp64 both sides


def test_pieces_of_pdsb_chi2(gpu):
    d_re, d_im, w, m_re, m_im = _arrays(4000, 3, seed=9)
    out = np.empty(4)
    _lib.check(gpu.pdsb_chi2(_lib.ptr(d_re), _lib.ptr(d_im), _lib.ptr(w), _lib.ptr(m_re), _lib.ptr(m_im), d_re.size,
                             _lib.HOST, _lib.ptr(out)))
    np.testing.assert_allclose(out[0], np.sum((d_re - m_re) ** 2 * w), rtol=1e-12)
    np.testing.assert_allclose(out[1], np.sum((d_im - m_im) ** 2 * w), rtol=1e-12)
    np.testing.assert_allclose(out[2], np.sum(np.log(w[w > 0] / (2 * np.pi))), rtol=1e-12)
    assert out[3] == -0.5 * out[0] - out[2] + -0.5 * out[1] - out[2]


def test_chisq_vs_live_reference_golden(gpu, fixture720):
    f = fixture720
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chisq_golden.npz"))
    d = Visibilities(f["u"], f["v"], f["freq"], f["real"], f["imag"], f["weights"])
    m = Visibilities(f["u"], f["v"], f["freq"], g["m_real"], g["m_imag"], f["weights"])
    got = chisq(d, m)
    # the reference returns a C float: equal after the same rounding (last-ulp ties aside)
    assert abs(got - float(g["chisq"])) <= 1.2e-7 * float(g["chisq"])
    assert got == np.float32(got)


def test_chisq_uses_channel_zero_only(gpu):
    d_re, d_im, w, m_re, m_im = _arrays(500, 4, seed=2)
    u, f = np.zeros(500), np.ones(4)
    got = chisq(Visibilities(u, u, f, d_re, d_im, w), Visibilities(u, u, f, m_re, m_im, w))
    ref = ol.chisq_c(d_re, d_im, w, m_re, m_im)
    assert abs(got - ref) <= 1.2e-7 * abs(ref)


@pytest.mark.parametrize("cfg,nuv", [("C1", None), ("C3", 8192)])
def test_fused_loglike_vs_oracle_chain(gpu, cfg, nuv):
    """image -> exact oracle visibilities -> verbatim emcee expression, against the fused GPU
    call where the model visibilities never leave the device."""
    c = synth.make_config(cfg, nuv=nuv)
    ref = od.exact_dft(c["u"], c["v"], c["model"].image, c["pixelsize"] * A, c["dRA"] * A, c["dDec"] * A)
    re, im, w = synth.synth_data(c["u"].size, c["nf"], model=(ref.real, ref.imag))
    data = Visibilities(c["u"], c["v"], c["freq"], re, im, w)
    ll, chi2 = loglike_image(data, c["model"], dRA=c["dRA"], dDec=c["dDec"])
    ll_ref = ol.lnlike_vis_numpy(re, im, w, ref.real, ref.imag)
    chi2_ref = ol.chi2_per_channel_numpy(re, im, w, ref.real, ref.imag)
    assert abs(ll - ll_ref) <= 1e-7 * abs(ll_ref)
    assert np.all(np.abs(chi2 - chi2_ref) <= 1e-6 * np.abs(chi2_ref))
    # and the unfused route through interpolate_model + visibility_lnlike agrees with the fused one
    vis = interpolate_model(c["u"], c["v"], c["freq"], c["model"], dRA=c["dRA"], dDec=c["dDec"])
    ll2 = utils.visibility_lnlike(data, vis)
    assert abs(ll - ll2) <= 1e-12 * abs(ll2)


def test_walker_batch_matches_single_calls(gpu):
    c = synth.make_config("C3", nuv=2048)
    rng = np.random.default_rng(0)
    re, im, w = synth.synth_data(2048, c["nf"])
    data = Visibilities(c["u"], c["v"], c["freq"], re, im, w)
    base = c["model"].image[:, :, :, 0]
    cubes = np.stack([base * s for s in (1.0, 0.7, 1.3)])
    dra, ddec = np.array([0.0, 0.01, -0.02]), np.array([0.0, -0.01, 0.03])
    px = c["model"].x[1] - c["model"].x[0]          # what interpolate_model derives dxy from (:20)
    ll = loglike_images(data, cubes, dra, ddec, pixelsize=px)
    for k in range(3):
        m = synth.SynthImage(np.ascontiguousarray(cubes[k][:, :, :, None]), c["pixelsize"], c["freq"])
        one, _ = loglike_image(data, m, dRA=dra[k], dDec=ddec[k])
        assert ll[k] == one


@pytest.mark.parametrize("kernel", ["fp32", "tcgen05", "nufft"])
@pytest.mark.parametrize("n,nf", [(256, 32), (64, 4)])
def test_walker_batch_pipeline_keeps_cubes_apart(gpu, kernel, n, nf):
    """pdsb_loglike_batch uploads cube k+1 on a copy stream while cube k is evaluated (two device buffers): seven
    different cubes, every likelihood bit-identical to its own single call.  256x256x32 cubes (16.8 MB) travel
    through the staged pinned ring, 64x64x4 ones directly."""
    import pdspy_b200
    rng = np.random.default_rng(5)
    nuv, W = 1500, 7
    px = 0.02
    u, v = synth.synth_uv(nuv, px * A, seed=3)
    freq = synth.synth_freq(nf)
    re, im, w = synth.synth_data(nuv, nf)
    data = Visibilities(u, v, freq, re, im, w)
    base = synth.synth_image(n, nf, px)[:, :, :, 0]
    cubes = np.stack([base * rng.uniform(0.5, 1.5) + 1e-3 * rng.random(base.shape) for _ in range(W)])
    dra, ddec = rng.uniform(-0.05, 0.05, W), rng.uniform(-0.05, 0.05, W)
    m0 = synth.SynthImage(np.ascontiguousarray(cubes[0][:, :, :, None]), px, freq)
    pxm = m0.x[1] - m0.x[0]                         # what the single call derives dxy from (an ulp from px)
    pdspy_b200.set_dft_kernel(kernel)
    try:
        ll = loglike_images(data, cubes, dra, ddec, pixelsize=pxm)
        again = loglike_images(data, cubes, dra, ddec, pixelsize=pxm)
        for k in range(W):
            m = synth.SynthImage(np.ascontiguousarray(cubes[k][:, :, :, None]), px, freq)
            one, _ = loglike_image(data, m, dRA=dra[k], dDec=ddec[k])
            assert ll[k] == one and again[k] == one
        assert len(set(ll.tolist())) == W
    finally:
        pdspy_b200.set_dft_kernel("fp32")


def test_lnlike_signature_with_injected_model_runner(gpu):
    """utils.emcee.lnlike keeps the reference's signature (emcee.py:6-9); the model runner
    (RADMC-3D orchestration, out of scope) is injected."""
    c = synth.make_config("C1", nuv=2000)
    re, im, w = synth.synth_data(2000, 1)
    data = Visibilities(c["u"], c["v"], c["freq"], re, im, w)
    visibilities = {"file": ["a"], "lam": ["1300"], "data": [data]}

    class M:
        pass

    def runner(vis, images, spectra, params, parameters, plot, **kw):
        m = M()
        m.visibilities = {"1300": interpolate_model(c["u"], c["v"], c["freq"], c["model"], dRA=params["x0"],
                                                    dDec=params["y0"], code=kw["ftcode"])}
        return m

    got = utils.emcee.lnlike({"x0": 0.05, "y0": -0.03}, visibilities, {"file": []}, {}, {}, False, run_model=runner)
    ref = od.exact_dft(c["u"], c["v"], c["model"].image, c["pixelsize"] * A, 0.05 * A, -0.03 * A)
    exp = ol.lnlike_vis_numpy(re, im, w, ref.real, ref.imag)
    assert abs(got - exp) <= 1e-7 * abs(exp)
    assert utils.emcee.lnlike({}, visibilities, {"file": []}, {}, {}, False, run_model=lambda *a, **k: 0.) == -np.inf
    pars = {"x0": {"fixed": False}, "y0": {"fixed": False}, "z": {"fixed": True}}
    got2 = utils.dynesty.lnlike([0.05, -0.03], visibilities, {"file": []}, {}, pars, False, run_model=runner)
    assert got2 == got


def test_model_modifiers_fused_into_the_epilogue(gpu):
    """flux_unc scaling, per-channel extinction and the free-free point source (SURVEY.md section 8f
    rank 2) applied on the device equal the reference's host passes around interpolate_model:
    image *= flux_unc (run_disk_model.py:319); image[:,:,i] *= extinction[i] (run_flared_model.py:286-299);
    real += F*exp(-2*3.14159*1j*(u*x0+v*y0)).real (run_disk_model.py:329-334)."""
    from pdspy_b200.interferometry import model_visibilities
    c = synth.make_config("C3", nuv=3000)
    nf = c["nf"]
    rng = np.random.default_rng(4)
    ext = np.exp(-rng.uniform(0, 2, nf))
    flux_unc, F, x0, y0 = 1.07, 0.0123, c["dRA"], c["dDec"]
    img = c["model"].image * flux_unc
    img = img * ext[None, None, :, None]
    ref = od.exact_dft(c["u"], c["v"], img, c["pixelsize"] * A, x0 * A, y0 * A)
    ff = F * np.exp(-2 * 3.14159 * (0 + 1j * (c["u"] * x0 * A + c["v"] * y0 * A)))
    ref_re = ref.real + ff.real[:, None]
    vis = model_visibilities(c["u"], c["v"], c["freq"], c["model"], dRA=x0, dDec=y0, flux_unc=flux_unc, extinction=ext,
                             freefree=F)
    scale = np.abs(ref).max(axis=0)
    assert (np.abs(vis.real - ref_re) / scale).max() < 1e-5 and (np.abs(vis.imag - ref.imag) / scale).max() < 1e-5
    re, im, w = synth.synth_data(3000, nf, model=(ref_re, ref.imag))
    data = Visibilities(c["u"], c["v"], c["freq"], re, im, w)
    ll, _ = loglike_image(data, c["model"], dRA=x0, dDec=y0, flux_unc=flux_unc, extinction=ext, freefree=F)
    exp = ol.lnlike_vis_numpy(re, im, w, ref_re, ref.imag)
    assert abs(ll - exp) <= 1e-7 * abs(exp)
    # and without modifiers nothing changes
    a = model_visibilities(c["u"], c["v"], c["freq"], c["model"], dRA=x0, dDec=y0)
    b = interpolate_model(c["u"], c["v"], c["freq"], c["model"], dRA=x0, dDec=y0)
    assert np.array_equal(a.real, b.real) and np.array_equal(a.imag, b.imag)


def test_sharded_likelihood_single_rank_and_cube_staging(gpu):
    """pdspy_b200.dist.ShardedLikelihood on one rank (no process group): equals the fused single call, for the
    host cube, the device cube and the staged-cube paths (which degenerate to one upload here); the N-rank
    behaviour is scripts/gpu_dist_cube.py under torchrun and the gloo tests of tests/test_dist.py."""
    import torch
    from pdspy_b200 import dist as pdist, DeviceBuffer
    n, nf, nuv = 96, 2, 3000
    u, v = synth.synth_uv(nuv, 0.03 * A)
    re, im, w = synth.synth_data(nuv, nf)
    data = Visibilities(u, v, synth.synth_freq(nf), re, im, w)
    img = synth.synth_image(n, nf, 0.03)
    m = synth.SynthImage(img, 0.03, synth.synth_freq(nf))
    ll0, _ = loglike_image(data, m, dRA=0.02, dDec=-0.01)
    try:
        like = pdist.ShardedLikelihood(pdist.shard_visibilities(data, 0, 1))
        cube = np.ascontiguousarray(img[:, :, :, 0])
        dxy = (m.x[1] - m.x[0]) * A
        vals = [like(cube, dxy, 0.02 * A, -0.01 * A, kind=0, cube=mode) for mode in (None, "sharded", "rank0")]
        vals.append(like(DeviceBuffer.from_numpy(cube), dxy, 0.02 * A, -0.01 * A, kind=1, shape=(n, n)))
        torch.cuda.synchronize()
    finally:
        _lib.check(gpu.pdsb_reset_stream())
    assert all(abs(x - ll0) <= 1e-12 * abs(ll0) for x in vals), (vals, ll0)


@pytest.mark.parametrize("kernel", ["fp32", "nufft"])
@pytest.mark.parametrize("world", [1, 3, 4])
def test_channel_partition_emulated_on_one_gpu(gpu, kernel, world):
    """ShardedLikelihood(channels=...) for every rank of a `world`-way channel split, run one after the other on
    this GPU: the nf + 1 doubles the ranks would all-reduce sum to the unsharded likelihood, each rank's chi^2 sits
    in its own channel window, and pdsb_channel_slice hands out exactly the cube's channels."""
    import pdspy_b200
    import torch
    from pdspy_b200 import dist as pdist
    n, nf, nuv = 64, 32, 4000
    u, v = synth.synth_uv(nuv, 0.03 * A)
    re, im, w = synth.synth_data(nuv, nf)
    freq = synth.synth_freq(nf)
    data = Visibilities(u, v, freq, re, im, w)
    img = synth.synth_image(n, nf, 0.03) + 1e-4
    m = synth.SynthImage(img, 0.03, freq)
    cube = np.ascontiguousarray(img[:, :, :, 0])
    dxy = (m.x[1] - m.x[0]) * A
    pdspy_b200.set_dft_kernel(kernel)
    try:
        ll0, chi2_0 = loglike_image(data, m, dRA=0.02, dDec=-0.01)
        total = torch.zeros(nf + 1, dtype=torch.float64, device="cuda")
        for r in range(world):
            c0, c1 = pdist.shard_channels(nf, r, world)
            like = pdist.ShardedLikelihood(pdist.shard_visibilities(data, r, world, by="channels"), channels=(c0, nf))
            buf = like.chi2_device(cube, n, n, _lib.HOST, dxy, 0.02 * A, -0.01 * A)
            got = buf.cpu().numpy()
            assert np.all(got[:c0] == 0) and np.all(got[c1:nf] == 0)
            np.testing.assert_allclose(got[c0:c1], chi2_0[c0:c1], rtol=1e-12)
            sl = like._slice.cpu().numpy().reshape(n, n, c1 - c0)
            assert np.array_equal(sl, cube[:, :, c0:c1])
            total += buf
            like.ds.destroy()
        t = total.cpu().numpy()
        ll = -0.5 * t[:-1].sum() - 2.0 * t[-1]
        assert abs(ll - ll0) <= 1e-12 * abs(ll0)
    finally:
        pdspy_b200.set_dft_kernel("fp32")
        _lib.check(gpu.pdsb_reset_stream())
    with pytest.raises(ValueError):
        pdist.ShardedLikelihood(pdist.shard_visibilities(data, 0, 2, by="channels"), channels=(20, nf))
    _lib.check(gpu.pdsb_reset_stream())


def test_unchanged_signature_chain_stays_on_the_device(gpu):
    """interpolate_model(...) -> utils.visibility_lnlike(data, model) (what utils.emcee.lnlike runs, emcee.py:31-43):
    the model visibilities are consumed from the device; reading them afterwards, or pickling the object, gives
    ordinary numpy arrays equal to what the likelihood saw."""
    import pickle
    from pdspy_b200 import device
    c = synth.make_config("C3", nuv=20000)
    re, im, w = synth.synth_data(c["u"].size, c["nf"])
    data = Visibilities(c["u"], c["v"], c["freq"], re, im, w)
    m = interpolate_model(c["u"], c["v"], c["freq"], c["model"], dRA=c["dRA"], dDec=c["dDec"])
    assert m._device_token is not None and device.model_buffers(m._device_token) is not None
    ll = utils.visibility_lnlike(data, m)
    assert m._device_token is not None                      # the likelihood did not pull the arrays to the host
    fused, _ = loglike_image(data, c["model"], dRA=c["dRA"], dDec=c["dDec"])
    assert abs(ll - fused) <= 1e-12 * abs(fused)
    clone = pickle.loads(pickle.dumps(m))                   # __reduce__ materialises
    assert m._device_token is None and type(clone).__name__ == "VisibilitiesObject"
    assert clone.real.shape == (c["u"].size, c["nf"]) and np.array_equal(clone.real, m.real) and np.all(clone.weights == 1)
    ref = ol.lnlike_vis_numpy(re, im, w, m.real, m.imag)
    assert abs(ll - ref) <= 1e-11 * abs(ref)
    assert abs(utils.visibility_lnlike(data, m) - ll) <= 1e-13 * abs(ll)      # host arrays now: same value
    # a result nobody reads gives its buffers back when it dies
    t = interpolate_model(c["u"], c["v"], c["freq"], c["model"])._device_token
    import gc
    gc.collect()
    assert device.model_buffers(t) is None


@pytest.mark.parametrize("policy", ["hash", "freeze"])
def test_in_place_edits_of_cached_arrays(gpu, policy):
    """The reference edits arrays in place (invert.py:15-47).  Default policy: any single-element edit of the data
    is seen by the next call.  "freeze": the arrays are read-only while cached, an edit raises instead of being
    missed, release() hands them back."""
    import pdspy_b200 as pb
    c = synth.make_config("C1", nuv=30000)
    re, im, w = synth.synth_data(c["u"].size, c["nf"], seed=5)
    data = Visibilities(c["u"], c["v"], c["freq"], re, im, w)
    pb.clear_cache()
    pb.set_cache_policy(policy)
    try:
        ll0, _ = loglike_image(data, c["model"])
        k = 12347                                            # not one of round 1's 4096 strided samples
        if policy == "freeze":
            assert not data.weights.flags.writeable
            with pytest.raises(ValueError):
                data.weights[k, 0] = 0.0
            pb.release(data.weights)
            assert data.weights.flags.writeable
        assert data.weights[k, 0] > 0
        data.weights[k, 0] = 0.0
        data.real[k + 1, 0] += 1.0
        ll1, _ = loglike_image(data, c["model"])
        fresh = Visibilities(c["u"].copy(), c["v"].copy(), c["freq"], re.copy(), im.copy(), w.copy())
        ll2, _ = loglike_image(fresh, c["model"])
        assert ll1 != ll0 and ll1 == ll2
    finally:
        pb.set_cache_policy("hash")
        pb.clear_cache()
