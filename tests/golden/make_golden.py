"""Generate the committed golden vectors from the LIVE reference (run in the build
container, where /root/reference exists):

    python tests/golden/make_golden.py

Outputs (all small):
  fixture_720.npz      the reference's only fixture, tests/testdata.hdf5, as float64 arrays
  grid_golden.npz      grid() of the live reference module (oracle/_ref, built from
                       pdspy/interferometry/libinterferometry.pyx with the reference's own
                       flags) on the fixture and on a seeded multi-channel set; outputs stored
                       sparsely (non-zero cells), index maps in full
  chisq_golden.npz     chisq() of the live reference
  dft_golden.npz       exact-DFT ORACLE outputs (oracle/dft.py) on the fixture's uv points.
                       NOT reference outputs: galario is unavailable (parity unpinned); kept as
                       a regression anchor for the restatement itself.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))          # tests/synth.py: the seeded synthetic inputs
sys.path.insert(0, HERE)

from hdf5_mini import read_hdf5_flat          # noqa: E402
from oracle import build_ref                  # noqa: E402
from oracle import dft as od                  # noqa: E402
import synth                  # noqa: E402

ARCSEC = 4.84813681e-6

GRID_CASES_FIXTURE = {
    # name: kwargs  (tests/test.py:9 runs invert(imsize=512, pixel_size=0.5, convolution="expsinc")
    # -> grid(gridsize=512, binsize=1/(0.5*512*arcsec), convolution="expsinc", imaging=True))
    "fx_pillbox_img256": dict(gridsize=256, binsize=1 / (0.5 * 256 * ARCSEC), convolution="pillbox", imaging=True),
    "fx_expsinc_img256": dict(gridsize=256, binsize=1 / (0.5 * 256 * ARCSEC), convolution="expsinc", imaging=True),
    "fx_expsinc_img512": dict(gridsize=512, binsize=1 / (0.5 * 512 * ARCSEC), convolution="expsinc", imaging=True),
    "fx_default": dict(gridsize=256, binsize=2000.0),
    "fx_uniform": dict(gridsize=256, binsize=500.0, convolution="expsinc", weighting="uniform"),
    "fx_superuniform": dict(gridsize=256, binsize=500.0, convolution="expsinc", weighting="superuniform"),
    "fx_robust": dict(gridsize=256, binsize=500.0, convolution="expsinc", weighting="robust", robust=2),
    "fx_odd_robust": dict(gridsize=255, binsize=500.0, convolution="pillbox", weighting="robust", robust=0.5,
                          npixels=1),
}
GRID_CASES_MULTI = {
    "mc_expsinc_line": dict(gridsize=128, binsize=8000.0, convolution="expsinc", mode="spectralline"),
    "mc_pillbox_line_uniform": dict(gridsize=128, binsize=8000.0, convolution="pillbox", mode="spectralline",
                                    weighting="uniform"),
    "mc_pillbox_cont": dict(gridsize=128, binsize=8000.0, convolution="pillbox", mode="continuum"),
    "mc_expsinc_mfs": dict(gridsize=128, binsize=8000.0, convolution="expsinc", mfs=True),
    "mc_expsinc_chan2_img": dict(gridsize=128, binsize=8000.0, convolution="expsinc", channel=2, imaging=True),
    "mc_pillbox_line_robust": dict(gridsize=64, binsize=8000.0, convolution="pillbox", mode="spectralline",
                                   weighting="robust", robust=0.5),
}


def multi_channel_set():
    """Seeded multi-channel data with zero and negative weights and points off the grid."""
    rng = np.random.default_rng(20240601)
    n, nf = 5000, 4
    u = rng.normal(0, 2e5, n)
    v = rng.normal(0, 2e5, n)
    freq = 230e9 + 1e8 * np.arange(nf)
    re = rng.normal(size=(n, nf))
    im = rng.normal(size=(n, nf))
    w = rng.uniform(0.5, 2, (n, nf))
    w[rng.random((n, nf)) < 0.01] = 0
    w[rng.random((n, nf)) < 0.001] *= -1
    re[17, 2] = 0.0
    im[17, 2] = 0.0                     # exercises the (real==0)&(imag==0) zeroing of :353
    return u, v, freq, re, im, w


SAMPLE_ABOVE = 4000      # cases with more non-zero cells than this store every 8th one


AVERAGE_CASES = {
    "av_cont": dict(gridsize=64, binsize=8000.0),
    "av_line_odd": dict(gridsize=65, binsize=8000.0, mode="spectralline"),
    "av_radial": dict(gridsize=20, binsize=40000.0, radial=True),
    "av_radial_log_line": dict(gridsize=20, radial=True, log=True, logmin=2400.0, logmax=1.2e6, mode="spectralline"),
    "av_mfs": dict(gridsize=64, binsize=8000.0, mfs=True),
}


def load_reference_python_module(ref, name):
    """Import one pure-Python module of the reference package (e.g. interferometry.center) without
    running pdspy/__init__.py (which pulls in galario, h5py, ...): fake parent packages whose
    __path__ points into /root/reference, with the compiled libinterferometry pre-seeded."""
    import importlib
    import types
    root = "/root/reference/pdspy"
    for pkg, path in (("pdspy", root), ("pdspy.interferometry", root + "/interferometry"),
                      ("pdspy.constants", root + "/constants")):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [path]
            sys.modules[pkg] = m
    sys.modules["pdspy.interferometry.libinterferometry"] = ref
    return importlib.import_module("pdspy." + name)


INVERT_CASES = {
    # tests/test.py:9 of the reference
    "test_py": dict(imsize=512, pixel_size=0.5, convolution="expsinc"),
    "pillbox_256": dict(imsize=256, pixel_size=0.5, convolution="pillbox"),
    "robust_centered_128": dict(imsize=128, pixel_size=1.0, convolution="expsinc", weighting="robust", robust=0.5,
                                centering=[0.3, -0.2, 1.0]),
    "beam_128": dict(imsize=128, pixel_size=1.0, convolution="expsinc", beam=True, uvtaper=30.0),
}


# image sizes that are not powers of two (the reference takes any: scipy's ifft2); kept in their own file,
# written by `python make_golden.py --invert-anysize`
INVERT_ANYSIZE_CASES = {
    "expsinc_300": dict(imsize=300, pixel_size=0.5, convolution="expsinc"),
    "pillbox_125": dict(imsize=125, pixel_size=1.0, convolution="pillbox"),
    "robust_centered_90": dict(imsize=90, pixel_size=1.0, convolution="expsinc", weighting="robust", robust=0.5,
                               centering=[0.3, -0.2, 1.0]),
    "beam_77": dict(imsize=77, pixel_size=1.5, convolution="expsinc", beam=True, uvtaper=30.0),
}


def load_reference_invert(ref):
    """The reference's own invert.py (scipy.fftpack) with the compiled grid(); pdspy.imaging is replaced
    by a minimal Image stub (the real one needs h5py/astropy)."""
    import importlib
    import types
    cen = load_reference_python_module(ref, "interferometry.center")
    sys.modules["pdspy.interferometry"].center = cen.center          # what `from . import center` resolves to
    if "pdspy.imaging" not in sys.modules:
        img = types.ModuleType("pdspy.imaging")

        class Image:
            def __init__(self, image, x=None, y=None, freq=None, **kw):
                self.image, self.x, self.y, self.freq = image, x, y, freq
        img.Image = Image
        sys.modules["pdspy.imaging"] = img
    return importlib.import_module("pdspy.interferometry.invert")


CLEAN_CASES = {
    "fx_expsinc_64": dict(imsize=64, pixel_size=1.0, convolution="expsinc", maxiter=60),
    "fx_pillbox_robust_64": dict(imsize=64, pixel_size=0.5, convolution="pillbox", maxiter=60, weighting="robust",
                                 robust=0.5),
    "two_sources_line_32": dict(imsize=32, pixel_size=0.5, convolution="expsinc", maxiter=40, mode="spectralline",
                                gain=0.2, nsigma=4.),
}


def two_source_set():
    """Two point sources with different spectra plus noise, 2 channels (for clean() in spectral-line mode)."""
    rng = np.random.default_rng(777)
    n = 3000
    u, v = rng.normal(0, 6e4, n), rng.normal(0, 6e4, n)
    freq = 230e9 + 1e8 * np.arange(2)
    vis = np.zeros((n, 2), dtype=complex)
    for (x0, y0), flux in (((1.5, -2.0), (1.0, 0.4)), ((-3.0, 1.0), (0.5, 0.8))):
        ph = np.exp(-2j * np.pi * (u * x0 * ARCSEC + v * y0 * ARCSEC))
        vis += ph[:, None] * np.array(flux)[None, :]
    vis += 0.3 * (rng.normal(size=vis.shape) + 1j * rng.normal(size=vis.shape))
    return u, v, freq, vis.real.copy(), vis.imag.copy(), np.ones((n, 2))


def load_reference_clean(ref):
    """The reference's own clean.py; astropy.stats.mad_std (astropy is absent) is supplied from its
    definition (oracle/clean.py:mad_std)."""
    import importlib
    import types
    from oracle import clean as ocl
    load_reference_invert(ref)
    st = types.ModuleType("astropy.stats")
    st.mad_std = ocl.mad_std
    sys.modules.setdefault("astropy", types.ModuleType("astropy"))
    sys.modules["astropy.stats"] = st
    return importlib.import_module("pdspy.interferometry.clean")


def sparse(a):
    """(indices, values) of the non-zero cells; sub-sampled for the dense expsinc maps so the
    committed file stays small.  The count and the plain sum of all non-zero cells are stored
    beside them (make_golden.run)."""
    flat = a.reshape(-1)
    idx = np.flatnonzero(flat)
    if idx.size > SAMPLE_ABOVE:
        idx = idx[::8]
    return idx.astype(np.int32), flat[idx]


def main():
    ref = build_ref.load()
    assert ref is not None, "reference tree not available"
    d = read_hdf5_flat("/root/reference/tests/testdata.hdf5")
    fx = {k: d[k].astype(np.float64) for k in ("u", "v", "freq", "real", "imag", "weights")}
    np.savez_compressed(os.path.join(HERE, "fixture_720.npz"), **fx)

    out = {}
    import contextlib
    import io

    def run(name, data, kw):
        with contextlib.redirect_stdout(io.StringIO()):
            g = ref.grid(data, **kw)
        for nm in ("real", "imag", "weights"):
            idx, val = sparse(getattr(g, nm))
            out["%s/%s_idx" % (name, nm)] = idx
            out["%s/%s_val" % (name, nm)] = val
            full = getattr(g, nm)
            out["%s/%s_nnz" % (name, nm)] = np.int64(np.count_nonzero(full))
            out["%s/%s_sum" % (name, nm)] = np.float64(full.sum())
        out["%s/shape" % name] = np.array(g.real.shape)
        out["%s/freq" % name] = g.freq
        out["%s/u_ends" % name] = g.u[[0, 1, -1]]

    data = ref.Visibilities(fx["u"], fx["v"], fx["freq"], fx["real"], fx["imag"], fx["weights"])
    for name, kw in GRID_CASES_FIXTURE.items():
        run(name, data, kw)
    u, v, freq, re, im, w = multi_channel_set()
    data = ref.Visibilities(u, v, freq, re, im, w)
    for name, kw in GRID_CASES_MULTI.items():
        run(name, data, kw)
    np.savez_compressed(os.path.join(HERE, "grid_golden.npz"), **out)

    # average() of the live reference on the multi-channel set (one exact-zero baseline added)
    av = {}
    u, v, freq, re, im, w = multi_channel_set()
    u[5] = 0.0
    v[5] = 0.0
    data = ref.Visibilities(u, v, freq, re, im, w)
    for name, kw in AVERAGE_CASES.items():
        with contextlib.redirect_stdout(io.StringIO()):
            a = ref.average(data, **kw)
        for nm in ("u", "v", "freq", "real", "imag", "weights"):
            av["%s/%s" % (name, nm)] = getattr(a, nm)
    # center() of the reference's own Python (center.py + model.py)
    cen = load_reference_python_module(ref, "interferometry.center")
    c = cen.center(data, [0.31, -0.17, 1.0])
    av["center/real"], av["center/imag"] = c.real, c.imag
    np.savez_compressed(os.path.join(HERE, "average_golden.npz"), **av)

    # invert() of the reference (invert.py + compiled grid + scipy.fftpack) on the fixture
    inv = load_reference_invert(ref)
    iv = {}
    for name, kw in INVERT_CASES.items():
        # fresh copies: invert(beam=True) overwrites the caller's real/imag arrays in place (:18-19)
        fxdata = ref.Visibilities(fx["u"].copy(), fx["v"].copy(), fx["freq"].copy(), fx["real"].copy(),
                                  fx["imag"].copy(), fx["weights"].copy())
        with contextlib.redirect_stdout(io.StringIO()):
            r = inv.invert(fxdata, **kw)
        im = r.image[:, :, 0, 0]
        iv[name + "/sample"] = im[::4, ::4].copy() if im.shape[0] > 128 else im.copy()     # keep the file small
        iv[name + "/stats"] = np.array([im.sum(), im.max(), im.min(), np.abs(im).sum()])
        iv[name + "/x"] = r.x
    np.savez_compressed(os.path.join(HERE, "invert_golden.npz"), **iv)

    # clean() of the reference (clean.py + invert.py + scipy fftconvolve / leastsq) on the fixture and on a
    # two-channel synthetic set: the loop's inputs (dirty image, dirty beam) and all its outputs
    cl = load_reference_clean(ref)
    cg = {}
    for name, kw in CLEAN_CASES.items():
        def fresh():
            a = two_source_set() if name.startswith("two") else [fx[k] for k in ("u", "v", "freq", "real", "imag", "weights")]
            return ref.Visibilities(*[np.array(x, dtype=np.float64) for x in a])
        with contextlib.redirect_stdout(io.StringIO()):
            ci, res, beam, model, mask = cl.clean(fresh(), **kw)
            inv_kw = {k: v for k, v in kw.items() if k not in ("maxiter", "gain", "nsigma")}
            dirty = inv.invert(fresh(), **inv_kw)
        cg[name + "/dirty"] = dirty.image[:, :, :, 0]
        cg[name + "/dirty_beam"] = beam.image[:, :, :, 0]
        cg[name + "/clean_image"] = ci.image[:, :, :, 0]
        cg[name + "/residuals"] = res.image[:, :, :, 0]
        cg[name + "/model"] = model.image[:, :, :, 0]
        cg[name + "/mask"] = mask.image[:, :, :, 0]
    np.savez_compressed(os.path.join(HERE, "clean_golden.npz"), **cg)

    # chisq of the live reference (channel 0, float return)
    rng = np.random.default_rng(99)
    m = ref.Visibilities(fx["u"], fx["v"], fx["freq"], fx["real"] * 0.9 + 0.01 * rng.normal(size=fx["real"].shape),
                         fx["imag"] * 1.1, fx["weights"])
    c1 = ref.chisq(ref.Visibilities(fx["u"], fx["v"], fx["freq"], fx["real"], fx["imag"], fx["weights"]), m)
    np.savez_compressed(os.path.join(HERE, "chisq_golden.npz"), m_real=m.real, m_imag=m.imag, chisq=np.float64(c1))

    # oracle DFT outputs on the fixture's uv (regression anchor; not reference output)
    img = synth.synth_image(64, 2, 0.5, kind="disk")
    vis = od.exact_dft_literal(fx["u"], fx["v"], img, 0.5 * ARCSEC, 0.05 * ARCSEC, -0.03 * ARCSEC)
    np.savez_compressed(os.path.join(HERE, "dft_golden.npz"), real=vis.real, imag=vis.imag)
    print("golden vectors written to", HERE)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(" ", f, os.path.getsize(os.path.join(HERE, f)))


def main_invert_anysize():
    import contextlib
    import io
    ref = build_ref.load()
    assert ref is not None, "reference tree not available"
    d = np.load(os.path.join(HERE, "fixture_720.npz"))
    fx = {k: d[k] for k in d.files}
    inv = load_reference_invert(ref)
    iv = {}
    for name, kw in INVERT_ANYSIZE_CASES.items():
        fxdata = ref.Visibilities(fx["u"].copy(), fx["v"].copy(), fx["freq"].copy(), fx["real"].copy(),
                                  fx["imag"].copy(), fx["weights"].copy())
        with contextlib.redirect_stdout(io.StringIO()):
            r = inv.invert(fxdata, **kw)
        im = r.image[:, :, 0, 0]
        iv[name + "/sample"] = im[::4, ::4].copy() if im.shape[0] > 128 else im.copy()
        iv[name + "/stats"] = np.array([im.sum(), im.max(), im.min(), np.abs(im).sum()])
        iv[name + "/x"] = r.x
    np.savez_compressed(os.path.join(HERE, "invert_anysize_golden.npz"), **iv)
    print("written", os.path.getsize(os.path.join(HERE, "invert_anysize_golden.npz")))


def main_galario():
    """The pin SURVEY.md 8c says is missing: run on a machine where `galario` IS installed (it is not in this
    container).  Writes galario_golden.npz = the literal statements of interpolate_model.py:18-27 on seeded inputs;
    tests/test_oracle_dft.py::test_galario_restatement_against_real_galario holds oracle/dft.py:galario_like to it."""
    import galario                                    # noqa: F401  (fails here: that is the point of the hook)
    sys.path.insert(0, os.path.join(os.path.dirname(HERE)))
    import synth
    arcsec = 4.84813681e-6                            # constants/astronomy.py:9
    out = {}
    d = np.load(os.path.join(HERE, "fixture_720.npz"))
    cases = {"fixture64": (d["u"], d["v"], synth.synth_image(64, 2, 0.5, kind="disk"), 0.5, 0.05, -0.03),
             "c1_256": (*synth.synth_uv(2000, 0.1 * arcsec), synth.synth_image(256, 1, 0.1, kind="disk"), 0.1, 0.05, -0.03),
             "random128": (*synth.synth_uv(500, 0.1 * arcsec), synth.synth_image(128, 3, 0.1, kind="random"), 0.1, 0.0, 0.0)}
    for name, (u, v, img, px, dra, ddec) in cases.items():
        u, v = np.ascontiguousarray(u, np.float64), np.ascontiguousarray(v, np.float64)
        dxy = px * arcsec
        real, imag = [], []
        galario.double.threads(1)                     # interpolate_model.py:18
        for i in range(img.shape[2]):                 # interpolate_model.py:22-27, verbatim
            vis = galario.double.sampleImage(img[::-1, :, i, 0].copy(order="C"), dxy, u, v, dRA=dra * arcsec,
                                             dDec=ddec * arcsec)
            real.append(vis.real.reshape((u.size, 1)))
            imag.append(-vis.imag.reshape((u.size, 1)))
        out[name + "/u"], out[name + "/v"] = u, v
        out[name + "/args"] = np.array([px, dra, ddec, img.shape[0], img.shape[2], 0 if name != "random128" else 1])
        out[name + "/real"], out[name + "/imag"] = np.concatenate(real, axis=1), np.concatenate(imag, axis=1)
    np.savez_compressed(os.path.join(HERE, "galario_golden.npz"), **out)
    print("written galario_golden.npz (galario %s)" % getattr(galario, "__version__", "?"))


if __name__ == "__main__":
    if "--invert-anysize" in sys.argv:
        main_invert_anysize()
    elif "--galario" in sys.argv:
        main_galario()
    else:
        main()
