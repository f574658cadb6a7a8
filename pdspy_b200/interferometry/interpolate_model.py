"""interpolate_model: image cube -> model visibilities at (u, v), on the GPU.

Drop-in for pdspy/interferometry/interpolate_model.py:11-57 (same name, positional and keyword
arguments, return type).  The reference hands the arithmetic to galario (FFT + bilinear
interpolation, :23-24); here `code="galario"` selects the exact direct transform on the GPU
(include/pdsb.h:pdsb_sample_image), which is what galario's interpolation approximates, in the
same conventions (row flip :23, dxy from model.x :20, arcsec-scaled dRA/dDec :24, imag -> -imag
:27, unit weights :57)."""
import ctypes

import numpy

from ..constants import arcsec
from .. import _lib
from ..device import dataset_for, register_model
from .visibilities import Visibilities
from .cube import postprocess_channels_device
from .unstructured import regrid


def _cube(model):
    image = model.image
    if image.ndim != 4:
        raise ValueError("model.image must be [ny, nx, nfreq, npol]")
    if image.shape[3] != 1:
        image = image[:, :, :, 0:1]          # the reference reads polarisation 0 (:23)
    return numpy.ascontiguousarray(image, dtype=numpy.float64)


def interpolate_model(u, v, freq, model, nthreads=1, dRA=0., dDec=0.,
                      code="galario", nxy=1024, dxy=0.01):

    if code == "galario":
        # nthreads is galario's OpenMP thread count (:18); the GPU path has no use for it.
        dxy = (model.x[1] - model.x[0]) * arcsec

        image = _cube(model)
        ny, nx, nf = image.shape[:3]
        if nf != len(model.freq):
            raise ValueError("model.image has %d channels but model.freq has %d" % (nf, len(model.freq)))

        u = numpy.ascontiguousarray(u, dtype=numpy.float64)
        v = numpy.ascontiguousarray(v, dtype=numpy.float64)
        ds = dataset_for(u, v)

        if u.size > 0:
            # the result stays on the device until somebody reads it (visibilities.py); utils.emcee.lnlike does not
            L = _lib.lib()
            token, dre, dim_ = register_model((u.size, nf))
            _lib.check(L.pdsb_sample_image(ds.handle, _lib.ptr(image), ny, nx, nf, _lib.HOST,
                                           float(dxy), float(dRA * arcsec), float(dDec * arcsec),
                                           _lib.ptr(dre), _lib.ptr(dim_), _lib.DEVICE))
            return Visibilities._from_device(u, v, freq, token, (u.size, nf))
        real = numpy.empty((u.size, nf))
        imag = numpy.empty((u.size, nf))

    elif code == "galario-fft":
        # EXTENSION (not a value the reference knows): galario's own algorithm - FFT of the image + bilinear
        # interpolation at the uv points (what code="galario" runs in the reference) - on the GPU, for numerical
        # continuity with galario-based results; it carries galario's interpolation error, which the exact
        # transform of code="galario" here does not.
        dxy = (model.x[1] - model.x[0]) * arcsec
        image = _cube(model)
        ny, nx, nf = image.shape[:3]
        if ny != nx:
            raise ValueError("the FFT path needs a square image (as galario does)")
        u = numpy.ascontiguousarray(u, dtype=numpy.float64)
        v = numpy.ascontiguousarray(v, dtype=numpy.float64)
        ds = dataset_for(u, v)
        real = numpy.empty((u.size, nf))
        imag = numpy.empty((u.size, nf))
        if u.size > 0:
            L = _lib.lib()
            _lib.check(L.pdsb_sample_image_fft(ds.handle, _lib.ptr(image), nx, nf, _lib.HOST, float(dxy),
                                               float(dRA * arcsec), float(dDec * arcsec), _lib.ptr(real),
                                               _lib.ptr(imag), _lib.HOST))

    elif code == "nufft":
        # EXTENSION: the exact transform of code="galario" here (what the reference's galario call approximates) by
        # a type-2 non-uniform FFT with an 8-point kernel: 4e-8 of max|V| instead of galario's 1e-3..4e-2, at the
        # cost of an FFT path, not of a direct sum (vis.cu).  The result stays on the device.
        dxy = (model.x[1] - model.x[0]) * arcsec
        image = _cube(model)
        ny, nx, nf = image.shape[:3]
        if ny % 2 or nx % 2:
            raise ValueError("the NUFFT path needs even image sides (odd sides: code=\"galario\", the direct transform)")
        u = numpy.ascontiguousarray(u, dtype=numpy.float64)
        v = numpy.ascontiguousarray(v, dtype=numpy.float64)
        ds = dataset_for(u, v)
        if u.size > 0:
            L = _lib.lib()
            token, dre, dim_ = register_model((u.size, nf))
            _lib.check(L.pdsb_sample_image_nufft(ds.handle, _lib.ptr(image), ny, nx, nf, _lib.HOST, float(dxy),
                                                 float(dRA * arcsec), float(dDec * arcsec), _lib.ptr(dre),
                                                 _lib.ptr(dim_), _lib.DEVICE))
            return Visibilities._from_device(u, v, freq, token, (u.size, nf))
        real = numpy.empty((u.size, nf))
        imag = numpy.empty((u.size, nf))

    elif code == "galario-unstructured":
        # scattered points -> piecewise-linear interpolant on the nxy x nxy grid of dxy arcsec (Jy/pixel),
        # then the same transform as above (unstructured.py; parity unpinned: galario fork not available)
        from ..imaging import Image
        cube = regrid(model, int(nxy), float(dxy))
        x = (numpy.arange(int(nxy)) - int(nxy) // 2) * float(dxy)
        return interpolate_model(u, v, freq, Image(cube, x=x, y=x.copy(), freq=numpy.asarray(model.freq)),
                                 nthreads=nthreads, dRA=dRA, dDec=dDec, code="galario")

    elif code == "trift":
        # the exact transform of the Delaunay piecewise-linear interpolant of the scattered points (what the
        # third-party `trift` package computes in its "extended" mode, :49-55; parity unpinned: trift.py)
        from . import trift as _trift
        u = numpy.ascontiguousarray(u, dtype=numpy.float64)
        v = numpy.ascontiguousarray(v, dtype=numpy.float64)
        ds = dataset_for(u, v)
        nf = numpy.shape(model.image)[1] if numpy.ndim(model.image) == 2 else 0
        if u.size > 0:
            token, dre, dim_ = register_model((u.size, nf))
            _trift.sample(ds, u, v, model, dRA, dDec, dre, dim_, _lib.DEVICE)
            return Visibilities._from_device(u, v, freq, token, (u.size, nf))
        real = numpy.empty((0, nf))
        imag = numpy.empty((0, nf))
    else:
        # the reference falls through to an UnboundLocalError on `real`; be explicit instead
        raise ValueError("unknown code %r" % (code,))

    return Visibilities(u, v, freq, real, imag, numpy.ones(real.shape))


def _mods(nf, flux_unc, extinction, freefree, dRA, dDec):
    scale = None
    if flux_unc != 1.0 or extinction is not None:
        scale = numpy.full(nf, float(flux_unc))
        if extinction is not None:
            scale = scale * numpy.asarray(extinction, dtype=numpy.float64)
        scale = numpy.ascontiguousarray(scale)
    ff = float(freefree) if freefree is not None else 0.0
    return scale, ff, float(dRA * arcsec), float(dDec * arcsec)


def _staged_cube(model, subsample, averaging, hanning, extinction):
    """(pointer, kind, ny, nx, nf, keepalive, extinction left for the epilogue) of the cube the transform reads:
    the host cube as is, or, when channel post-processing is asked for, its post-processed copy left on the
    device.  The reference multiplies the RENDERED channels by extinction[i] (run_flared_model.py:286-299) before
    the sub-sample mean / Hanning smoothing / binning (:308-366); the two do not commute, so with
    post-processing the extinction (one entry per rendered channel) goes into that pass, and only without it
    can it ride on the transform's epilogue as a per-channel factor."""
    image = _cube(model)
    ny, nx, nf = image.shape[:3]
    if subsample * averaging > 1 or hanning:                  # run_flared_model.py:308-309
        buf, nf = postprocess_channels_device(image, subsample, averaging, hanning, in_scale=extinction)
        return _lib.ptr(buf), _lib.DEVICE, ny, nx, nf, buf, None
    if extinction is not None and numpy.shape(extinction) != (nf,):
        raise ValueError("extinction must have one entry per cube channel (%d)" % nf)
    return _lib.ptr(image), _lib.HOST, ny, nx, nf, image, extinction


def model_visibilities(u, v, freq, model, dRA=0., dDec=0., flux_unc=1.0, extinction=None, freefree=None,
                       subsample=1, averaging=1, hanning=False):
    """interpolate_model plus the host passes the reference wraps around it, in one device pass:
    image *= flux_unc (run_disk_model.py:319), image[:,:,i] *= extinction[i] (run_flared_model.py:286-299),
    the sub-sample mean / Hanning smoothing / channel binning of run_flared_model.py:308-366 (`freq` is then
    the data's channel list, as there), real += point-source free-free flux at (dRA, dDec)
    (run_disk_model.py:329-334).  Returns Visibilities."""
    dxy = (model.x[1] - model.x[0]) * arcsec
    image, image_kind, ny, nx, nf, _keep, extinction = _staged_cube(model, subsample, averaging, hanning, extinction)
    u = numpy.ascontiguousarray(u, dtype=numpy.float64)
    v = numpy.ascontiguousarray(v, dtype=numpy.float64)
    ds = dataset_for(u, v)
    real, imag = numpy.empty((u.size, nf)), numpy.empty((u.size, nf))
    scale, ff, x0, y0 = _mods(nf, flux_unc, extinction, freefree, dRA, dDec)
    if u.size > 0:
        L = _lib.lib()
        _lib.check(L.pdsb_sample_image_ex(ds.handle, image, ny, nx, nf, image_kind, float(dxy), x0, y0,
                                          _lib.ptr(scale), ff, x0, y0, _lib.ptr(real), _lib.ptr(imag), _lib.HOST))
    return Visibilities(u, v, freq, real, imag, numpy.ones(real.shape))


def loglike_image_fft(data, model, dRA=0., dDec=0.):
    """interpolate_model(code="galario-fft") + the visibility log-likelihood term in one device pass
    (pdsb_loglike_fft).  Returns (lnlike, chi2_real, chi2_imag)."""
    dxy = (model.x[1] - model.x[0]) * arcsec
    image = _cube(model)
    ny, nx, nf = image.shape[:3]
    if ny != nx:
        raise ValueError("the FFT path needs a square image (as galario does)")
    ds = dataset_for(data.u, data.v, (data.real, data.imag, data.weights))
    out = numpy.empty(4)
    _lib.check(_lib.lib().pdsb_loglike_fft(ds.handle, _lib.ptr(image), nx, nf, _lib.HOST, float(dxy),
                                          float(dRA * arcsec), float(dDec * arcsec), _lib.ptr(out)))
    return float(out[3]), float(out[0]), float(out[1])


def loglike_image_nufft(data, model, dRA=0., dDec=0.):
    """interpolate_model(code="nufft") + the visibility log-likelihood term in one device pass
    (pdsb_loglike_nufft).  Returns (lnlike, chi2_real, chi2_imag)."""
    dxy = (model.x[1] - model.x[0]) * arcsec
    image = _cube(model)
    ny, nx, nf = image.shape[:3]
    if ny % 2 or nx % 2:
        raise ValueError("the NUFFT path needs even image sides")
    ds = dataset_for(data.u, data.v, (data.real, data.imag, data.weights))
    out = numpy.empty(4)
    _lib.check(_lib.lib().pdsb_loglike_nufft(ds.handle, _lib.ptr(image), ny, nx, nf, _lib.HOST, float(dxy),
                                            float(dRA * arcsec), float(dDec * arcsec), _lib.ptr(out)))
    return float(out[3]), float(out[0]), float(out[1])


def loglike_image(data, model, dRA=0., dDec=0., flux_unc=1.0, extinction=None, freefree=None,
                  subsample=1, averaging=1, hanning=False):
    """Fused interpolate_model + visibility log-likelihood term (pdspy/utils/emcee.py:31-43) for
    one dataset: the model visibilities never leave the GPU.  `data` is a Visibilities with
    u, v, real, imag, weights; `model` an Image.  flux_unc / extinction / freefree as in
    model_visibilities (as are subsample / averaging / hanning).  Returns (lnlike, chi2_per_channel)."""
    dxy = (model.x[1] - model.x[0]) * arcsec
    image, image_kind, ny, nx, nf, _keep, extinction = _staged_cube(model, subsample, averaging, hanning, extinction)
    ds = dataset_for(data.u, data.v, (data.real, data.imag, data.weights))
    chi2 = numpy.empty(nf)
    out = ctypes.c_double()
    scale, ff, x0, y0 = _mods(nf, flux_unc, extinction, freefree, dRA, dDec)
    L = _lib.lib()
    _lib.check(L.pdsb_loglike_ex(ds.handle, image, ny, nx, nf, image_kind, float(dxy), x0, y0,
                                 _lib.ptr(scale), ff, x0, y0, _lib.ptr(chi2),
                                 ctypes.cast(ctypes.byref(out), ctypes.c_void_p)))
    return out.value, chi2


def loglike_images(data, models, dRA, dDec, pixelsize=None):
    """Walker batch: `models` is a sequence of Images of identical shape, or one array
    [W, ny, nx, nf] with `pixelsize` (arcsec) given; dRA, dDec per walker (arcsec).
    Returns lnlike[W]."""
    if isinstance(models, numpy.ndarray):
        cubes = numpy.ascontiguousarray(models, dtype=numpy.float64)
        if cubes.ndim != 4:
            raise ValueError("cube batch must be [W, ny, nx, nf]")
        if pixelsize is None:
            raise ValueError("pixelsize (arcsec) is required with a raw cube batch")
        dxy = pixelsize * arcsec
    else:
        cubes = numpy.stack([_cube(m)[:, :, :, 0] for m in models])
        dxy = (models[0].x[1] - models[0].x[0]) * arcsec
    return _loglike_batch(data, cubes, dxy, dRA, dDec)


def _loglike_batch(data, cubes, dxy, dRA, dDec):
    W, ny, nx, nf = cubes.shape
    dRA = numpy.ascontiguousarray(numpy.broadcast_to(numpy.asarray(dRA, dtype=numpy.float64) * arcsec, (W,)))
    dDec = numpy.ascontiguousarray(numpy.broadcast_to(numpy.asarray(dDec, dtype=numpy.float64) * arcsec, (W,)))
    ds = dataset_for(data.u, data.v, (data.real, data.imag, data.weights))
    out = numpy.empty(W)
    L = _lib.lib()
    _lib.check(L.pdsb_loglike_batch(ds.handle, _lib.ptr(cubes), W, ny, nx, nf, _lib.HOST, float(dxy),
                                    _lib.ptr(dRA), _lib.ptr(dDec), _lib.ptr(out)))
    return out
