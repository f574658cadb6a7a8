"""clean(): Hogbom CLEAN of the dirty image, GPU drop-in for pdspy/interferometry/clean.py:8-127
(used by the reference's plotting scripts only).

invert() of the data and of the beam (twice the size) are the GPU calls of invert.py here; the clean-beam
fit stays the reference's scipy.optimize.leastsq on the host (a 3-parameter fit); the iteration (masked
maximum, shifted-beam subtraction, mad_std stopping rule, mask updates) and the final restore run in
libpdsb (pdsb_clean_loop, pdsb_clean_restore).  Same signature and return tuple as the reference:
(clean_image, residuals, beam, model, mask)."""
import ctypes

import numpy

from .. import _lib
from ..imaging import Image
from .invert import invert


def mad_std(x):
    """astropy.stats.mad_std (clean.py:1) on the GPU: 1.482602218505602 * median(|x - median(x)|)."""
    a = numpy.ascontiguousarray(x, dtype=numpy.float64).ravel()
    out = ctypes.c_double()
    _lib.check(_lib.lib().pdsb_mad_std(_lib.ptr(a), a.size, _lib.HOST, ctypes.byref(out)))
    return out.value


def fit_clean_beam(dirty_beam):
    """clean.py:36-55: an elliptical Gaussian fitted to the main lobe (> 0.4) of each channel's beam."""
    from scipy.optimize import leastsq
    clean_beam = numpy.zeros(dirty_beam.shape)
    ny, nx, nfreq = dirty_beam.shape
    x, y = numpy.meshgrid(numpy.arange(nx) - nx / 2 + 1, numpy.arange(ny) - ny / 2)
    for i in range(nfreq):
        fitfunc = lambda p, x, y: numpy.exp(-(x * numpy.cos(p[2]) - y * numpy.sin(p[2])) ** 2 / (2 * p[0] ** 2) -
                                            (x * numpy.sin(p[2]) + y * numpy.cos(p[2])) ** 2 / (2 * p[1] ** 2))
        errfunc = lambda p, x, y, z, w: numpy.ravel((fitfunc(p, x, y) - z) * w)
        p0 = [0.5, 0.5, 0.]
        weights = numpy.abs(dirty_beam[:, :, i]) * (dirty_beam[:, :, i] > 0.4)
        p, success = leastsq(errfunc, p0, args=(x, y, dirty_beam[:, :, i], weights))
        clean_beam[:, :, i] = fitfunc(p, x, y)
    return clean_beam


def clean_arrays(dirty, dirty_beam, clean_beam, gain=0.1, maxiter=1000, nsigma=5.):
    """The loop and the restore on arrays: dirty [ny, nx, nf], beams [2ny, 2nx, nf].
    Returns (clean_image, residuals, model, mask, niter, threshold)."""
    dirty = numpy.array(dirty, dtype=numpy.float64, order="C")          # becomes the residuals
    dirty_beam = numpy.ascontiguousarray(dirty_beam, dtype=numpy.float64)
    clean_beam = numpy.ascontiguousarray(clean_beam, dtype=numpy.float64)
    ny, nx, nf = dirty.shape
    if dirty_beam.shape != (2 * ny, 2 * nx, nf) or clean_beam.shape != dirty_beam.shape:
        raise ValueError("beams must be [2*ny, 2*nx, nf]")
    model, mask = numpy.empty(dirty.shape), numpy.empty(dirty.shape)
    niter, thr = ctypes.c_int(), ctypes.c_double()
    L = _lib.lib()
    _lib.check(L.pdsb_clean_loop(_lib.ptr(dirty), _lib.ptr(dirty_beam), ny, nx, nf,
                                 float((dirty_beam - clean_beam).max()), float(gain), int(maxiter), float(nsigma),
                                 _lib.HOST, _lib.ptr(model), _lib.ptr(mask), ctypes.byref(niter), ctypes.byref(thr)))
    clean_image = numpy.empty(dirty.shape)
    _lib.check(L.pdsb_clean_restore(_lib.ptr(model), _lib.ptr(clean_beam), _lib.ptr(dirty), ny, nx, nf, _lib.HOST,
                                    _lib.ptr(clean_image)))
    return clean_image, dirty, model, mask, niter.value, thr.value


def clean(data, imsize=256, pixel_size=0.25, convolution="pillbox", mfs=False, weighting="natural", robust=2,
          npixels=0, centering=None, mode='continuum', gain=0.1, maxiter=1000, threshold=0.001, uvtaper=None,
          nsigma=5.):

    # First make the image, and an image of the beam (clean.py:15-27; the beam call does not pass uvtaper).
    image = invert(data, imsize=imsize, pixel_size=pixel_size, convolution=convolution, mfs=mfs, weighting=weighting,
                   robust=robust, npixels=npixels, centering=centering, mode=mode, uvtaper=uvtaper)
    beam = invert(data, imsize=2 * imsize, pixel_size=pixel_size, convolution=convolution, mfs=mfs,
                  weighting=weighting, robust=robust, npixels=npixels, centering=centering, mode=mode, beam=True)

    dirty = image.image[:, :, :, 0]
    dirty_beam = beam.image[:, :, :, 0]
    clean_beam = fit_clean_beam(dirty_beam)

    # `threshold` is overwritten before use in the reference as well (clean.py:57-58).
    clean_image, residuals, model, mask, n, thr = clean_arrays(dirty, dirty_beam, clean_beam, gain=gain,
                                                               maxiter=maxiter, nsigma=nsigma)
    print("Cleaning to a threshold of ", thr)

    shape4 = model.shape + (1,)
    model = Image(model.reshape(shape4), freq=data.freq)
    residuals = Image(residuals.reshape(shape4), freq=data.freq)
    clean_image = Image(clean_image.reshape(shape4), freq=data.freq)
    mask = Image(mask.astype(float).reshape(shape4), freq=data.freq)
    return clean_image, residuals, beam, model, mask
