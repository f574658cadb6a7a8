"""invert(): dirty imaging, GPU drop-in for pdspy/interferometry/invert.py:8-92 (SURVEY.md section 8f
rank 3; the reference's own smoke test tests/test.py:9 is exactly this call).

grid(..., imaging=True), center() and the per-channel 2-D inverse FFT + gridding correction + x flip
run in libpdsb (pdsb_grid, pdsb_center, pdsb_invert_image: a hand-written fp64 FFT); the pieces that are
O(imsize^2) numpy set-up in the reference (image axes from fftfreq, the convolution function sampled on
the uv grid :94-121) stay the reference's numpy expressions.  imsize must be a power of two."""
import numpy

from .. import _lib
from ..constants import arcsec
from ..imaging import Image
from .average import center
from .grid import grid


def invert(data, imsize=256, pixel_size=0.25, convolution="pillbox", mfs=False, weighting="natural",
           robust=2, npixels=0, centering=None, mode='continuum', beam=False, uvtaper=None,
           deterministic=True):

    if imsize & (imsize - 1) or imsize > 4096:
        raise NotImplementedError("the GPU invert() needs imsize to be a power of two <= 4096")

    # If we are calculating the beam, set all of the real values to 1 and the imaginary data to 0
    # (in place, restored below: exactly what the reference does, invert.py:15-20,40-42).
    if beam:
        real = data.real.copy()
        imag = data.imag.copy()
        data.real[:, :] = 1.
        data.imag[:, :] = 0.

    if type(uvtaper) != type(None):
        taper = numpy.exp(-0.5 * data.uvdist ** 2 / (uvtaper * 1e3) ** 2)
        weights = data.weights.copy()
        for i in range(data.freq.size):
            data.weights[:, i] *= taper

    binsize = 1.0 / (pixel_size * imsize * arcsec)
    try:
        gridded_data = grid(data, gridsize=imsize, binsize=binsize, convolution=convolution, mfs=mfs, imaging=True,
                            weighting=weighting, robust=robust, npixels=npixels, mode=mode,
                            deterministic=deterministic)
    finally:
        if beam:
            data.real = real
            data.imag = imag
        if type(uvtaper) != type(None):
            data.weights = weights

    if type(centering) != type(None):
        gridded_data = center(gridded_data, centering)

    # fftfreq(n, d) = [0, 1, ..., n/2-1, -n/2, ..., -1] / (d*n)   (scipy.fftpack.fftfreq)
    x = numpy.fft.fftshift(numpy.fft.fftfreq(imsize, binsize)) / arcsec
    y = numpy.fft.fftshift(numpy.fft.fftfreq(imsize, binsize)) / arcsec

    u = gridded_data.u.reshape((imsize, imsize))
    v = gridded_data.v.reshape((imsize, imsize))
    conv_func = {"pillbox": pillbox, "expsinc": exp_sinc}[convolution]
    conv = numpy.ascontiguousarray(conv_func(u, v, binsize, binsize), dtype=numpy.float64)

    nch = gridded_data.real.shape[1]
    image = numpy.empty((imsize, imsize, nch, 1))
    L = _lib.lib()
    _lib.check(L.pdsb_invert_image(_lib.ptr(_lib.f64(gridded_data.real)), _lib.ptr(_lib.f64(gridded_data.imag)),
                                   _lib.ptr(conv), imsize, nch, _lib.HOST, _lib.ptr(image)))

    # Make sure the beam peaks at exactly 1 (invert.py:84-87).
    if beam:
        for i in range(nch):
            image[:, :, i, 0] /= image[:, :, i, 0].max()

    return Image(image, x=x, y=y, freq=gridded_data.freq)


def pillbox(u, v, delta_u, delta_v):
    """invert.py:94-103."""
    m = 1
    arr = numpy.ones(u.shape, dtype=float) * u.size
    arr[numpy.abs(u) >= m * delta_u / 2] = 0
    arr[numpy.abs(v) >= m * delta_v / 2] = 0
    return arr


def exp_sinc(u, v, delta_u, delta_v):
    """invert.py:105-121."""
    alpha1 = 1.55
    alpha2 = 2.52
    m = 6
    arr = numpy.sinc(u / (alpha1 * delta_u)) * numpy.exp(-1 * (u / (alpha2 * delta_u)) ** 2) * \
        numpy.sinc(v / (alpha1 * delta_v)) * numpy.exp(-1 * (v / (alpha2 * delta_v)) ** 2)
    arr[numpy.abs(u) >= m * delta_u / 2] = 0
    arr[numpy.abs(v) >= m * delta_v / 2] = 0
    arr = arr / arr.sum() * arr.size
    return arr
