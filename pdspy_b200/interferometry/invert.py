"""invert(): dirty imaging on the GPU, drop-in for pdspy/interferometry/invert.py:8-92 (SURVEY.md section 8f
rank 3; the reference's own smoke test tests/test.py:9 is exactly this call).

Pipeline: grid(..., imaging=True) -> optional center() -> per-channel 2-D inverse FFT, division by the
transform of the gridding kernel, x flip (pdsb_grid, pdsb_center, pdsb_invert_image: a hand-written fp64
FFT).  Host side only prepares small things: the image axes and the gridding kernel sampled on the uv grid.

Differences in HOW (not in what) from the reference: the beam / uv-taper variants of the data are built as
a separate Visibilities instead of overwriting the caller's arrays and restoring them afterwards
(invert.py:15-20, 24-29, 40-47), and the separable exp*sinc kernel map is formed as an outer product of
two 1-D profiles.  Any imsize up to 4096 (powers of two run the radix-2 transform directly, other sizes go
through Bluestein's algorithm on the same transform)."""
import numpy

from .. import _lib
from ..constants import arcsec
from ..imaging import Image
from .average import center
from .grid import grid
from .visibilities import Visibilities

_SINC_WIDTH, _EXP_WIDTH = 1.55, 2.52          # exp*sinc kernel of the gridder, in cells (invert.py:107-108)


def _imaging_data(data, beam, uvtaper):
    """The visibilities that actually get gridded: unit real part / zero imaginary part for the beam
    (invert.py:15-20), weights times a Gaussian taper of width `uvtaper` kilo-lambda (invert.py:24-29)."""
    if not beam and uvtaper is None:
        return data
    real, imag, weights = data.real, data.imag, data.weights
    if beam:
        real, imag = numpy.ones_like(data.real), numpy.zeros_like(data.imag)
    if uvtaper is not None:
        sigma = uvtaper * 1e3
        weights = data.weights * numpy.exp(-0.5 * data.uvdist ** 2 / sigma ** 2)[:, numpy.newaxis]
    return Visibilities(data.u, data.v, data.freq, real, imag, weights)


def _axis(imsize, cell):
    """Sky offsets of the image pixels in arcsec: fftshift(fftfreq(imsize, cell)) / arcsec (invert.py:57-58)."""
    return numpy.fft.fftshift(numpy.fft.fftfreq(imsize, cell)) / arcsec


def _profile_1d(x, cell, convolution):
    """One-dimensional factor of the gridding kernel at uv offsets x (invert.py:94-121)."""
    if convolution == "pillbox":
        return numpy.where(numpy.abs(x) < 0.5 * cell, 1.0, 0.0)
    g = numpy.sinc(x / (_SINC_WIDTH * cell)) * numpy.exp(-(x / (_EXP_WIDTH * cell)) ** 2)
    return numpy.where(numpy.abs(x) < 3.0 * cell, g, 0.0)


def kernel_map(convolution, uu, vv, cell):
    """The gridding kernel sampled at the uv-grid points (rows follow vv, columns uu), scaled as the
    reference scales it: pillbox = n_cells inside the central cell, exp*sinc normalised to sum n_cells."""
    if convolution not in ("pillbox", "expsinc"):
        raise ValueError("unknown convolution %r" % (convolution,))
    m = numpy.outer(_profile_1d(vv, cell, convolution), _profile_1d(uu, cell, convolution))
    if convolution == "pillbox":
        return m * m.size
    return m / m.sum() * m.size


def invert(data, imsize=256, pixel_size=0.25, convolution="pillbox", mfs=False, weighting="natural",
           robust=2, npixels=0, centering=None, mode='continuum', beam=False, uvtaper=None,
           deterministic=True):
    if imsize < 2 or imsize > 4096:
        raise NotImplementedError("the GPU invert() handles 2 <= imsize <= 4096 (one image row per thread block)")

    cell = 1.0 / (pixel_size * imsize * arcsec)              # uv cell of an image of imsize pixels (invert.py:33)
    gridded = grid(_imaging_data(data, beam, uvtaper), gridsize=imsize, binsize=cell, convolution=convolution,
                   mfs=mfs, imaging=True, weighting=weighting, robust=robust, npixels=npixels, mode=mode,
                   deterministic=deterministic)
    if centering is not None:
        gridded = center(gridded, centering)

    # grid() returns u, v flattened from meshgrid(uu, vv): the first row holds uu, the first column vv
    uu = gridded.u[:imsize]
    vv = gridded.v[::imsize]
    conv = numpy.ascontiguousarray(kernel_map(convolution, uu, vv, cell))

    nch = gridded.real.shape[1]
    image = numpy.empty((imsize, imsize, nch, 1))
    _lib.check(_lib.lib().pdsb_invert_image(_lib.ptr(_lib.f64(gridded.real)), _lib.ptr(_lib.f64(gridded.imag)),
                                            _lib.ptr(conv), imsize, nch, _lib.HOST, _lib.ptr(image)))
    if beam:                                                 # the beam peaks at exactly 1 (invert.py:84-87)
        image /= image.max(axis=(0, 1), keepdims=True)

    axis = _axis(imsize, cell)
    return Image(image, x=axis, y=axis.copy(), freq=gridded.freq)
