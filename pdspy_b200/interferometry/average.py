"""average() and center(): GPU drop-ins for pdspy/interferometry/libinterferometry.pyx:151-311 and
pdspy/interferometry/center.py:5-25.

average(): the numpy preamble (weight clamp, uvdist / grid filters, bin indices with numpy.round
and log10) and the final normalisation + compaction are the reference's own numpy expressions, kept
on the host because their rounding is part of the result; the Cython accumulation loop
(:262-277) - the only O(nuv*nfreq) interpreted part - runs in libpdsb as ordered sums
(include/pdsb.h:pdsb_bin_average), bit-exact against the reference."""
import numpy

from .. import _lib
from ..constants import arcsec
from .grid import freqcorrect
from .visibilities import Visibilities

_WARNING = ("WARNING: uv.grid was supplied with a gridsize and binsize that do not cover the full range of "
            "the input data in the uv-plane and is cutting baselines that are outside of this grid. Make sure "
            "to check your results carefully.")


def average(data, gridsize=256, binsize=None, radial=False, log=False, logmin=None, logmax=None, mfs=False,
            mode="continuum"):
    if mfs:
        vis = freqcorrect(data)
        u, v, uvdist, freq = vis.u, vis.v, vis.uvdist, vis.freq
        real, imag, weights = vis.real, vis.imag, vis.weights
    else:
        u, v, uvdist, freq = data.u.copy(), data.v.copy(), data.uvdist.copy(), data.freq.copy()
        real, imag, weights = data.real.copy(), data.imag.copy(), data.weights

    # Set the weights equal to 0 when the point is flagged (i.e. weight < 0)   (:179)
    weights = numpy.where(weights < 0, 0.0, weights)
    # Set the weights equal to 0 when the real and imaginary parts are both 0   (:181)
    weights[(real == 0) & (imag == 0)] = 0.0

    good_data = uvdist != 0.0
    u, v, uvdist = u[good_data], v[good_data], uvdist[good_data]
    real, imag, weights = real[good_data, :], imag[good_data, :], weights[good_data, :]

    nfreq = freq.size
    if mode == "continuum":
        nchannels = 1
    elif mode == "spectralline":
        nchannels = freq.size
    else:
        raise ValueError("mode must be 'continuum' or 'spectralline'")

    # bins (:205-248)
    if radial:
        if log:
            temp = numpy.linspace(numpy.log10(logmin), numpy.log10(logmax), gridsize + 1)
            new_u = 10 ** ((temp[1:] + temp[0:-1]) / 2)
        else:
            new_u = numpy.linspace(binsize / 2, (gridsize - 0.5) * binsize, gridsize)
        new_u = new_u.reshape((1, gridsize))
        new_v = numpy.zeros((1, gridsize))
        if log:
            dtemp = temp[1] - temp[0]
            i = numpy.round((numpy.log10(uvdist) - numpy.log10(logmin)) / dtemp - 0.5).astype(numpy.uint32)
            j = numpy.zeros(uvdist.size).astype(numpy.uint32)
        else:
            i = numpy.round(uvdist / binsize).astype(numpy.uint32)
            j = numpy.zeros(uvdist.size).astype(numpy.uint32)
    else:
        if gridsize % 2 == 0:
            i = numpy.round(u / binsize + gridsize / 2.).astype(numpy.uint32)
            j = numpy.round(v / binsize + gridsize / 2.).astype(numpy.uint32)
        else:
            i = numpy.round(u / binsize + (gridsize - 1) / 2.).astype(numpy.uint32)
            j = numpy.round(v / binsize + (gridsize - 1) / 2.).astype(numpy.uint32)

    good_i = numpy.logical_and(i >= 0, i < gridsize)
    good_j = numpy.logical_and(j >= 0, j < gridsize)
    good = numpy.logical_and(good_i, good_j)
    if good.sum() < good.size:
        print(_WARNING)

    u, v = u[good], v[good]
    real, imag, weights = real[good, :], imag[good, :], weights[good, :]
    i, j = i[good], j[good]
    nuv = u.size

    gj = 1 if radial else gridsize
    shape = (gj, gridsize, nchannels)
    new_real, new_imag, new_weights = numpy.empty(shape), numpy.empty(shape), numpy.empty(shape)
    acc_u = None if radial else numpy.empty(shape)
    acc_v = None if radial else numpy.empty(shape)

    # the accumulation loop (:262-277) on the GPU, in the reference's (k, n) order
    L = _lib.lib()
    _lib.check(L.pdsb_bin_average(_lib.ptr(numpy.ascontiguousarray(i)), _lib.ptr(numpy.ascontiguousarray(j)),
                                  _lib.ptr(_lib.f64(u)), _lib.ptr(_lib.f64(v)), _lib.ptr(_lib.f64(real)),
                                  _lib.ptr(_lib.f64(imag)), _lib.ptr(_lib.f64(weights)), nuv, nfreq, gridsize, gj,
                                  1 if mode == "spectralline" else 0, 1 if radial else 0, _lib.HOST,
                                  _lib.ptr(acc_u), _lib.ptr(acc_v), _lib.ptr(new_real), _lib.ptr(new_imag),
                                  _lib.ptr(new_weights), _lib.HOST))
    if not radial:
        new_u, new_v = acc_u, acc_v

    # normalisation and compaction (:279-311), the reference's numpy expressions
    good_data = new_weights != 0.0
    new_real[good_data] = new_real[good_data] / new_weights[good_data]
    new_imag[good_data] = new_imag[good_data] / new_weights[good_data]
    if not radial:
        new_u[good_data] = new_u[good_data] / new_weights[good_data]
        new_v[good_data] = new_v[good_data] / new_weights[good_data]

    good_data = numpy.any(good_data, axis=2)
    if not radial:
        new_u = (new_u * new_weights).sum(axis=2)[good_data] / new_weights.sum(axis=2)[good_data]
        new_v = (new_v * new_weights).sum(axis=2)[good_data] / new_weights.sum(axis=2)[good_data]
    else:
        new_u = new_u[good_data]
        new_v = new_v[good_data]

    good_data = numpy.dstack([good_data for m in range(nchannels)])

    if mode == "continuum":
        freq = numpy.array([data.freq.sum() / data.freq.size])

    return Visibilities(new_u, new_v, freq,
                        new_real[good_data].reshape((new_u.size, nchannels)),
                        new_imag[good_data].reshape((new_u.size, nchannels)),
                        new_weights[good_data].reshape((new_u.size, nchannels)))


def center(data, params):
    """center.py:5-25: multiply the data by the conjugate of a unit point source at
    (params[0], params[1]) arcsec, per channel at u*freq[i]/mean(freq)."""
    if type(params) == list:
        params = numpy.array(params)
    x0, y0 = float(params[0]) * arcsec, float(params[1]) * arcsec      # model.py:60-61 `par *= arcsec`
    nuv, nf = data.real.shape
    real, imag = numpy.empty((nuv, nf)), numpy.empty((nuv, nf))
    if nuv > 0:
        L = _lib.lib()
        _lib.check(L.pdsb_center(_lib.ptr(_lib.f64(data.u)), _lib.ptr(_lib.f64(data.v)), _lib.ptr(_lib.f64(data.freq)),
                                 _lib.ptr(_lib.f64(data.real)), _lib.ptr(_lib.f64(data.imag)), nuv, nf,
                                 float(data.freq.mean()), x0, y0, _lib.HOST, _lib.ptr(real), _lib.ptr(imag)))
    return Visibilities(data.u.copy(), data.v.copy(), data.freq.copy(), real, imag, data.weights.copy())
