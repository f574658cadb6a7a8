"""average() and center(): GPU drop-ins for pdspy/interferometry/libinterferometry.pyx:151-311 and
pdspy/interferometry/center.py:5-25.

average(): one call into libpdsb (include/pdsb.h:pdsb_average).  The weight clamp, the uvdist and on-grid
filters, the numpy.round bin indices, the ordered (k, n) accumulation, the normalisation, the channel-weighted
mean positions and the compaction of the non-empty cells all run on the device, with numpy's operation order
and casts, so every output is bit-identical to the reference.  What stays here is O(gridsize) or O(nuv) and
transcendental: the bin centres, and for log-radial bins log10(uvdist), whose value must be the host libm's."""
import ctypes

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
    if mode not in ("continuum", "spectralline"):
        raise ValueError("mode must be 'continuum' or 'spectralline'")
    gridsize = int(gridsize)
    src = data
    if mfs and radial and log:
        src, mfs = freqcorrect(data), False          # log10 of the corrected uvdist has to come from the host libm
    nf = src.freq.size
    kind, log_uvdist, log_min, dtemp, centres = 0, None, 0.0, 0.0, None
    if radial and log:
        edges = numpy.linspace(numpy.log10(logmin), numpy.log10(logmax), gridsize + 1)
        centres = 10 ** ((edges[1:] + edges[0:-1]) / 2)
        kind, log_min, dtemp = 2, float(numpy.log10(logmin)), float(edges[1] - edges[0])
        with numpy.errstate(divide="ignore"):
            log_uvdist = numpy.log10(_lib.f64(src.uvdist))
    elif radial:
        kind, centres = 1, numpy.linspace(binsize / 2, (gridsize - 0.5) * binsize, gridsize)
    nch = (1 if mfs else nf) if mode == "spectralline" else 1
    npos = gridsize if radial else gridsize * gridsize
    out_u, out_v = numpy.empty(npos), numpy.empty(npos)
    out_re, out_im, out_w = (numpy.empty((npos, nch)) for _ in range(3))
    n_out, n_drop = ctypes.c_int64(), ctypes.c_int64()
    _lib.check(_lib.lib().pdsb_average(
        _lib.ptr(_lib.f64(src.u)), _lib.ptr(_lib.f64(src.v)), None if mfs else _lib.ptr(_lib.f64(src.uvdist)),
        _lib.ptr(log_uvdist), _lib.ptr(_lib.f64(src.freq)), _lib.ptr(_lib.f64(src.real)), _lib.ptr(_lib.f64(src.imag)),
        _lib.ptr(_lib.f64(src.weights)), src.u.size, nf, 1 if mfs else 0, float(src.freq.mean()) if mfs else 0.0,
        gridsize, float(binsize) if binsize is not None else 0.0, kind, log_min, dtemp,
        _lib.ptr(numpy.ascontiguousarray(centres)) if centres is not None else None,
        1 if mode == "spectralline" else 0, _lib.ptr(out_u), _lib.ptr(out_v), _lib.ptr(out_re), _lib.ptr(out_im),
        _lib.ptr(out_w), ctypes.byref(n_out), ctypes.byref(n_drop)))
    if n_drop.value > 0:
        print(_WARNING)
    n = n_out.value
    if mode == "continuum":
        freq = numpy.array([data.freq.sum() / data.freq.size])
    elif mfs:
        freq = numpy.array([data.freq.mean()])
    else:
        freq = src.freq.copy()
    return Visibilities(out_u[:n].copy(), out_v[:n].copy(), freq, out_re[:n].copy(), out_im[:n].copy(),
                        out_w[:n].copy())


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
