"""grid(), freqcorrect(), chisq(): GPU drop-ins for the Cython functions of
pdspy/interferometry/libinterferometry.pyx (:313-541, :587-608, :610-633).

Same names, arguments, defaults, return types and the same warning text.  The numpy parts
whose rounding is part of the reference result (numpy.linspace cell centres :370-381) stay
numpy on the host; everything per-visibility runs in libpdsb (include/pdsb.h:pdsb_grid)."""
import ctypes
import threading

import numpy

from .. import _lib
from .visibilities import Visibilities

_WARNING = ("WARNING: uv.grid was supplied with a gridsize and binsize that do not cover the full range of "
            "the input data in the uv-plane and is cutting baselines that are outside of this grid. Make sure "
            "to check your results carefully.")


def freqcorrect(data, freq=None):
    """libinterferometry.pyx:587-608."""
    if freq != None:                                     # noqa: E711  (as the reference writes it)
        new_freq = numpy.array([freq], dtype=numpy.float64)
    else:
        new_freq = numpy.array([data.freq.mean()])

    nuv, nf = data.u.size, data.freq.size
    new_u = numpy.empty(nuv * nf)
    new_v = numpy.empty(nuv * nf)
    if nuv > 0:
        L = _lib.lib()
        _lib.check(L.pdsb_freqcorrect(_lib.ptr(_lib.f64(data.u)), _lib.ptr(_lib.f64(data.v)),
                                      _lib.ptr(_lib.f64(data.freq)), nuv, nf, float(new_freq[0]), _lib.HOST,
                                      _lib.ptr(new_u), _lib.ptr(new_v)))

    new_real = data.real.reshape((data.real.size, 1))
    new_imag = data.imag.reshape((data.imag.size, 1))
    new_weights = data.weights.reshape((data.weights.size, 1))

    return Visibilities(new_u, new_v, new_freq, new_real, new_imag, new_weights)


class MeshFiller:
    """numpy.meshgrid(uu, vv) of the result's coordinates (libinterferometry.pyx:383), optionally on a second host
    thread while the device works; an exception raised there (MemoryError) resurfaces in result()."""

    def __init__(self, uu, vv, threaded=True):
        self._out, self._err = None, None
        self._thread = threading.Thread(target=self._run, args=(uu, vv)) if threaded else None
        if self._thread is not None:
            self._thread.start()
        else:
            self._run(uu, vv)

    def _run(self, uu, vv):
        try:
            self._out = numpy.meshgrid(uu, vv)
        except BaseException as e:                   # noqa: BLE001 - handed to the caller's thread
            self._err = e

    def join(self):
        if self._thread is not None:
            self._thread.join()

    def result(self):
        self.join()
        if self._err is not None:
            raise self._err
        return self._out


def grid(data, gridsize=256, binsize=2000.0, convolution="pillbox", mfs=False, channel=None,
         imaging=False, weighting="natural", robust=2, npixels=0, mode="continuum",
         deterministic=True, return_maps=False):
    """Convolutional gridding.  Arguments as the reference's grid().  Two extra keywords:
    deterministic (default True: reference summation order, bit-exact pillbox maps; False:
    fp64 atomics, faster, order-dependent in the last bits) and return_maps (also return the
    uint32 index maps i, j and the re-weighted weights)."""
    if mfs:
        vis = freqcorrect(data)
        u, v, freq = vis.u, vis.v, vis.freq
        real, imag, weights = vis.real, vis.imag, vis.weights
    else:
        u, v = data.u, data.v
        if channel != None:                              # noqa: E711
            freq = numpy.array([data.freq[channel]])
            real = data.real[:, channel].reshape((data.real.shape[0], 1))
            imag = data.imag[:, channel].reshape((data.real.shape[0], 1))
            weights = data.weights[:, channel].reshape((data.real.shape[0], 1))
        else:
            freq, real, imag, weights = data.freq, data.real, data.imag, data.weights

    if convolution not in _lib.CONV:
        raise ValueError("convolution must be 'pillbox' or 'expsinc'")
    if mode not in _lib.MODE:
        raise ValueError("mode must be 'continuum' or 'spectralline'")
    wt = _lib.WEIGHTING.get(weighting, 0)                # unknown strings fall through as natural (:429)

    nuv, nf = u.size, freq.size
    nchannels = 1 if mode == "continuum" else nf

    # cell centres: numpy.linspace, exactly as :370-381
    if gridsize % 2 == 0:
        uu = numpy.linspace(-gridsize * binsize / 2, (gridsize / 2 - 1) * binsize, gridsize)
        vv = numpy.linspace(-gridsize * binsize / 2, (gridsize / 2 - 1) * binsize, gridsize)
    else:
        uu = numpy.linspace(-(gridsize - 1) * binsize / 2, (gridsize - 1) * binsize / 2, gridsize)
        vv = numpy.linspace(-(gridsize - 1) * binsize / 2, (gridsize - 1) * binsize / 2, gridsize)
    G2 = gridsize ** 2
    # the [G, G] coordinate arrays of the result (:383): for large grids they are filled on a second host thread while
    # pdsb_grid runs (ctypes releases the GIL; at G = 2048 the 67 MB cost as much host time as the whole device step)
    filler = MeshFiller(uu, vv, threaded=G2 >= 1 << 18)

    new_real = numpy.empty((G2, nchannels))
    new_imag = numpy.empty((G2, nchannels))
    new_weights = numpy.empty((G2, nchannels))
    gi = numpy.empty((nuv, nf), dtype=numpy.uint32) if return_maps else None
    gj = numpy.empty((nuv, nf), dtype=numpy.uint32) if return_maps else None
    wmod = numpy.empty((nuv, nf)) if return_maps else None
    n_out = ctypes.c_int64(0)

    L = _lib.lib()
    try:
        _lib.check(L.pdsb_grid(_lib.ptr(_lib.f64(u)), _lib.ptr(_lib.f64(v)), _lib.ptr(_lib.f64(freq)),
                               _lib.ptr(_lib.f64(real)), _lib.ptr(_lib.f64(imag)), _lib.ptr(_lib.f64(weights)),
                               nuv, nf, _lib.HOST, int(gridsize), float(binsize), _lib.ptr(uu), _lib.ptr(vv),
                               _lib.CONV[convolution], wt, float(robust), int(npixels), _lib.MODE[mode],
                               1 if imaging else 0, 1 if deterministic else 0,
                               _lib.ptr(new_real), _lib.ptr(new_imag), _lib.ptr(new_weights),
                               _lib.ptr(gi), _lib.ptr(gj), _lib.ptr(wmod), _lib.HOST, ctypes.byref(n_out)))
    finally:
        filler.join()
    new_u, new_v = filler.result()
    if n_out.value > 0:
        print(_WARNING)

    if mode == "continuum":
        freq = numpy.array([data.freq.sum() / data.freq.size])

    out = Visibilities(new_u.reshape(G2), new_v.reshape(G2), freq, new_real, new_imag, new_weights)
    if return_maps:
        return out, gi, gj, wmod
    return out


def chisq(data, model):
    """libinterferometry.pyx:610-633: sum over uv of |d-m|^2 w on channel 0, returned through a
    C float (the reference's `cdef float chisq_calc`)."""
    out = ctypes.c_float()
    nuv, nf = data.real.shape
    L = _lib.lib()
    _lib.check(L.pdsb_chisq(_lib.ptr(_lib.f64(data.real)), _lib.ptr(_lib.f64(data.imag)),
                            _lib.ptr(_lib.f64(data.weights)), _lib.ptr(_lib.f64(model.real)),
                            _lib.ptr(_lib.f64(model.imag)), nuv, nf, _lib.HOST, ctypes.byref(out)))
    return out.value
