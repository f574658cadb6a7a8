"""Channel post-processing of a model cube on the GPU (pdspy/modeling/run_flared_model.py:308-366)."""
import numpy

from .. import _lib


def _check(nf_in, subsample, averaging):
    subsample, averaging = int(subsample), int(averaging)
    if subsample < 1 or averaging < 1:
        raise ValueError("subsample and averaging must be >= 1")
    if nf_in % (subsample * averaging) != 0:
        raise ValueError("the channel axis (%d) must be nf_out * averaging (%d) * subsample (%d)"
                         % (nf_in, averaging, subsample))
    return subsample, averaging


def _in_scale(in_scale, nf_in):
    if in_scale is None:
        return None
    s = numpy.ascontiguousarray(in_scale, dtype=numpy.float64)
    if s.shape != (nf_in,):
        raise ValueError("the per-channel factor (extinction) must have one entry per cube channel before "
                         "post-processing: expected %d, got shape %s" % (nf_in, s.shape))
    return s


def postprocess_channels(image, subsample=1, averaging=1, hanning=False, in_scale=None):
    """What the reference does to a freshly rendered cube before interpolate_model: mean over blocks of
    `subsample` sub-channels, Hanning smoothing along the channel axis when `hanning`
    (numpy.hanning(5)/sum through scipy.signal.fftconvolve(mode="same"), i.e. taps 1/4, 1/2, 1/4 with
    zero padding), mean over blocks of `averaging` channels.  in_scale [nf_in]: multiplied into the input channels
    first (the reference's extinction, run_flared_model.py:286-299, which precedes all of this).

    image: [ny, nx, nf_in, 1] (regular cube) or [npts, nf_in] (unstructured image); returns the same
    rank with nf_in / subsample / averaging channels."""
    a = numpy.ascontiguousarray(image, dtype=numpy.float64)
    if a.ndim == 4:
        if a.shape[3] != 1:
            raise ValueError("regular cubes are [ny, nx, nf, 1]")
        nf_in, npix = a.shape[2], a.shape[0] * a.shape[1]
    elif a.ndim == 2:
        nf_in, npix = a.shape[1], a.shape[0]
    else:
        raise ValueError("expected [ny, nx, nf, 1] or [npts, nf]")
    subsample, averaging = _check(nf_in, subsample, averaging)
    nf_out = nf_in // subsample // averaging
    out = numpy.empty(a.shape[:2] + (nf_out, 1) if a.ndim == 4 else (npix, nf_out))
    scale = _in_scale(in_scale, nf_in)
    if npix > 0:
        _lib.check(_lib.lib().pdsb_channel_postprocess_scaled(_lib.ptr(a), npix, nf_in, subsample, 1 if hanning else 0,
                                                              averaging, _lib.ptr(scale), _lib.HOST, _lib.ptr(out)))
    return out


def postprocess_channels_device(image, subsample=1, averaging=1, hanning=False, in_scale=None):
    """The same, host cube in, DeviceBuffer out (for chaining into pdsb_sample_image / pdsb_loglike without
    a round trip).  image: [ny, nx, nf_in, 1].  Returns (buffer, nf_out)."""
    a = numpy.ascontiguousarray(image, dtype=numpy.float64)
    ny, nx, nf_in = a.shape[:3]
    subsample, averaging = _check(nf_in, subsample, averaging)
    nf_out = nf_in // subsample // averaging
    scale = _in_scale(in_scale, nf_in)
    src = _lib.DeviceBuffer.from_numpy(a)
    dst = _lib.DeviceBuffer(max(1, ny * nx * nf_out) * 8)
    if ny * nx > 0:
        _lib.check(_lib.lib().pdsb_channel_postprocess_scaled(_lib.ptr(src), ny * nx, nf_in, subsample,
                                                              1 if hanning else 0, averaging, _lib.ptr(scale),
                                                              _lib.DEVICE, _lib.ptr(dst)))
    return dst, nf_out
