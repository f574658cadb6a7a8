"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the channel post-processing of a model cube.

Restates pdspy/modeling/run_flared_model.py:308-366 (regular cubes :311-339, unstructured images :341-366)
with numpy; `literal()` is the reference's own sequence of expressions, scipy.signal.fftconvolve included,
and pins `post()` in tests/test_oracle_pinning.py."""
import numpy


def post(image, subsample=1, averaging=1, hanning=False):
    """image [..., nf_in] with the channel axis LAST -> [..., nf_in / subsample / averaging]."""
    a = numpy.asarray(image, dtype=numpy.float64)
    nf_in = a.shape[-1]
    nf_mid = nf_in // subsample
    # :316-320 -- mean of each block of `subsample` sub-channels
    rec = numpy.empty(a.shape[:-1] + (nf_mid,))
    for i in range(nf_mid):
        rec[..., i] = a[..., i * subsample:(i + 1) * subsample].mean(axis=-1)
    # :324-329 -- numpy.hanning(5)/sum = [0, 1/4, 1/2, 1/4, 0], mode="same" (zero padded)
    if hanning:
        pad = numpy.zeros(a.shape[:-1] + (nf_mid + 2,))
        pad[..., 1:-1] = rec
        rec = 0.5 * pad[..., 1:-1] + 0.25 * pad[..., :-2] + 0.25 * pad[..., 2:]
    # :333-339 -- mean of each block of `averaging` channels
    nf_out = nf_mid // averaging
    out = numpy.empty(a.shape[:-1] + (nf_out,))
    for i in range(nf_out):
        out[..., i] = rec[..., i * averaging:(i + 1) * averaging].mean(axis=-1)
    return out


def literal(image4, nfreq_data, subsample, averaging, hanning):
    """The reference's expressions for a regular cube [npix, npix, nf_in, 1], verbatim in structure
    (run_flared_model.py:311-339); needs scipy."""
    import scipy.signal
    npix = image4.shape[0]
    recombined = numpy.empty((npix, npix, nfreq_data * averaging, 1))
    for i in range(image4.shape[2] // subsample):
        recombined[:, :, i, 0] = image4[:, :, i * subsample:(i + 1) * subsample, 0].mean(axis=2)
    if hanning:
        hanning_window = numpy.hanning(5) / numpy.hanning(5).sum()
        recombined = scipy.signal.fftconvolve(recombined, hanning_window.reshape((1, 1, hanning_window.size, 1)),
                                              axes=2, mode="same")
    binned = numpy.empty((npix, npix, nfreq_data, 1))
    for i in range(nfreq_data):
        binned[:, :, i, 0] = recombined[:, :, i * averaging:(i + 1) * averaging, 0].mean(axis=2)
    return binned
