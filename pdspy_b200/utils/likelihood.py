"""Visibility term of the log-likelihood, on the GPU.

Mirrors pdspy/utils/emcee.py:31-43 and its identical copy pdspy/utils/dynesty.py:47-59:

    good = data.weights > 0
    -0.5*sum((data.real-model.real)**2 * data.weights) - sum(log(data.weights[good]/(2*pi)))
    + -0.5*sum((data.imag-model.imag)**2 * data.weights) - sum(log(data.weights[good]/(2*pi)))

(the log term enters twice and with a minus sign: replicated verbatim)."""
import numpy

from .. import _lib, device


def visibility_lnlike(data, model):
    """One dataset: `data`, `model` are Visibilities with [nuv, nf] arrays.  When `model` is interpolate_model's
    result and nobody has read it yet, its arrays are still on the device and are compared there with the
    cached device copy of the data (pdsb_chi2_dataset): nothing but the four sums crosses PCIe."""
    out = numpy.empty(4)
    L = _lib.lib()
    token = getattr(model, "_device_token", None)
    ent = device.model_buffers(token) if token is not None else None
    if ent is not None:
        m_re, m_im, shape = ent
        if tuple(shape) != tuple(data.real.shape):
            raise ValueError("model and data visibilities have different shapes")
        ds = device.dataset_for(data.u, data.v, (data.real, data.imag, data.weights))
        _lib.check(L.pdsb_chi2_dataset(ds.handle, _lib.ptr(m_re), _lib.ptr(m_im), _lib.DEVICE, _lib.ptr(out)))
        return float(out[3])
    d_real, d_imag, w = _lib.f64(data.real), _lib.f64(data.imag), _lib.f64(data.weights)
    m_real, m_imag = _lib.f64(model.real), _lib.f64(model.imag)
    if m_real.shape != d_real.shape or m_imag.shape != d_real.shape:
        raise ValueError("model and data visibilities have different shapes")
    _lib.check(L.pdsb_chi2(_lib.ptr(d_real), _lib.ptr(d_imag), _lib.ptr(w), _lib.ptr(m_real), _lib.ptr(m_imag),
                           d_real.size, _lib.HOST, _lib.ptr(out)))
    return float(out[3])


def lnlike_visibilities(visibilities, m):
    """The visibility loop of lnlike (emcee.py:31-43): one term per dataset j, comparing
    visibilities["data"][j] with m.visibilities[visibilities["lam"][j]].  Returns the list the
    reference appends to `chisq`."""
    chisq = []
    for j in range(len(visibilities["file"])):
        chisq.append(visibility_lnlike(visibilities["data"][j], m.visibilities[visibilities["lam"][j]]))
    return chisq
