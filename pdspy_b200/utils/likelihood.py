"""Visibility term of the log-likelihood, on the GPU.

Mirrors pdspy/utils/emcee.py:31-43 and its identical copy pdspy/utils/dynesty.py:47-59:

    good = data.weights > 0
    -0.5*sum((data.real-model.real)**2 * data.weights) - sum(log(data.weights[good]/(2*pi)))
    + -0.5*sum((data.imag-model.imag)**2 * data.weights) - sum(log(data.weights[good]/(2*pi)))

(the log term enters twice and with a minus sign: replicated verbatim)."""
import numpy

from .. import _lib


def visibility_lnlike(data, model):
    """One dataset: `data`, `model` are Visibilities with [nuv, nf] arrays."""
    out = numpy.empty(4)
    L = _lib.lib()
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
