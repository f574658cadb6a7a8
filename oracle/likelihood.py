"""CPU oracle for the visibility term of the log-likelihood.  TEST INFRASTRUCTURE ONLY.

Follows pdspy/utils/emcee.py:31-43 (identical copy pdspy/utils/dynesty.py:47-59) as a
verbatim numpy expression, and pdspy/interferometry/libinterferometry.pyx:610-633 for
chisq().  Pinned against the live reference module (oracle/_ref) for chisq(); the
emcee expression is numpy-only in the reference, so the restatement IS the reference
expression.
"""
import ctypes
import numpy as np


def lnlike_vis_numpy(d_real, d_imag, d_weights, m_real, m_imag):
    """emcee.py:31-43 for one dataset, written exactly as the reference writes it."""
    good = d_weights > 0
    return (-0.5 * np.sum((d_real - m_real) ** 2 * d_weights)
            - np.sum(np.log(d_weights[good] / (2 * np.pi)))
            + -0.5 * np.sum((d_imag - m_imag) ** 2 * d_weights)
            - np.sum(np.log(d_weights[good] / (2 * np.pi))))


def chi2_per_channel_numpy(d_real, d_imag, d_weights, m_real, m_imag):
    """sum over uv of w*|d-m|^2, per channel (the channel reduction of section 8e)."""
    return np.sum(((d_real - m_real) ** 2 + (d_imag - m_imag) ** 2) * d_weights, axis=0)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def lnlike_vis_c(d_real, d_imag, d_weights, m_real, m_imag):
    from .build import lib
    arrs = [np.ascontiguousarray(a, np.float64) for a in (d_real, d_imag, d_weights, m_real, m_imag)]
    return lib().oracle_loglike(*[_p(a) for a in arrs], arrs[0].size, None, None, None)


def chisq_c(d_real, d_imag, d_weights, m_real, m_imag):
    """libinterferometry.pyx:610-633: channel 0 only, returned through a C float.
    The reference passes nuv = data.real.size (:612) and indexes [i,0], which runs
    past the array when nf > 1 (boundscheck is off, :616); it is only defined for
    nf == 1, where real.size == number of rows.  The restatement loops over rows."""
    from .build import lib
    arrs = [np.ascontiguousarray(a, np.float64) for a in (d_real, d_imag, d_weights, m_real, m_imag)]
    nuv, nf = arrs[0].shape
    return float(lib().oracle_chisq(*[_p(a) for a in arrs], nuv, nf))
