"""CPU oracle for convolutional gridding.  TEST INFRASTRUCTURE ONLY.

Restates grid() / freqcorrect() of pdspy/interferometry/libinterferometry.pyx:313-541,
587-608.  The numpy preamble is written with the same numpy expressions, in the same
order, as the reference (index maps and cell centres are bit-exact by construction);
the loops are in oracle/c/oracle.c:oracle_grid_core.  Pinned against the live
reference module (oracle/_ref) in tests/test_oracle_pinning.py and against
tests/golden/grid_*.npz.
"""
import ctypes
import numpy as np


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def freqcorrect(u, v, freq, real, imag, weights, new_freq=None):
    """libinterferometry.pyx:587-608."""
    nf = np.array([freq.mean()]) if new_freq is None else np.array([new_freq], dtype=np.float64)
    inv_freq = 1. / nf[0]
    scale = freq * inv_freq
    nu = (u.reshape((u.size, 1)) * scale).reshape((real.size,))
    nv = (v.reshape((v.size, 1)) * scale).reshape((real.size,))
    return (nu, nv, nf, real.reshape((real.size, 1)), imag.reshape((imag.size, 1)),
            weights.reshape((weights.size, 1)))


def cell_centres(gridsize, binsize):
    """libinterferometry.pyx:370-381."""
    if gridsize % 2 == 0:
        uu = np.linspace(-gridsize * binsize / 2, (gridsize / 2 - 1) * binsize, gridsize)
    else:
        uu = np.linspace(-(gridsize - 1) * binsize / 2, (gridsize - 1) * binsize / 2, gridsize)
    return uu, uu.copy()


def index_maps(u, v, freq, gridsize, binsize):
    """libinterferometry.pyx:388-403: left-to-right fp64, then numpy's cast to uint32."""
    mean_freq = np.mean(freq)
    inv_freq = 1. / mean_freq
    if gridsize % 2 == 0:
        i = np.array(u.reshape((u.size, 1)) * freq * inv_freq / binsize + gridsize / 2., dtype=np.uint32)
        j = np.array(v.reshape((v.size, 1)) * freq * inv_freq / binsize + gridsize / 2., dtype=np.uint32)
    else:
        i = np.array(u.reshape((u.size, 1)) * freq * inv_freq / binsize + (gridsize - 1) / 2., dtype=np.uint32)
        j = np.array(v.reshape((v.size, 1)) * freq * inv_freq / binsize + (gridsize - 1) / 2., dtype=np.uint32)
    return i, j


def grid(u, v, freq, real, imag, weights, gridsize=256, binsize=2000.0, convolution="pillbox",
         mfs=False, channel=None, imaging=False, weighting="natural", robust=2, npixels=0,
         mode="continuum", return_maps=False):
    """Returns (new_u[G*G], new_v[G*G], freq_out, real[G*G,nch], imag, weights) and,
    with return_maps, also (i, j, good)."""
    from .build import lib
    data_freq = freq                                                # `data.freq` of :536
    if mfs:
        u, v, freq, real, imag, weights = freqcorrect(u, v, freq, real, imag, weights)
    else:
        if channel is not None:
            freq = np.array([freq[channel]])
            real = real[:, channel].reshape((real.shape[0], 1))
            imag = imag[:, channel].reshape((imag.shape[0], 1))
            weights = weights[:, channel].reshape((weights.shape[0], 1))
    weights = np.where(weights < 0, 0.0, weights)                   # :351
    weights[(real == 0) & (imag == 0)] = 0.0                        # :353
    nuv, nf = u.size, freq.size
    nch = 1 if mode == "continuum" else nf
    uu, vv = cell_centres(gridsize, binsize)
    new_u, new_v = np.meshgrid(uu, vv)
    i, j = index_maps(u, v, freq, gridsize, binsize)
    good = np.logical_and(np.logical_and(i >= 0, i < gridsize), np.logical_and(j >= 0, j < gridsize))
    G = gridsize
    nr = np.zeros((G, G, nch))
    ni = np.zeros((G, G, nch))
    nw = np.zeros((G, G, nch))
    conv = {"pillbox": 0, "expsinc": 1}[convolution]
    wt = {"natural": 0, "uniform": 1, "superuniform": 2, "robust": 3}[weighting]
    uc, vc, fc = (np.ascontiguousarray(a, np.float64) for a in (u, v, freq))
    rc, ic = (np.ascontiguousarray(a, np.float64) for a in (real, imag))
    wc = np.ascontiguousarray(weights, np.float64)
    i32 = np.ascontiguousarray(i, np.uint32)
    j32 = np.ascontiguousarray(j, np.uint32)
    g8 = np.ascontiguousarray(good, np.uint8)
    lib().oracle_grid_core(_p(uc), _p(vc), _p(fc), _p(rc), _p(ic), _p(wc), nuv, nf, _p(i32), _p(j32),
                           _p(g8), G, float(binsize), _p(uu), _p(vv), conv, wt, float(robust),
                           int(npixels), int(mode != "continuum"), 2, _p(nr), _p(ni), _p(nw))
    if imaging:                                                     # :525-529, numpy sums
        for n in range(nch):
            nr[:, :, n] /= nw[:, :, n].sum()
            ni[:, :, n] /= nw[:, :, n].sum()
            nw[:, :, n] /= nw[:, :, n].sum()
    else:                                                           # :531-533
        g = nw > 0
        nr[g] = nr[g] / nw[g]
        ni[g] = ni[g] / nw[g]
    if mode == "continuum":
        freq_out = np.array([data_freq.sum() / data_freq.size])
    else:
        freq_out = freq
    out = (new_u.reshape(G ** 2), new_v.reshape(G ** 2), freq_out, nr.reshape((G ** 2, nch)),
           ni.reshape((G ** 2, nch)), nw.reshape((G ** 2, nch)))
    if return_maps:
        return out + (i, j, good, wc)
    return out
