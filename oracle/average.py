"""CPU oracle for average() and center().  TEST INFRASTRUCTURE ONLY.

average: restates pdspy/interferometry/libinterferometry.pyx:151-311 - numpy preamble and epilogue
with the reference's own expressions, the accumulation loop (:262-277) as a plain Python/numpy
ordered loop (numpy.add.at applies additions in index order, which IS the (k, n) order).
center: restates pdspy/interferometry/center.py:5-25 with point_model of model.py:102-104.
Pinned against the live reference module (oracle/_ref) in tests/test_oracle_pinning.py."""
import numpy

ARCSEC = 4.84813681e-6


def average(u, v, freq, real, imag, weights, gridsize=256, binsize=None, radial=False, log=False, logmin=None,
            logmax=None, mfs=False, mode="continuum"):
    from .grid import freqcorrect
    data_freq = freq
    if mfs:
        u, v, freq, real, imag, weights = freqcorrect(u, v, freq, real, imag, weights)
        uvdist = numpy.sqrt(u ** 2 + v ** 2)
    else:
        uvdist = numpy.sqrt(u ** 2 + v ** 2)
        u, v, freq, real, imag = u.copy(), v.copy(), freq.copy(), real.copy(), imag.copy()
    weights = numpy.where(weights < 0, 0.0, weights)
    weights[(real == 0) & (imag == 0)] = 0.0
    g = uvdist != 0.0
    u, v, uvdist, real, imag, weights = u[g], v[g], uvdist[g], real[g, :], imag[g, :], weights[g, :]
    nfreq = freq.size
    nch = 1 if mode == "continuum" else nfreq
    if radial:
        if log:
            temp = numpy.linspace(numpy.log10(logmin), numpy.log10(logmax), gridsize + 1)
            new_u = 10 ** ((temp[1:] + temp[0:-1]) / 2)
            dtemp = temp[1] - temp[0]
            i = numpy.round((numpy.log10(uvdist) - numpy.log10(logmin)) / dtemp - 0.5).astype(numpy.uint32)
        else:
            new_u = numpy.linspace(binsize / 2, (gridsize - 0.5) * binsize, gridsize)
            i = numpy.round(uvdist / binsize).astype(numpy.uint32)
        j = numpy.zeros(uvdist.size).astype(numpy.uint32)
        new_u = new_u.reshape((1, gridsize))
        new_v = numpy.zeros((1, gridsize))
        shape = (1, gridsize, nch)
    else:
        half = gridsize / 2. if gridsize % 2 == 0 else (gridsize - 1) / 2.
        i = numpy.round(u / binsize + half).astype(numpy.uint32)
        j = numpy.round(v / binsize + half).astype(numpy.uint32)
        shape = (gridsize, gridsize, nch)
        new_u, new_v = numpy.zeros(shape), numpy.zeros(shape)
    good = numpy.logical_and(i < gridsize, j < gridsize)
    u, v, real, imag, weights, i, j = u[good], v[good], real[good, :], imag[good, :], weights[good, :], i[good], j[good]
    new_real, new_imag, new_weights = numpy.zeros(shape), numpy.zeros(shape), numpy.zeros(shape)
    nuv = u.size
    # ordered accumulation: flat index per (k, n) in loop order; numpy.add.at is unbuffered and
    # applies the additions one by one in that order
    kk, nn = numpy.meshgrid(numpy.arange(nuv), numpy.arange(nfreq), indexing="ij")
    cell = (j[kk].astype(numpy.int64) * shape[1] + i[kk]) * nch + (nn if nch > 1 else 0)
    cell = cell.ravel()
    w = weights.ravel()
    numpy.add.at(new_real.reshape(-1), cell, (real * weights).ravel())
    numpy.add.at(new_imag.reshape(-1), cell, (imag * weights).ravel())
    numpy.add.at(new_weights.reshape(-1), cell, w)
    if not radial:
        numpy.add.at(new_u.reshape(-1), cell, (u[:, None] * weights).ravel())
        numpy.add.at(new_v.reshape(-1), cell, (v[:, None] * weights).ravel())
    gd = new_weights != 0.0
    new_real[gd] = new_real[gd] / new_weights[gd]
    new_imag[gd] = new_imag[gd] / new_weights[gd]
    if not radial:
        new_u[gd] = new_u[gd] / new_weights[gd]
        new_v[gd] = new_v[gd] / new_weights[gd]
    gd = numpy.any(gd, axis=2)
    if not radial:
        new_u = (new_u * new_weights).sum(axis=2)[gd] / new_weights.sum(axis=2)[gd]
        new_v = (new_v * new_weights).sum(axis=2)[gd] / new_weights.sum(axis=2)[gd]
    else:
        new_u, new_v = new_u[gd], new_v[gd]
    gd = numpy.dstack([gd for m in range(nch)])
    if mode == "continuum":
        freq = numpy.array([data_freq.sum() / data_freq.size])
    return (new_u, new_v, freq, new_real[gd].reshape((new_u.size, nch)), new_imag[gd].reshape((new_u.size, nch)),
            new_weights[gd].reshape((new_u.size, nch)))


def center(u, v, freq, real, imag, x0, y0):
    data_complex = real + 1j * imag
    model_complex = numpy.empty(real.shape, dtype=complex)
    for i in range(len(freq)):
        us, vs = u * freq[i] / freq.mean(), v * freq[i] / freq.mean()
        m = 1j * numpy.zeros(us.size)
        m += 1. * numpy.exp(-2 * 3.14159 * (0 + 1j * (us * (x0 * ARCSEC) + vs * (y0 * ARCSEC))))   # model.py:102-104
        model_complex[:, i] = (m.real + 1j * m.imag)
    c = data_complex * model_complex.conj()
    return c.real, c.imag
