"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the Hogbom loop of pdspy/interferometry/clean.py.

`mad_std` restates astropy.stats.mad_std (absent here: astropy is not installed) from its definition,
1.482602218505602 * median(|x - median(x)|).  `loop` restates clean.py:27-34, 57-113 with the
fftconvolve of a one-pixel image written as the shifted beam it equals; tests/golden/make_golden.py runs
the reference's own clean.py (scipy fftconvolve and all) to pin it (clean_golden.npz)."""
import numpy


def mad_std(x):
    x = numpy.asarray(x, dtype=numpy.float64)
    return 1.482602218505602 * numpy.median(numpy.abs(x - numpy.median(x)))


def loop(dirty, dirty_beam, clean_beam, gain=0.1, maxiter=1000, nsigma=5.):
    """Returns (clean_image, residuals, model, mask, niter, threshold)."""
    dirty = numpy.array(dirty, dtype=numpy.float64)
    ny, nx, nf = dirty.shape
    wherezero = dirty == 0                                                   # :31
    model = numpy.zeros(dirty.shape)                                         # :34
    bres = (dirty_beam - clean_beam).max()
    threshold = max(bres * dirty.max(), 5. * mad_std(dirty))                 # :57-58
    mask = numpy.zeros(dirty.shape)
    mask[dirty > threshold] = 1.0                                            # :59-60
    n, stop = 0, False
    while n < maxiter and not stop:                                          # :67
        if (dirty * mask).max() < threshold:                                 # :70-76
            threshold = max(bres * dirty.max(), 5. * mad_std(dirty))
            new_mask = numpy.zeros(dirty.shape)
            new_mask[dirty > threshold] = 1.0
            mask = numpy.logical_or(mask, new_mask)
        dm = dirty * mask
        if not dm.max() > 0:                                                 # see include/pdsb.h:pdsb_clean_loop
            break
        p = int(numpy.argmax(dm))                                            # :80, first of equal maxima
        y0, x0, ch = numpy.unravel_index(p, dirty.shape)
        val = dirty[y0, x0, ch] * gain
        model[y0, x0, ch] += val                                             # :84
        # :88-93  fftconvolve(delta at (y0, x0) * val, beam, "same")[y, x] = val * beam[y + ny-1 - y0, x + nx-1 - x0]
        dirty[:, :, ch] = dirty[:, :, ch] - val * dirty_beam[ny - 1 - y0:2 * ny - 1 - y0, nx - 1 - x0:2 * nx - 1 - x0, ch]
        dirty[wherezero] = 0.                                                # :94
        stop = (dirty * mask).max() < nsigma * mad_std(dirty)                # :98
        n += 1
    clean_image = numpy.zeros(dirty.shape)                                   # :108-113
    for p in numpy.flatnonzero(model):
        y0, x0, ch = numpy.unravel_index(p, dirty.shape)
        clean_image[:, :, ch] += model[y0, x0, ch] * clean_beam[ny - 1 - y0:2 * ny - 1 - y0, nx - 1 - x0:2 * nx - 1 - x0, ch]
    clean_image += dirty
    return clean_image, dirty, model, numpy.asarray(mask, dtype=float), n, threshold
