"""CPU oracle (TEST INFRASTRUCTURE ONLY) for interpolate_model(code="galario-unstructured").

PARITY UNPINNED: the arithmetic of the reference lives in `galario.double.sampleUnstructuredImage` of
psheehan's galario fork (pdspy/interferometry/interpolate_model.py:32-47, docs/installation.rst:22-27),
which is absent here.  This restates the definition the call stands for with an independent tool:
scipy.interpolate.LinearNDInterpolator (its own Delaunay triangulation) on the regular grid of nxy pixels
of dxy, zero outside the hull, times the pixel solid angle; then oracle/dft.py's exact transform.  The
frame is fixed by the reference's call: points (model.x, -model.y), the regular image handed to galario
row-flipped (:23) and the result conjugated (:27, :45)."""
import numpy

ARCSEC = 4.84813681e-6


def regrid(x_arcsec, y_arcsec, image, nxy, dxy_arcsec):
    """[nxy, nxy, nf, 1] Jy/pixel in the reference's row order from scattered Jy/sr intensities [npts, nf]."""
    from scipy.interpolate import LinearNDInterpolator
    xg, yg = numpy.asarray(x_arcsec) * ARCSEC, -numpy.asarray(y_arcsec) * ARCSEC
    d = dxy_arcsec * ARCSEC
    ax = (numpy.arange(nxy) - nxy // 2) * d
    gx, gy = numpy.meshgrid(ax, ax)                       # galario frame: row r <-> y_g = (r - nxy/2) d
    s = max(numpy.abs(xg).max(), numpy.abs(yg).max())
    f = LinearNDInterpolator(numpy.column_stack([xg, yg]) / s, numpy.asarray(image, dtype=numpy.float64), fill_value=0.0)
    a = f(gx.ravel() / s, gy.ravel() / s).reshape(nxy, nxy, -1) * d * d
    return numpy.ascontiguousarray(a[::-1][:, :, :, None])
