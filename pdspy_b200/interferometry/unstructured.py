"""Unstructured (scattered-point) model images -> the regular pixel grid the transform reads.

interpolate_model(code="galario-unstructured") of the reference (interpolate_model.py:32-47) hands the
points (model.x, -model.y), their intensities and a target grid (nxy pixels of dxy arcsec) to
`galario.double.sampleUnstructuredImage` of psheehan's galario fork - NOT installed here and not vendored
(docs/installation.rst:22-27), so, as for sampleImage, PARITY IS UNPINNED.  What is implemented is the
definition that call stands for: the piecewise-linear (Delaunay) interpolant of the scattered
intensities (RADMC-3D circular images, Jy/sr, Model.py:536-558), sampled on the regular grid, times the
pixel solid angle, then the same transform as code="galario".

The triangulation and the per-pixel (triangle, barycentric weights) depend on the geometry only; they are
computed once on the host (scipy.spatial.Delaunay, cached per geometry) and each call then runs one gather
kernel on the GPU (pdsb_regrid_linear)."""
from collections import OrderedDict

import numpy

from .. import _lib
from ..constants import arcsec

_CACHE = OrderedDict()
_CACHE_MAX = 4


def grid_axes(nxy, dxy_rad):
    """Pixel-centre coordinates of the target grid in galario's frame: x_g[c] = (c - nxy/2) dxy along columns,
    y_g[r] = (r - nxy/2) dxy along ITS rows r (its row 0 is the reference image's last row, the [::-1] of
    interpolate_model.py:23)."""
    ax = (numpy.arange(nxy) - nxy // 2) * dxy_rad
    return ax, ax.copy()


def triangulate(xg, yg, nxy, dxy_rad):
    """(tri [nxy*nxy, 3] int32, bary [nxy*nxy, 3] float64) for the grid of grid_axes, row-major in galario's
    (r, c) order; tri[p, 0] = -1 outside the convex hull of the points."""
    from scipy.spatial import Delaunay
    key = (xg.tobytes(), yg.tobytes(), int(nxy), float(dxy_rad))
    hit = _CACHE.get(key)
    if hit is not None:
        _CACHE.move_to_end(key)
        return hit
    pts = numpy.column_stack([xg, yg])
    scale = numpy.abs(pts).max() or 1.0           # Qhull works in units of the field, not of radians (1e-6)
    dl = Delaunay(pts / scale)
    ax, ay = grid_axes(nxy, dxy_rad)
    gx, gy = numpy.meshgrid(ax / scale, ay / scale)
    q = numpy.column_stack([gx.ravel(), gy.ravel()])
    simplex = dl.find_simplex(q)
    inside = simplex >= 0
    tri = numpy.full((q.shape[0], 3), -1, dtype=numpy.int32)
    bary = numpy.zeros((q.shape[0], 3))
    s = simplex[inside]
    T = dl.transform[s]                            # [n, 3, 2]: inverse edge matrix and origin vertex
    b2 = numpy.einsum("nij,nj->ni", T[:, :2, :], q[inside] - T[:, 2, :])
    bary[inside, 0:2] = b2
    bary[inside, 2] = 1.0 - b2.sum(axis=1)
    tri[inside] = dl.simplices[s].astype(numpy.int32)
    out = (numpy.ascontiguousarray(tri), numpy.ascontiguousarray(bary))
    _CACHE[key] = out
    if len(_CACHE) > _CACHE_MAX:
        _CACHE.popitem(last=False)
    return out


def regrid(model, nxy, dxy):
    """The regular cube [nxy, nxy, nf, 1] (Jy/pixel, reference row order: row j is galario row nxy-1-j) of an
    unstructured image: model.x, model.y in arcsec, model.image [npts, nf] in Jy/sr; dxy in arcsec."""
    values = numpy.ascontiguousarray(model.image, dtype=numpy.float64)
    if values.ndim != 2:
        raise ValueError("an unstructured image is [npts, nfreq]")
    npts, nf = values.shape
    xg = numpy.ascontiguousarray(model.x, dtype=numpy.float64) * arcsec
    yg = -numpy.ascontiguousarray(model.y, dtype=numpy.float64) * arcsec        # interpolate_model.py:39
    if xg.size != npts or yg.size != npts:
        raise ValueError("x, y and image disagree on the number of points")
    dxy_rad = dxy * arcsec
    tri, bary = triangulate(xg, yg, int(nxy), dxy_rad)
    out = numpy.empty((nxy * nxy, nf))
    _lib.check(_lib.lib().pdsb_regrid_linear(_lib.ptr(values), npts, _lib.ptr(tri), _lib.ptr(bary), nxy * nxy, nf,
                                             float(dxy_rad * dxy_rad), _lib.HOST, _lib.ptr(out)))
    cube = out.reshape(nxy, nxy, nf)[::-1]                                      # galario rows -> reference rows
    return numpy.ascontiguousarray(cube[:, :, :, None])
