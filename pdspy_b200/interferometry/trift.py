"""code="trift" of interpolate_model (pdspy/interferometry/interpolate_model.py:49-55): host-side geometry.

The reference calls `trift.cpu.trift(x, y, image, u, v, dRA, dDec, mode="extended")` of the third-party `trift`
package (absent here; PARITY UNPINNED, see oracle/trift.py): the exact Fourier transform of the Delaunay
piecewise-linear interpolant of the scattered points.  The transform itself runs on the GPU (pdsb_sample_triangles,
csrc/trift.cu).  What is done here depends on the geometry only and is cached per geometry: the triangulation
(scipy.spatial.Delaunay) and a subdivision of the triangles that are too large for the device's series - a
triangle is cut into four through its edge midpoints until its vertices lie within PHASE_SPAN / q_max of their
centroid (q_max = 2 pi times the longest baseline).  A sub-triangle's vertex values are barycentric combinations of
its parent's three points (the interpolant is linear on the parent), carried as a 3 x 3 matrix."""
from collections import OrderedDict

import numpy

from .. import _lib
from ..constants import arcsec

PHASE_SPAN = 4.0          # largest |vertex phase - mean phase| of a triangle the device series is asked to handle
_RECORD = numpy.dtype([("x", "<f8", 3), ("y", "<f8", 3), ("B", "<f8", 9), ("area2", "<f8"), ("idx", "<i4", 3),
                       ("pad", "<i4")])
_CACHE = OrderedDict()
_CACHE_MAX = 4


def triangle_records(x_rad, y_rad, qmax):
    """The packed triangle records of the point set for baselines up to qmax / (2 pi) wavelengths."""
    assert _lib.load().pdsb_triangle_record_bytes() == _RECORD.itemsize
    from scipy.spatial import Delaunay
    limit = PHASE_SPAN / max(float(qmax), 1e-300)
    # bucket the limit (powers of 2^(1/4)) so that slightly different uv sets share a cache entry
    limit = 2.0 ** (numpy.floor(4 * numpy.log2(limit)) / 4)
    key = (x_rad.tobytes(), y_rad.tobytes(), float(limit))
    hit = _CACHE.get(key)
    if hit is not None:
        _CACHE.move_to_end(key)
        return hit
    pts = numpy.column_stack([x_rad, y_rad])
    scale = numpy.abs(pts).max() or 1.0           # Qhull works in units of the field, not of radians (1e-6)
    simplices = Delaunay(pts / scale).simplices.astype(numpy.int32)
    ntri = simplices.shape[0]
    V = pts[simplices]                                           # [ntri, 3, 2] vertex coordinates
    B = numpy.broadcast_to(numpy.eye(3), (ntri, 3, 3)).copy()    # vertex a = sum_k B[a, k] * parent point k
    parent = simplices
    done_V, done_B, done_P = [], [], []
    for _ in range(24):
        cen = V.mean(axis=1, keepdims=True)
        big = numpy.sqrt(((V - cen) ** 2).sum(axis=2)).max(axis=1) > limit
        done_V.append(V[~big])
        done_B.append(B[~big])
        done_P.append(parent[~big])
        if not big.any():
            break
        V, B, parent = V[big], B[big], parent[big]
        # midpoints 01, 12, 20 and the four children (0, m01, m20), (m01, 1, m12), (m20, m12, 2), (m01, m12, m20)
        mV = 0.5 * (V + numpy.roll(V, -1, axis=1))
        mB = 0.5 * (B + numpy.roll(B, -1, axis=1))
        corner, mid = (V, B), (mV, mB)
        kids = (((corner, 0), (mid, 0), (mid, 2)), ((mid, 0), (corner, 1), (mid, 1)), ((mid, 2), (mid, 1), (corner, 2)),
                ((mid, 0), (mid, 1), (mid, 2)))
        V = numpy.concatenate([numpy.stack([src[0][:, j] for src, j in kid], axis=1) for kid in kids])
        B = numpy.concatenate([numpy.stack([src[1][:, j] for src, j in kid], axis=1) for kid in kids])
        parent = numpy.concatenate([parent] * 4)
    else:
        raise ValueError("triangles too large for the requested baselines (more than 24 subdivision levels)")
    V, B, parent = numpy.concatenate(done_V), numpy.concatenate(done_B), numpy.concatenate(done_P)
    rec = numpy.zeros(V.shape[0], dtype=_RECORD)
    rec["x"], rec["y"] = V[:, :, 0], V[:, :, 1]
    rec["B"] = B.reshape(-1, 9)
    rec["area2"] = numpy.abs((V[:, 1, 0] - V[:, 0, 0]) * (V[:, 2, 1] - V[:, 0, 1]) -
                             (V[:, 2, 0] - V[:, 0, 0]) * (V[:, 1, 1] - V[:, 0, 1]))
    rec["idx"] = parent
    rec = rec[rec["area2"] > 0]
    _CACHE[key] = rec
    if len(_CACHE) > _CACHE_MAX:
        _CACHE.popitem(last=False)
    return rec


def sample(ds, u, v, model, dRA, dDec, out_real, out_imag, out_kind):
    """Fill out_real / out_imag [nuv, nf] (host arrays or device buffers) with the triangle transform of `model`
    (an UnstructuredImage: x, y arcsec, image [npts, nf] Jy/sr) at the uv points of dataset handle `ds`."""
    values = numpy.ascontiguousarray(model.image, dtype=numpy.float64)
    if values.ndim != 2:
        raise ValueError("an unstructured image is [npts, nfreq]")
    npts, nf = values.shape
    x = numpy.ascontiguousarray(model.x, dtype=numpy.float64) * arcsec
    y = numpy.ascontiguousarray(model.y, dtype=numpy.float64) * arcsec
    if x.size != npts or y.size != npts:
        raise ValueError("x, y and image disagree on the number of points")
    qmax = 2 * numpy.pi * float(numpy.sqrt(u * u + v * v).max()) if u.size else 0.0
    rec = triangle_records(x, y, qmax)
    _lib.check(_lib.lib().pdsb_sample_triangles(ds.handle, _lib.ptr(rec.view(numpy.uint8)) if rec.size else None, rec.size,
                                                _lib.ptr(values), npts, nf, _lib.HOST, float(dRA * arcsec),
                                                float(dDec * arcsec), _lib.ptr(out_real), _lib.ptr(out_imag), out_kind))
    return nf
