"""CPU oracle (TEST INFRASTRUCTURE ONLY) for interpolate_model(code="trift").

PARITY UNPINNED.  The reference hands this case to `trift.cpu.trift(x, y, image, u, v, dRA, dDec, nthreads=,
mode="extended")` (pdspy/interferometry/interpolate_model.py:49-55) of the third-party package `trift`
(psheehan/trift; optional import, :6-9; no version pinned anywhere in the reference), which is absent here.
trift's published algorithm: Delaunay-triangulate the scattered points and sum the ANALYTIC Fourier transform of
every triangle, "extended" mode taking the intensity to vary linearly over the triangle between its three
vertex values.  This file restates that: the exact transform of the piecewise-linear Delaunay interpolant.
Sign / frame conventions cannot be read off the missing package; they are anchored on the reference's other
two codes for the same image (`ftcode` is a free switch in run_flared_model.py:370-375, so the same sky must give
the same visibilities): V(u,v) = exp(-2 pi i (u dRA + v dDec)) * Int I(x, y) exp(+2 pi i (u x - v y)) dx dy with
x, y = model.x, model.y in radians - the limit of code="galario-unstructured" for vanishing pixel size.

Formulation here (deliberately not the device's): per triangle with vertex phases p_k = qx x_k + qy y_k,
  Int_T lambda_a exp(i q.r) dA = 2 A exp[i p_a, i p_a, i p_b, i p_c]      (Hermite-Genocchi)
with the divided differences of exp evaluated in numpy.longdouble: written out explicitly where the three phases
are well separated (error 1e-19 / d^3), and from their Taylor expansion about the mean phase,
sum_k h_k(x_a, x_a, x_b, x_c) / (k + 3)!  (h_k: complete homogeneous symmetric polynomials, built by polynomial
multiplication rather than the device's running recurrences), where two of them come close and the spread is small."""
import numpy

ARCSEC = 4.84813681e-6


def delaunay(x_rad, y_rad):
    from scipy.spatial import Delaunay
    pts = numpy.column_stack([x_rad, y_rad])
    scale = numpy.abs(pts).max() or 1.0
    return Delaunay(pts / scale).simplices


def _dd_confluent(za, zb, zc):
    """exp[za, za, zb, zc] for distinct complex za, zb, zc (explicit formula, longdouble)."""
    ea, eb, ec = numpy.exp(za), numpy.exp(zb), numpy.exp(zc)
    ab, ac, bc = za - zb, za - zc, zb - zc
    # d/dza of exp[za, zb, zc] = ea/(ab ac) - eb/(ab bc) + ec/(ac bc)
    return ea / (ab * ac) - ea * (1 / (ab * ab * ac) + 1 / (ab * ac * ac)) + eb / (ab * ab * bc) - ec / (ac * ac * bc)


def _dd_series(za, zb, zc, nterm=48):
    """The same by Taylor expansion about the mean node (small spread)."""
    zm = (za + zb + zc) / 3
    xa, xb, xc = za - zm, zb - zm, zc - zm
    cld = numpy.clongdouble
    # h_k(xa, xa, xb, xc) = coefficient of t^k in 1 / ((1 - xa t)^2 (1 - xb t)(1 - xc t)): multiply the four series
    def geom(xv):
        return [xv ** k for k in range(nterm)]
    def mul(p, q):
        return [sum(p[j] * q[k - j] for j in range(k + 1)) for k in range(nterm)]
    h = mul(mul(geom(xa), geom(xa)), mul(geom(xb), geom(xc)))
    fact = numpy.longdouble(6)
    tot = numpy.zeros(za.shape, dtype=cld)
    for k in range(nterm):
        tot = tot + h[k] / fact
        fact = fact * (k + 4)
    return numpy.exp(zm) * tot


def trift(x_arcsec, y_arcsec, image, u, v, dRA_rad=0.0, dDec_rad=0.0, simplices=None, return_min_separation=False):
    """[nuv, nf] complex128.  image [npts, nf] Jy/sr; x, y arcsec; u, v in wavelengths."""
    ld, cld = numpy.longdouble, numpy.clongdouble
    x = numpy.asarray(x_arcsec, dtype=ld) * ld(ARCSEC)
    y = numpy.asarray(y_arcsec, dtype=ld) * ld(ARCSEC)
    img = numpy.asarray(image, dtype=ld)
    tri = delaunay(numpy.asarray(x, dtype=float), numpy.asarray(y, dtype=float)) if simplices is None else simplices
    twopi = 2 * numpy.pi.__class__(numpy.pi) if False else ld(2) * numpy.arccos(ld(-1))
    qx = twopi * numpy.asarray(u, dtype=ld)
    qy = -twopi * numpy.asarray(v, dtype=ld)
    out = numpy.zeros((qx.size, img.shape[1]), dtype=cld)
    minsep = numpy.inf
    for t in tri:
        xa, ya = x[t], y[t]
        area2 = abs((xa[1] - xa[0]) * (ya[2] - ya[0]) - (xa[2] - xa[0]) * (ya[1] - ya[0]))
        p = qx[:, None] * xa[None, :] + qy[:, None] * ya[None, :]            # [nuv, 3]
        far = numpy.abs(p - p.mean(axis=1, keepdims=True)).max(axis=1) > 6
        if far.any():             # only these go through the explicit formula whatever their separation
            minsep = min(minsep, float(numpy.abs(p[far][:, [0, 0, 1]] - p[far][:, [1, 2, 2]]).min()))
        z = 1j * p.astype(cld)
        sep = numpy.abs(p[:, [0, 0, 1]] - p[:, [1, 2, 2]]).min(axis=1)
        spread = numpy.abs(p - p.mean(axis=1, keepdims=True)).max(axis=1)
        ser = (spread <= 6) & ((sep < 0.3) | (spread < 1.0))
        for a in range(3):
            b, c = (a + 1) % 3, (a + 2) % 3
            w = numpy.empty(qx.size, dtype=cld)
            if ser.any():
                w[ser] = _dd_series(z[ser, a], z[ser, b], z[ser, c])
            if (~ser).any():
                w[~ser] = _dd_confluent(z[~ser, a], z[~ser, b], z[~ser, c])
            out += (area2 * w)[:, None] * img[t[a]][None, :]
    shift = numpy.exp(-1j * (twopi * (numpy.asarray(u, dtype=ld) * ld(dRA_rad) + numpy.asarray(v, dtype=ld) * ld(dDec_rad))).astype(cld))
    res = (out * shift[:, None]).astype(numpy.complex128)
    return (res, minsep) if return_min_separation else res
