"""CPU oracle for image cube -> model visibilities.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product (pdspy_b200/) never does.

PARITY UNPINNED at this boundary: the reference delegates the arithmetic of
pdspy/interferometry/interpolate_model.py:11-57 to the third-party package
`galario` (`from galario import double`, interpolate_model.py:5; call at :23-24),
which is neither vendored in /root/reference, nor version-pinned (meta.yaml:36 lists
a bare `galario`; setup.py:82-84 omits it), nor installable here (no network).  The
reference holds no golden visibilities for this path (tests/test.py never calls
interpolate_model).  What follows restates galario's published algorithm
(Tazzari, Beaujean & Testi 2018, MNRAS 476, 4527: shift, real FFT, bilinear
interpolation, phase factor) and, beside it, the exact transform that algorithm
approximates.  Conventions are anchored on the reference's own call site and on its
analytic models (pdspy/interferometry/model.py:102-147), which fix the sign of
u, v, x0, y0 and the conjugation: a point of flux F at x0 (east, i.e. towards
smaller column index) and y0 (north) must give F*exp(-2*pi*i*(u*x0+v*y0)).

Reference lines followed:
  interpolate_model.py:20     dxy = (model.x[1]-model.x[0])*arcsec
  interpolate_model.py:23-24  sampleImage(model.image[::-1,:,i,0].copy(), dxy, u, v,
                              dRA=dRA*arcsec, dDec=dDec*arcsec)   -- per channel i
  interpolate_model.py:26-27  real = vis.real ; imag = -vis.imag  (conjugation)
  interpolate_model.py:57     Visibilities(u, v, freq, real, imag, ones)
  constants/astronomy.py:9    arcsec = 4.84813681e-6  (the truncated literal, kept)

Exact transform in the reference's final convention (n = image size, even):

  V_i(u,v) = sum_{j,c} image[j,c,i,0]
             * exp(+2*pi*i*dxy*( u*(c - nx//2) + v*(ny//2 - 1 - j) ))
             * exp(-2*pi*i*( u*dRA + v*dDec ))

The `ny//2 - 1 - j` is the row flip [::-1] of an even-sized image seen from
galario's phase centre at pixel (n/2, n/2): it is a one-pixel Dec offset which the
reference has and which is therefore kept.
"""
import numpy as np

ARCSEC = 4.84813681e-6      # pdspy/constants/astronomy.py:9


def pixel_coords(ny, nx, dxy):
    """x_c and y_j (radians) that multiply u and v in the exact transform."""
    xc = dxy * (np.arange(nx, dtype=np.float64) - nx // 2)
    yj = dxy * (ny // 2 - 1 - np.arange(ny, dtype=np.float64))
    return xc, yj


def exact_dft_literal(u, v, image, dxy, dRA=0.0, dDec=0.0, block=256):
    """The exact transform summed pixel by pixel, one complex exponential per
    (pixel, uv) pair, in fp64.  O(nuv*ny*nx) exponentials: small cases only.

    image: [ny, nx, nf, 1] (reference layout, imaging/libimaging.pyx:11)
    dxy, dRA, dDec in radians.  Returns complex128 [nuv, nf] (real, imag as the
    reference stores them after its conjugation)."""
    u = np.asarray(u, np.float64)
    v = np.asarray(v, np.float64)
    ny, nx, nf = image.shape[:3]
    xc, yj = pixel_coords(ny, nx, dxy)
    X = np.broadcast_to(xc[None, :], (ny, nx)).reshape(-1)
    Y = np.broadcast_to(yj[:, None], (ny, nx)).reshape(-1)
    img = np.asarray(image, np.float64)[:, :, :, 0].reshape(ny * nx, nf)
    out = np.empty((u.size, nf), np.complex128)
    for s in range(0, u.size, block):
        uu = u[s:s + block, None]
        vv = v[s:s + block, None]
        ph = 2.0 * np.pi * (uu * (X[None, :] - dRA) + vv * (Y[None, :] - dDec))
        out[s:s + block] = np.exp(1j * ph) @ img
    return out


def exact_dft(u, v, image, dxy, dRA=0.0, dDec=0.0, block=4096):
    """Same exact transform, evaluated through the separability of the phase
    (exp(i(a+b)) = exp(ia)exp(ib)): per uv point one row of nx column factors and
    one row of ny row factors.  Used where exact_dft_literal would take minutes;
    tests check the two agree to ~1e-13."""
    u = np.asarray(u, np.float64)
    v = np.asarray(v, np.float64)
    ny, nx, nf = image.shape[:3]
    xc, yj = pixel_coords(ny, nx, dxy)
    img = np.asarray(image, np.float64)[:, :, :, 0]
    out = np.empty((u.size, nf), np.complex128)
    for s in range(0, u.size, block):
        uu = u[s:s + block, None]
        vv = v[s:s + block, None]
        Eu = np.exp(2j * np.pi * uu * (xc[None, :] - dRA))      # [b, nx]
        Ev = np.exp(2j * np.pi * vv * (yj[None, :] - dDec))      # [b, ny]
        for i in range(nf):
            T = Eu @ img[:, :, i].T                              # [b, ny]
            out[s:s + block, i] = np.sum(T * Ev, axis=1)
    return out


def galario_like(u, v, image, dxy, dRA=0.0, dDec=0.0):
    """Restatement of galario's sampleImage algorithm as the reference calls it
    (interpolate_model.py:23-27), including the reference's row flip and final
    conjugation: A = image[::-1,:,i,0]; F = fftshift_rows(rfft2(fftshift(A)));
    bilinear interpolation of F at (n/2 + v/du, |u|/du) with du = 1/(n*dxy), the
    mirrored row and the conjugate for u < 0; multiplication by
    exp(+2*pi*i*(u*dRA+v*dDec)); then imag -> -imag.

    On FFT grid points it equals exact_dft to rounding; off the grid it carries the
    bilinear interpolation error (reported by tests/bench as the 'method gap').
    Square, even-sized images only (galario's own requirement)."""
    u = np.asarray(u, np.float64)
    v = np.asarray(v, np.float64)
    ny, nx, nf = image.shape[:3]
    assert ny == nx and nx % 2 == 0, "galario requires a square, even-sized image"
    n = nx
    du = 1.0 / (n * dxy)
    out = np.empty((u.size, nf), np.complex128)
    uneg = u < 0.0
    indu = np.abs(u) / du
    indv = n / 2.0 + np.where(uneg, -v, v) / du
    fu = np.floor(indu).astype(np.int64)
    fv = np.floor(indv).astype(np.int64)
    t = indu - fu
    s = indv - fv
    fu1 = np.minimum(fu + 1, n // 2)
    fv1 = np.minimum(fv + 1, n)            # row n is the periodic copy of row 0
    phase = np.exp(2j * np.pi * (u * dRA + v * dDec))
    for i in range(nf):
        A = np.asarray(image, np.float64)[::-1, :, i, 0]
        F = np.fft.fftshift(np.fft.rfft2(np.fft.fftshift(A)), axes=0)
        F = np.concatenate([F, F[:1]], axis=0)
        val = ((1 - t) * (1 - s) * F[fv, fu] + t * (1 - s) * F[fv, fu1]
               + (1 - t) * s * F[fv1, fu] + t * s * F[fv1, fu1])
        val = np.where(uneg, np.conj(val), val)
        out[:, i] = np.conj(val * phase)
    return out


def interpolate_model(u, v, freq, model, dRA=0.0, dDec=0.0, method="exact"):
    """Oracle with the reference's argument meaning (dRA, dDec in arcsec; dxy from
    model.x).  Returns (real, imag, weights) as [nuv, nf] fp64 arrays."""
    dxy = (model.x[1] - model.x[0]) * ARCSEC
    fn = {"exact": exact_dft, "literal": exact_dft_literal, "galario": galario_like}[method]
    vis = fn(u, v, model.image, dxy, dRA * ARCSEC, dDec * ARCSEC)
    return vis.real.copy(), vis.imag.copy(), np.ones(vis.shape)
