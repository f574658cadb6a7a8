"""Synthetic inputs of the shapes BASELINE.json names.

Used by bench.py, the tests and the golden-vector generator so that all three see
the same seeded data.  Pure numpy; no GPU, no oracle.
"""
import numpy as np

ARCSEC = 4.84813681e-6      # pdspy/constants/astronomy.py:9


def synth_uv(nuv, dxy_rad, seed=1234, uvmin=1.5e4):
    """N/2 baselines, length log-uniform in [uvmin, 0.45/dxy], angle uniform in
    [0, pi), then Hermitian-doubled exactly as the reference's readers do
    (pdspy/interferometry/readuvfits.py:68-73): u=[u,-u], v=[v,-v]."""
    rng = np.random.default_rng(seed)
    half = nuv // 2
    uvmax = 0.45 / dxy_rad
    r = np.exp(rng.uniform(np.log(uvmin), np.log(max(uvmax, 2 * uvmin)), half))
    th = rng.uniform(0.0, np.pi, half)
    u = r * np.cos(th)
    v = r * np.sin(th)
    u = np.concatenate([u, -u])
    v = np.concatenate([v, -v])
    if u.size < nuv:                       # odd request: pad with one extra point
        u = np.concatenate([u, [uvmin]])
        v = np.concatenate([v, [0.0]])
    return u, v


def synth_freq(nf):
    return 230e9 + 1e5 * np.arange(nf, dtype=np.float64)


def synth_image(n, nf, pixelsize_arcsec, kind="disk", seed=5678):
    """fp64 [n, n, nf, 1] Jy/pixel, non-negative.  kind='disk': inclined Gaussian
    disk (sigma_major 0.4", axis ratio 0.6, PA 30 deg) + ring at 0.8", centred on
    pixel (n/2, n/2), with a Keplerian-like butterfly channel mask for cubes;
    kind='random': uniform random (the adversarial no-structure case)."""
    x = (np.arange(n) - n / 2) * pixelsize_arcsec
    if kind == "random":
        rng = np.random.default_rng(seed)
        return rng.random((n, n, nf, 1))
    # scale the structure to the field of view so every config has a resolved source
    fov = n * pixelsize_arcsec
    s = min(1.0, fov / 4.0)
    X, Y = np.meshgrid(x, x)
    pa = np.deg2rad(30.0)
    xr = X * np.cos(pa) + Y * np.sin(pa)
    yr = -X * np.sin(pa) + Y * np.cos(pa)
    q = 0.6
    rr = np.sqrt(xr ** 2 + (yr / q) ** 2)
    base = np.exp(-0.5 * (rr / (0.4 * s)) ** 2) + 0.3 * np.exp(-0.5 * ((rr - 0.8 * s) / (0.08 * s)) ** 2)
    img = np.empty((n, n, nf, 1))
    if nf == 1:
        img[:, :, 0, 0] = base
    else:
        vel = xr / np.maximum(rr, 1e-3 * s) / np.sqrt(np.maximum(rr, 0.05 * s) / s)   # ~ cos(phi)/sqrt(r)
        vch = np.linspace(-3.0, 3.0, nf)
        dv = (vch[1] - vch[0]) * 1.5
        for i in range(nf):
            img[:, :, i, 0] = base * np.exp(-0.5 * ((vel - vch[i]) / dv) ** 2)
    img *= 1.0 / max(img[:, :, nf // 2, 0].sum(), 1e-30)     # ~1 Jy in the central channel
    return img


def synth_data(nuv, nf, seed=4321, model=None):
    """re, im ~ N(model, 1/sqrt(w)) (or N(0,1)); w ~ U[0.5,2] with 1% zeros and 0.1%
    negative (exercises the `good` / clamp paths of emcee.py:32 and
    libinterferometry.pyx:351).  Hermitian-doubled like synth_uv."""
    rng = np.random.default_rng(seed)
    half = nuv // 2
    w = rng.uniform(0.5, 2.0, (half, nf))
    w[rng.random((half, nf)) < 0.01] = 0.0
    neg = rng.random((half, nf)) < 0.001
    w[neg] = -w[neg]
    sig = 1.0 / np.sqrt(np.where(w > 0, w, 1.0))
    re = rng.normal(size=(half, nf)) * sig
    im = rng.normal(size=(half, nf)) * sig
    if model is not None:
        re += model[0][:half]
        im += model[1][:half]
    re = np.concatenate([re, re])
    im = np.concatenate([im, -im])
    w = np.concatenate([w, w])
    if re.shape[0] < nuv:
        pad = nuv - re.shape[0]
        re = np.concatenate([re, np.zeros((pad, nf))])
        im = np.concatenate([im, np.zeros((pad, nf))])
        w = np.concatenate([w, np.ones((pad, nf))])
    return re, im, w


_BLOCK = 8192


def synth_data_half_rows(nuv, nf, start, stop, seed=4321):
    """Rows [start, stop) of the FIRST half of a Hermitian-doubled synthetic data set of nuv points, generated
    block by block from (seed, block) so that any partition of the rows over ranks sees the same values:
    bench.py shards ONE data set this way.  Returns re, im, w of shape [stop - start, nf] (statistics as
    synth_data, not the same numbers)."""
    half = nuv // 2
    assert 0 <= start <= stop <= half
    re = np.empty((stop - start, nf))
    im = np.empty((stop - start, nf))
    w = np.empty((stop - start, nf))
    for b in range(start // _BLOCK, (max(stop, 1) - 1) // _BLOCK + 1):
        lo, hi = b * _BLOCK, min((b + 1) * _BLOCK, half)
        rng = np.random.default_rng([seed, b])
        wb = rng.uniform(0.5, 2.0, (hi - lo, nf))
        wb[rng.random((hi - lo, nf)) < 0.01] = 0.0
        neg = rng.random((hi - lo, nf)) < 0.001
        wb[neg] = -wb[neg]
        sig = 1.0 / np.sqrt(np.where(wb > 0, wb, 1.0))
        rb = rng.normal(size=(hi - lo, nf)) * sig
        ib = rng.normal(size=(hi - lo, nf)) * sig
        a, z = max(lo, start), min(hi, stop)
        if z > a:
            re[a - start:z - start] = rb[a - lo:z - lo]
            im[a - start:z - start] = ib[a - lo:z - lo]
            w[a - start:z - start] = wb[a - lo:z - lo]
    return re, im, w


def synth_data_shard(nuv, nf, rank, world, seed=4321):
    """(rows, re, im, w) of rank's shard of the Hermitian-doubled data set (both halves of its baselines, as
    pdspy_b200.dist.shard_rows cuts them); the union over ranks is the world == 1 data set, bit for bit."""
    half = nuv // 2
    base, rem = divmod(half, world)
    s = rank * base + min(rank, rem)
    e = s + base + (1 if rank < rem else 0)
    re, im, w = synth_data_half_rows(nuv, nf, s, e, seed)
    rows = np.concatenate([np.arange(s, e), np.arange(half + s, half + e)])
    return rows, np.concatenate([re, re]), np.concatenate([im, -im]), np.concatenate([w, w])


class SynthImage:
    """Duck-typed stand-in for pdspy.imaging.Image (imaging/libimaging.pyx:7-48):
    only the attributes the hot path reads (image, x, y, freq)."""

    def __init__(self, image, pixelsize_arcsec, freq):
        n = image.shape[1]
        self.image = image
        self.x = (np.arange(n) - n / 2) * pixelsize_arcsec
        self.y = (np.arange(image.shape[0]) - image.shape[0] / 2) * pixelsize_arcsec
        self.freq = freq


CONFIGS = {
    # name: (npix, nchan, nuv, pixelsize["], dRA["], dDec["])
    "C1": (256, 1, 50_000, 0.1, 0.05, -0.03),
    "C2": (1024, 1, 1_000_000, 0.01, 0.13, -0.07),
    "C3": (512, 64, 1_000_000, 0.02, 0.02, -0.01),
}


def make_config(name, kind="disk", nuv=None, seed=1234):
    n, nf, nuv0, px, dra, ddec = CONFIGS[name]
    nuv = nuv0 if nuv is None else nuv
    u, v = synth_uv(nuv, px * ARCSEC, seed=seed)
    freq = synth_freq(nf)
    model = SynthImage(synth_image(n, nf, px, kind=kind), px, freq)
    return dict(u=u, v=v, freq=freq, model=model, dRA=dra, dDec=ddec, npix=n, nf=nf, pixelsize=px)
