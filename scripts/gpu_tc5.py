"""Bring-up of the experimental tcgen05 DFT variant (200): parity on small shapes, then timing."""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pdspy_b200 as pb
from pdspy_b200 import _lib, synth
from pdspy_b200.interferometry import interpolate_model
from oracle import dft as od
A = synth.ARCSEC
L = _lib.lib()
var = int(sys.argv[1]) if len(sys.argv) > 1 else 200
L.pdsb_set_dft_variant(var)
for (ny, nx, nf, nuv, herm) in [(64, 64, 1, 100, False), (64, 64, 3, 400, True), (63, 65, 2, 333, False), (130, 34, 9, 64, True),
                                (256, 256, 1, 5000, True)]:
    rng = np.random.default_rng(ny * 1000 + nx)
    img = rng.random((ny, nx, nf, 1))
    if herm:
        u, v = synth.synth_uv(nuv, 0.1 * A)
    else:
        u = rng.normal(0, 3e5, nuv); v = rng.normal(0, 3e5, nuv)
    m = synth.SynthImage(img, 0.1, synth.synth_freq(nf))
    ref = od.exact_dft(u, v, img, 0.1 * A, 0.05 * A, -0.03 * A)
    for split in (1, 0):
        L.pdsb_set_dft_split(split)
        vis = interpolate_model(u, v, m.freq, m, dRA=0.05, dDec=-0.03)
        e = np.abs(vis.real + 1j * vis.imag - ref).max() / np.abs(ref).max()
        print((ny, nx, nf, nuv, herm), "variant", var, "split", split, "relerr %.2e" % e, "" if e < 1e-5 else " <<<<< FAIL", flush=True)
L.pdsb_set_dft_split(0)
for wl, nuv in (("C2", None), ("C3", 250_000)):
    c = synth.make_config(wl, nuv=nuv)
    img = np.ascontiguousarray(c["model"].image[:, :, :, 0])
    ds = pb.Dataset(c["u"], c["v"])
    dimg = pb.DeviceBuffer.from_numpy(img)
    n, nf, nuv = c["npix"], c["nf"], c["u"].size
    dre, dim_ = pb.DeviceBuffer(nuv * nf * 8), pb.DeviceBuffer(nuv * nf * 8)
    pairs = float(n) * n * nf * nuv
    sub = np.random.default_rng(1).choice(nuv, 256, replace=False)
    ref = od.exact_dft(c["u"][sub], c["v"][sub], c["model"].image, c["pixelsize"] * A, c["dRA"] * A, c["dDec"] * A)
    ts = []
    for rep in range(3):
        _lib.check(L.pdsb_profile_reset()); _lib.check(L.pdsb_profile_enable(1))
        _lib.check(L.pdsb_sample_image(ds.handle, _lib.ptr(dimg), n, n, nf, _lib.DEVICE, c["pixelsize"] * A, c["dRA"] * A,
                                       c["dDec"] * A, _lib.ptr(dre), _lib.ptr(dim_), _lib.DEVICE))
        ms, cnt = ctypes.c_double(), ctypes.c_int64()
        _lib.check(L.pdsb_profile_get(b"dft_", ctypes.byref(ms), ctypes.byref(cnt)))
        ts.append(ms.value)
    _lib.check(L.pdsb_profile_enable(0))
    V = dre.download((nuv, nf)) + 1j * dim_.download((nuv, nf))
    err = (np.abs(V[sub] - ref) / np.abs(V).max(axis=0)).max()
    print(wl, "variant", var, "dft kernel ms", [round(x, 2) for x in ts], "pairs/s %.3e" % (pairs / min(ts[1:]) * 1e3), "relerr %.1e" % err, flush=True)
