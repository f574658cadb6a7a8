"""Time every DFT kernel variant on C2 and on the C3 shape (reduced uv), check parity on a subset."""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pdspy_b200 as pb
from pdspy_b200 import _lib, synth
from oracle import dft as od
A = synth.ARCSEC
L = _lib.lib()
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
    for var in (11, 100, 101, 102, 103, 104):
        L.pdsb_set_dft_variant(var)
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
        best = min(ts[1:])
        print(wl, "variant %2d" % var, "dft kernel ms", [round(x, 2) for x in ts], "pairs/s %.3e" % (pairs / best * 1e3),
              "executed frac %.3f" % (pairs / 2 * 2 / best * 1e3 / 74.45e12), "relerr %.1e" % err, flush=True)
