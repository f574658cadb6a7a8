"""Per-kernel time of one likelihood step with the tcgen05 variant (in-library events)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pdspy_b200 as pb
from pdspy_b200 import _lib, synth
from pdspy_b200.interferometry import Visibilities
A = synth.ARCSEC
L = _lib.lib()
for wl, nuv in (("C2", None), ("C3", 250_000)):
    c = synth.make_config(wl, nuv=nuv)
    re, im, w = synth.synth_data(c["u"].size, c["nf"])
    ds = pb.Dataset(c["u"], c["v"]); ds.set_data(re, im, w)
    img = np.ascontiguousarray(c["model"].image[:, :, :, 0]); dimg = pb.DeviceBuffer.from_numpy(img)
    n, nf = c["npix"], c["nf"]
    chi2 = np.empty(nf); ll = ctypes.c_double()
    for var in (0, 200):
        L.pdsb_set_dft_variant(var)
        for rep in range(3):
            _lib.check(L.pdsb_profile_reset()); _lib.check(L.pdsb_profile_enable(1))
            _lib.check(L.pdsb_loglike(ds.handle, _lib.ptr(dimg), n, n, nf, _lib.DEVICE, c["pixelsize"] * A, c["dRA"] * A, c["dDec"] * A,
                                      _lib.ptr(chi2), ctypes.cast(ctypes.byref(ll), ctypes.c_void_p)))
        out = {}
        for name in (b"plane_absmax", b"plane_scale", b"fold_", b"dft_", b"loglike_epi", b"reduce", b""):
            ms, cnt = ctypes.c_double(), ctypes.c_int64()
            _lib.check(L.pdsb_profile_get(name, ctypes.byref(ms), ctypes.byref(cnt)))
            out[name.decode() or "ALL"] = (round(ms.value, 3), cnt.value)
        _lib.check(L.pdsb_profile_enable(0))
        print(wl, "variant", var, out, flush=True)
