"""Driver for ncu: one C3 likelihood through the galario-algorithm path (pdsb_loglike_nufft), device-resident cube."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import pdspy_b200 as pb
import synth
from pdspy_b200 import _lib
A = synth.ARCSEC
c = synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "C3")
re, im, w = synth.synth_data(c["u"].size, c["nf"])
ds = pb.Dataset(c["u"], c["v"]); ds.set_data(re, im, w)
cube = pb.DeviceBuffer.from_numpy(np.ascontiguousarray(c["model"].image[:, :, :, 0]))
out = np.empty(4)
L = _lib.lib()
import ctypes
for rep in range(3):
    if rep == 2:
        _lib.check(L.pdsb_profile_reset()); _lib.check(L.pdsb_profile_enable(1))
    _lib.check(L.pdsb_loglike_nufft(ds.handle, _lib.ptr(cube), c["npix"], c["npix"], c["nf"], _lib.DEVICE, c["pixelsize"] * A, c["dRA"] * A,
                                  c["dDec"] * A, _lib.ptr(out)))
print("lnlike", out[3])
_lib.check(L.pdsb_profile_enable(0))
for name in (b"rfft2_planes", b"nufft_chi2", b"reduce_columns", b"fft_twiddle"):
    t, n = ctypes.c_double(), ctypes.c_int64()
    L.pdsb_profile_get(name, ctypes.byref(t), ctypes.byref(n))
    if n.value:
        print(name.decode(), "ms %.3f" % t.value, "launch scopes", n.value)
