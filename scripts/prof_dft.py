"""Small driver for ncu: runs the device-resident image->visibility path a few times.
usage: prof_dft.py WORKLOAD VARIANT REPS [NUV]"""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import pdspy_b200 as pb
import synth
from pdspy_b200 import _lib
A = synth.ARCSEC
wl, variant, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
nuv = int(sys.argv[4]) if len(sys.argv) > 4 else None
c = synth.make_config(wl, nuv=nuv)
L = _lib.lib()
L.pdsb_set_dft_variant(variant)
img = np.ascontiguousarray(c["model"].image[:, :, :, 0])
ds = pb.Dataset(c["u"], c["v"])
dimg = pb.DeviceBuffer.from_numpy(img)
n, nf, nuv = c["npix"], c["nf"], c["u"].size
dre, dim_ = pb.DeviceBuffer(nuv * nf * 8), pb.DeviceBuffer(nuv * nf * 8)
for r in range(reps):
    _lib.check(L.pdsb_timer_start())
    _lib.check(L.pdsb_sample_image(ds.handle, _lib.ptr(dimg), n, n, nf, _lib.DEVICE, c["pixelsize"] * A, c["dRA"] * A,
                                   c["dDec"] * A, _lib.ptr(dre), _lib.ptr(dim_), _lib.DEVICE))
    ms = ctypes.c_double(); _lib.check(L.pdsb_timer_stop(ctypes.byref(ms)))
    print(wl, "variant", variant, "rep", r, "ms %.3f" % ms.value, "pairs/s %.3e" % (float(n) * n * nf * nuv / ms.value * 1e3), flush=True)
