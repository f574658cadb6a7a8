"""Driver for ncu: grid() of BASELINE.json configs[3] (10M visibilities -> 2048^2, exp*sinc, fast mode), device-resident
inputs.  usage: prof_grid.py [REPS]"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import pdspy_b200 as pb
import synth
from pdspy_b200 import _lib
from oracle import grid as og
L = _lib.lib()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nvis, G, px = 10_000_000, 2048, 0.01
u, v = synth.synth_uv(nvis, px * synth.ARCSEC)
_, re, im, w = synth.synth_data_shard(nvis, 1, 0, 1, seed=777)
freq = synth.synth_freq(1)
binsize = 2.2 * np.hypot(u, v).max() / G
uu, vv = og.cell_centres(G, binsize)
d = {k: pb.DeviceBuffer.from_numpy(a) for k, a in dict(u=u, v=v, freq=freq, re=re, im=im, w=w, uu=uu, vv=vv).items()}
o = [pb.DeviceBuffer(G * G * 8) for _ in range(3)]
for r in range(reps):
    _lib.check(L.pdsb_timer_start())
    _lib.check(L.pdsb_grid(_lib.ptr(d["u"]), _lib.ptr(d["v"]), _lib.ptr(d["freq"]), _lib.ptr(d["re"]), _lib.ptr(d["im"]),
                           _lib.ptr(d["w"]), nvis, 1, _lib.DEVICE, G, float(binsize), _lib.ptr(d["uu"]), _lib.ptr(d["vv"]),
                           _lib.CONV["expsinc"], _lib.WEIGHTING["natural"], 2.0, 0, 0, 0, 0,
                           _lib.ptr(o[0]), _lib.ptr(o[1]), _lib.ptr(o[2]), None, None, None, _lib.DEVICE, None))
    ms = ctypes.c_double(); _lib.check(L.pdsb_timer_stop(ctypes.byref(ms)))
    print("grid rep", r, "ms %.3f" % ms.value, flush=True)
