"""Gridding timing on BASELINE.json configs[3]: 10M visibilities -> 2048^2 (device-resident inputs)."""
import ctypes, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import pdspy_b200 as pb
import synth
from pdspy_b200 import _lib
from oracle import grid as og
L = _lib.lib()
nvis = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
G = 2048
px = 0.01
u, v = synth.synth_uv(nvis, px * synth.ARCSEC)
re, im, w = synth.synth_data(nvis, 1)
freq = synth.synth_freq(1)
binsize = 2.2 * np.hypot(u, v).max() / G
uu, vv = og.cell_centres(G, binsize)
d = {k: pb.DeviceBuffer.from_numpy(a) for k, a in dict(u=u, v=v, freq=freq, re=re, im=im, w=w, uu=uu, vv=vv).items()}
o_re, o_im, o_w = (pb.DeviceBuffer(G * G * 8) for _ in range(3))
nout = ctypes.c_int64()
alg_bytes = nvis * 40 + G * G * 24
for conv in ("pillbox", "expsinc"):
    for wt in ("natural", "robust"):
        for det in ((0,) if os.environ.get("FAST_ONLY") else (0, 1)):
            ts = []
            for rep in range(3):
                _lib.check(L.pdsb_profile_reset()); _lib.check(L.pdsb_profile_enable(1))
                _lib.check(L.pdsb_timer_start())
                _lib.check(L.pdsb_grid(_lib.ptr(d["u"]), _lib.ptr(d["v"]), _lib.ptr(d["freq"]), _lib.ptr(d["re"]), _lib.ptr(d["im"]),
                                       _lib.ptr(d["w"]), nvis, 1, _lib.DEVICE, G, float(binsize), _lib.ptr(d["uu"]), _lib.ptr(d["vv"]),
                                       _lib.CONV[conv], _lib.WEIGHTING[wt], 0.5, 0, 0, 0, det,
                                       _lib.ptr(o_re), _lib.ptr(o_im), _lib.ptr(o_w), None, None, None, _lib.DEVICE, ctypes.byref(nout)))
                ms = ctypes.c_double(); _lib.check(L.pdsb_timer_stop(ctypes.byref(ms))); ts.append(ms.value)
            _lib.check(L.pdsb_profile_enable(0))
            parts = {}
            for name in (b"grid_prep", b"grid_tile_hist", b"grid_tile_scan", b"grid_tile_records", b"grid_tile_accum", b"grid_sort_hist", b"grid_sort_scan", b"grid_sort_scatter", b"grid_scatter_atomic", b"grid_values", b"grid_ordered_sum",
                         b"grid_emit_keys", b"grid_normalise", b"grid_sum", b"grid_reweight", b"grid_fill"):
                t, n = ctypes.c_double(), ctypes.c_int64()
                L.pdsb_profile_get(name, ctypes.byref(t), ctypes.byref(n))
                if n.value:
                    parts[name.decode()] = (round(t.value, 3), n.value)
            best = min(ts[1:])
            print(conv, wt, "det" if det else "atomic", "ms", [round(x, 2) for x in ts], "Mvis/s %.1f" % (nvis / best / 1e3),
                  "alg GB/s %.1f" % (alg_bytes / best / 1e6), parts, flush=True)
