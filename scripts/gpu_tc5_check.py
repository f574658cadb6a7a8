"""tcgen05 kernel against the fp64 kernel on full C2 / C3 (+ a high-dynamic-range cube): lnlike, per-channel
chi^2, max visibility error, kernel time (GPU)."""
import ctypes
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import synth
import pdspy_b200 as pb
from pdspy_b200 import _lib
from pdspy_b200.interferometry import loglike_image, Visibilities, interpolate_model

L = _lib.lib()
variants = [("fp64", 300), ("fp32", 0), ("tc5", 200), ("tc5-order0", 201)]
for wl, hdr in (("C2", False), ("C3", False), ("C3", True)):
    c = synth.make_config(wl)
    if hdr:   # a star pixel 1e5 x the disk peak in every channel, one empty channel
        img = c["model"].image
        n = img.shape[0]
        img[n // 2, n // 2, :, 0] = 1e5 * img.max()
        img[:, :, 3, 0] = 0.0
    re, im, w = synth.synth_data(c["u"].size, c["nf"])
    d = Visibilities(c["u"], c["v"], c["freq"], re, im, w)
    out, vis = {}, {}
    for name, var in variants:
        _lib.check(L.pdsb_set_dft_variant(var))
        out[name] = loglike_image(d, c["model"], dRA=c["dRA"], dDec=c["dDec"])
        _lib.check(L.pdsb_profile_reset()); _lib.check(L.pdsb_profile_enable(1))
        for _ in range(3):
            loglike_image(d, c["model"], dRA=c["dRA"], dDec=c["dDec"])
        _lib.check(L.pdsb_profile_enable(0))
        ms, nl = ctypes.c_double(), ctypes.c_int64()
        _lib.check(L.pdsb_profile_get(b"dft_", ctypes.byref(ms), ctypes.byref(nl)))
        if wl == "C2" or hdr:
            v = interpolate_model(c["u"], c["v"], c["freq"], c["model"], dRA=c["dRA"], dDec=c["dDec"])
            vis[name] = v.real + 1j * v.imag
        out[name] = out[name] + (ms.value / max(nl.value, 1),)
    _lib.check(L.pdsb_set_dft_variant(0))
    ll0, chi0, t0 = out["fp64"]
    for name, _ in variants[1:]:
        ll, chi, t = out[name]
        line = "%s%s %-11s lnlike rel %.2e  chi2/channel max rel %.2e  dft kernel %.3f ms" % (
            wl, " HDR" if hdr else "", name, abs(ll - ll0) / abs(ll0), np.nanmax(np.abs(chi - chi0) / np.where(chi0 != 0, np.abs(chi0), np.nan)), t)
        if name in vis:
            sc = np.abs(vis["fp64"]).max(axis=0)
            sc[sc == 0] = 1.0
            line += "  max |dV|/max|V| %.2e" % (np.abs(vis[name] - vis["fp64"]) / sc).max()
        print(line, flush=True)
