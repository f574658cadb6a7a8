"""Per-call latency of the public API on the small configuration (C1: 256^2, 50k uv): what an MCMC step pays."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import pdspy_b200 as pb
import synth
from pdspy_b200.interferometry import interpolate_model, loglike_image, Visibilities
for kern in ("fp32", "tcgen05", "nufft"):
    pb.set_dft_kernel(kern)
    for wl in ("C1", "C2"):
        c = synth.make_config(wl, nuv=None if wl == "C1" else 200_000)
        re, im, w = synth.synth_data(c["u"].size, c["nf"])
        data = Visibilities(c["u"], c["v"], c["freq"], re, im, w)
        for name, fn in (("interpolate_model", lambda: interpolate_model(c["u"], c["v"], c["freq"], c["model"], dRA=c["dRA"], dDec=c["dDec"])),
                         ("loglike_image", lambda: loglike_image(data, c["model"], dRA=c["dRA"], dDec=c["dDec"]))):
            for _ in range(3):
                fn()
            n = 30
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            dt = (time.perf_counter() - t0) / n
            print("%-8s %s nuv=%d %-18s %.3f ms per call" % (kern, wl, c["u"].size, name, dt * 1e3), flush=True)
