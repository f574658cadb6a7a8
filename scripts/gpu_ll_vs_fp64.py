import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
import pdspy_b200 as pb
from pdspy_b200 import synth
from pdspy_b200.interferometry import loglike_image, Visibilities
for wl in ("C2", "C3"):
    c = synth.make_config(wl)
    re, im, w = synth.synth_data(c["u"].size, c["nf"])
    d = Visibilities(c["u"], c["v"], c["freq"], re, im, w)
    out = {}
    for k in ("fp64", "fp32", "tcgen05", "mma"):
        pb.set_dft_kernel(k)
        out[k] = loglike_image(d, c["model"], dRA=c["dRA"], dDec=c["dDec"])
    pb.set_dft_kernel("fp32")
    for k in ("fp32", "tcgen05", "mma"):
        print(wl, k, "lnlike rel diff vs fp64 %.2e" % (abs(out[k][0] - out["fp64"][0]) / abs(out["fp64"][0])),
              "chi2/channel max rel %.2e" % (np.abs(out[k][1] - out["fp64"][1]) / np.abs(out["fp64"][1])).max(), flush=True)
