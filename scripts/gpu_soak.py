"""Soak: repeated likelihood calls are bit-reproducible and leak nothing (device memory in use stays flat)."""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import pdspy_b200 as pb
import synth
from pdspy_b200.interferometry import loglike_image, interpolate_model, Visibilities, grid
import torch


def used():
    free, total = torch.cuda.mem_get_info()
    return (total - free) / 2 ** 20


for kern in ("fp32", "tcgen05", "fp64", "nufft", "nufft-cube"):
    pb.set_dft_kernel(kern.split("-")[0])
    c = synth.make_config("C1") if kern != "nufft-cube" else synth.make_config("C3", nuv=100_000)      # cube: the tiled sampler
    re, im, w = synth.synth_data(c["u"].size, c["nf"])
    d = Visibilities(c["u"], c["v"], c["freq"], re, im, w)
    first, m0 = None, None
    t0 = time.time()
    for it in range(1500 if kern != "fp64" else 300):
        ll, chi = loglike_image(d, c["model"], dRA=c["dRA"], dDec=c["dDec"])
        if first is None:
            first = (ll, chi.copy())
        assert ll == first[0] and np.array_equal(chi, first[1]), (kern, it)
        if it == 50:
            m0 = used()
    print("%-8s %d identical likelihoods in %.1f s, device memory %.0f -> %.0f MiB" % (kern, it + 1, time.time() - t0, m0, used()), flush=True)
pb.set_dft_kernel("fp32")
u, v = synth.synth_uv(200_000, 0.01 * synth.ARCSEC)
re, im, w = synth.synth_data(200_000, 1)
d = Visibilities(u, v, synth.synth_freq(1), re, im, w)
ref = None
for it in range(50):
    g = grid(d, gridsize=512, binsize=2.2 * np.hypot(u, v).max() / 512, convolution="expsinc")
    if ref is None:
        ref = g.real.copy()
    assert np.array_equal(g.real, ref)
print("ordered grid: 50 identical results; device memory %.0f MiB" % used(), flush=True)
# the fast gridding mode is NOT bit-reproducible (atomic flush order) but must stay within rounding and leak nothing
ref, m0 = None, None
for it in range(200):
    g = grid(d, gridsize=512, binsize=2.2 * np.hypot(u, v).max() / 512, convolution="expsinc", deterministic=False)
    if ref is None:
        ref = g.real.copy()
    assert np.abs(g.real - ref).max() <= 1e-12 * np.abs(ref).max()
    if it == 20:
        m0 = used()
print("fast grid: 200 results within 1e-12; device memory %.0f -> %.0f MiB" % (m0, used()), flush=True)
