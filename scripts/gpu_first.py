"""First GPU bring-up: correctness of every entry point against the oracle + variant timings."""
import ctypes, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pdspy_b200 as pb
from pdspy_b200 import _lib, synth
from pdspy_b200.interferometry import interpolate_model, grid, chisq, Visibilities, loglike_image
from oracle import dft as od, grid as og, likelihood as ol

A = synth.ARCSEC
L = _lib.lib()

def relerr(a, b):
    return np.abs(a - b).max() / np.abs(b).max()

print("== device", flush=True)
sm, khz, mem, maj, mnr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int64(), ctypes.c_int(), ctypes.c_int()
_lib.check(L.pdsb_device_info(ctypes.byref(sm), ctypes.byref(khz), ctypes.byref(mem), ctypes.byref(maj), ctypes.byref(mnr)))
print("SMs", sm.value, "clock kHz", khz.value, "mem GB", mem.value / 1e9, "cc", maj.value, mnr.value)

print("== fma microbench")
for var in (0, 1):
    tf, ms = ctypes.c_double(), ctypes.c_double()
    for rep in range(3):
        _lib.check(L.pdsb_bench_fma(var, 20000, ctypes.byref(tf), ctypes.byref(ms)))
    print("variant", var, "TFLOP/s", round(tf.value, 2), "ms", round(ms.value, 3), flush=True)

print("== small DFT parity, all variants")
for (ny, nx, nf, nuv, herm) in [(64, 64, 3, 400, True), (63, 65, 2, 333, False), (40, 72, 33, 130, True), (8, 8, 1, 5, False)]:
    px = 0.1
    rng = np.random.default_rng(ny * 1000 + nx)
    img = rng.random((ny, nx, nf, 1))
    if herm:
        u, v = synth.synth_uv(nuv, px * A)
    else:
        u = rng.normal(0, 3e5, nuv); v = rng.normal(0, 3e5, nuv)
    model = synth.SynthImage(img, px, synth.synth_freq(nf))
    ref = od.exact_dft(u, v, img, px * A, 0.05 * A, -0.03 * A)
    for var in range(1, 7):
        L.pdsb_set_dft_variant(var)
        for split in (0, 1, 3):
            L.pdsb_set_dft_split(split)
            vis = interpolate_model(u, v, model.freq, model, dRA=0.05, dDec=-0.03)
            e = relerr(vis.real + 1j * vis.imag, ref)
            flag = "" if e < 1e-5 else "  <<<<<< FAIL"
            print((ny, nx, nf, nuv, herm), "variant", var, "split", split, "relerr %.2e" % e, flag, flush=True)
L.pdsb_set_dft_variant(0); L.pdsb_set_dft_split(0)

print("== C1 parity (256^2, 50k) + fixture-like")
c1 = synth.make_config("C1")
t = time.time(); ref = od.exact_dft(c1["u"], c1["v"], c1["model"].image, c1["pixelsize"] * A, c1["dRA"] * A, c1["dDec"] * A); t_or = time.time() - t
for var in range(1, 7):
    L.pdsb_set_dft_variant(var)
    t = time.time(); vis = interpolate_model(c1["u"], c1["v"], c1["freq"], c1["model"], dRA=c1["dRA"], dDec=c1["dDec"]); t1 = time.time() - t
    t = time.time(); vis = interpolate_model(c1["u"], c1["v"], c1["freq"], c1["model"], dRA=c1["dRA"], dDec=c1["dDec"]); t2 = time.time() - t
    print("C1 variant", var, "relerr %.2e" % relerr(vis.real + 1j * vis.imag, ref), "wall first %.3f s second %.4f s (oracle separable %.1f s)" % (t1, t2, t_or), flush=True)
L.pdsb_set_dft_variant(0)
c1r = synth.make_config("C1", kind="random")
ref = od.exact_dft(c1r["u"], c1r["v"], c1r["model"].image, c1r["pixelsize"] * A, c1r["dRA"] * A, c1r["dDec"] * A)
vis = interpolate_model(c1r["u"], c1r["v"], c1r["freq"], c1r["model"], dRA=c1r["dRA"], dDec=c1r["dDec"])
print("C1 random image relerr %.2e" % relerr(vis.real + 1j * vis.imag, ref), flush=True)

print("== likelihood")
re, im, w = synth.synth_data(c1["u"].size, 1, model=(ref.real * 0 + vis.real, vis.imag))
data = Visibilities(c1["u"], c1["v"], c1["freq"], re, im, w)
ll, chi2 = loglike_image(data, c1r["model"], dRA=c1r["dRA"], dDec=c1r["dDec"])
ll_ref = ol.lnlike_vis_numpy(re, im, w, ref.real, ref.imag)
print("fused lnlike", ll, "oracle(exact model)", ll_ref, "rel %.2e" % (abs(ll - ll_ref) / abs(ll_ref)))
ll2 = pb.utils.visibility_lnlike(data, vis)
ll2_ref = ol.lnlike_vis_numpy(re, im, w, vis.real, vis.imag)
print("unfused lnlike", ll2, "numpy same arrays", ll2_ref, "rel %.2e" % (abs(ll2 - ll2_ref) / abs(ll2_ref)))
print("fused vs unfused rel %.2e" % (abs(ll - ll2) / abs(ll2)))
print("chisq", chisq(data, vis), ol.chisq_c(re, im, w, vis.real, vis.imag))

print("== grid parity vs oracle")
rng = np.random.default_rng(7)
n, nf = 20000, 3
u = rng.normal(0, 2e5, n); v = rng.normal(0, 2e5, n); freq = 230e9 + 1e8 * np.arange(nf)
re = rng.normal(size=(n, nf)); im = rng.normal(size=(n, nf)); w = rng.uniform(0.5, 2, (n, nf))
w[rng.random((n, nf)) < 0.01] = 0; w[rng.random((n, nf)) < 0.001] *= -1
d = Visibilities(u, v, freq, re, im, w)
cases = [dict(gridsize=128, binsize=8000., convolution="pillbox"),
         dict(gridsize=128, binsize=8000., convolution="expsinc"),
         dict(gridsize=128, binsize=8000., convolution="pillbox", mode="spectralline"),
         dict(gridsize=129, binsize=8000., convolution="expsinc", mode="spectralline", imaging=True),
         dict(gridsize=128, binsize=8000., convolution="pillbox", weighting="uniform", mode="spectralline"),
         dict(gridsize=128, binsize=8000., convolution="expsinc", weighting="superuniform", mode="spectralline"),
         dict(gridsize=64, binsize=8000., convolution="pillbox", weighting="robust", robust=0.5, mode="spectralline"),
         dict(gridsize=128, binsize=8000., convolution="expsinc", mfs=True),
         dict(gridsize=128, binsize=8000., convolution="pillbox", channel=1, imaging=True)]
for kw in cases:
    o = og.grid(u, v, freq, re, im, w, return_maps=True, **kw)
    for det in (True, False):
        t = time.time()
        g, gi, gj, wm = grid(d, deterministic=det, return_maps=True, **kw)
        dt = time.time() - t
        names = ["real", "imag", "weights"]
        res = []
        for nm, ob in zip(names, o[3:6]):
            gb = getattr(g, nm)
            res.append("%s %s %.1e" % (nm, "EXACT" if np.array_equal(gb, ob) else "diff", np.abs(gb - ob).max() / max(np.abs(ob).max(), 1e-300)))
        maps = np.array_equal(gi, o[6]) and np.array_equal(gj, o[7])
        print(kw, "det" if det else "atomic", "maps", "EXACT" if maps else "DIFF", res, "wmod", np.array_equal(wm, o[9]), "%.3fs" % dt, flush=True)

print("== variant timing on C2 (1024^2 x 1M uv), device-resident")
c2 = synth.make_config("C2")
img = np.ascontiguousarray(c2["model"].image[:, :, :, 0])
ds = pb.Dataset(c2["u"], c2["v"])
dimg = pb.DeviceBuffer.from_numpy(img)
nuv = c2["u"].size
dre = pb.DeviceBuffer(nuv * 8); dim_ = pb.DeviceBuffer(nuv * 8)
pairs = 1024.0 * 1024.0 * nuv
for var in range(1, 7):
    L.pdsb_set_dft_variant(var)
    for split in ((0, 1, 2, 4, 8) if var in (1, 3) else (0,)):
        L.pdsb_set_dft_split(split)
        ts = []
        for rep in range(4):
            _lib.check(L.pdsb_timer_start())
            _lib.check(L.pdsb_sample_image(ds.handle, _lib.ptr(dimg), 1024, 1024, 1, _lib.DEVICE, 0.01 * A, 0.13 * A, -0.07 * A, _lib.ptr(dre), _lib.ptr(dim_), _lib.DEVICE))
            ms = ctypes.c_double(); _lib.check(L.pdsb_timer_stop(ctypes.byref(ms))); ts.append(ms.value)
        best = min(ts[1:])
        print("C2 variant", var, "split", split, "ms", [round(x, 2) for x in ts], "pairs/s %.3e" % (pairs / best * 1e3), "alg TFLOP/s %.1f" % (pairs * 4 / best * 1e3 / 1e12), flush=True)
L.pdsb_set_dft_variant(0); L.pdsb_set_dft_split(0)
vre = dre.download((nuv, 1)); vim = dim_.download((nuv, 1))
sub = np.random.default_rng(1).choice(nuv, 2048, replace=False)
ref = od.exact_dft(c2["u"][sub], c2["v"][sub], c2["model"].image, 0.01 * A, 0.13 * A, -0.07 * A)
full_max = np.abs(vre + 1j * vim).max()
print("C2 subset parity relerr(max|V| of all) %.2e" % (np.abs((vre + 1j * vim)[sub] - ref).max() / full_max))
print("DONE")
