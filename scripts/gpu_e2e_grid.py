"""Where the end-to-end time of grid() on numpy arrays goes (BASELINE configs[3]: 10M visibilities -> 2048^2)."""
import contextlib, cProfile, io, os, pstats, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import synth
from pdspy_b200.interferometry import grid, Visibilities
nvis, G = 10_000_000, 2048
u, v = synth.synth_uv(nvis, 0.01 * synth.ARCSEC)
re, im, w = synth.synth_data(nvis, 1)
data = Visibilities(u, v, synth.synth_freq(1), re, im, w)
binsize = 2.2 * np.hypot(u, v).max() / G
def step():
    with contextlib.redirect_stdout(io.StringIO()):
        return grid(data, gridsize=G, binsize=binsize, convolution="expsinc", deterministic=False)
for _ in range(2):
    step()
t0 = time.perf_counter()
for _ in range(5):
    g = step()
print("e2e ms per grid()", (time.perf_counter() - t0) / 5 * 1e3, "weight sum", g.weights.sum())
pr = cProfile.Profile(); pr.enable(); step(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(14); print(s.getvalue()[:2500])
