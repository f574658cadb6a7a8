import sys, os, io, contextlib
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from test_gpu_grid import multi_channel_set, _quiet
from pdspy_b200.interferometry import grid, Visibilities
from pdspy_b200 import dist as pdist, _lib
u, v, freq, re, im, w = multi_channel_set()
d = Visibilities(u, v, freq, re, im, w)
for kw in [dict(weighting="uniform", npixels=1, convolution="expsinc"), dict(weighting="uniform", npixels=1, convolution="pillbox"), dict(weighting="natural", convolution="expsinc")]:
    ref, _ = _quiet(grid, d, gridsize=128, binsize=8000., deterministic=True, **kw)
    fast, _ = _quiet(grid, d, gridsize=128, binsize=8000., deterministic=False, **kw)
    got, _ = _quiet(pdist.sharded_grid, d, gridsize=128, binsize=8000., **kw)
    for nm in ("real", "imag", "weights"):
        a, b, c = getattr(got, nm), getattr(ref, nm), getattr(fast, nm)
        print(kw, nm, "sharded-vs-ref %.3e fast-vs-ref %.3e" % (np.abs(a - b).max() / np.abs(b).max(), np.abs(c - b).max() / np.abs(b).max()), "nnz", np.count_nonzero(a), np.count_nonzero(b))
