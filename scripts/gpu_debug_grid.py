import sys, os, io, contextlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
from make_golden import GRID_CASES_FIXTURE
from oracle import grid as og
from pdspy_b200.interferometry import grid, Visibilities
d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "fixture_720.npz"))
f = {k: d[k] for k in d.files}
data = Visibilities(f["u"], f["v"], f["freq"], f["real"], f["imag"], f["weights"])
for name in ("fx_superuniform", "fx_uniform", "fx_robust", "fx_expsinc_img256"):
    kw = GRID_CASES_FIXTURE[name]
    with contextlib.redirect_stdout(io.StringIO()):
        o = og.grid(f["u"], f["v"], f["freq"], f["real"], f["imag"], f["weights"], return_maps=True, **kw)
    for det in (True, False):
        with contextlib.redirect_stdout(io.StringIO()):
            g, gi, gj, wm = grid(data, deterministic=det, return_maps=True, **kw)
        print(name, "det" if det else "atomic", "wmod equal", np.array_equal(wm, o[9]), "max wmod diff", np.abs(wm - o[9]).max())
        for nm, ob in zip(("real", "imag", "weights"), o[3:6]):
            gb = getattr(g, nm)
            nz_g, nz_o = gb != 0, ob != 0
            bad = np.flatnonzero(nz_g != nz_o)
            print("  ", nm, "nnz gpu", nz_g.sum(), "oracle", nz_o.sum(), "maxdiff %.2e" % (np.abs(gb - ob).max() / np.abs(ob).max()), "pattern diffs at", bad[:10])
            for q in bad[:3]:
                l, m = divmod(int(q), kw["gridsize"])
                print("      cell", (l, m), "gpu", gb[q], "oracle", ob[q], "weights gpu/oracle", g.weights[q], o[5][q])
                # contributions from oracle's point of view
                i, j = o[6][:, 0], o[7][:, 0]
                near = np.flatnonzero((np.abs(i.astype(int) - m) <= 3) & (np.abs(j.astype(int) - l) <= 3))
                for k in near:
                    print("         k", k, "home", (int(j[k]), int(i[k])), "u,v", f["u"][k], f["v"][k], "re,im", f["real"][k, 0], f["imag"][k, 0], "w", o[9][k, 0], "wgpu", wm[k, 0])
