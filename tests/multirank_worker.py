"""Worker of tests/test_gpu_multirank.py: run under torch.distributed.run with N >= 2 ranks, one GPU each.
Every check compares the N-rank result with the same computation on rank 0's GPU alone; prints one line per
check and exits non-zero on the first failure."""
import contextlib
import ctypes
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch                                   # noqa: E402
import torch.distributed as dist               # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ["PDSB_DEVICE"] = str(lr)
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))

import synth                                   # noqa: E402
from pdspy_b200 import _lib, dist as pdist, Dataset     # noqa: E402
from pdspy_b200.interferometry import Visibilities, grid, loglike_image     # noqa: E402

A = synth.ARCSEC
failures = []


def check(name, ok, detail=""):
    if rank == 0:
        print("%s %s %s" % ("PASS" if ok else "FAIL", name, detail), flush=True)
    if not ok:
        failures.append(name)


def quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


# ---- (e1) the likelihood of ONE data set, uv points sharded, against the unsharded call ----
c = synth.make_config("C3", nuv=60_000)
re, im, w = synth.synth_data(c["u"].size, c["nf"])
data = Visibilities(c["u"], c["v"], c["freq"], re, im, w)
like = pdist.ShardedLikelihood(pdist.shard_visibilities(data, rank, world))
cube = np.ascontiguousarray(c["model"].image[:, :, :, 0])
dxy = c["pixelsize"] * A
for kernel in ("fp32", "tcgen05", "nufft"):
    _lib.check(_lib.lib().pdsb_set_dft_variant({"fp32": 0, "tcgen05": 200, "nufft": 400}[kernel]))
    vals = {mode: like(cube, dxy, c["dRA"] * A, c["dDec"] * A, kind=0, cube=mode) for mode in (None, "sharded", "rank0")}
    single, _ = loglike_image(data, c["model"], dRA=c["dRA"], dDec=c["dDec"])
    for mode, val in vals.items():
        check("sharded likelihood (%s, cube=%s) == single GPU" % (kernel, mode), abs(val - single) <= 1e-12 * abs(single),
              "rel %.1e" % (abs(val - single) / abs(single)))
# ---- (e1') the same data set split by frequency channels (the partition that also divides the FFT-based kernels' transform) ----
c0, c1 = pdist.shard_channels(c["nf"], rank, world)
likec = pdist.ShardedLikelihood(pdist.shard_visibilities(data, rank, world, by="channels"), channels=(c0, c["nf"]))
for kernel in ("fp32", "nufft"):
    _lib.check(_lib.lib().pdsb_set_dft_variant({"fp32": 0, "nufft": 400}[kernel]))
    single, _ = loglike_image(data, c["model"], dRA=c["dRA"], dDec=c["dDec"])
    for mode in (None, "sharded"):
        val = likec(cube, dxy, c["dRA"] * A, c["dDec"] * A, kind=0, cube=mode)
        check("channel-sharded likelihood (%s, cube=%s) == single GPU" % (kernel, mode), abs(val - single) <= 1e-12 * abs(single),
              "rel %.1e" % (abs(val - single) / abs(single)))
likec.ds.destroy()
_lib.check(_lib.lib().pdsb_set_dft_variant(0))

# ---- (e2) walker batch: walkers split over the ranks, gathered, against single calls ----
nwalk = 6
ds = Dataset(c["u"], c["v"])
ds.set_data(re, im, w)
s, e = pdist.shard_walkers(nwalk, rank, world)
cubes = np.ascontiguousarray(np.stack([cube * (1 + 0.01 * k) for k in range(s, e)])) if e > s else np.zeros((0,) + cube.shape)
dra = np.ascontiguousarray(np.array([0.01 * k for k in range(s, e)]) * A)
ddec = np.ascontiguousarray(np.array([-0.02 * k for k in range(s, e)]) * A)
mine = np.zeros(max(e - s, 1))
if e > s:
    _lib.check(_lib.lib().pdsb_loglike_batch(ds.handle, _lib.ptr(cubes), e - s, cube.shape[0], cube.shape[1], c["nf"], _lib.HOST,
                                             float(dxy), _lib.ptr(dra), _lib.ptr(ddec), _lib.ptr(mine)))
gathered = [None] * world
dist.all_gather_object(gathered, mine[:e - s].tolist())
allw = np.array([x for part in gathered for x in part])
if rank == 0:
    exp = []
    for k in range(nwalk):
        one = ctypes.c_double()
        _lib.check(_lib.lib().pdsb_loglike(ds.handle, _lib.ptr(np.ascontiguousarray(cube * (1 + 0.01 * k))), cube.shape[0],
                                           cube.shape[1], c["nf"], _lib.HOST, float(dxy), 0.01 * k * A, -0.02 * k * A, None,
                                           ctypes.cast(ctypes.byref(one), ctypes.c_void_p)))
        exp.append(one.value)
    check("walker batch over %d ranks == single calls" % world, allw.size == nwalk and np.array_equal(allw, np.array(exp)))

# ---- (e3) gridding: data split (throughput mode) and output rows split (bit-exact mode), with re-weighting ----
u, v = synth.synth_uv(300_000, 0.01 * A)
gre, gim, gw = synth.synth_data(300_000, 2)
freq = synth.synth_freq(2)
gd = Visibilities(u, v, freq, gre, gim, gw)
G = 512
binsize = 2.2 * np.hypot(u, v).max() / G
for kw in (dict(convolution="expsinc", mode="spectralline", imaging=True),
           dict(convolution="pillbox", weighting="uniform", npixels=1),
           dict(convolution="expsinc", weighting="robust", robust=0.5, npixels=2, mode="spectralline")):
    ref = quiet(grid, gd, gridsize=G, binsize=binsize, deterministic=True, **kw) if rank == 0 else None
    g = quiet(pdist.sharded_grid, pdist.shard_visibilities(gd, rank, world), gridsize=G, binsize=binsize, **kw)
    if rank == 0:
        err = max(np.abs(getattr(g, nm) - getattr(ref, nm)).max() / np.abs(getattr(ref, nm)).max() for nm in ("real", "imag", "weights"))
        check("sharded_grid %s" % kw, err <= 1e-11, "max rel %.1e" % err)
    else:
        check("sharded_grid %s" % kw, True)
    b = quiet(pdist.banded_grid, gd, gridsize=G, binsize=binsize, **kw)
    if rank == 0:
        check("banded_grid %s bit-identical" % kw, all(np.array_equal(getattr(b, nm), getattr(ref, nm)) for nm in ("real", "imag", "weights")))
    else:
        check("banded_grid %s bit-identical" % kw, True)

dist.barrier()
dist.destroy_process_group()
sys.exit(1 if failures else 0)
