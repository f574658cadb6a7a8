"""torchrun check of pdspy_b200.dist.sharded_grid: N ranks, each a shard; result == single-GPU grid()."""
import contextlib, io, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, torch.distributed as dist
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ["PDSB_DEVICE"] = str(lr)
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import synth
from pdspy_b200 import dist as pdist
from pdspy_b200.interferometry import Visibilities, grid
u, v = synth.synth_uv(2_000_000, 0.01 * synth.ARCSEC)
re, im, w = synth.synth_data(2_000_000, 2)
freq = synth.synth_freq(2)
d = Visibilities(u, v, freq, re, im, w)
G = 1024
binsize = 2.2 * np.hypot(u, v).max() / G
for kw in (dict(convolution="expsinc", mode="spectralline", imaging=True), dict(convolution="pillbox")):
    shard = pdist.shard_visibilities(d, rank, world)
    with contextlib.redirect_stdout(io.StringIO()):
        g = pdist.sharded_grid(shard, gridsize=G, binsize=binsize, **kw)
        if rank == 0:
            ref = grid(d, gridsize=G, binsize=binsize, deterministic=True, **kw)
    if rank == 0:
        for nm in ("real", "imag", "weights"):
            a, b = getattr(g, nm), getattr(ref, nm)
            print(world, kw, nm, "max rel diff %.2e" % (np.abs(a - b).max() / np.abs(b).max()), flush=True)
# bit-exact band mode: every rank sees all the data and owns a band of output rows
for kw in (dict(convolution="pillbox"), dict(convolution="expsinc", mode="spectralline", imaging=True)):
    import time
    with contextlib.redirect_stdout(io.StringIO()):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        g = pdist.banded_grid(d, gridsize=G, binsize=binsize, **kw)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        if rank == 0:
            ref = grid(d, gridsize=G, binsize=binsize, deterministic=True, **kw)
    if rank == 0:
        print(world, "banded", kw, "bit-identical:", all(np.array_equal(getattr(g, nm), getattr(ref, nm)) for nm in ("real", "imag", "weights")),
              "%.1f ms incl. host staging" % ((t1 - t0) * 1e3), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
