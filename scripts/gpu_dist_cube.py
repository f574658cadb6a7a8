"""torchrun check of ShardedLikelihood.stage_cube: lnlike with cube="sharded" / "rank0" == the per-rank
full upload == the single-process value, on N ranks."""
import os, sys
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
from pdspy_b200 import dist as pdist, PinnedArray
from pdspy_b200.interferometry import Visibilities
A = synth.ARCSEC
n, nf, nuv = 128, 3, 20000
u, v = synth.synth_uv(nuv, 0.02 * A)
re, im, w = synth.synth_data(nuv, nf)
d = Visibilities(u, v, synth.synth_freq(nf), re, im, w)
cube = np.ascontiguousarray(synth.synth_image(n, nf, 0.02)[:, :, :, 0])
pin = PinnedArray(cube.shape)
pin.array[...] = cube
like = pdist.ShardedLikelihood(pdist.shard_visibilities(d, rank, world))
vals = {}
for mode in (None, "sharded", "rank0"):
    src = pin.array if (mode != "rank0" or rank == 0) else np.zeros_like(cube)     # only rank 0's copy may be read
    vals[mode] = like(src, 0.02 * A, 0.01 * A, -0.02 * A, kind=0, cube=mode)
if rank == 0:
    print(world, "lnlike full %.12e sharded %.12e rank0 %.12e" % (vals[None], vals["sharded"], vals["rank0"]),
          "OK" if vals[None] == vals["sharded"] == vals["rank0"] else "<<<<< MISMATCH", flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
