#!/usr/bin/env python
"""Headline benchmark: the fused model-to-visibility likelihood on synthetic data of the shapes
BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dft fp32|tcgen05] [--workload C3|C2|C4|C5]
                    [--impl reference] [--no-extras] [--no-cpu-baseline]

N > 1 is launched by the driver as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
one rank per GPU.  ONE synthetic data set is sharded over the ranks by uv points (strong scaling: the
likelihood of one data set), the image cube is replicated, chi^2[nf] + the log term are all-reduced over
NCCL, and rank 0 checks the sharded result against the same likelihood evaluated unsharded on its own GPU.

A step = one likelihood evaluation: fold the fp64 cube -> direct Fourier sampling at every uv point ->
weighted chi^2 against the data -> scalar (pdspy: interpolate_model.py:11-57 followed by
utils/emcee.py:31-43).  Rank 0 prints ONE JSON line: value / e2e / roofline of the default workload
(C3 = BASELINE.json configs[2]) and, under `extras.workloads`, the other BASELINE configurations
(C2 = configs[1], C4 = configs[3] gridding, C5 = configs[4] walker batch) with value, e2e and roofline
each.  See DESIGN.md "Measurement".
"""
import argparse
import contextlib
import ctypes
import io
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))          # tests/synth.py: the seeded synthetic inputs

import synth                      # noqa: E402

A = synth.ARCSEC
METRIC = "pixel_visibility_pairs_per_s"
UNIT = "pairs/s"
FLOP_PER_PAIR = 4.0          # algorithmic: 2 FMA per (pixel, uv, channel) pair
L2_FLUSH_BYTES = 256 << 20   # > 126 MB L2
DFT_VARIANT = {"fp32": 0, "tcgen05": 200, "nufft": 400}
DTYPE = {"fp32": "f32 products, f64 phase seeds and accumulation",
         "tcgen05": "f16 x2 lattice-split operands on tcgen05 (22 bits), f32 accumulate, f64 partial sums",
         "nufft": "f64 (type-2 non-uniform FFT: f64 transform, 8 x 8 f64 taps per visibility and channel)"}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def ncu_traffic(key):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` captures
    (profiles/traffic.json, written by scripts/ncu_traffic.py from the .ncu-rep files); None where not captured."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t.get(key)
    except Exception:
        return None


def workload_config(name, nuv_override=None):
    if name == "C4":
        return dict(name="C4", npix=2048, nf=1, nuv=nuv_override or 10_000_000, pixelsize=0.01, dRA=0.0, dDec=0.0)
    n, nf, nuv, px, dra, ddec = synth.CONFIGS["C3" if name == "C5" else name]
    if nuv_override:
        nuv = nuv_override
    return dict(name=name, npix=n, nf=nf, nuv=nuv, pixelsize=px, dRA=dra, dDec=ddec)


WORKLOAD_NAMES = {
    "C3": "spectral-line cube 512x512x64 channels, 1M uv per channel, full chi^2 log-likelihood (BASELINE.json configs[2])",
    "C2": "continuum 1024x1024 image onto 1M uv points with dRA/dDec offset + chi^2 (BASELINE.json configs[1])",
    "C1": "256x256 single channel onto 50k uv points (BASELINE.json configs[0])",
    "C4": "grid() of 10M visibilities onto 2048x2048, exp*sinc convolution kernel, natural weights (BASELINE.json configs[3])",
    "C5": "batched likelihood for emcee walkers x 512x512x64-channel cubes, walkers sharded over the GPUs "
          "(BASELINE.json configs[4])"}


def describe(cfg):
    """config of the JSON line: identical for every N (the driver compares it across its runs)."""
    return {"workload": WORKLOAD_NAMES[cfg["name"]], "npix": cfg["npix"], "channels": cfg["nf"], "nuv": cfg["nuv"],
            "pairs_per_step": float(cfg["npix"]) ** 2 * cfg["nuv"] * cfg["nf"],
            "partition": "one data set, uv points sharded over the ranks (Hermitian pairs kept together), cube "
                         "replicated, all-reduce of nf+1 doubles",
            "cache": "L2 flushed (256 MB memset) before every timed step",
            "uv": "Hermitian-doubled synthetic ALMA-like list, seed 1234 (tests/synth.py)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons, power = [], [], set(), []
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                   "samples": len(sm), "power_w_max": float(max(power))}
        return out


# ------------------------------------------------------------------------------------------
# reference arm
def run_reference(args, cfg, rank, world):
    """Reference arm: the reference's CPU path for this workload on the host cores.  galario (the
    library pdspy calls, interpolate_model.py:23-24) is not installable here, so its algorithm is
    timed from the restatement oracle/dft.py:galario_like (rfft2 + bilinear + phase, per channel as
    interpolate_model.py:22 loops), followed by the verbatim numpy likelihood (emcee.py:31-43).
    Bounded sample per step: one channel per host thread, all uv points, all pixels."""
    if rank != 0:
        return
    if cfg["name"] == "C4":
        # the reference's own compiled grid() (oracle/_ref), serial by construction
        from oracle import build_ref
        ref = build_ref.load()
        if ref is None:
            emit({"impl": "reference", "unavailable": "oracle/_ref (the reference's compiled grid) is not built"})
            return
        nvis, G = cfg["nuv"], cfg["npix"]
        ns = min(nvis, 300_000)
        u, v = synth.synth_uv(nvis, cfg["pixelsize"] * A)
        re, im, w = synth.synth_data(ns, 1)
        binsize = 2.2 * np.hypot(u, v).max() / G
        dd = ref.Visibilities(u[:ns].copy(), v[:ns].copy(), synth.synth_freq(1), re, im, w)
        for _ in range(args.warmup):
            ref.grid(dd, gridsize=G, binsize=binsize, convolution="expsinc")
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ref.grid(dd, gridsize=G, binsize=binsize, convolution="expsinc")
        dt = time.perf_counter() - t0
        value = ns * args.steps / dt
        emit({"impl": "reference", "metric": "gridded_visibilities_per_s", "value": value, "unit": "vis/s",
              "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
              "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
              "data": "synthetic", "config": gridding_config(cfg),
              "cpu_baseline": {"value": value, "unit": "vis/s", "cores": 1, "kind": "reference",
                               "sample": "first %d of %d visibilities per step" % (ns, nvis)},
              "e2e": {"value": value, "unit": "vis/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return
    from concurrent.futures import ThreadPoolExecutor
    from oracle import dft as od, likelihood as ol
    cores = os.cpu_count() or 1
    nch = min(cfg["nf"], max(8, cores))        # one channel per host thread
    u, v = synth.synth_uv(cfg["nuv"], cfg["pixelsize"] * A)
    img = synth.synth_image(cfg["npix"], cfg["nf"], cfg["pixelsize"])[:, :, :nch, :]
    re, im, w = synth.synth_data(cfg["nuv"], nch)
    dxy = cfg["pixelsize"] * A
    threads = min(cores, nch)

    def one_channel(i):
        return od.galario_like(u, v, img[:, :, i:i + 1, :], dxy, cfg["dRA"] * A, cfg["dDec"] * A)[:, 0]

    def step():
        with ThreadPoolExecutor(threads) as ex:
            cols = list(ex.map(one_channel, range(nch)))
        vis = np.stack(cols, axis=1)
        return ol.lnlike_vis_numpy(re, im, w, vis.real, vis.imag)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    pairs = float(cfg["npix"]) ** 2 * cfg["nuv"] * nch * args.steps
    value = pairs / dt
    sample = "%d of %d channels, all %d uv points, all pixels per step" % (nch, cfg["nf"], cfg["nuv"])
    cfgd = describe(cfg)
    if cfg["name"] == "C5":
        cfgd = walker_config(cfg, args.walkers)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfgd,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "what": "galario algorithm (rfft2 + bilinear interpolation, restated: galario itself is "
                                     "not installable offline) + numpy likelihood; pairs/s counts the pixel-visibility "
                                     "pairs the result represents, not operations executed (the FFT path is "
                                     "O(n^2 log n + nuv)); a stand-in for galario's C++/OpenMP library, which would "
                                     "be several times faster: ratios against this line are upper bounds"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "likelihood_evals_per_s_equiv": args.steps / dt * nch / cfg["nf"]}
    emit(line)


def cpu_baseline_port(cfg):
    """The like-for-like algorithm on the host: exact fp64 direct DFT (oracle/c/oracle.c, OpenMP, all
    cores) + likelihood on a bounded sample of the uv points (all pixels, all channels)."""
    from oracle.build import lib as olib
    from oracle import likelihood as ol
    L = olib()
    cores = L.oracle_num_threads()
    nuv_s = {"C3": 6144, "C2": 98304, "C1": 50000}[cfg["name"]]          # about 10 s of work on 16 host cores
    nuv_s = min(nuv_s, cfg["nuv"])
    u, v = synth.synth_uv(cfg["nuv"], cfg["pixelsize"] * A)
    u, v = np.ascontiguousarray(u[:nuv_s]), np.ascontiguousarray(v[:nuv_s])
    img = np.ascontiguousarray(synth.synth_image(cfg["npix"], cfg["nf"], cfg["pixelsize"])[:, :, :, 0])
    re, im, w = synth.synth_data(nuv_s, cfg["nf"])
    ore, oim = np.empty((nuv_s, cfg["nf"])), np.empty((nuv_s, cfg["nf"]))
    p = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    t0 = time.perf_counter()
    L.oracle_dft(p(u), p(v), nuv_s, p(img), cfg["npix"], cfg["npix"], cfg["nf"], cfg["pixelsize"] * A,
                 cfg["dRA"] * A, cfg["dDec"] * A, p(ore), p(oim))
    ol.lnlike_vis_numpy(re, im, w, ore, oim)
    dt = time.perf_counter() - t0
    pairs = float(cfg["npix"]) ** 2 * nuv_s * cfg["nf"]
    return {"value": pairs / dt, "unit": UNIT, "cores": cores, "kind": "port", "seconds": dt,
            "sample": "first %d of %d uv points, all pixels, all %d channels, one pass (exact fp64 direct DFT, "
                      "one sincos per pixel-visibility pair, OpenMP)" % (nuv_s, cfg["nuv"], cfg["nf"])}


# ------------------------------------------------------------------------------------------
class Env:
    """Process-wide handles of one bench run."""

    def __init__(self, rank, local_rank, world):
        import torch
        import torch.distributed as dist
        from pdspy_b200 import _lib
        self.rank, self.local_rank, self.world = rank, local_rank, world
        self.torch, self.dist, self._lib = torch, dist, _lib
        self.L = _lib.lib()
        _lib.check(self.L.pdsb_set_stream(torch.cuda.current_stream().cuda_stream))
        self.flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def launches(self):
        n = ctypes.c_int64()
        self.L.pdsb_launch_count(ctypes.byref(n))
        return int(n.value)

    def profile(self, prefix):
        ms, n = ctypes.c_double(), ctypes.c_int64()
        self._lib.check(self.L.pdsb_profile_get(prefix, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, int(n.value)


def timed_device_steps(env, step, steps, warmup, profile_prefix=None, sample_clocks=False):
    """`warmup` untimed + `steps` timed calls of step(), each timed with CUDA events on the stream the kernels
    run on, L2 flushed before every one, barrier + synchronize on both sides, max over ranks.
    Returns (total ms, kernel ms per launch of `profile_prefix` (max over ranks), launches of it on this rank,
    libpdsb launches in the timed region, clocks, last result)."""
    torch, _lib, L = env.torch, env._lib, env.L
    out = None
    for _ in range(warmup):
        env.flush.zero_()
        out = step()
    env.barrier()
    sampler = ClockSampler(env.local_rank) if sample_clocks and env.rank == 0 else None
    if sampler:
        sampler.start()
    _lib.check(L.pdsb_profile_reset())
    _lib.check(L.pdsb_profile_enable(1))
    n0 = env.launches()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    env.barrier()
    for e0, e1 in evs:
        env.flush.zero_()
        e0.record()
        out = step()
        e1.record()
    env.barrier()
    n1 = env.launches()
    _lib.check(L.pdsb_profile_enable(0))
    clocks = sampler.stop() if sampler else None
    total_ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    k_ms, k_n = env.profile(profile_prefix) if profile_prefix else (0.0, 0)
    total_ms, k_ms_max = env.max_over_ranks([total_ms, k_ms])
    return total_ms, (k_ms_max / k_n if k_n else None), k_n, n1 - n0, clocks, out


def timed_host_steps(env, step, steps, warmup=2):
    """End to end: host wall clock around `steps` synchronous calls, barrier on both sides, max over ranks."""
    out = None
    for _ in range(warmup):
        out = step()
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = step()
    env.barrier()
    dt = time.perf_counter() - t0
    return env.max_over_ranks([dt])[0], out


# ------------------------------------------------------------------------------------------
def likelihood_leg(env, cfg, dft, steps, warmup, main=False, cpu_baseline=False):
    """C2 / C3: the likelihood of ONE data set sharded over the ranks, on DFT kernel `dft`."""
    from pdspy_b200 import DeviceBuffer, PinnedArray, dist as pdist
    from pdspy_b200.device import Dataset
    from pdspy_b200.dist import ShardedLikelihood
    from pdspy_b200.interferometry import Visibilities
    _lib, L, rank, world = env._lib, env.L, env.rank, env.world
    n, nf, nuv = cfg["npix"], cfg["nf"], cfg["nuv"]
    u, v = synth.synth_uv(nuv, cfg["pixelsize"] * A)
    # the FFT-based kernel transforms the cube per channel: over several GPUs its channels are split as far as the
    # fused sampler allows (>= 16 channels per rank) and the uv points over what is left of the ranks; the
    # direct-sum kernels split the uv points only (their per-uv phase work is shared by all channels)
    cw = 1
    if dft == "nufft":
        while cw * 2 <= world and world % (cw * 2) == 0 and nf % (cw * 2) == 0 and nf // (cw * 2) >= 16:
            cw *= 2
    uw, crank, urank = world // cw, rank % cw, rank // cw
    rows, re, im, w = synth.synth_data_shard(nuv, nf, urank, uw)
    assert np.array_equal(rows, pdist.shard_rows(u, v, urank, uw))
    freq = synth.synth_freq(nf)
    c0, c1 = pdist.shard_channels(nf, crank, cw)
    if cw > 1:
        re, im, w = (np.ascontiguousarray(x[:, c0:c1]) for x in (re, im, w))
    shard = Visibilities(np.ascontiguousarray(u[rows]), np.ascontiguousarray(v[rows]), freq[c0:c1], re, im, w)
    like = ShardedLikelihood(shard, channels=(c0, nf) if cw > 1 else None)
    del re, im, w, shard
    cube = np.ascontiguousarray(synth.synth_image(n, nf, cfg["pixelsize"])[:, :, :, 0])     # [n, n, nf] fp64
    dxy, dra, ddec = cfg["pixelsize"] * A, cfg["dRA"] * A, cfg["dDec"] * A
    dcube = DeviceBuffer.from_numpy(cube)
    pinned = PinnedArray(cube.shape)
    pinned.array[...] = cube
    pairs_step = float(n) * n * nuv * nf
    shard_upload = world > 1 and n % world == 0

    def step_device():
        return like(dcube, dxy, dra, ddec, kind=_lib.DEVICE, shape=(n, n))

    # end to end: on several GPUs every rank uploads its 1/world slab of the host cube and the slabs are
    # all-gathered over NVLink (ShardedLikelihood.stage_cube); on one GPU the cube is uploaded whole
    def step_e2e():
        return like(pinned.array, dxy, dra, ddec, kind=_lib.HOST, cube="sharded" if shard_upload else None)

    _lib.check(L.pdsb_set_dft_variant(DFT_VARIANT[dft]))
    prefix = {"tcgen05": b"dft_tc5", "nufft": b"nufft_chi2"}.get(dft, b"dft_f2")
    total_ms, k_ms, k_n, launches, clocks, ll = timed_device_steps(env, step_device, steps, warmup, prefix,
                                                                   sample_clocks=main)
    e2e_s, ll_e2e = timed_host_steps(env, step_e2e, steps)
    # parity of the timed step itself: the same sharded likelihood on the all-fp64 kernel (1e-11 from the CPU oracle)
    _lib.check(L.pdsb_set_dft_variant(300))
    ll_f64 = step_device()
    _lib.check(L.pdsb_set_dft_variant(DFT_VARIANT[dft]))
    # ... and the sharded sum against the SAME data set evaluated unsharded on rank 0's GPU
    ll_single = None
    if world > 1 and rank == 0:
        _, fre, fim, fw = synth.synth_data_shard(nuv, nf, 0, 1)
        full = Dataset(u, v)
        full.set_data(fre, fim, fw)
        del fre, fim, fw
        out = ctypes.c_double()
        _lib.check(L.pdsb_loglike(full.handle, _lib.ptr(dcube), n, n, nf, _lib.DEVICE, float(dxy), float(dra), float(ddec),
                                  None, ctypes.cast(ctypes.byref(out), ctypes.c_void_p)))
        ll_single = out.value
        full.destroy()
    env.barrier()
    _lib.check(L.pdsb_set_dft_variant(0))
    res = None
    if rank == 0:
        pk = peaks()
        hermitian = like.ds.hermitian
        pairs_launch = float(n) * n * nf * like.ds.nuv                    # this rank's share, per launch
        h2d = int(cube.nbytes) if shard_upload or world == 1 else int(cube.nbytes) * world
        res = {"dft_kernel": dft,
               "partition_used": ("%d channel group(s) x %d uv shard(s)" % (cw, uw)) if cw > 1 else "%d uv shard(s)" % world,
               "ms_per_step": total_ms / steps, "value": pairs_step * steps / (total_ms * 1e-3),
               "unit": UNIT, "likelihood_evals_per_s": steps / (total_ms * 1e-3), "steps": steps,
               "lnlike": ll, "lnlike_rel_diff_vs_fp64_kernel": abs(ll - ll_f64) / abs(ll_f64),
               "lnlike_rel_diff_vs_single_rank": (abs(ll - ll_single) / abs(ll_single)) if ll_single is not None else None,
               "lnlike_rel_diff_e2e_vs_device": abs(ll_e2e - ll) / abs(ll),
               "e2e": {"value": pairs_step * steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": int((nf + 1) * 8) * world,
                       "h2d_bytes_per_step_per_rank": int(cube.nbytes) // world if shard_upload else int(cube.nbytes),
                       "ms_per_step": e2e_s / steps * 1e3, "likelihood_evals_per_s": steps / e2e_s,
                       "api": "pdspy_b200.dist.ShardedLikelihood.__call__ (host fp64 cube in, host scalar out" +
                              ("; each rank uploads 1/N of the cube, NCCL all-gather over NVLink)" if shard_upload else ")"),
                       "timer": "host wall clock around K synchronous calls, max over ranks"},
               "gpu_launches": launches, "clocks": clocks}
        if k_ms:
            if dft == "fp32":
                sm, khz = ctypes.c_int(), ctypes.c_int()
                L.pdsb_device_info(ctypes.byref(sm), ctypes.byref(khz), None, None, None)
                sm_max_mhz = pk.get("sm_max_mhz") or khz.value / 1e3
                fp32_peak = sm.value * 128 * 2 * sm_max_mhz * 1e6 / 1e12          # TFLOP/s, non-tensor FMA pipe
                achieved = pairs_launch * FLOP_PER_PAIR / (k_ms * 1e-3) / 1e12
                executed = pairs_launch * (0.5 if hermitian else 1.0) * 2.0 / (k_ms * 1e-3) / 1e12
                tf, ms = ctypes.c_double(), ctypes.c_double()
                _lib.check(L.pdsb_bench_fma(1, 20000, ctypes.byref(tf), ctypes.byref(ms)))
                res["roofline"] = {
                    "kernel": "dft_kernel (direct Fourier sampling, FP32 FMA pipe; no tensor cores)",
                    "bound": "fp32_fma", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                    "frac": achieved / fp32_peak,
                    "peak_source": "nominal non-tensor FP32: SMs x 128 lanes x 2 x sm_max_mhz (MEASURED_PEAKS.json has "
                                   "no FP32 entry; its sm_max_mhz is used)",
                    "peak_measured_fma_microbench": tf.value, "algorithmic_flop_per_pair": FLOP_PER_PAIR,
                    "executed": executed, "executed_frac": executed / fp32_peak,
                    "executed_note": "FMA-pipe flops the inner loop actually issues per launch: the real-image mirror "
                                     "fold needs 1 FMA per pixel per uv point instead of 2, and a Hermitian-doubled uv "
                                     "list is evaluated for one half; frac > 1 is those two algorithmic savings, "
                                     "executed_frac is the pipe utilisation",
                    "launch_ms": k_ms, "launches": k_n}
            elif dft == "nufft":
                hbm = pk.get("hbm_gbs", 6650.0)
                t_ms, t_n = env.profile(b"rfft2_planes")
                path_ms = k_ms + (t_ms / t_n if t_n else 0.0)
                alg = float(cube.nbytes) * like.nf_local / nf + 24.0 * like.ds.nuv * like.nf_local
                res["roofline"] = {
                    "kernel": "NUFFT path: rfft2_planes_padded (deapodised, zero-padded half-spectrum FFT of every channel) + "
                              "nufft_chi2_tiled_kernel<2> (8 x 8 taps per unique uv point and channel from shared memory, "
                              "chi^2 per channel summed in the same kernel)",
                    "bound": "hbm", "achieved": alg / (path_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                    "frac": alg / (path_ms * 1e-3) / 1e9 / hbm, "peak_source": "MEASURED_PEAKS.json hbm_gbs",
                    "algorithmic_bytes_note": "this rank's channels of the fp64 cube read once + real, imag, weights of "
                                              "its shard read once",
                    "launch_ms": path_ms, "sampler_ms": k_ms, "transform_ms": (t_ms / t_n if t_n else None), "launches": k_n}
            else:
                # 3 fp16 MACs (hi.lo, lo.hi, hi.hi) per pixel-visibility pair actually evaluated
                macs = 3.0 * float(n) * n * nf * like.ds.nuv_unique
                ach = 2.0 * macs / (k_ms * 1e-3) / 1e12
                peak = pk.get("bf16_tflops_sustained") or pk.get("bf16_tflops")
                res["roofline"] = {
                    "kernel": "dft_tc5_kernel (tcgen05.mma kind::f16, accumulators and A operand in TMEM, B through a "
                              "bulk-TMA ring)",
                    "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak if peak else None,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS bf16 GEMM back to back: the kernel is "
                                   "timed inside a long step)",
                    "counts": "2 flop x 3 split-product MACs per evaluated pixel-visibility pair (mirror-folded image, "
                              "Hermitian half of the uv list), per rank",
                    "algorithmic_flops_over_peak": (pairs_launch * FLOP_PER_PAIR / (k_ms * 1e-3) / 1e12 / peak) if peak else None,
                    "launch_ms": k_ms, "launches": k_n}
            res["roofline"]["traffic"] = ncu_traffic("%s:%s:%d" % (cfg["name"], dft, world))
            res["roofline"]["traffic_source"] = ("dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` "
                                                 "capture of this launch, read from profiles/traffic.json; null where "
                                                 "that (workload, kernel, N) was not captured")
            res["roofline"]["algorithmic_bytes"] = (float(cube.nbytes) * like.nf_local / nf + 24.0 * like.ds.nuv * like.nf_local
                                                    if dft == "nufft" else
                                                    float(n) * n * nf * 4 + like.ds.nuv_unique * 16.0 * (1 + nf))
        if cpu_baseline:
            res["cpu_baseline"] = cpu_baseline_port(cfg)
    if main:
        return res, (like, dcube, pinned, cube, (dxy, dra, ddec))
    like.ds.destroy()
    dcube.free()
    pinned.free()
    return res, None


def chain_leg(env, cfg, dft, steps):
    """The reference's own two calls, unchanged signatures, on this rank's shard:
        m = interpolate_model(u, v, freq, model, dRA=, dDec=)          (interpolate_model.py:11-57)
        visibility_lnlike(data, m)                                      (the emcee.py:31-43 term of utils.emcee.lnlike)
    with a pageable host cube, as pdspy holds it.  The model visibilities stay on the device between the two calls
    (lazy Visibilities); timed under both handle-cache policies (pdspy_b200/device.py)."""
    import pdspy_b200 as pb
    from pdspy_b200 import dist as pdist, utils
    from pdspy_b200.interferometry import interpolate_model, Visibilities
    _lib, L, rank, world = env._lib, env.L, env.rank, env.world
    n, nf, nuv = cfg["npix"], cfg["nf"], cfg["nuv"]
    u, v = synth.synth_uv(nuv, cfg["pixelsize"] * A)
    rows, re, im, w = synth.synth_data_shard(nuv, nf, rank, world)
    freq = synth.synth_freq(nf)
    data = Visibilities(np.ascontiguousarray(u[rows]), np.ascontiguousarray(v[rows]), freq, re, im, w)
    model = synth.SynthImage(synth.synth_image(n, nf, cfg["pixelsize"]), cfg["pixelsize"], freq)
    _lib.check(L.pdsb_set_dft_variant(DFT_VARIANT[dft]))

    def step():
        m = interpolate_model(data.u, data.v, freq, model, dRA=cfg["dRA"], dDec=cfg["dDec"])
        return utils.visibility_lnlike(data, m)
    out = {}
    for policy in ("hash", "freeze"):
        pb.clear_cache()
        pb.set_cache_policy(policy)
        s, ll = timed_host_steps(env, step, steps, warmup=2)
        out[policy] = (s, ll)
    pb.set_cache_policy("hash")
    pb.clear_cache()
    _lib.check(L.pdsb_set_dft_variant(0))
    if rank != 0:
        return None
    pairs_step = float(n) * n * nuv * nf
    return {"what": "interpolate_model(u, v, freq, model, dRA=, dDec=) -> utils.visibility_lnlike(data, m): the reference's "
                    "call signatures (what utils.emcee.lnlike runs), pageable host cube in, scalar out, each rank on its uv "
                    "shard (no all-reduce); the [nuv, nf] model visibilities never leave the device",
            "dft_kernel": dft, "lnlike_shard": out["hash"][1],
            "cache_policy_hash": {"ms_per_step": out["hash"][0] / steps * 1e3, "value": pairs_step * steps / out["hash"][0],
                                  "unit": UNIT, "note": "default: u, v, real, imag, weights re-hashed on the host every call"},
            "cache_policy_freeze": {"ms_per_step": out["freeze"][0] / steps * 1e3,
                                    "value": pairs_step * steps / out["freeze"][0], "unit": UNIT,
                                    "note": "cached arrays made read-only instead of re-hashed"},
            "h2d_bytes_per_step": int(model.image.nbytes) * world, "d2h_bytes_per_step": 32 * world}


def galario_fft_leg(env, cfg, steps, handles, nufft=False):
    """extra: the reference's own algorithm (galario: FFT + bilinear interpolation) on the GPU, this rank's shard;
    nufft=True: the exact transform through the 8-point non-uniform FFT instead (same FFT path, 8 x 8 taps)."""
    like, dcube, pinned, cube, (dxy, dra, ddec) = handles
    _lib, L = env._lib, env.L
    n, nf = cfg["npix"], cfg["nf"]
    fft_out = np.empty(4)
    entry = L.pdsb_loglike_nufft if nufft else L.pdsb_loglike_fft

    def step(image, kind):
        size = (n, n) if nufft else (n,)
        _lib.check(entry(like.ds.handle, _lib.ptr(image), *size, nf, kind, float(dxy), float(dra), float(ddec),
                         _lib.ptr(fft_out)))
        return float(fft_out[3])
    ms, _, _, _, _, _ = timed_device_steps(env, lambda: step(dcube, _lib.DEVICE), steps, 2)
    e2e_s, _ = timed_host_steps(env, lambda: step(pinned.array, _lib.HOST), steps)
    if env.rank != 0:
        return None
    pairs_step = float(n) * n * cfg["nuv"] * nf
    hbm = peaks().get("hbm_gbs", 6650.0)
    alg_bytes = float(cube.nbytes) + 24.0 * like.ds.nuv * nf          # per rank: whole cube + its uv shard's data
    what = ("the reference's OWN algorithm for this step on the GPU - galario's FFT + bilinear interpolation "
            "(pdsb_loglike_fft, fp64, restated from galario's published algorithm) + the same chi^2; it "
            "carries galario's interpolation error (1e-3..4e-2 of max|V|), which the direct transform of "
            "value / e2e does not; pairs/s counts the pairs the result represents, as in --impl reference; "
            "NOT used for value / e2e; chi^2 of this rank's uv shard, no all-reduce")
    if nufft:
        what = ("the EXACT transform of value / e2e through a type-2 non-uniform FFT (pdsb_loglike_nufft, fp64: image / "
                "kernel transform, zero-padded to 2n, FFT per channel, 8 x 8 exponential-of-semicircle taps per visibility "
                "and channel) + the same chi^2: 4e-8 of max|V| from the exact oracle (tests/test_gpu_nufft.py), i.e. inside "
                "the tolerances of the direct sum at the cost of an FFT path; pairs/s counts the pairs the result "
                "represents; NOT used for value / e2e (the north star prescribes the direct sum); chi^2 of this rank's "
                "uv shard, no all-reduce")
    return {"what": what,
            "ms_per_step": ms / steps, "value": pairs_step * steps / (ms * 1e-3), "unit": UNIT,
            "e2e": {"value": pairs_step * steps / e2e_s, "unit": UNIT, "ms_per_step": e2e_s / steps * 1e3,
                    "h2d_bytes_per_step": int(cube.nbytes) * env.world, "d2h_bytes_per_step": 32 * env.world},
            "roofline": {"kernel": ("rfft2_planes_padded (half-spectrum FFT at 2n) + nufft_chi2_kernel (8 x 8 taps fused with the "
                                    "chi^2 sums; one evaluation per Hermitian pair)") if nufft else
                                   ("rfft2_planes (half-spectrum FFT of every channel) + fft_chi2_kernel (bilinear gather fused "
                                    "with the chi^2 sums)"), "bound": "hbm",
                         "algorithmic_bytes": alg_bytes, "achieved": alg_bytes / (ms / steps * 1e-3) / 1e9,
                         "peak": hbm, "unit": "GB/s", "frac": alg_bytes / (ms / steps * 1e-3) / 1e9 / hbm,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs",
                         "algorithmic_bytes_note": "fp64 cube read once + real, imag, weights of this rank's shard read once "
                                                   "(24 B per visibility and channel); the transformed half spectrum is "
                                                   "scratch",
                         "traffic": ncu_traffic("%s:%s:%d" % (cfg["name"], "nufft" if nufft else "galario_fft", env.world)),
                         "traffic_source": "sum of dram__bytes_read + dram__bytes_write over the path's launches in one "
                                           "`ncu --set full` capture (profiles/traffic.json); null where not captured"},
            "lnlike_shard": float(fft_out[3])}


# ------------------------------------------------------------------------------------------
def gridding_config(cfg):
    return {"workload": WORKLOAD_NAMES["C4"] + ", fast (sorted-tile) mode", "nvis": cfg["nuv"], "gridsize": cfg["npix"],
            "partition": "visibilities sharded over the ranks, private raw-sum maps, all-reduce of 3 x G^2 doubles, "
                         "normalise (pdspy_b200.dist.sharded_grid)",
            "cache": "L2 flushed (256 MB memset) before every timed step"}


def gridding_leg(env, cfg, steps, warmup, cpu_baseline=False, sample_clocks=False):
    """BASELINE.json configs[3]: grid() of 10M visibilities onto 2048^2 with the exp*sinc convolution
    kernel and weights (fast mode).  Metric: visibilities/s; roofline: HBM on the algorithmic bytes
    (40 B per visibility + 24 B per cell).  world > 1: the visibilities are sharded,
    every rank grids its shard into private raw-sum maps, NCCL all-reduce, normalise."""
    from pdspy_b200 import DeviceBuffer, dist as pdist
    from pdspy_b200.interferometry import grid as grid_py, Visibilities
    from oracle import grid as og
    torch, dist, _lib, L, rank, world = env.torch, env.dist, env._lib, env.L, env.rank, env.world
    nvis, G = cfg["nuv"], cfg["npix"]
    u, v = synth.synth_uv(nvis, cfg["pixelsize"] * A)
    binsize = 2.2 * np.hypot(u, v).max() / G
    rows, re, im, w = synth.synth_data_shard(nvis, 1, rank, world, seed=777)
    us, vs = np.ascontiguousarray(u[rows]), np.ascontiguousarray(v[rows])
    freq = synth.synth_freq(1)
    uu, vv = og.cell_centres(G, binsize)               # numpy.linspace of libinterferometry.pyx:370-381
    d = {k: DeviceBuffer.from_numpy(a) for k, a in dict(u=us, v=vs, freq=freq, re=re, im=im, w=w, uu=uu, vv=vv).items()}
    maps = torch.zeros((3, G * G, 1), dtype=torch.float64, device="cuda")
    nloc = us.size

    def step():
        _lib.check(L.pdsb_grid(_lib.ptr(d["u"]), _lib.ptr(d["v"]), _lib.ptr(d["freq"]), _lib.ptr(d["re"]), _lib.ptr(d["im"]),
                               _lib.ptr(d["w"]), nloc, 1, _lib.DEVICE, G, float(binsize), _lib.ptr(d["uu"]), _lib.ptr(d["vv"]),
                               _lib.CONV["expsinc"], _lib.WEIGHTING["natural"], 2.0, 0, 0, 2 if world > 1 else 0, 0,
                               maps[0].data_ptr(), maps[1].data_ptr(), maps[2].data_ptr(), None, None, None, _lib.DEVICE, None))
        if world > 1:
            dist.all_reduce(maps, op=dist.ReduceOp.SUM)
            _lib.check(L.pdsb_grid_normalise(maps[0].data_ptr(), maps[1].data_ptr(), maps[2].data_ptr(), G, 1, 0))
        return None

    total_ms, k_ms, k_n, launches, clocks, _ = timed_device_steps(env, step, steps, warmup, b"grid_tile",
                                                                  sample_clocks=sample_clocks)
    wsum = float(maps[2].sum().item())
    phases = {}
    for name in (b"grid_tile_hist", b"grid_tile_scan", b"grid_tile_records", b"grid_tile_accum", b"grid_normalise"):
        t_ms, t_n = env.profile(name)
        if t_n:
            phases[name.decode()] = t_ms / t_n
    FP64_OPS_PER_VIS = 12 * 33 + 126
    fp64_peak = 148 * 64 * peaks().get("sm_max_mhz", 1965.0) * 1e6 / 1e12
    # end to end through the Python mirror: host numpy arrays in, gridded Visibilities out
    shard = Visibilities(us, vs, freq, re, im, w)

    def step_e2e():
        with contextlib.redirect_stdout(io.StringIO()):
            if world > 1:
                return pdist.sharded_grid(shard, gridsize=G, binsize=binsize, convolution="expsinc")
            return grid_py(shard, gridsize=G, binsize=binsize, convolution="expsinc", deterministic=False)
    ne2e = max(1, steps // 2)
    e2e_s, g = timed_host_steps(env, step_e2e, ne2e, warmup=1)
    e2e_s /= ne2e
    res = None
    if rank == 0:
        hbm = peaks().get("hbm_gbs", 6650.0)
        alg_bytes = nvis * 40.0 + G * G * 24.0
        step_ms = total_ms / steps
        res = {"metric": "gridded_visibilities_per_s", "value": nvis / (step_ms * 1e-3), "unit": "vis/s",
               "ms_per_step": step_ms, "steps": steps, "dtype": "f64",
               "config": gridding_config(cfg),
               "e2e": {"value": nvis / e2e_s, "unit": "vis/s", "ms_per_step": e2e_s * 1e3,
                       "h2d_bytes_per_step": int(nvis * 40 + 2 * G * 8), "d2h_bytes_per_step": int(3 * G * G * 8) * world,
                       "api": "pdspy_b200.dist.sharded_grid(shard, ...)" if world > 1 else
                              "pdspy_b200.interferometry.grid(data, ..., deterministic=False) (numpy in, Visibilities out)"},
               "gpu_launches": launches, "clocks": clocks, "weight_sum": wsum,
               "weight_sum_e2e_rel_diff": abs(float(g.weights.sum()) - wsum) / abs(wsum),
               "roofline": {"kernel": "whole grid() step (prep + tile sort + tile kernel + normalise" +
                                      (" + NCCL all-reduce of the maps)" if world > 1 else ")"),
                            "bound": "hbm", "achieved": alg_bytes / (step_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                            "frac": alg_bytes / (step_ms * 1e-3) / 1e9 / hbm,
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs",
                            "algorithmic_bytes": alg_bytes,
                            "algorithmic_bytes_note": "40 B per visibility + 24 B per cell (SURVEY.md 8d), whole job",
                            "phase_ms": phases,
                            "traffic": ncu_traffic("C4:grid:%d" % world),
                            "traffic_source": "sum of dram__bytes_read + dram__bytes_write over the step's launches in one "
                                              "`ncu --set full` capture (profiles/traffic.json); null where not captured",
                            "dominant_kernel": {
                                "kernel": "gf_tile_kernel<6,0> (records -> register accumulation -> one fp64 atomic per region cell)",
                                "bound": "fp64 pipe", "launch_ms": phases.get("grid_tile_accum"),
                                "useful_fp64_ops_per_vis": FP64_OPS_PER_VIS,
                                "achieved": (nloc * FP64_OPS_PER_VIS / (phases["grid_tile_accum"] * 1e-3) / 1e12
                                             if phases.get("grid_tile_accum") else None),
                                "peak": fp64_peak, "unit": "T fp64 instructions/s (FMA = 1)",
                                "frac": (nloc * FP64_OPS_PER_VIS / (phases["grid_tile_accum"] * 1e-3) / 1e12 / fp64_peak
                                         if phases.get("grid_tile_accum") else None),
                                "peak_source": "nominal: SMs x 64 fp64 lanes x sm_max_mhz (a dense DFMA stream measured 0.91 of it)",
                                "note": "12 exp*sinc factors per visibility in the reference's operation order (33 fp64 "
                                        "instructions each) + the 6 x 6 x 3 outer-product accumulation (126): the "
                                        "arithmetic alone is 0.28 ms of fp64 pipe per 10M visibilities, so HBM on the "
                                        "algorithmic bytes (SURVEY.md 8d) is not the limiter of this step; ncu: "
                                        "sm__pipe_fp64_cycles_active 54 %, DRAM 8 % (profiles/r02_grid_ncu.md)"}}}
        if cpu_baseline:
            from oracle import build_ref
            ref = build_ref.load()
            ns = 500_000
            if ref is not None:
                cu, cv = np.ascontiguousarray(u[:ns]), np.ascontiguousarray(v[:ns])
                cre, cim, cw = synth.synth_data(ns, 1)
                dd = ref.Visibilities(cu, cv, freq, cre, cim, cw)
                t0 = time.perf_counter()
                with contextlib.redirect_stdout(io.StringIO()):
                    ref.grid(dd, gridsize=G, binsize=binsize, convolution="expsinc")
                dt = time.perf_counter() - t0
                res["cpu_baseline"] = {"value": ns / dt, "unit": "vis/s", "cores": 1, "kind": "reference", "seconds": dt,
                                       "sample": "first %d of %d visibilities, the reference's own compiled grid() "
                                                 "(oracle/_ref), serial by construction" % (ns, nvis)}
    for b in d.values():
        b.free()
    del maps
    return res


# ------------------------------------------------------------------------------------------
def walker_config(cfg, walkers):
    cfgd = describe(cfg)
    cfgd.update({"workload": WORKLOAD_NAMES["C5"], "walkers": walkers,
                 "partition": "walkers split over the ranks, data set replicated, no data-path collective",
                 "cache": "each cube (134 MB fp64) exceeds the L2"})
    return cfgd


def walker_leg(env, cfg, walkers, dft, steps):
    """BASELINE.json configs[4]: W walkers' cubes against ONE data set; walkers are split over the ranks,
    every rank holds the whole data set, no collective on the data path (one gather of W doubles)."""
    from pdspy_b200 import DeviceBuffer, PinnedArray, dist as pdist
    from pdspy_b200.device import Dataset
    torch, _lib, L, rank, world = env.torch, env._lib, env.L, env.rank, env.world
    n, nf = cfg["npix"], cfg["nf"]
    u, v = synth.synth_uv(cfg["nuv"], cfg["pixelsize"] * A)
    _, re, im, w = synth.synth_data_shard(cfg["nuv"], nf, 0, 1)
    ds = Dataset(u, v)
    ds.set_data(re, im, w)
    del re, im, w
    ws, we = pdist.shard_walkers(walkers, rank, world)
    nw = we - ws
    base = np.ascontiguousarray(synth.synth_image(n, nf, cfg["pixelsize"])[:, :, :, 0])
    pinned = PinnedArray((max(nw, 1), n, n, nf))
    for k in range(nw):
        pinned.array[k] = base * (1.0 + 0.01 * (ws + k))
    dcubes = DeviceBuffer.from_numpy(pinned.array[:max(nw, 1)])
    dxy = cfg["pixelsize"] * A
    dra = np.ascontiguousarray(np.full(max(nw, 1), cfg["dRA"] * A))
    ddec = np.ascontiguousarray(np.full(max(nw, 1), cfg["dDec"] * A))
    out = np.zeros(max(nw, 1))

    def step(cubes, kind, count=None):
        cnt = nw if count is None else min(count, nw)
        if cnt > 0:
            _lib.check(L.pdsb_loglike_batch(ds.handle, _lib.ptr(cubes), cnt, n, n, nf, kind, float(dxy), _lib.ptr(dra),
                                            _lib.ptr(ddec), _lib.ptr(out)))
        return float(out[0])

    _lib.check(L.pdsb_set_dft_variant(DFT_VARIANT[dft]))
    step(dcubes, _lib.DEVICE, 1)                    # warm-up: one walker
    env.barrier()
    n0 = env.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        lnlike0 = step(dcubes, _lib.DEVICE)
    e1.record()
    env.barrier()
    n1 = env.launches()
    dev_ms = e0.elapsed_time(e1)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(pinned.array, _lib.HOST)
    env.barrier()
    e2e_s = time.perf_counter() - t0
    # the first walker of rank 0 once more on the all-fp64 kernel
    _lib.check(L.pdsb_set_dft_variant(300))
    ll_f64 = step(dcubes, _lib.DEVICE, 1)
    _lib.check(L.pdsb_set_dft_variant(0))
    dev_ms, e2e_s = env.max_over_ranks([dev_ms, e2e_s])
    res = None
    if rank == 0:
        pairs = float(n) * n * cfg["nuv"] * nf * walkers * steps
        res = {"metric": METRIC, "dft_kernel": dft, "value": pairs / (dev_ms * 1e-3), "unit": UNIT, "steps": steps,
               "ms_per_step": dev_ms / steps, "dtype": DTYPE[dft], "config": walker_config(cfg, walkers),
               "walkers_per_rank": nw, "likelihood_evals_per_s": walkers * steps / (dev_ms * 1e-3),
               "e2e": {"value": pairs / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(nw * base.nbytes) * world,
                       "d2h_bytes_per_step": int(nw * 8) * world, "ms_per_step": e2e_s / steps * 1e3,
                       "likelihood_evals_per_s": walkers * steps / e2e_s,
                       "api": "pdsb_loglike_batch (host fp64 cubes in pinned memory, host lnlike[W] out; cube k+1 uploads on a copy stream while cube k is evaluated)"},
               "gpu_launches": n1 - n0, "lnlike0": lnlike0,
               "lnlike0_rel_diff_vs_fp64_kernel": abs(lnlike0 - ll_f64) / abs(ll_f64)}
    ds.destroy()
    dcubes.free()
    pinned.free()
    return res


# ------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Keep stdout to the one JSON line: NCCL prints its version banner (and libraries their notices) on
    # fd 1, so fd 1 is pointed at stderr for the whole run and the result goes to a saved copy.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--dft", default="fp32", choices=["fp32", "tcgen05", "nufft"],
                    help="DFT kernel of value / e2e / roofline: the FP32-pipe kernel BASELINE.json's north star "
                         "prescribes (default) or the tcgen05 tensor-core kernel")
    ap.add_argument("--walkers", type=int, default=128, help="C5: total emcee walkers (sharded over ranks)")
    ap.add_argument("--nuv", type=int, default=0, help="override the uv count (testing)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the headline workload (no extras.workloads)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = workload_config(args.workload, args.nuv or None)

    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return

    import torch
    import torch.distributed as dist
    os.environ["PDSB_DEVICE"] = str(local_rank)
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    env = Env(rank, local_rank, world)
    cpu_base = world == 1 and not args.no_cpu_baseline
    common = {"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
              "vs_baseline": None, "data": "synthetic"}
    line = None

    if args.workload == "C4":
        res = gridding_leg(env, cfg, args.steps, args.warmup, cpu_baseline=cpu_base, sample_clocks=True)
        if rank == 0:
            line = dict(res, **common, scaling="strong")
    elif args.workload == "C5":
        res = walker_leg(env, cfg, args.walkers, args.dft, max(1, min(args.steps, 2)))
        if rank == 0:
            line = dict(res, **common, scaling="strong")
            line["steps"], line["warmup"] = res["steps"], 1
    else:
        res, keep = likelihood_leg(env, cfg, args.dft, args.steps, args.warmup, main=True, cpu_baseline=cpu_base)
        if rank == 0:
            line = {"metric": METRIC, "value": res.pop("value"), "unit": res.pop("unit"), **common,
                    "ms_per_step": res.pop("ms_per_step"), "scaling": "strong", "dtype": DTYPE[args.dft],
                    "config": describe(cfg)}
            res.pop("steps")
            line.update(res)
        extras = {"note": "context for the headline: NOT part of value / e2e",
                  "reference_arm_caveat": "`--impl reference` times a numpy restatement of galario's algorithm (galario is "
                                          "not installable offline); galario's C++/OpenMP library would be several times "
                                          "faster, so ratios against that arm are upper bounds"}
        ksteps = max(3, min(args.steps, 5))
        if not args.no_extras:
            ff = galario_fft_leg(env, cfg, ksteps, keep)
            if rank == 0:
                ff["lnlike_gap_note"] = ("the direct transform (value / e2e) and galario's FFT + bilinear interpolation differ "
                                         "by galario's interpolation error: at N=1 compare lnlike_shard with the headline lnlike")
                extras["galario_fft_algorithm"] = ff
            nu = galario_fft_leg(env, cfg, ksteps, keep, nufft=True)
            if rank == 0:
                if world == 1 and line.get("lnlike"):
                    nu["lnlike_rel_diff_vs_headline"] = abs(nu["lnlike_shard"] - line["lnlike"]) / abs(line["lnlike"])
                extras["nufft_exact_transform"] = nu
        keep[0].ds.destroy()
        keep[1].free()
        keep[2].free()
        keep = None
        if not args.no_extras:
            ch = chain_leg(env, cfg, args.dft, ksteps)
            if rank == 0:
                ch["e2e_ms_per_step_ShardedLikelihood"] = line["e2e"]["ms_per_step"]
                extras["dropin_chain"] = ch
            others = [k for k in ("fp32", "tcgen05", "nufft") if k != args.dft]
            for other in others:
                r2, _ = likelihood_leg(env, cfg, other, ksteps, 3)
                if rank == 0:
                    r2["what"] = ("the same step on the %s kernel (`bench.py --dft %s` makes it the headline); NOT used for "
                                  "value / e2e" % (other, other))
                    r2["speedup_vs_headline_kernel"] = line["ms_per_step"] / r2["ms_per_step"]
                    extras[{"tcgen05": "tensor_core_variant", "fp32": "fp32_pipe_variant", "nufft": "nufft_variant"}[other]] = r2
            workloads = {}
            if args.workload == "C3":
                c2 = workload_config("C2")
                for k in [args.dft] + others:
                    r, _ = likelihood_leg(env, c2, k, ksteps, 3)
                    if rank == 0:
                        r["config"] = describe(c2)
                        workloads["C2" if k == args.dft else "C2_" + k] = r
            g = gridding_leg(env, workload_config("C4"), ksteps, 3, cpu_baseline=cpu_base)
            if rank == 0:
                workloads["C4"] = g
            c5 = workload_config("C5")
            for k in [args.dft] + others:
                r = walker_leg(env, c5, 16 * world, k, 1)
                if rank == 0:
                    r["note"] = ("16 walkers per GPU x %d GPU(s): configs[4]'s per-GPU share (128 walkers on 8 GPUs); "
                                 "`--workload C5` runs all 128 at any N" % world)
                    workloads["C5" if k == args.dft else "C5_" + k] = r
            if rank == 0:
                extras["workloads"] = workloads
                line["extras"] = extras
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
