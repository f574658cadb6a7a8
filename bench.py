#!/usr/bin/env python
"""Headline benchmark: the fused model-to-visibility likelihood on synthetic data of the shapes
BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C3|C2] [--impl reference]

N > 1 is launched by the driver as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
one rank per GPU; the uv list is sharded over ranks (strong scaling: the likelihood of ONE
dataset), the image cube is replicated, and chi^2[nf] + the log term are all-reduced over NCCL.

A step = one likelihood evaluation: fold the fp64 cube -> direct Fourier sampling at every uv
point -> weighted chi^2 against the data -> scalar (pdspy: interpolate_model.py:11-57 followed by
utils/emcee.py:31-43).  Rank 0 prints ONE JSON line.  See DESIGN.md "Measurement".
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
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
# DRAM bytes per DFT launch measured by ncu (bytes), keyed by (workload, n_gpus)
NCU_TRAFFIC_BYTES = {("C3", 1): 571.6e6, ("C2", 1): 12.27e6}


def workload_config(name, nuv_override=None):
    if name == "C4":
        return dict(name="C4", npix=2048, nf=1, nuv=nuv_override or 10_000_000, pixelsize=0.01, dRA=0.0, dDec=0.0)
    n, nf, nuv, px, dra, ddec = synth.CONFIGS["C3" if name == "C5" else name]
    if nuv_override:
        nuv = nuv_override
    return dict(name=name, npix=n, nf=nf, nuv=nuv, pixelsize=px, dRA=dra, dDec=ddec)


def describe(cfg, world):
    names = {"C3": "spectral-line cube 512x512x64 channels, 1M uv per channel, full chi^2 log-likelihood "
                   "(BASELINE.json configs[2])",
             "C2": "continuum 1024x1024 image onto 1M uv points with dRA/dDec offset + chi^2 "
                   "(BASELINE.json configs[1])",
             "C1": "256x256 single channel onto 50k uv points (BASELINE.json configs[0])",
             "C5": "batched likelihood for emcee walkers x 512x512x64-channel cubes, walkers sharded over the GPUs "
                   "(BASELINE.json configs[4])"}
    return {"workload": names[cfg["name"]], "npix": cfg["npix"], "channels": cfg["nf"], "nuv": cfg["nuv"],
            "pairs_per_step": float(cfg["npix"]) ** 2 * cfg["nuv"] * cfg["nf"],
            "partition": "uv points sharded over %d rank(s), cube replicated, all-reduce of nf+1 doubles" % world,
            "cache": "L2 flushed (256 MB memset) before every timed step",
            "uv": "Hermitian-doubled synthetic ALMA-like list, seed 1234 (pdspy_b200/synth.py)"}


def make_local_data(cfg, rank, world):
    """This rank's uv shard + synthetic data for it (noise only; the value is irrelevant to timing)."""
    from pdspy_b200 import dist
    u, v = synth.synth_uv(cfg["nuv"], cfg["pixelsize"] * A)
    rows = dist.shard_rows(u, v, rank, world)
    us, vs = np.ascontiguousarray(u[rows]), np.ascontiguousarray(v[rows])
    re, im, w = synth.synth_data(us.size, cfg["nf"], seed=4321 + 1000 * rank + world)
    freq = synth.synth_freq(cfg["nf"])
    from pdspy_b200.interferometry import Visibilities
    return Visibilities(us, vs, freq, re, im, w), (u, v)


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
            # "under load": samples at or above the median (the idle edges are below it)
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                   "samples": len(sm), "power_w_max": float(max(power))}
        return out


# ------------------------------------------------------------------------------------------
def run_reference(args, cfg, rank, world):
    """Reference arm: the reference's CPU path for this workload on the host cores.  galario (the
    library pdspy calls, interpolate_model.py:23-24) is not installable here, so its algorithm is
    timed from the restatement oracle/dft.py:galario_like (rfft2 + bilinear + phase, per channel as
    interpolate_model.py:22 loops), followed by the verbatim numpy likelihood (emcee.py:31-43).
    Bounded sample per step: NCH_SAMPLE of the channels, all uv points, all pixels."""
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
              "higher_is_better": True, "scaling": "replicas only (not sharded)", "vs_baseline": None, "dtype": "f64",
              "data": "synthetic",
              "config": {"workload": "grid() of 10M visibilities onto 2048x2048, exp*sinc convolution kernel, natural "
                                     "weights (BASELINE.json configs[3])", "nvis": nvis, "gridsize": G},
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
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": describe(cfg, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "what": "galario algorithm (rfft2 + bilinear interpolation, restated: galario itself is "
                                     "not installable offline) + numpy likelihood; pairs/s counts the pixel-visibility "
                                     "pairs the result represents, not operations executed (the FFT path is "
                                     "O(n^2 log n + nuv))"},
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


def run_gridding(args, cfg, rank, world):
    """BASELINE.json configs[3]: grid() of 10M visibilities onto 2048^2 with the exp*sinc convolution
    kernel and weights (fast mode).  Metric: visibilities/s; roofline: HBM on the algorithmic bytes
    (40 B per visibility + 24 B per cell)."""
    if rank != 0:
        return                                        # the path is not sharded: one GPU
    from pdspy_b200 import _lib, DeviceBuffer
    from pdspy_b200.interferometry import grid as grid_py, Visibilities
    from oracle import grid as og
    L = _lib.lib()
    nvis, G = cfg["nuv"], cfg["npix"]
    u, v = synth.synth_uv(nvis, cfg["pixelsize"] * A)
    re, im, w = synth.synth_data(nvis, 1)
    freq = synth.synth_freq(1)
    binsize = 2.2 * np.hypot(u, v).max() / G
    uu, vv = og.cell_centres(G, binsize)               # numpy.linspace of libinterferometry.pyx:370-381
    d = {k: DeviceBuffer.from_numpy(a) for k, a in dict(u=u, v=v, freq=freq, re=re, im=im, w=w, uu=uu, vv=vv).items()}
    o_re, o_im, o_w = (DeviceBuffer(G * G * 8) for _ in range(3))
    flush = DeviceBuffer(L2_FLUSH_BYTES)
    nout = ctypes.c_int64()

    def step():
        _lib.check(L.pdsb_grid(_lib.ptr(d["u"]), _lib.ptr(d["v"]), _lib.ptr(d["freq"]), _lib.ptr(d["re"]), _lib.ptr(d["im"]),
                               _lib.ptr(d["w"]), nvis, 1, _lib.DEVICE, G, float(binsize), _lib.ptr(d["uu"]), _lib.ptr(d["vv"]),
                               _lib.CONV["expsinc"], _lib.WEIGHTING["natural"], 2.0, 0, 0, 0, 0,
                               _lib.ptr(o_re), _lib.ptr(o_im), _lib.ptr(o_w), None, None, None, _lib.DEVICE, None))

    for _ in range(args.warmup):
        step()
    _lib.check(L.pdsb_profile_reset())
    _lib.check(L.pdsb_profile_enable(1))
    sampler = ClockSampler(0)
    sampler.start()
    n0 = ctypes.c_int64()
    L.pdsb_launch_count(ctypes.byref(n0))
    total_ms = 0.0
    for _ in range(args.steps):
        _lib.check(L.pdsb_memset(_lib.ptr(flush), 0, L2_FLUSH_BYTES))
        _lib.check(L.pdsb_timer_start())
        step()
        ms = ctypes.c_double()
        _lib.check(L.pdsb_timer_stop(ctypes.byref(ms)))
        total_ms += ms.value
    n1 = ctypes.c_int64()
    L.pdsb_launch_count(ctypes.byref(n1))
    _lib.check(L.pdsb_profile_enable(0))
    clocks = sampler.stop()
    tile_ms, tile_n = ctypes.c_double(), ctypes.c_int64()
    _lib.check(L.pdsb_profile_get(b"grid_tile_accum", ctypes.byref(tile_ms), ctypes.byref(tile_n)))
    # end to end through the Python mirror: host numpy arrays in, gridded Visibilities out
    data = Visibilities(u, v, freq, re, im, w)
    grid_py(data, gridsize=G, binsize=binsize, convolution="expsinc", deterministic=False)
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps // 2)):
        grid_py(data, gridsize=G, binsize=binsize, convolution="expsinc", deterministic=False)
    e2e_s = (time.perf_counter() - t0) / max(1, args.steps // 2)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    alg_bytes = nvis * 40.0 + G * G * 24.0
    step_ms = total_ms / args.steps
    line = {"metric": "gridded_visibilities_per_s", "value": nvis / (step_ms * 1e-3), "unit": "vis/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "replicas only (not sharded)", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "grid() of 10M visibilities onto 2048x2048, exp*sinc convolution kernel, natural "
                                   "weights, fast (sorted-tile) mode (BASELINE.json configs[3])",
                       "nvis": nvis, "gridsize": G, "cache": "L2 flushed (256 MB memset) before every timed step"},
            "e2e": {"value": nvis / e2e_s, "unit": "vis/s", "ms_per_step": e2e_s * 1e3,
                    "h2d_bytes_per_step": int(nvis * 40 + 2 * G * 8), "d2h_bytes_per_step": int(3 * G * G * 8),
                    "api": "pdspy_b200.interferometry.grid(data, ..., deterministic=False) (numpy in, Visibilities out)"},
            "gpu_launches": int(n1.value - n0.value), "clocks": clocks,
            "roofline": {"kernel": "whole grid() step (prep + tile sort + grid_tile2_kernel + normalise)", "bound": "hbm",
                         "achieved": alg_bytes / (step_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                         "frac": alg_bytes / (step_ms * 1e-3) / 1e9 / hbm,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                         "algorithmic_bytes": alg_bytes, "tile_kernel_ms": tile_ms.value / max(tile_n.value, 1),
                         "traffic": None,
                         "note": "36 fp64 read-modify-writes per visibility make on-chip accumulation, not HBM, the "
                                 "limiter (SURVEY.md section 8d caveat)"}}
    if not args.no_cpu_baseline:
        from oracle import build_ref
        ref = build_ref.load()
        ns = 500_000
        if ref is not None:
            dd = ref.Visibilities(u[:ns].copy(), v[:ns].copy(), freq, re[:ns].copy(), im[:ns].copy(), w[:ns].copy())
            t0 = time.perf_counter()
            ref.grid(dd, gridsize=G, binsize=binsize, convolution="expsinc")
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": ns / dt, "unit": "vis/s", "cores": 1, "kind": "reference", "seconds": dt,
                                    "sample": "first %d of %d visibilities, the reference's own compiled grid() "
                                              "(oracle/_ref), serial by construction" % (ns, nvis)}
    emit(line)


def run_walker_batch(args, cfg, rank, local_rank, world):
    """BASELINE.json configs[4]: W walkers' cubes against ONE dataset; walkers are split over the ranks,
    every rank holds the whole dataset, no collective on the data path (one gather of W doubles)."""
    import torch
    import torch.distributed as dist
    from pdspy_b200 import _lib, DeviceBuffer, PinnedArray, dist as pdist
    from pdspy_b200.device import Dataset
    L = _lib.lib()
    _lib.check(L.pdsb_set_stream(torch.cuda.current_stream().cuda_stream))
    n, nf = cfg["npix"], cfg["nf"]
    u, v = synth.synth_uv(cfg["nuv"], cfg["pixelsize"] * A)
    re, im, w = synth.synth_data(cfg["nuv"], nf)
    ds = Dataset(u, v)
    ds.set_data(re, im, w)
    del re, im, w
    ws, we = pdist.shard_walkers(args.walkers, rank, world)
    nw = we - ws
    base = np.ascontiguousarray(synth.synth_image(n, nf, cfg["pixelsize"])[:, :, :, 0])
    pinned = PinnedArray((max(nw, 1), n, n, nf))
    for k in range(nw):
        pinned.array[k] = base * (1.0 + 0.01 * (ws + k))
    dcubes = DeviceBuffer.from_numpy(pinned.array[:max(nw, 1)])
    dxy = cfg["pixelsize"] * A
    dra = np.ascontiguousarray(np.full(max(nw, 1), cfg["dRA"] * A))
    ddec = np.ascontiguousarray(np.full(max(nw, 1), cfg["dDec"] * A))
    out = np.empty(max(nw, 1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(cubes, kind):
        if nw > 0:
            _lib.check(L.pdsb_loglike_batch(ds.handle, _lib.ptr(cubes), nw, n, n, nf, kind, float(dxy), _lib.ptr(dra),
                                            _lib.ptr(ddec), _lib.ptr(out)))

    step(dcubes, _lib.DEVICE)                       # warm-up: one full batch
    barrier()
    n0 = ctypes.c_int64()
    L.pdsb_launch_count(ctypes.byref(n0))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(dcubes, _lib.DEVICE)
    e1.record()
    barrier()
    n1 = ctypes.c_int64()
    L.pdsb_launch_count(ctypes.byref(n1))
    dev_ms = e0.elapsed_time(e1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(pinned.array, _lib.HOST)
    barrier()
    e2e_s = time.perf_counter() - t0
    # extra (not in value / e2e): the same batch on the opt-in tcgen05 tensor-core kernel
    lnlike0 = float(out[0])
    _lib.check(L.pdsb_set_dft_variant(200))
    step(dcubes, _lib.DEVICE)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(dcubes, _lib.DEVICE)
    e1.record()
    barrier()
    tc_ms = e0.elapsed_time(e1)
    lnlike0_tc = float(out[0])
    _lib.check(L.pdsb_set_dft_variant(0))
    t = torch.tensor([dev_ms, e2e_s, tc_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s, tc_ms = float(t[0]), float(t[1]), float(t[2])
    if rank == 0:
        pairs = float(n) * n * cfg["nuv"] * nf * args.walkers * args.steps
        cfgd = describe(cfg, world)
        cfgd.update({"walkers": args.walkers, "walkers_per_rank": nw,
                     "partition": "%d walkers split over %d rank(s), dataset replicated, no data-path collective"
                                  % (args.walkers, world), "cache": "each cube (134 MB fp64) exceeds the L2"})
        emit({
            "metric": METRIC, "value": pairs / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": 1, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 products, f64 phase seeds and accumulation", "data": "synthetic",
            "config": cfgd, "likelihood_evals_per_s": args.walkers * args.steps / (dev_ms * 1e-3),
            "e2e": {"value": pairs / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(nw * base.nbytes),
                    "d2h_bytes_per_step": int(nw * nf * 8), "ms_per_step": e2e_s / args.steps * 1e3,
                    "likelihood_evals_per_s": args.walkers * args.steps / e2e_s,
                    "api": "pdsb_loglike_batch (host fp64 cubes in, host lnlike[W] out)"},
            "gpu_launches": int(n1.value - n0.value), "lnlike0": lnlike0,
            "extras": {"tensor_core_variant": {
                "what": "same batch with the experimental opt-in tcgen05 DFT kernel (pdsb_set_dft_variant(200)); "
                        "NOT used for value / e2e",
                "ms_per_step": tc_ms / args.steps, "value": pairs / (tc_ms * 1e-3), "unit": UNIT,
                "likelihood_evals_per_s": args.walkers * args.steps / (tc_ms * 1e-3),
                "speedup_vs_default": dev_ms / tc_ms, "lnlike0": lnlike0_tc,
                "lnlike0_rel_diff_vs_default": abs(lnlike0_tc - lnlike0) / abs(lnlike0)}}})


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
    ap.add_argument("--walkers", type=int, default=128, help="C5: total emcee walkers (sharded over ranks)")
    ap.add_argument("--nuv", type=int, default=0, help="override the uv count (testing)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = workload_config(args.workload, args.nuv or None)

    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return
    if args.workload == "C4":
        os.environ["PDSB_DEVICE"] = str(local_rank)
        run_gridding(args, cfg, rank, world)
        return

    import torch
    import torch.distributed as dist
    os.environ["PDSB_DEVICE"] = str(local_rank)
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from pdspy_b200 import _lib, DeviceBuffer, PinnedArray
    from pdspy_b200.dist import ShardedLikelihood
    L = _lib.lib()
    if args.workload == "C5":
        run_walker_batch(args, cfg, rank, local_rank, world)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    data, _ = make_local_data(cfg, rank, world)
    like = ShardedLikelihood(data)            # uploads the shard; libpdsb now runs on torch's stream
    n, nf = cfg["npix"], cfg["nf"]
    cube = np.ascontiguousarray(synth.synth_image(n, nf, cfg["pixelsize"])[:, :, :, 0])     # [n, n, nf] fp64
    dxy, dra, ddec = cfg["pixelsize"] * A, cfg["dRA"] * A, cfg["dDec"] * A
    dcube = DeviceBuffer.from_numpy(cube)
    pinned = PinnedArray(cube.shape)
    pinned.array[...] = cube
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
    pairs_step = float(n) * n * cfg["nuv"] * nf

    def step_device():
        return like(dcube, dxy, dra, ddec, kind=_lib.DEVICE, shape=(n, n))

    # end to end: on several GPUs every rank uploads its 1/world slab of the host cube and the slabs are
    # all-gathered over NVLink (ShardedLikelihood.stage_cube); on one GPU the cube is uploaded whole
    shard_upload = world > 1 and n % world == 0

    def step_e2e():
        return like(pinned.array, dxy, dra, ddec, kind=_lib.HOST, cube="sharded" if shard_upload else None)

    # ---- device-resident leg: `value` ----
    for _ in range(args.warmup):
        flush.zero_()
        ll = step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _lib.check(L.pdsb_profile_reset())
    _lib.check(L.pdsb_profile_enable(1))
    n0 = ctypes.c_int64()
    L.pdsb_launch_count(ctypes.byref(n0))
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for e0, e1 in evs:
        flush.zero_()
        e0.record()
        ll = step_device()
        e1.record()
    barrier()
    n1 = ctypes.c_int64()
    L.pdsb_launch_count(ctypes.byref(n1))
    _lib.check(L.pdsb_profile_enable(0))
    clocks = sampler.stop() if rank == 0 else None
    total_ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    dft_ms, dft_n = ctypes.c_double(), ctypes.c_int64()
    _lib.check(L.pdsb_profile_get(b"dft_", ctypes.byref(dft_ms), ctypes.byref(dft_n)))
    t = torch.tensor([total_ms, dft_ms.value], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, dft_ms_max = float(t[0]), float(t[1])

    # ---- end-to-end leg: host cube (pinned) -> H2D -> fold/DFT/chi^2 -> all-reduce -> host scalar ----
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ll_e2e = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t[0])

    # ---- extras (not part of value / e2e): the experimental tensor-core variants on the same step ----
    # 200: tcgen05.mma with TMEM accumulators and a TMEM A operand, bulk-TMA B ring (pdspy_b200/csrc/dft_tc5.cu);
    # 103: mma.sync m16n8k16 (pdspy_b200/csrc/dft_mma.cu).  Both: fp16 hi+lo split operands, 3 MMAs per product,
    # fp32 accumulate; same data flow and outputs, parity <= 1.1e-6 of max|V| (tests/test_gpu_dft.py).  Reported
    # for context: the default and the headline stay on the FP32 pipe, as BASELINE.json's north star prescribes.
    def time_variant(variant, prefix):
        """(device-resident ms, e2e seconds, DFT-kernel ms from in-library events, lnlike) over args.steps steps."""
        _lib.check(L.pdsb_set_dft_variant(variant))
        for _ in range(2):
            ll_v = step_device()
        barrier()
        _lib.check(L.pdsb_profile_reset())
        _lib.check(L.pdsb_profile_enable(1))
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for e0, e1 in evs:
            flush.zero_()
            e0.record()
            ll_v = step_device()
            e1.record()
        barrier()
        _lib.check(L.pdsb_profile_enable(0))
        k_ms, k_n = ctypes.c_double(), ctypes.c_int64()
        _lib.check(L.pdsb_profile_get(prefix, ctypes.byref(k_ms), ctypes.byref(k_n)))
        step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        e2e = time.perf_counter() - t0
        _lib.check(L.pdsb_set_dft_variant(0))
        ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
        tt = torch.tensor([ms, e2e, k_ms.value], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt[0]), float(tt[1]), float(tt[2]) / max(1, k_n.value), ll_v

    # ---- extra: the reference's own algorithm (galario: FFT + bilinear interpolation) on the GPU, this rank's shard ----
    fft_out = np.empty(4)

    def step_fft(image, kind):
        _lib.check(L.pdsb_loglike_fft(like.ds.handle, _lib.ptr(image), n, nf, kind, float(dxy), float(dra), float(ddec),
                                      _lib.ptr(fft_out)))
    for _ in range(2):
        step_fft(dcube, _lib.DEVICE)
    barrier()
    fevs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for e0, e1 in fevs:
        flush.zero_()
        e0.record()
        step_fft(dcube, _lib.DEVICE)
        e1.record()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_fft(pinned.array, _lib.HOST)
    barrier()
    fft_e2e_s = time.perf_counter() - t0
    tt = torch.tensor([sum(e0.elapsed_time(e1) for e0, e1 in fevs), fft_e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    fft_ms, fft_e2e_s = float(tt[0]), float(tt[1])

    # ---- parity of the timed step itself: the same likelihood on the all-fp64 kernel (dft_f64.cu, 1e-11 from the CPU oracle) ----
    _lib.check(L.pdsb_set_dft_variant(300))
    ll_f64 = step_device()
    _lib.check(L.pdsb_set_dft_variant(0))

    tc5_ms, tc5_e2e_s, tc5_kernel_ms, ll_tc5 = time_variant(200, b"dft_tc5")
    tc_ms, tc_e2e_s, tc_kernel_ms, ll_tc = time_variant(103, b"dft_mma")

    if rank == 0:
        sm, khz = ctypes.c_int(), ctypes.c_int()
        L.pdsb_device_info(ctypes.byref(sm), ctypes.byref(khz), None, None, None)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sm_max_mhz = peaks.get("sm_max_mhz") or khz.value / 1e3
        fp32_peak = sm.value * 128 * 2 * sm_max_mhz * 1e6 / 1e12          # TFLOP/s, non-tensor FMA pipe
        tf, ms = ctypes.c_double(), ctypes.c_double()
        _lib.check(L.pdsb_bench_fma(1, 20000, ctypes.byref(tf), ctypes.byref(ms)))
        hermitian = like.ds.hermitian
        pairs_launch = float(n) * n * nf * like.ds.nuv                    # this rank's share, per launch
        dft_avg_ms = dft_ms_max / max(dft_n.value, 1)
        achieved = pairs_launch * FLOP_PER_PAIR / (dft_avg_ms * 1e-3) / 1e12
        executed_flop = pairs_launch * (0.5 if hermitian else 1.0) * 2.0   # 1 FMA per pixel per unique uv
        executed = executed_flop / (dft_avg_ms * 1e-3) / 1e12
        value = pairs_step * args.steps / (total_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32 products, f64 phase seeds and accumulation",
            "data": "synthetic", "config": describe(cfg, world),
            "likelihood_evals_per_s": args.steps / (total_ms * 1e-3),
            "lnlike": ll,
            "lnlike_rel_diff_vs_fp64_kernel": abs(ll - ll_f64) / abs(ll_f64),
            "e2e": {"value": pairs_step * args.steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(cube.nbytes) if shard_upload or world == 1 else int(cube.nbytes) * world,
                    "d2h_bytes_per_step": int((nf + 1) * 8) * world,
                    "h2d_bytes_per_step_per_rank": int(cube.nbytes) // world if shard_upload else int(cube.nbytes),
                    "ms_per_step": e2e_s / args.steps * 1e3, "likelihood_evals_per_s": args.steps / e2e_s,
                    "api": "pdspy_b200.dist.ShardedLikelihood.__call__ (host fp64 cube in, host scalar out" +
                           ("; each rank uploads 1/%d of the cube, NCCL all-gather over NVLink)" % world if shard_upload else ")"),
                    "timer": "host wall clock around K synchronous calls, max over ranks"},
            "gpu_launches": int(n1.value - n0.value),
            "clocks": clocks,
            "roofline": {
                "kernel": "dft_kernel (direct Fourier sampling, FP32 FMA pipe; no tensor cores)",
                "bound": "fp32_fma", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": achieved / fp32_peak,
                "peak_source": "nominal non-tensor FP32: SMs x 128 lanes x 2 x sm_max_mhz (MEASURED_PEAKS.json has "
                               "no FP32 entry; its sm_max_mhz is used)",
                "peak_measured_fma_microbench": tf.value,
                "algorithmic_flop_per_pair": FLOP_PER_PAIR,
                "executed": executed, "executed_frac": executed / fp32_peak,
                "executed_note": "FMA-pipe flops the inner loop actually issues per launch: the real-image mirror "
                                 "fold needs 1 FMA per pixel per uv point instead of 2, and a Hermitian-doubled uv "
                                 "list is evaluated for one half; frac > 1 is those two algorithmic savings, "
                                 "executed_frac is the pipe utilisation",
                "launch_ms": dft_avg_ms, "launches": int(dft_n.value),
                "traffic": NCU_TRAFFIC_BYTES.get((cfg["name"], world)),
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this "
                                  "launch (profiles/r01_dft_ncu_c3_default.md, r01_dft_ncu.md); null where not captured",
                "algorithmic_bytes": float(n) * n * nf * 4 + like.ds.nuv_unique * 16.0 * (1 + nf)},
        }
        def tc_entry(what, ms, e2e_s_, kernel_ms, llv, peak_key):
            # the tensor cores execute 3 fp16 MACs (hi*hi, hi*lo, lo*hi) per pixel-visibility pair actually
            # evaluated (the Hermitian half of the list, per rank)
            macs = 3.0 * float(n) * n * nf * like.ds.nuv_unique
            ach = 2.0 * macs / (kernel_ms * 1e-3) / 1e12
            peak = peaks.get(peak_key)
            return {"what": what + "; NOT used for value / e2e", "ms_per_step": ms / args.steps,
                    "value": pairs_step * args.steps / (ms * 1e-3), "unit": UNIT, "speedup_vs_default": total_ms / ms,
                    "e2e": {"value": pairs_step * args.steps / e2e_s_, "unit": UNIT, "ms_per_step": e2e_s_ / args.steps * 1e3,
                            "h2d_bytes_per_step": int(cube.nbytes) if shard_upload or world == 1 else int(cube.nbytes) * world,
                            "d2h_bytes_per_step": 8 * (nf + 1) * world},
                    "roofline": {"bound": "tensor", "kernel_ms": kernel_ms, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                                 "frac": ach / peak if peak else None,
                                 "peak_source": "MEASURED_PEAKS.json %s (cuBLAS bf16 GEMM, sustained figure: kernel timed "
                                                "inside a long step)" % peak_key,
                                 "counts": "2 flop x 3 split-product MACs per evaluated pixel-visibility pair, per rank"},
                    "lnlike": llv, "lnlike_rel_diff_vs_default": abs(llv - ll) / abs(ll),
                    "lnlike_rel_diff_vs_fp64_kernel": abs(llv - ll_f64) / abs(ll_f64)}
        line["extras"] = {
            "tensor_core_variant": tc_entry(
                "same step with the experimental opt-in DFT kernel on the 5th-generation tensor cores (tcgen05.mma, "
                "accumulators and A operand in TMEM, B through a bulk-TMA ring, warp-specialised; fp16 hi+lo split "
                "operands, 3 MMAs per product, fp32 accumulate; dft_tc5.cu, pdsb_set_dft_variant(200))",
                tc5_ms, tc5_e2e_s, tc5_kernel_ms, ll_tc5, "bf16_tflops_sustained"),
            "galario_fft_algorithm": {
                "what": "the reference's OWN algorithm for this step on the GPU - galario's FFT + bilinear interpolation "
                        "(pdsb_loglike_fft, fp64, restated from galario's published algorithm) + the same chi^2; it "
                        "carries galario's interpolation error (1e-3..4e-2 of max|V|), which the direct transform of "
                        "value / e2e does not; pairs/s counts the pairs the result represents, as in --impl reference; "
                        "NOT used for value / e2e; chi^2 of this rank's uv shard, no all-reduce",
                "ms_per_step": fft_ms / args.steps, "value": pairs_step * args.steps / (fft_ms * 1e-3), "unit": UNIT,
                "e2e": {"value": pairs_step * args.steps / fft_e2e_s, "unit": UNIT, "ms_per_step": fft_e2e_s / args.steps * 1e3,
                        "h2d_bytes_per_step": int(cube.nbytes) * world, "d2h_bytes_per_step": 32 * world},
                "lnlike_shard": float(fft_out[3])},
            "mma_sync_variant": tc_entry(
                "same step with the experimental opt-in DFT kernel on the warp-level tensor-core path (mma.sync "
                "m16n8k16, same operand split; dft_mma.cu, pdsb_set_dft_variant(103))",
                tc_ms, tc_e2e_s, tc_kernel_ms, ll_tc, "bf16_tflops_sustained")}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_port(cfg)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
