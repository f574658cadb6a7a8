"""Multi-GPU partitioning of the likelihood.

One process per GPU.  The uv list is split across ranks; every rank holds the whole image cube
(it is small: 67 MB fp32 for 512^2 x 64) and computes chi^2[nf] over its uv shard; the only
exchange is one all-reduce of nf + 1 doubles (chi^2 per channel and the data-only log term).
Alternatively the FREQUENCY CHANNELS are split (by="channels"): every rank holds all uv points of its
channels and evaluates only its channels of the cube - the partition that also divides the per-cube
work of the FFT-based kernels (the transform is per channel), which the uv split replicates on
every rank.  The exchange is the same all-reduce.
Walker batches are split by walker and need no reduction at all.

Host logic only: works with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU
tests)."""
import numpy as np


def is_hermitian_doubled(u, v):
    n = u.size
    if n < 2 or n % 2:
        return False
    h = n // 2
    return bool(np.array_equal(u[h:], -u[:h]) and np.array_equal(v[h:], -v[:h]))


def shard_bounds(n, rank, world):
    """Contiguous, balanced [start, end) of n items for `rank` of `world`."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_rows(u, v, rank, world):
    """Row indices of this rank's uv shard.  A Hermitian-doubled list (readuvfits.py:68-73) is
    split by baseline so that each shard is again [half, -half] and the device can keep folding
    it; anything else is split into contiguous chunks."""
    n = u.size
    if is_hermitian_doubled(u, v):
        h = n // 2
        s, e = shard_bounds(h, rank, world)
        return np.concatenate([np.arange(s, e), np.arange(h + s, h + e)])
    s, e = shard_bounds(n, rank, world)
    return np.arange(s, e)


def shard_channels(nf, rank, world):
    """[c0, c1) of this rank's contiguous block of the nf frequency channels (may be empty when world > nf)."""
    return shard_bounds(nf, rank, world)


def shard_visibilities(data, rank, world, by="uv"):
    """This rank's part of a Visibilities-like object (u, v, freq, real, imag, weights): a block of uv rows
    (by="uv", Hermitian pairs kept together) or all rows of a block of channels (by="channels")."""
    from .interferometry import Visibilities
    if by == "channels":
        c0, c1 = shard_channels(len(data.freq), rank, world)
        return Visibilities(np.ascontiguousarray(data.u), np.ascontiguousarray(data.v),
                            np.ascontiguousarray(data.freq[c0:c1]), np.ascontiguousarray(data.real[:, c0:c1]),
                            np.ascontiguousarray(data.imag[:, c0:c1]), np.ascontiguousarray(data.weights[:, c0:c1]))
    if by != "uv":
        raise ValueError("by must be 'uv' or 'channels'")
    rows = shard_rows(data.u, data.v, rank, world)
    return Visibilities(np.ascontiguousarray(data.u[rows]), np.ascontiguousarray(data.v[rows]), data.freq,
                        np.ascontiguousarray(data.real[rows]), np.ascontiguousarray(data.imag[rows]),
                        np.ascontiguousarray(data.weights[rows]))


def shard_walkers(nwalkers, rank, world):
    return shard_bounds(nwalkers, rank, world)


def combine_lnlike(chi2_and_logsum, group=None):
    """All-reduce (sum) a tensor holding [chi2_0..chi2_{nf-1}, logsum] over ranks and return the
    emcee.py:31-43 value  -0.5*sum(chi2) - 2*logsum.  `chi2_and_logsum` is a torch tensor on the
    device the process group communicates from (CUDA for NCCL, CPU for gloo)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(chi2_and_logsum, op=dist.ReduceOp.SUM, group=group)
    t = chi2_and_logsum.double().cpu()
    return float(-0.5 * t[:-1].sum() - 2.0 * t[-1])


class ShardedLikelihood:
    """Fused image -> visibilities -> chi^2 over this rank's uv shard, then one all-reduce of
    nf+1 doubles.  torch supplies the NCCL process group and the tiny device tensor that is
    reduced; all arithmetic is in libpdsb (pdsb_loglike_device).

        like = ShardedLikelihood(shard_visibilities(data, rank, world))
        lnlike = like(image_cube, dxy_rad, dRA_rad, dDec_rad)      # same value on every rank

    Channel partition: the shard holds all uv points of channels [c0, c0 + nf_local) of cubes with nf_total
    channels; every call slices those channels out of the cube on the device (pdsb_channel_slice) and writes
    its chi^2 into its own window of the nf_total + 1 doubles that are all-reduced:

        c0, c1 = shard_channels(nf_total, rank, world)
        like = ShardedLikelihood(shard_visibilities(data, rank, world, by="channels"), channels=(c0, nf_total))
    """

    def __init__(self, data_shard, group=None, channels=None):
        import ctypes
        import torch
        from . import _lib
        from .device import Dataset
        self._lib = _lib
        self.L = _lib.lib()
        self.group = group
        self.torch = torch
        # run libpdsb on torch's current stream so the collective is ordered behind the kernels
        _lib.check(self.L.pdsb_set_stream(torch.cuda.current_stream().cuda_stream))
        self.ds = Dataset(data_shard.u, data_shard.v)
        if data_shard.real.shape[1] > 0:                     # (a channel shard is empty when world > nf)
            self.ds.set_data(data_shard.real, data_shard.imag, data_shard.weights)
        self.nf_local = data_shard.real.shape[1]
        self.c0, self.nf = (0, self.nf_local) if channels is None else (int(channels[0]), int(channels[1]))
        self.channels = channels is not None
        if self.c0 < 0 or self.c0 + self.nf_local > self.nf:
            raise ValueError("channel window [%d, %d) outside the cube's %d channels" % (self.c0, self.c0 + self.nf_local,
                                                                                         self.nf))
        self._slice = None
        ls = ctypes.c_double(0.0)
        if self.nf_local > 0:
            _lib.check(self.L.pdsb_dataset_logsum(self.ds.handle, ctypes.byref(ls)))
        self.logsum = ls.value
        self.buf = torch.zeros(self.nf + 1, dtype=torch.float64, device="cuda")
        self.logsum_dev = torch.tensor([ls.value], dtype=torch.float64, device="cuda")

    def chi2_device(self, image, ny, nx, kind, dxy, dRA, dDec):
        """Launch the local part; result (chi2[nf], logsum) stays in self.buf on the device."""
        _lib = self._lib
        if self.channels:
            # this rank's channels of the cube, compact on the device; chi^2 into its window of the reduced vector
            self.buf.zero_()
            if self.nf_local > 0:
                if self._slice is None or self._slice.numel() != ny * nx * self.nf_local:
                    self._slice = self.torch.empty(ny * nx * self.nf_local, dtype=self.torch.float64, device="cuda")
                _lib.check(self.L.pdsb_channel_slice(_lib.ptr(image), kind, ny * nx, self.nf, self.c0, self.nf_local,
                                                     self._slice.data_ptr()))
                _lib.check(self.L.pdsb_loglike_device(self.ds.handle, self._slice.data_ptr(), ny, nx, self.nf_local,
                                                      _lib.DEVICE, float(dxy), float(dRA), float(dDec),
                                                      self.buf.data_ptr() + 8 * self.c0))
            self.buf[self.nf:].copy_(self.logsum_dev)
            return self.buf
        _lib.check(self.L.pdsb_loglike_device(self.ds.handle, _lib.ptr(image), ny, nx, self.nf, kind, float(dxy),
                                              float(dRA), float(dDec), self.buf.data_ptr()))
        self.buf[self.nf:].copy_(self.logsum_dev)
        return self.buf

    def stage_cube(self, image, cube="sharded", shape=None):
        """Put a HOST cube [ny, nx, nf] (fp64, C order; pinned memory makes the copies asynchronous DMA) on
        every rank's GPU and return the device tensor.

        cube="sharded": every rank holds the host cube (the usual MPI-pool situation after a broadcast of
            the parameters, or a cube read from shared storage); each uploads only its 1/world slab of rows
            and the slabs are exchanged over NVLink with one NCCL all-gather - world times less PCIe traffic
            per rank than uploading the whole cube everywhere.
        cube="rank0": only rank 0's `image` is read; the other ranks may pass image=None together with
            shape=(ny, nx, nf) (needed on their first call; later calls remember it).  Rank 0 uploads the cube
            and NCCL broadcasts it.
        Falls back to a plain full upload on one rank / without a process group."""
        import torch.distributed as dist
        torch = self.torch
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
        if image is not None:
            shape = tuple(image.shape[:3])
        elif shape is not None:
            shape = tuple(int(x) for x in shape)
        else:
            shape = getattr(self, "_cube_shape", None)
        if shape is None or len(shape) != 3:
            # every rank raises before any collective is entered: nobody is left waiting in a broadcast
            raise ValueError("stage_cube needs the cube, or shape=(ny, nx, nf) when image is None")
        if image is None and cube != "rank0":
            raise ValueError("image=None is only meaningful with cube='rank0'")
        if getattr(self, "_cube_dev", None) is None or tuple(self._cube_dev.shape) != shape:
            self._cube_dev = torch.empty(shape, dtype=torch.float64, device="cuda")
            self._cube_shape = shape
        dev = self._cube_dev
        if not multi:
            dev.copy_(torch.from_numpy(image.reshape(shape)), non_blocking=True)
            return dev
        rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if cube == "rank0":
            if rank == 0:
                dev.copy_(torch.from_numpy(image.reshape(shape)), non_blocking=True)
            dist.broadcast(dev, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
            return dev
        if cube != "sharded":
            raise ValueError("cube must be 'sharded' or 'rank0'")
        ny = shape[0]
        if ny % world:                       # uneven slabs: every rank uploads everything
            dev.copy_(torch.from_numpy(image.reshape(shape)), non_blocking=True)
            return dev
        s, e = shard_bounds(ny, rank, world)
        dev[s:e].copy_(torch.from_numpy(image.reshape(shape)[s:e]), non_blocking=True)
        dist.all_gather_into_tensor(dev, dev[s:e], group=self.group)
        return dev

    def __call__(self, image, dxy, dRA=0.0, dDec=0.0, kind=0, shape=None, cube=None):
        """kind 0: host cube, 1: device pointer / DeviceBuffer (then `shape` = (ny, nx)).  With a host cube,
        cube="sharded" | "rank0" stages it through stage_cube (with "rank0" the other ranks may pass image=None
        and shape=(ny, nx)); None keeps the per-rank full upload."""
        if kind == 0 and cube is not None:
            dev = self.stage_cube(image, cube, shape=(tuple(shape) + (self.nf,)) if shape is not None and len(shape) == 2
                                  else shape)
            ny, nx = dev.shape[0], dev.shape[1]
            return combine_lnlike(self.chi2_device(int(dev.data_ptr()), ny, nx, 1, dxy, dRA, dDec), self.group)
        if shape is None:
            ny, nx = image.shape[0], image.shape[1]
        else:
            ny, nx = shape
        return combine_lnlike(self.chi2_device(image, ny, nx, kind, dxy, dRA, dDec), self.group)


def _cell_centres(G, binsize):
    """numpy.linspace of libinterferometry.pyx:370-381."""
    if G % 2 == 0:
        uu = np.linspace(-G * binsize / 2, (G / 2 - 1) * binsize, G)
    else:
        uu = np.linspace(-(G - 1) * binsize / 2, (G - 1) * binsize / 2, G)
    return uu, uu.copy()


def _gridded_to_host(L, _lib, maps, G, nch, uu, vv):
    """The three normalised device maps [3, G*G, nch] as host arrays (read back through libpdsb's multi-threaded pinned
    ring straight into their final arrays) and the flattened coordinate arrays of the result (numpy.meshgrid,
    libinterferometry.pyx:383, filled on a second host thread meanwhile)."""
    from .interferometry.grid import MeshFiller
    filler = MeshFiller(uu, vv)
    host = [np.empty((G * G, nch)) for _ in range(3)]
    try:
        for k in range(3):
            _lib.check(L.pdsb_memcpy(_lib.ptr(host[k]), _lib.HOST, maps[k].data_ptr(), _lib.DEVICE, host[k].nbytes))
        _lib.check(L.pdsb_synchronize())
    finally:
        filler.join()
    new_u, new_v = filler.result()
    return new_u.reshape(G * G), new_v.reshape(G * G), host


def _reduced_weight_map(L, _lib, torch, dist, multi, group, arrays, nuv, nf, G, binsize, uu, vv, weighting, npixels, mode,
                        deterministic, nch):
    """First half of the re-weighting (libinterferometry.pyx:429-461) over all ranks: every rank box-sums the clamped
    weights of ITS visibilities (pdsb_grid_weights_map; inside a row band when pdsb_set_grid_band is set), the maps
    and the per-channel weight sums are all-reduced, and the ones of :430 are added.  Returns (device map, sumw)."""
    import ctypes
    binned = torch.zeros(G * G * nch, dtype=torch.float64, device="cuda")
    sumw = np.zeros(nf)
    n_out = ctypes.c_int64(0)
    u, v, freq, re, im, w = arrays
    _lib.check(L.pdsb_grid_weights_map(_lib.ptr(u), _lib.ptr(v), _lib.ptr(freq), _lib.ptr(re), _lib.ptr(im), _lib.ptr(w),
                                       nuv, nf, _lib.HOST, G, float(binsize), _lib.ptr(uu), _lib.ptr(vv),
                                       _lib.WEIGHTING[weighting], int(npixels), _lib.MODE[mode], 1 if deterministic else 0,
                                       0, binned.data_ptr(), _lib.ptr(sumw), ctypes.byref(n_out)))
    return binned, sumw


def sharded_grid(data_shard, gridsize=256, binsize=2000.0, convolution="pillbox", imaging=False, weighting="natural",
                 robust=2, npixels=0, mode="continuum", group=None):
    """grid() of a data set whose visibilities are split over the ranks (SURVEY.md section 8e,
    throughput mode): each rank grids its shard into private raw-sum maps on its GPU (fast mode), the
    three maps are all-reduced over NCCL, then normalised on the device.  With uniform / superuniform / robust
    weighting the binned-weight map (and, robust, the weight sums) are made per shard and all-reduced first
    (libinterferometry.pyx:429-485 needs them for the whole data set).  Every rank returns the full
    gridded Visibilities.  Summation order differs from the serial reference: 1e-15-level differences,
    like the single-GPU fast mode.  The mean frequency must be the same on every rank (it is: the
    shards share `freq`)."""
    import ctypes
    import torch
    import torch.distributed as dist
    from . import _lib
    from .interferometry import Visibilities
    from .interferometry.grid import _WARNING
    L = _lib.lib()
    _lib.check(L.pdsb_set_stream(torch.cuda.current_stream().cuda_stream))
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    u, v, freq = _lib.f64(data_shard.u), _lib.f64(data_shard.v), _lib.f64(data_shard.freq)
    re, im, w = _lib.f64(data_shard.real), _lib.f64(data_shard.imag), _lib.f64(data_shard.weights)
    nuv, nf = u.size, freq.size
    nch = 1 if mode == "continuum" else nf
    G = int(gridsize)
    uu, vv = _cell_centres(G, binsize)
    maps = torch.zeros((3, G * G, nch), dtype=torch.float64, device="cuda")
    n_out = ctypes.c_int64(0)
    binned = None
    if weighting != "natural":
        binned, sumw = _reduced_weight_map(L, _lib, torch, dist, multi, group, (u, v, freq, re, im, w), nuv, nf, G, binsize,
                                           uu, vv, weighting, npixels, mode, False, nch)
        sw = torch.from_numpy(sumw).cuda()
        if multi:
            dist.all_reduce(binned, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(sw, op=dist.ReduceOp.SUM, group=group)
        binned += 1.0                                       # numpy.ones :430
        sumw = np.ascontiguousarray(sw.cpu().numpy())
        torch.cuda.synchronize()
        _lib.check(L.pdsb_set_grid_reweight(binned.data_ptr(), binned.numel(), _lib.ptr(sumw), nf))
    try:
        _lib.check(L.pdsb_grid(_lib.ptr(u), _lib.ptr(v), _lib.ptr(freq), _lib.ptr(re), _lib.ptr(im), _lib.ptr(w), nuv, nf,
                               _lib.HOST, G, float(binsize), _lib.ptr(uu), _lib.ptr(vv), _lib.CONV[convolution],
                               _lib.WEIGHTING[weighting], float(robust), int(npixels), _lib.MODE[mode],
                               2, 0,                         # imaging = 2: raw sums; fast mode
                               maps[0].data_ptr(), maps[1].data_ptr(), maps[2].data_ptr(), None, None, None,
                               _lib.DEVICE, ctypes.byref(n_out)))
    finally:
        if binned is not None:
            _lib.check(L.pdsb_set_grid_reweight(None, 0, None, 0))
    nout = torch.tensor([n_out.value], dtype=torch.int64, device="cuda")
    if multi:
        dist.all_reduce(maps, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(nout, op=dist.ReduceOp.SUM, group=group)
    _lib.check(L.pdsb_grid_normalise(maps[0].data_ptr(), maps[1].data_ptr(), maps[2].data_ptr(), G, nch,
                                     1 if imaging else 0))
    new_u, new_v, host = _gridded_to_host(L, _lib, maps, G, nch, uu, vv)
    if int(nout.item()) > 0:
        print(_WARNING)
    out_freq = np.array([freq.sum() / freq.size]) if mode == "continuum" else freq
    return Visibilities(new_u, new_v, out_freq, host[0], host[1], host[2])


def band_rows(u, v, freq, gridsize, binsize, row_lo, row_hi, lo, hi, include_outside=False):
    """Ascending indices of the visibility rows whose footprint can reach output rows [row_lo, row_hi) of
    grid() (footprint rows j - lo .. j + hi around the home row j, libinterferometry.pyx:493-507).
    Conservative by two cells: the exact row test is made on the device (pdsb_set_grid_band), this only
    keeps a rank from streaming visibilities that cannot matter.  include_outside adds the rows with an
    element off the grid, so that exactly one rank reports them (the WARNING of :424)."""
    G = int(gridsize)
    scale = np.asarray(freq, dtype=np.float64) / (np.sum(freq) / len(freq)) / binsize
    jf = np.floor(np.asarray(v, dtype=np.float64)[:, None] * scale[None, :] + G / 2.)
    sel = ((jf >= row_lo - hi - 2) & (jf < row_hi + lo + 2)).any(axis=1)
    if include_outside:
        xf = np.floor(np.asarray(u, dtype=np.float64)[:, None] * scale[None, :] + G / 2.)
        sel |= ((jf < 0) | (jf >= G) | (xf < 0) | (xf >= G)).any(axis=1)
    return np.flatnonzero(sel)


def banded_grid(data, gridsize=256, binsize=2000.0, convolution="pillbox", imaging=False, weighting="natural", robust=2,
                npixels=0, mode="continuum", group=None, bands=None, weight_map=None):
    """grid() over several GPUs with results BIT-IDENTICAL to the single-GPU ordered mode (SURVEY.md section
    8e (ii)).  Every rank sees the whole data set and owns a band of gridsize/world output rows: it streams
    only the visibilities whose footprint can reach its band (in their original order, which is what fixes
    the rounding), accumulates that band in the reference's (k, n) order (pdsb_set_grid_band + ordered
    pdsb_grid with raw sums), the bands are summed over NCCL (every cell is non-zero on one rank only, so
    the sum adds exact zeros) and normalised on the device.

    Uniform / superuniform / robust weighting (libinterferometry.pyx:429-485): the binned-weight map is made the
    same way first - every rank box-sums the weights into ITS band of the map in (k, n) order from the visibilities
    within +-npixels rows of it, the bands are all-reduced (exact zeros elsewhere) - so every rank holds the map
    the single-GPU run computes, bit for bit; robust's two global sums are then taken over the full map and the
    full weight array exactly as on one GPU.

    bands=(index, count) computes ONE band without a process group and returns (raw device maps [3, G*G,
    nch], n_outside) for the caller to sum over all indices and pass to pdsb_grid_normalise (this is how
    the single-GPU test emulates N ranks); with a re-weighting scheme `weight_map` must then be the summed
    result of banded_weight_map over all bands.  Normally leave both None."""
    import ctypes
    import torch
    import torch.distributed as dist
    from . import _lib
    from .interferometry import Visibilities
    from .interferometry.grid import _WARNING
    L = _lib.lib()
    _lib.check(L.pdsb_set_stream(torch.cuda.current_stream().cuda_stream))
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if bands is not None:
        rank, world = bands
    elif multi:
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    u, v, freq = _lib.f64(data.u), _lib.f64(data.v), _lib.f64(data.freq)
    nf = freq.size
    nch = 1 if mode == "continuum" else nf
    G = int(gridsize)
    uu, vv = _cell_centres(G, binsize)
    row_lo, row_hi = shard_bounds(G, rank, world)
    lo, hi = (2, 3) if convolution == "expsinc" else (1, 1)          # nmin, nmax of :405-417
    binned = sumw = None
    if weighting != "natural":
        if bands is not None:
            if weight_map is None:
                raise ValueError("bands=... with re-weighting needs weight_map (sum of banded_weight_map over the bands)")
            binned, sumw = weight_map
        else:
            binned, sumw = banded_weight_map(data, G, binsize, weighting, npixels, mode, (rank, world))
            if multi:
                dist.all_reduce(binned, op=dist.ReduceOp.SUM, group=group)
    maps = torch.zeros((3, G * G, nch), dtype=torch.float64, device="cuda")
    n_out = ctypes.c_int64(0)
    if row_hi > row_lo:
        rows = band_rows(u, v, freq, G, binsize, row_lo, row_hi, lo, hi, include_outside=(rank == 0))
        take = (lambda a: np.ascontiguousarray(_lib.f64(a)[rows]))
        _lib.check(L.pdsb_set_grid_band(row_lo, row_hi))
        if binned is not None:
            torch.cuda.synchronize()
            _lib.check(L.pdsb_set_grid_reweight(binned.data_ptr(), binned.numel(), _lib.ptr(sumw), nf))
        try:
            _lib.check(L.pdsb_grid(_lib.ptr(take(u)), _lib.ptr(take(v)), _lib.ptr(freq), _lib.ptr(take(data.real)),
                                   _lib.ptr(take(data.imag)), _lib.ptr(take(data.weights)), rows.size, nf, _lib.HOST,
                                   G, float(binsize), _lib.ptr(uu), _lib.ptr(vv), _lib.CONV[convolution],
                                   _lib.WEIGHTING[weighting], float(robust), int(npixels),
                                   _lib.MODE[mode], 2, 1,            # imaging = 2: raw sums; ordered mode
                                   maps[0].data_ptr(), maps[1].data_ptr(), maps[2].data_ptr(), None, None, None,
                                   _lib.DEVICE, ctypes.byref(n_out)))
        finally:
            _lib.check(L.pdsb_set_grid_band(0, 0))
            _lib.check(L.pdsb_set_grid_reweight(None, 0, None, 0))
    if bands is not None:
        return maps, n_out.value                     # raw band maps on the device, for the caller to combine
    nout = torch.tensor([n_out.value], dtype=torch.int64, device="cuda")
    if multi:
        dist.all_reduce(maps, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(nout, op=dist.ReduceOp.SUM, group=group)
    _lib.check(L.pdsb_grid_normalise(maps[0].data_ptr(), maps[1].data_ptr(), maps[2].data_ptr(), G, nch,
                                     1 if imaging else 0))
    new_u, new_v, host = _gridded_to_host(L, _lib, maps, G, nch, uu, vv)
    if int(nout.item()) > 0:
        print(_WARNING)
    out_freq = np.array([freq.sum() / freq.size]) if mode == "continuum" else freq
    return Visibilities(new_u, new_v, out_freq, host[0], host[1], host[2])


def banded_weight_map(data, gridsize, binsize, weighting, npixels, mode, band):
    """This band's rows of the binned-weight map (the ones of :430 plus the box sums added in the reference's (k, n)
    order; exact zeros outside the band), and the per-channel sums of the clamped weights over the WHOLE data set
    (the same on every rank, and the same arithmetic as the single-GPU run).  band = (index, count)."""
    import ctypes
    import torch
    from . import _lib
    L = _lib.lib()
    rank, world = band
    u, v, freq = _lib.f64(data.u), _lib.f64(data.v), _lib.f64(data.freq)
    re, im, w = _lib.f64(data.real), _lib.f64(data.imag), _lib.f64(data.weights)
    nuv, nf = u.size, freq.size
    nch = 1 if mode == "continuum" else nf
    G = int(gridsize)
    uu, vv = _cell_centres(G, binsize)
    row_lo, row_hi = shard_bounds(G, rank, world)
    binned = torch.zeros(G * G * nch, dtype=torch.float64, device="cuda")
    sumw = np.zeros(nf)
    if row_hi > row_lo:
        _lib.check(L.pdsb_set_grid_band(row_lo, row_hi))
    try:
        n_out = ctypes.c_int64(0)
        if row_hi > row_lo:
            _lib.check(L.pdsb_grid_weights_map(_lib.ptr(u), _lib.ptr(v), _lib.ptr(freq), _lib.ptr(re), _lib.ptr(im),
                                               _lib.ptr(w), nuv, nf, _lib.HOST, G, float(binsize), _lib.ptr(uu),
                                               _lib.ptr(vv), _lib.WEIGHTING[weighting], int(npixels), _lib.MODE[mode], 1,
                                               1, binned.data_ptr(), _lib.ptr(sumw), ctypes.byref(n_out)))
    finally:
        _lib.check(L.pdsb_set_grid_band(0, 0))
    return binned, sumw
