"""Multi-rank host logic on CPU: uv sharding and the nf+1 all-reduce over gloo, world_size 2.
The per-shard chi^2 here comes from the oracle (no GPU in this test); the GPU path plugs into the
same combine step (pdspy_b200/dist.py:ShardedLikelihood)."""
import os
import socket

import numpy as np
import pytest

import synth
from pdspy_b200 import dist as pdist
from pdspy_b200.interferometry import Visibilities


def _data(n=2000, nf=3, hermitian=True):
    u, v = synth.synth_uv(n, 0.1 * synth.ARCSEC)
    if not hermitian:
        u = u + 1.0
    re, im, w = synth.synth_data(n, nf)
    return Visibilities(u, v, synth.synth_freq(nf), re, im, w)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("hermitian", [True, False])
def test_shards_partition_the_rows_and_stay_hermitian(world, hermitian):
    d = _data(2000 if hermitian else 1999 + 1, 2, hermitian)
    seen = []
    for r in range(world):
        rows = pdist.shard_rows(d.u, d.v, r, world)
        seen.append(rows)
        s = pdist.shard_visibilities(d, r, world)
        assert s.real.shape == (rows.size, 2)
        if hermitian:
            assert pdist.is_hermitian_doubled(s.u, s.v)       # each shard can still be folded on the device
            h = rows.size // 2
            np.testing.assert_array_equal(s.imag[h:], -s.imag[:h])
    allrows = np.sort(np.concatenate(seen))
    np.testing.assert_array_equal(allrows, np.arange(d.u.size))
    sizes = [r.size for r in seen]
    assert max(sizes) - min(sizes) <= 2


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_channel_shards_partition_the_channels(world):
    d = _data(400, 5)
    cols = []
    for r in range(world):
        c0, c1 = pdist.shard_channels(5, r, world)
        s = pdist.shard_visibilities(d, r, world, by="channels")
        assert s.real.shape == (400, c1 - c0) and s.real.flags.c_contiguous
        np.testing.assert_array_equal(s.u, d.u)
        np.testing.assert_array_equal(s.freq, d.freq[c0:c1])
        np.testing.assert_array_equal(s.weights, d.weights[:, c0:c1])
        cols.extend(range(c0, c1))
    assert cols == list(range(5))                 # world = 8 > nf = 5: three ranks hold nothing
    with pytest.raises(ValueError):
        pdist.shard_visibilities(d, 0, world, by="rows")


def test_shard_bounds_edge_cases():
    assert pdist.shard_bounds(0, 0, 4) == (0, 0)
    assert [pdist.shard_bounds(5, r, 4) for r in range(4)] == [(0, 2), (2, 3), (3, 4), (4, 5)]
    assert pdist.shard_walkers(128, 7, 8) == (112, 128)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from oracle import likelihood as ol
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = _data()
    rng = np.random.default_rng(11)
    m_re, m_im = rng.normal(size=d.real.shape), rng.normal(size=d.real.shape)
    rows = pdist.shard_rows(d.u, d.v, rank, world)
    s = pdist.shard_visibilities(d, rank, world)
    chi2 = ol.chi2_per_channel_numpy(s.real, s.imag, s.weights, m_re[rows], m_im[rows])
    good = s.weights > 0
    logsum = np.sum(np.log(s.weights[good] / (2 * np.pi)))
    t = torch.tensor(np.concatenate([chi2, [logsum]]))
    ll = pdist.combine_lnlike(t)
    full = ol.lnlike_vis_numpy(d.real, d.imag, d.weights, m_re, m_im)
    # the same data set split by channels: every rank fills its window of the nf + 1 reduced doubles
    c0, c1 = pdist.shard_channels(d.real.shape[1], rank, world)
    sc = pdist.shard_visibilities(d, rank, world, by="channels")
    tc = np.zeros(d.real.shape[1] + 1)
    tc[c0:c1] = ol.chi2_per_channel_numpy(sc.real, sc.imag, sc.weights, m_re[:, c0:c1], m_im[:, c0:c1])
    goodc = sc.weights > 0
    tc[-1] = np.sum(np.log(sc.weights[goodc] / (2 * np.pi)))
    llc = pdist.combine_lnlike(torch.tensor(tc))
    assert abs(llc - full) <= 1e-12 * abs(full)
    q.put((rank, ll, full))
    dist.destroy_process_group()


def test_two_rank_all_reduce_reproduces_the_full_likelihood():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ll, full in res:
        assert abs(ll - full) <= 1e-12 * abs(full)
    assert res[0][1] == res[1][1]            # every rank holds the same value
