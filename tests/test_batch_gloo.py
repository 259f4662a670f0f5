"""Multi-rank host logic on CPU (gloo, world_size 2): shard partition + histogram / moment all-reduce equal the
single-process result over the whole batch (the merge_histos semantic, PyPOLAR.py:2318-2384)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import restate as rs
from polarimetry_b200 import batch


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 9, 1024):
        for world in (1, 2, 3, 8):
            got = [i for r in range(world) for i in batch.shard_range(n, r, world)]
            assert got == list(range(n))
            sizes = [len(batch.shard_range(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        batch.shard_range(4, 2, 2)


def _rho_map(seed):
    rng = np.random.default_rng(seed)
    rho = rng.uniform(0, 180, size=(24, 32))
    rho[rng.random(rho.shape) < 0.3] = np.nan
    return rho


def _worker(rank, world, port, n_items, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    hist = torch.zeros(60, dtype=torch.int64)
    mom = torch.zeros(3, dtype=torch.float64)
    for k in batch.shard_range(n_items, rank, world):
        rho = _rho_map(k)
        hist += torch.from_numpy(rs.histo(rho, None, 'Rho', 0, 180, 7.0, 60, 'polar1')[0])
        mom += batch.circular_moments(torch.from_numpy(rho))
    batch.allreduce_sum_(hist)
    batch.allreduce_sum_(mom)
    if rank == 0:
        q.put((hist.numpy(), mom.numpy()))
    dist.destroy_process_group()


def test_two_rank_histogram_allreduce_matches_merged_histogram():
    n_items, world = 7, 2
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    hist, mom = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    allv = np.concatenate([_rho_map(k).ravel() for k in range(n_items)])
    want = rs.histo(allv, None, 'Rho', 0, 180, 7.0, 60, 'polar1')[0]      # concatenate-then-histogram, as merge_histos does
    assert np.array_equal(hist, want)
    fin = allv[np.isfinite(allv)]
    assert int(mom[0]) == fin.size
    assert abs(batch.circular_mean_from_moments(torch.from_numpy(mom)) - rs.circularmean(fin)) < 1e-9


def _frame_hist(k):
    rng = np.random.default_rng(900 + k)
    return torch.from_numpy(rng.integers(0, 1000, size=(2, 3, 16), dtype=np.int64))


def _ts_worker(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    own = batch.shard_range(n_frames, rank, world)
    local = torch.stack([_frame_hist(k) for k in own]) if len(own) else torch.zeros((0, 2, 3, 16), dtype=torch.int64)
    full = batch.place_frames(local, own, n_frames)
    batch.allreduce_sum_(full)
    q.put((rank, full.numpy()))
    dist.destroy_process_group()


@pytest.mark.parametrize('n_frames', [5, 1])
def test_two_rank_time_series_frames(n_frames):
    """configs[4] semantics on CPU (gloo, world 2): per-frame histograms placed in a zero buffer and all-reduced once give every rank
    the whole series, also when a rank owns no frame."""
    world = 2
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_ts_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get() for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = np.stack([_frame_hist(k).numpy() for k in range(n_frames)])
    for r in range(world):
        assert np.array_equal(got[r], want)
    with pytest.raises(ValueError):
        batch.place_frames(torch.zeros((2, 4)), range(0, 3), 5)
