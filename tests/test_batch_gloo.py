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


# ---- batch ROI statistics: the two-pass moment sums of pb_stats_batch, all-reduced after each pass ------------------
def _stat_maps(k):
    rng = np.random.default_rng(500 + k)
    rho = rng.uniform(0, 180, size=(24, 32))
    rho[rng.random(rho.shape) < 0.3] = np.nan
    psi = np.where(np.isfinite(rho), rng.uniform(40, 180, size=rho.shape), np.nan)
    inten = np.where(np.isfinite(rho), rng.uniform(1000, 4000, size=rho.shape), np.nan)
    return rho, psi, inten


def _host_pass(maps, npass, acc1, rot_fig):
    """NumPy statement of one pb_stats_batch pass over one stack's maps (what the kernel accumulates)."""
    rho, psi, inten = maps
    keep = np.isfinite(rho)
    d = np.mod(2 * (rho[keep] + rot_fig), 360) / 2
    out = np.zeros((1, 5, 4 if npass == 1 else 2))
    if npass == 1:
        out[0, :3, 0] = keep.sum()
        out[0, 0, 1], out[0, 0, 2] = np.cos(np.deg2rad(2 * d)).sum(), np.sin(np.deg2rad(2 * d)).sum()
        out[0, 1, 1], out[0, 2, 1] = psi[keep].sum(), inten[keep].sum()
    else:
        n = acc1[0, 0, 0]
        m0 = np.mod(np.degrees(np.arctan2(acc1[0, 0, 2] / n, acc1[0, 0, 1] / n)), 360) / 2
        dl = rs.wrapto180(2 * (d - m0)) / 2
        out[0, 0] = dl.sum(), (dl * dl).sum()
        for v, arr in ((1, psi), (2, inten)):
            dv = arr[keep] - acc1[0, v, 1] / n
            out[0, v] = dv.sum(), (dv * dv).sum()
    return out


def _stats_worker(rank, world, port, n_items, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    own = [_stat_maps(k) for k in batch.shard_range(n_items, rank, world)]
    acc1 = torch.from_numpy(sum((_host_pass(m, 1, None, 7.0) for m in own), np.zeros((1, 5, 4))))
    batch.allreduce_sum_(acc1)                                   # the exchange after pass 1
    acc2 = torch.from_numpy(sum((_host_pass(m, 2, acc1.numpy(), 7.0) for m in own), np.zeros((1, 5, 2))))
    batch.allreduce_sum_(acc2)                                   # ... and after pass 2
    if rank == 0:
        q.put(batch.finalize_stats(acc1, acc2, ['Rho', 'Psi']))
    dist.destroy_process_group()


def test_two_rank_batch_statistics_equal_the_concatenated_statistics():
    n_items, world = 5, 2
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_stats_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    rows = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    maps = [_stat_maps(k) for k in range(n_items)]
    res = {'vars': {'Rho': np.concatenate([m[0].ravel() for m in maps])[None], 'Psi': np.concatenate([m[1].ravel() for m in maps])[None]},
           'intmap': np.concatenate([m[2].ravel() for m in maps])[None]}
    want = rs.vecstats(res, np.ones((1,) + res['intmap'].shape, dtype=np.int32), rs.Params(rotation=(0.0, 7.0, 0.0)))
    row = rows[0]
    assert row['N'] == want['N']
    for k in ('MeanRho', 'StdRho', 'MeanDeltaRho', 'MeanPsi', 'StdPsi', 'MeanInt', 'StdInt'):
        assert rs.fmt3(float(row[k])) == rs.fmt3(float(want[k])), (k, row[k], want[k])
        assert abs(row[k] - want[k]) < 1e-9 * max(1.0, abs(want[k]))
