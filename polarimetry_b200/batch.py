"""Batch / time-series mode: many independent stacks (or frames) sharded over the GPUs of one box.

One process per GPU (torchrun). Stacks are independent, so the data path needs no collective: each rank
analyses its shard with its own `Engine`; only the per-ROI histograms (and the moment sums of the ROI
statistics) are summed across ranks, with one all-reduce per batch over NCCL / NVLink. That sum is what the
reference's `merge_histos` (PyPOLAR.py:2318-2384) computes when it concatenates the per-file values and
histograms them once with the same bin edges.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous block partition of range(n_items): the first n_items % world ranks get one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError('bad rank / world')
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def allreduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM over the ranks (no-op without an initialised process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def circular_moments(rho_deg: torch.Tensor, sel: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[N, sum cos 2rho, sum sin 2rho] in float64: the additive sufficient statistics of circularmean (PC:55-56)."""
    r = rho_deg.to(torch.float64)
    ok = torch.isfinite(r) if sel is None else (torch.isfinite(r) & sel.bool())
    a = torch.deg2rad(2.0 * r[ok])
    return torch.stack([ok.sum().to(torch.float64), torch.cos(a).sum(), torch.sin(a).sum()])


def circular_mean_from_moments(m: torch.Tensor) -> float:
    """mod(angle(mean(exp(2j rho))), 360) / 2 in degrees, from (reduced) moments."""
    n, c, s = (float(x) for x in m)
    if n == 0:
        return float('nan')
    return float(np.mod(np.degrees(np.arctan2(s / n, c / n)), 360.0) / 2.0)


def finalize_stats(acc1, acc2, names: Sequence[str]) -> list:
    """Accumulators of pb_stats_batch (summed over the ranks) -> per histogram group the numeric columns of return_vecexcel
    (PP:2739-2779): N, MeanRho (circularmean, PC:55-56), StdRho / MeanDeltaRho (of wrapto180(2(d - mean))/2), Mean / Std of the
    other variables and of the intensity. acc1 [G, V+1, 4] = N, sum d | sum cos 2d, sum sin 2d; acc2 [G, V+1, 2] = sum delta,
    sum delta^2 (np.std is the population standard deviation: sqrt(mean(delta^2) - mean(delta)^2) with delta = d - mean)."""
    a1 = acc1.detach().cpu().numpy() if isinstance(acc1, torch.Tensor) else np.asarray(acc1)
    a2 = acc2.detach().cpu().numpy() if isinstance(acc2, torch.Tensor) else np.asarray(acc2)
    rows = []
    for g in range(a1.shape[0]):
        n = float(a1[g, 0, 0])
        row = {'N': int(n)}
        with np.errstate(invalid='ignore', divide='ignore'):
            for v, name in enumerate(list(names) + ['Int']):
                s1, s2 = a1[g, v, 1], a1[g, v, 2]
                if v == 0:
                    mean = float(np.mod(np.degrees(np.arctan2(s2 / n, s1 / n)), 360.0) / 2.0) if n else float('nan')
                else:
                    mean = float(s1 / n) if n else float('nan')
                md = a2[g, v, 0] / n if n else float('nan')
                var = a2[g, v, 1] / n - md * md if n else float('nan')
                std = float(np.sqrt(max(var, 0.0))) if n else float('nan')
                row['Mean' + name], row['Std' + name] = mean, std
                if v == 0:
                    row['MeanDeltaRho'] = float(md)
        rows.append(row)
    return rows


class BatchAnalyzer:
    """Analyse this rank's shard of a batch and reduce the histograms over all ranks."""

    def __init__(self, engine, params, calib=None, rois: Optional[Sequence[dict]] = None, per_roi: bool = False, group=None):
        self.engine, self.params, self.calib, self.rois, self.per_roi, self.group = engine, params, calib, rois, per_roi, group
        self._out = {}

    def run(self, stacks_local, reduce: bool = True) -> dict:
        """stacks_local: this rank's [B,N,H,W] stacks (device tensor). Returns the engine's output dict; o['hist'] holds the
        histograms summed over the batch and, if reduce, over all ranks (int64 [1, n_groups, n_vars, nbins])."""
        o = self.engine.analyze_batch(stacks_local, self.params, self.calib, self.rois, self.per_roi, want_intensity=False, out=self._out)
        if reduce:
            allreduce_sum_(o['hist'], self.group)
        return o

    def stats(self, o: dict, reduce: bool = True) -> list:
        """ROI statistics of the WHOLE batch (all ranks' stacks, as if their values were concatenated: the reference's merge
        semantics): moment sums on the device, one small all-reduce after each of the two passes, finalised on the host."""
        from .api import VAR_NAMES
        fn = (lambda t: allreduce_sum_(t, self.group)) if reduce else None
        acc1, acc2 = self.engine.batch_stats_accumulate(o, self.params, fn)
        return finalize_stats(acc1, acc2, VAR_NAMES[self.params.method])


def place_frames(hist_local: torch.Tensor, own: range, n_frames: int) -> torch.Tensor:
    """Per-frame results of this rank's frames `own` in a [n_frames, ...] buffer that is zero for the frames of the other ranks:
    the SUM all-reduce of these buffers is the whole series."""
    if hist_local.shape[0] != len(own):
        raise ValueError(f'{hist_local.shape[0]} local frames for a shard of {len(own)}')
    full = torch.zeros((n_frames,) + tuple(hist_local.shape[1:]), dtype=hist_local.dtype, device=hist_local.device)
    if len(own):
        full[own.start:own.stop] = hist_local
    return full


class TimeSeriesAnalyzer:
    """Time-series mode (BASELINE configs[4]): the frames of a series are sharded over the ranks (shard_range); every frame keeps its
    own histograms; one all-reduce of the [n_frames, groups, vars, nbins] buffer gives every rank the whole series."""

    def __init__(self, engine, params, calib=None, rois: Optional[Sequence[dict]] = None, per_roi: bool = False, group=None):
        self.engine, self.params, self.calib, self.rois, self.per_roi, self.group = engine, params, calib, rois, per_roi, group
        self._out = {}          # output maps are allocated once and reused by every call on frames of the same shape

    def run(self, frames_local, n_frames: int, rank: int, world: int, reduce: bool = True) -> dict:
        """frames_local: this rank's frames [len(shard_range(n_frames, rank, world)), N, H, W] (device tensor). Returns the engine's
        output dict (maps of the local frames) with o['hist_frames']: int64 [n_frames, n_groups, n_vars, nbins]."""
        own = shard_range(n_frames, rank, world)
        if self._out and tuple(self._out['mask'].shape) != (frames_local.shape[0],) + tuple(frames_local.shape[-2:]):
            self._out = {}
        o = self.engine.analyze_batch(frames_local, self.params, self.calib, self.rois, self.per_roi, want_intensity=False,
                                      hist_per_stack=True, out=self._out)
        full = place_frames(o['hist'], own, n_frames)
        if reduce:
            allreduce_sum_(full, self.group)
        o['hist_frames'] = full
        return o
