"""Folder mode with overlap: TIFF pages -> pinned ring -> device -> analysis, pipelined.

The reference's folder loop (PyPOLAR.py:2124-2131) is serial: `open_file` -> `define_stack` (`cv2.imreadmulti`, PP:1923-1942)
-> `set_dark` (PP:2247-2252) -> `analyze_stack`, one file after the other on one thread. Here the same statements run as a
pipeline over a small ring of page-locked staging buffers:

    decode thread(s):  cv2.imreadmulti(file, start=page, count=1, IMREAD_ANYDEPTH) page by page (the reference's reader, same
                       flags) straight into the slot's pinned [N,H,W] buffer; each page is sent with its own asynchronous copy
                       on the copy stream as soon as it is decoded
    copy stream:       the per-page cudaMemcpyAsync's of file k+1 run while ...
    compute stream:    ... file k is rescaled (float TIFFs, PP:1929-1930: pb_rescale_stack), its dark value computed
                       (Stack.compute_dark, PC:266-278: pb_compute_dark) and analysed (pb_analyze, the fused kernels)

A slot is reused only after the analysis that read its device buffer has finished (event). Results are identical to
`Engine.define_stack` + `Engine.analyze` file by file (tests/test_ingest.py). `timeline` records host time stamps of the decode
and CUDA-event times of copy / compute per file, from which the overlap is reported.
"""
from __future__ import annotations

import queue
import threading
import time
from pathlib import Path
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from .api import AnalysisParams, Calibration, Engine, Result


def _probe(file) -> tuple:
    """(n_pages, height, width, numpy dtype) of a multi-page TIFF, through the reference's reader."""
    import cv2
    n = cv2.imcount(str(file), cv2.IMREAD_ANYDEPTH)
    ok, pg = cv2.imreadmulti(str(file), 0, 1, flags=cv2.IMREAD_ANYDEPTH)
    if n < 1 or not ok or not len(pg):
        raise FileNotFoundError(f'cannot read {file}')
    return n, pg[0].shape[0], pg[0].shape[1], pg[0].dtype, pg[0]


class FolderAnalyzer:
    """Analyse a sequence of multi-page TIFF stacks with decode / H2D / compute overlapped (slots >= 2 staging buffers)."""

    def __init__(self, engine: Engine, params: AnalysisParams, calib: Optional[Calibration] = None, rois: Optional[Sequence[dict]] = None,
                 per_roi: bool = False, slots: int = 3, decode_threads: int = 2, default_dark: Optional[float] = None):
        if slots < 2:
            raise ValueError('FolderAnalyzer: at least 2 staging slots are needed for any overlap')
        self.eng, self.p, self.calib, self.rois, self.per_roi = engine, params, calib, rois, per_roi
        self.slots, self.decode_threads, self.default_dark = slots, max(1, decode_threads), default_dark
        self.timeline: List[dict] = []

    def run(self, files: Sequence, on_result: Optional[Callable[[int, Path, Result, float], None]] = None) -> List[dict]:
        """Analyse `files` in order. `on_result(index, file, result, dark)` is called on the calling thread as each file completes
        (its maps live on the device until the callback returns; the next files are being decoded / copied meanwhile).
        Returns one record per file: dict(file, dark, hist, n_valid, plan)."""
        import cv2
        eng, dev = self.eng, self.eng.device
        files = [Path(f) for f in files]
        if not files:
            return []
        copy_stream = torch.cuda.Stream(device=dev)
        compute_stream = torch.cuda.current_stream(dev)
        free_slots: 'queue.Queue[int]' = queue.Queue()
        ready: 'queue.Queue' = queue.Queue()
        slot_host = [None] * self.slots          # pinned staging, (re)allocated when a file's shape / dtype changes
        slot_dev = [None] * self.slots
        slot_done = [None] * self.slots          # event: the analysis that read slot_dev[s] has finished
        for s in range(self.slots):
            free_slots.put(s)
        t_origin = time.perf_counter()
        lock = threading.Lock()
        next_index = [0]
        pending = {}

        def decode_worker():
            torch.cuda.set_device(dev)
            while True:
                with lock:
                    # index and staging slot are taken together, in file order: a file never waits for a slot held by a later one
                    k = next_index[0]
                    if k >= len(files):
                        return
                    next_index[0] += 1
                    s = free_slots.get()
                try:
                    n, h, w, dt, first = _probe(files[k])
                    if slot_done[s] is not None:
                        slot_done[s].synchronize()           # the device buffer of this slot is no longer read
                    tdt = {np.dtype('uint8'): torch.uint8, np.dtype('uint16'): torch.int16, np.dtype('float32'): torch.float32,
                           np.dtype('float64'): torch.float64}.get(np.dtype(dt))
                    if tdt is None:
                        raise TypeError(f'{files[k]}: {dt} pages are not supported (uint8 / uint16 / float32 / float64)')
                    if slot_host[s] is None or tuple(slot_host[s].shape) != (n, h, w) or slot_host[s].dtype != tdt:
                        slot_host[s] = torch.empty((n, h, w), dtype=tdt).pin_memory()
                        slot_dev[s] = torch.empty((n, h, w), dtype=tdt, device=dev)
                    host_np = slot_host[s].numpy()
                    rec = {'index': k, 'file': str(files[k]), 'slot': s, 'decode_t0': time.perf_counter() - t_origin}
                    ev0 = torch.cuda.Event(enable_timing=True)
                    ev1 = torch.cuda.Event(enable_timing=True)
                    for page in range(n):
                        if page == 0:
                            pg = first
                        else:
                            ok, pgs = cv2.imreadmulti(str(files[k]), page, 1, flags=cv2.IMREAD_ANYDEPTH)
                            if not ok or not len(pgs):
                                raise IOError(f'{files[k]}: page {page} unreadable')
                            pg = pgs[0]
                        if pg.shape != (h, w) or pg.dtype != dt:
                            raise ValueError(f'{files[k]}: page {page} differs in shape / type from page 0')
                        host_np[page] = pg.view(np.int16) if dt == np.uint16 else pg
                        with torch.cuda.stream(copy_stream):
                            if page == 0:
                                ev0.record(copy_stream)
                            slot_dev[s][page].copy_(slot_host[s][page], non_blocking=True)       # one cudaMemcpyAsync per page
                    with torch.cuda.stream(copy_stream):
                        ev1.record(copy_stream)
                    rec['decode_t1'] = time.perf_counter() - t_origin
                    rec.update(h2d_ev=(ev0, ev1), shape=(n, h, w), dtype=str(np.dtype(dt)))
                    ready.put((k, rec, None))
                except Exception as e:                                             # surfaces on the calling thread
                    free_slots.put(s)
                    ready.put((k, None, e))

        workers = [threading.Thread(target=decode_worker, daemon=True) for _ in range(min(self.decode_threads, self.slots - 1, len(files)))]
        for t in workers:
            t.start()
        records, timeline = [], []
        ev_origin = torch.cuda.Event(enable_timing=True)
        ev_origin.record(compute_stream)
        for want in range(len(files)):
            while want not in pending:
                k, rec, err = ready.get()
                if err is not None:
                    raise err
                pending[k] = rec
            rec = pending.pop(want)
            s = rec['slot']
            ev0, ev1 = rec['h2d_ev']
            compute_stream.wait_event(ev1)                                     # the pages of this file are on the device
            c0 = torch.cuda.Event(enable_timing=True)
            c1 = torch.cuda.Event(enable_timing=True)
            c0.record(compute_stream)
            values = slot_dev[s]
            if values.dtype == torch.int16:
                values = values.view(torch.uint16)
            st = eng.define_stack(values, default_dark=self.default_dark)       # rescale (float pages) + dark, on the device
            if self.default_dark is None:
                # set_dark (PP:2247-2252): the GUI's dark entry receives the calculated dark through stack.display = '{:.0f}',
                # and analyze_stack reads it back with float() (PP:2951)
                shown = '{:.0f}'.format(st['dark'])
                st['dark'] = float(shown) if shown not in ('nan', 'inf', '-inf') else 0.0
            p = AnalysisParams(**{**self.p.__dict__, 'dark': st['dark']}) if self.default_dark is None else self.p
            res = eng.analyze(st['values'], p, self.calib, self.rois, self.per_roi)
            c1.record(compute_stream)
            done = torch.cuda.Event()
            done.record(compute_stream)
            if on_result is not None:
                on_result(want, files[want], res, st['dark'])
            slot_done[s] = done
            free_slots.put(s)
            records.append({'file': str(files[want]), 'dark': st['dark'], 'hist': res.hist, 'n_valid': int(res.mask.sum().item()), 'plan': res.plan})
            timeline.append((rec, c0, c1))
        for t in workers:
            t.join()
        torch.cuda.synchronize(dev)
        self.timeline = []
        for rec, c0, c1 in timeline:
            ev0, ev1 = rec.pop('h2d_ev')
            self.timeline.append({**rec, 'h2d_ms': (ev_origin.elapsed_time(ev0), ev_origin.elapsed_time(ev1)),
                                  'compute_ms': (ev_origin.elapsed_time(c0), ev_origin.elapsed_time(c1)),
                                  'decode_ms': (rec['decode_t0'] * 1e3, rec['decode_t1'] * 1e3)})
        return records

    def overlap_report(self) -> dict:
        """Busy time of each stage and how much of the copy / compute time ran while another file was being decoded."""
        if not self.timeline:
            return {}

        def union(iv):
            iv = sorted(iv)
            tot, cur0, cur1 = 0.0, None, None
            for a, b in iv:
                if cur1 is None or a > cur1:
                    if cur1 is not None:
                        tot += cur1 - cur0
                    cur0, cur1 = a, b
                else:
                    cur1 = max(cur1, b)
            return tot + (cur1 - cur0 if cur1 is not None else 0.0)

        def inter(iv1, iv2):
            tot = 0.0
            for a, b in iv1:
                for c, d in iv2:
                    tot += max(0.0, min(b, d) - max(a, c))
            return tot

        dec = [t['decode_ms'] for t in self.timeline]
        h2d = [t['h2d_ms'] for t in self.timeline]
        cmp_ = [t['compute_ms'] for t in self.timeline]
        # NB: the event clock starts at the first compute-stream record, the host clock at run(): both within a millisecond
        wall = max(max(b for _, b in dec), max(b for _, b in cmp_)) - min(a for a, _ in dec)
        return {'files': len(self.timeline), 'wall_ms': wall, 'decode_busy_ms': union(dec), 'h2d_busy_ms': union(h2d), 'compute_busy_ms': union(cmp_),
                'compute_overlapped_with_decode_ms': inter(cmp_, dec), 'h2d_overlapped_with_decode_ms': inter(h2d, dec),
                'serial_sum_ms': sum(b - a for a, b in dec) + sum(b - a for a, b in cmp_)}
