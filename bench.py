#!/usr/bin/env python
"""bench.py -- Mpixel-angles/s of the fused PyPOLAR analysis path on B200 (BASELINE.json's metric).

Workload (config.workload): batch mode of BASELINE.json configs[3] on one GPU's share: synthetic 1PF stacks
2048x2048x18 uint16 (generator G1), dark 480, 3x3 binning, disk-cone LUT, intensity threshold,
rho/psi/intensity maps (float32) + mask + rho/psi histograms (60 bins, from the fp64 values). A "step" is one
pass of the hot path over the rank's batch of stacks resident in HBM (batch >> L2, so no flush is needed).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--bin B] [--batch S] [--impl reference]

N > 1: launched under torchrun, one rank per GPU, stacks sharded across ranks (weak scaling: fixed per-GPU
batch), no data-path collective; the per-rank histograms are summed with one NCCL all-reduce per step,
inside the timed region. Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

H = int(os.environ.get('PB_BENCH_H', '2048'))          # development knob only: the driver runs the 2048 x 2048 default
W = int(os.environ.get('PB_BENCH_W', str(H)))
N = 18
WORKLOADS = {'1pf': ('1PF', 18), 'shg': ('SHG', 36), '4polar3d': ('4POLAR 3D', 4), '1pf90': ('1PF', 90)}   # BASELINE configs[3] (headline), [1], [2]
DISK = ROOT / 'tests' / 'golden' / 'calib' / 'Disk_Ga0_Pa0_Ta0_Gb0_Pb0_Tb0_Gc0_Pc0_Tc0.npz'
DARK = 480.0
ILOW_PERCENTILE = 20.0        # SURVEY 8d generator G1: threshold at the 20th percentile of the total-intensity image ('{:.0f}'-rounded)
OUT_BYTES_PER_PIXEL = 13      # rho f32 + psi f32 + intmap f32 + mask u8


def g1_ilow(intensity: np.ndarray) -> float:
    """ILow = 20th percentile of the intensity image, rounded as the reference GUI formats it (PP:644, PP:2148)."""
    return float('{:.0f}'.format(np.percentile(np.asarray(intensity, dtype=np.float64), ILOW_PERCENTILE)))


def digest(res_maps):
    """Order-independent-free fingerprint of one stack's results: (valid pixels, crc of the mask, crc of each f32 map with NaN -> -1)."""
    import zlib
    out = []
    for m in res_maps:
        a = np.ascontiguousarray(m)
        if a.dtype.kind == 'f':
            a = np.nan_to_num(a.astype(np.float32), nan=-1.0)
        out.append(zlib.crc32(a.tobytes()))
    return out


def measured_peak_gbs():
    f = ROOT / 'MEASURED_PEAKS.json'
    if f.exists():
        return float(json.loads(f.read_text())['hbm_gbs']), 'measured'
    return 6650.0, 'fallback'


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md recipe)."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k] == 'Active' for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace('.', '').isdigit()]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'power_w_max': max(pw) if pw else None, 'samples': len(sm)}


# ---- CPU side: the oracle restatement (== the reference's NumPy/SciPy statements) ---------------------------
# Only this leg (cpu_baseline / --impl reference) may execute oracle/. One worker process per host core, capped by RAM
# (2.8 GB peak RSS per 2048x2048x18 stack); the pool stays alive across steps; the timed work per stack is
# what the GPU arm's step produces: threshold image, maps, mask and the rho/psi histograms (no statistics rows).
_CPU = {}


def _cpu_prepare(args):
    """Worker side, untimed: generate this worker's stacks and load the calibration table."""
    seeds, h, w, n = args
    from oracle import restate as rs
    from polarimetry_b200 import synth
    d = np.load(DISK)
    _CPU['cal'] = rs.Calib.from_disk(d['RoTest'], d['PsiTest'], int(d['NbMapValues']))
    _CPU['stacks'] = [(sd, synth.stack_1pf(h, w, n, seed=sd)) for sd in seeds]
    rs.analyze_stack(synth.stack_1pf(64, 64, n, seed=0), rs.Params(method='1PF', dark=DARK, bins=(3, 3), ilow=1000.0), _CPU['cal'], want_stats=False)
    return len(seeds)


def _cpu_run(args):
    """Worker side, timed: the reference's analysis path (oracle/restate.py) on every stack of this worker."""
    bins, ilow, want_digest = args
    from oracle import restate as rs
    out = []
    for sd, v in _CPU['stacks']:
        r = rs.analyze_stack(v, rs.Params(method='1PF', dark=DARK, bins=bins, ilow=ilow), _CPU['cal'], want_stats=False)
        rec = {'seed': sd, 'valid': int(r['mask'].sum())}
        if want_digest:
            rec['hist'] = [r['hist'][('Rho', None)].tolist(), r['hist'][('Psi', None)].tolist()]
            rec['crc'] = digest([r['mask'].astype(np.uint8), r['vars']['Rho'], r['vars']['Psi'], r['intmap']])
        out.append(rec)
    return out


def host_workers(bytes_per_worker: float = 3.2e9):
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    try:
        avail = next(int(l.split()[1]) * 1024 for l in open('/proc/meminfo') if l.startswith('MemAvailable'))
        n = min(n, max(1, int(0.7 * avail / bytes_per_worker)))
    except Exception:
        pass
    return max(1, min(n, 32))


class CpuArm:
    """A pool of worker processes holding pre-generated stacks (seeds 0, 1, ... : worker k owns seeds [k*per, (k+1)*per))."""

    def __init__(self, workers: int, per_worker: int, n: int = 18, h: int = H, w: int = W):
        import multiprocessing as mp
        self.workers, self.per, self.shape = workers, per_worker, (n, h, w)
        self.pool = mp.get_context('spawn').Pool(workers)
        jobs = [([k * per_worker + j for j in range(per_worker)], h, w, n) for k in range(workers)]
        self.pool.map(_cpu_prepare, jobs, chunksize=1)

    def step(self, bins, ilow, want_digest=False):
        """One pass over every worker's stacks: (Mpx-angles/s, seconds, per-stack records)."""
        t0 = time.perf_counter()
        recs = self.pool.map(_cpu_run, [(bins, ilow, want_digest)] * self.workers, chunksize=1)
        dt = time.perf_counter() - t0
        flat = [r for w in recs for r in w]
        n, h, w = self.shape
        return len(flat) * h * w * n / dt / 1e6, dt, flat

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_ilow(bins, n: int = 18) -> float:
    """G1 threshold from the oracle's threshold image of the seed-0 stack (the reference arm must not touch the GPU)."""
    from oracle import restate as rs
    from polarimetry_b200 import synth
    return g1_ilow(rs.get_intensity(synth.stack_1pf(H, W, n, seed=0), DARK, bins))


def run_reference_arm(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    if a.workload != '1pf':
        print(json.dumps({'impl': 'reference', 'unavailable': 'the CPU arm times the headline 1pf workload only'}))
        return
    bins = (a.bin, a.bin)
    ilow = a.ilow if a.ilow is not None else cpu_ilow(bins)
    workers = host_workers()
    arm = CpuArm(workers, 1)
    vals, valid = [], 0
    for s in range(a.warmup + a.steps):
        v, dt, recs = arm.step(bins, ilow)
        valid = sum(r['valid'] for r in recs) / (len(recs) * H * W)
        if s >= a.warmup:
            vals.append((v, dt))
    arm.close()
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals]) * 1e3)
    sample = (f'{workers} stacks {H}x{W}x18 uint16 per step, one per worker process, {workers} processes kept alive across steps '
              f'(oracle/restate.py = the reference\'s NumPy/SciPy statements: threshold image, maps, mask, rho/psi histograms)')
    line = {'impl': 'reference', 'metric': 'Mpixel-angles/s', 'value': value, 'unit': 'Mpx-angles/s', 'n_gpus': a.gpus, 'steps': a.steps,
            'warmup': a.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': workload_config(a, ilow), 'stacks_per_step': workers, 'valid_pixel_fraction': round(valid, 4),
            'cpu_baseline': {'value': value, 'unit': 'Mpx-angles/s', 'cores': workers, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': 'Mpx-angles/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    print(json.dumps(line))


def workload_config(a, ilow):
    """The workload both arms run -- identical in the two JSON lines (per-run facts such as the resident batch live outside)."""
    wl = getattr(a, 'workload', '1pf')
    text = {'1pf': f'1PF batch of {H}x{W}x18 uint16 stacks (BASELINE configs[3] per-GPU share; generator G1), dark {DARK:.0f}, bin {a.bin}x{a.bin}, '
                   f'disk-cone LUT (no distortion), threshold at the 20th percentile of the intensity image, rho/psi/intensity maps f32 + mask + rho/psi histograms (60 bins)',
            'shg': f'SHG batch of {H}x{W}x36 uint16 stacks (BASELINE configs[1]), bin {a.bin}x{a.bin}, rho/S2/S4/S_SHG/intensity maps f32 + mask + histograms',
            '4polar3d': f'4POLAR 3D batch of 4x{H}x{W} uint16 (BASELINE configs[2]), bin {a.bin}x{a.bin}, rho/psi/eta/intensity maps f32 + mask + histograms',
            '1pf90': f'1PF batch of {H}x{W}x90 uint16 stacks (the per-frame shape of BASELINE configs[4] at 1/16 of its field), bin {a.bin}x{a.bin}, maps f32 + mask + histograms'}[wl]
    return {'workload': text, 'ilow': ilow, 'bin': [a.bin, a.bin], 'l2': 'inputs larger than L2 (no flush needed)',
            'precision_mode': 'exact (fp64 replay, bit-exact bins)'}


def pin_rank_cpus(local: int, world: int):
    """Give each rank of one box its own slice of the CPUs this process may use (the copy engines' staging threads and the Python
    driver of a rank then stay put). The GPU boxes of this pool are VMs with one visible NUMA node (SCALE_r01 topology: every GPU
    reports CPU affinity 0-31, NUMA 0, GPU NUMA ID N/A), so there is no per-GPU NUMA node to bind to."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
        if world > 1 and len(cpus) >= 2 * world:
            per = len(cpus) // world
            os.sched_setaffinity(0, cpus[local * per:(local + 1) * per])
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def device_g1_frame(frame, seed: int, dark: int = 480):
    """Generator G1 (the formulas of polarimetry_b200.synth.stack_1pf) on the device, plane by plane, for frames too
    large to draw on the host: fills frame [N,H,W] (int16 storage of the uint16 bit pattern)."""
    import math
    import torch
    N, H, W = frame.shape
    dev = frame.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    x = torch.arange(W, device=dev, dtype=torch.float32)[None, :]
    y = torch.arange(H, device=dev, dtype=torch.float32)[:, None]
    rho0 = torch.remainder(180.0 * x / W, 180.0) * (math.pi / 180.0)
    mod = 0.05 + 0.9 * y / H
    I0 = 2000.0 + 1500.0 * torch.sin(x / 37.0) * torch.cos(y / 53.0)
    for k in range(N):
        lam = I0 * (1.0 + mod * torch.cos(2.0 * (-math.pi * k / N - rho0)))
        v = torch.poisson(lam.clamp_(min=0.0), generator=g) + float(dark)
        frame[k] = v.clamp_(0, 65535).to(torch.int32).to(torch.int16)
        del lam, v


def run_timeseries(a):
    """BASELINE configs[4]: large-field time series, 64 frames of 8192x8192x90 uint16, bin 5x5, per-frame rho/psi histograms.
    Frames are independent: sharded over the ranks (polarimetry_b200.batch.TimeSeriesAnalyzer), every frame keeps its own
    histograms, ONE all-reduce of the [n_frames, groups, vars, nbins] buffer per step gives every rank the whole series.
    Weak scaling: each GPU holds `--frames-per-gpu` resident frames (8 = its share of the 64 at 8 GPUs)."""
    import torch
    import torch.distributed as dist
    from polarimetry_b200 import AnalysisParams, Calibration, Engine
    from polarimetry_b200.batch import TimeSeriesAnalyzer, shard_range
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    ncpu = pin_rank_cpus(local, world)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)
    eng = Engine(local)
    F, N, Hf = a.frames_per_gpu, 90, a.field
    n_frames = F * world
    own = shard_range(n_frames, rank, world)
    d = np.load(DISK)
    cal = Calibration(d['RoTest'], d['PsiTest'], int(d['NbMapValues']))
    frames = torch.empty((F, N, Hf, Hf), dtype=torch.int16, device=dev)
    for k, f in enumerate(own):
        device_g1_frame(frames[k], seed=f)
    frames = frames.view(torch.uint16)
    bins = (a.bin, a.bin)
    ilow = a.ilow
    if ilow is None:       # G1 threshold from the threshold image of frame 0 (same value on every rank: frame 0 is regenerated locally)
        f0 = torch.empty((N, Hf, Hf), dtype=torch.int16, device=dev)
        device_g1_frame(f0, seed=0)
        ilow = g1_ilow(eng.get_intensity(f0.view(torch.uint16), DARK, bins).cpu().numpy())
        del f0
    p = AnalysisParams(method='1PF', dark=DARK, bins=bins, ilow=ilow, out_dtype='float32')
    ts = TimeSeriesAnalyzer(eng, p, cal)
    launches0 = eng.launch_count()
    W3 = max(a.warmup, 3)
    for _ in range(W3):
        o = ts.run(frames, n_frames, rank, world, reduce=world > 1)
    torch.cuda.synchronize()
    launches_per_step = (eng.launch_count() - launches0) // W3
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    ev[0].record()
    for s in range(a.steps):
        o = ts.run(frames, n_frames, rank, world, reduce=world > 1)
        ev[s + 1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[s].elapsed_time(ev[s + 1]) for s in range(a.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    clocks = sampler.stop() if sampler else None
    value = n_frames * Hf * Hf * N * a.steps / (total_ms * 1e-3) / 1e6
    # checks on the timed outputs: every frame of the series has its histogram (also the other ranks' frames after the reduction),
    # each frame's rho histogram counts its valid pixels, and the local frames' counts equal their masks
    hf = o['hist_frames'].cpu().numpy()
    assert hf.shape[0] == n_frames and all(int(hf[f, 0, 0].sum()) > 0 for f in range(n_frames))
    for k, f in enumerate(own):
        assert int(hf[f, 0, 0].sum()) == int(o['mask'][k].sum().item()), f
    valid_frac = float(o['mask'].float().mean().item())
    peak, peak_src = measured_peak_gbs()
    alg_bytes = F * (Hf * Hf * N * 2 + Hf * Hf * OUT_BYTES_PER_PIXEL)
    kern_ms = float(np.median(step_ms))
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': None, 'peak_source': peak_src,
                'kernel': o['plan'], 'algorithmic_bytes_per_launch': alg_bytes, 'note': 'one launch = this rank\'s resident frames; duration = median CUDA-event step time'}
    # e2e: one frame per call from pinned host memory (12.1 GB in), maps + mask + per-frame histograms back
    e2e = None
    if not a.no_e2e:
        host = frames[:1].view(torch.int16).cpu().pin_memory().view(torch.uint16)
        hout = {}
        eng.analyze_host(host, p, cal, out=hout)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        reps = 2
        t0 = time.perf_counter()
        for _ in range(reps):
            ho = eng.analyze_host(host, p, cal, out=hout)
            _ = int(ho['hist'].sum())
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {'value': world * Hf * Hf * N * reps / float(te.item()) / 1e6, 'unit': 'Mpx-angles/s', 'h2d_bytes_per_step': Hf * Hf * N * 2,
               'd2h_bytes_per_step': Hf * Hf * OUT_BYTES_PER_PIXEL + 2 * 60 * 8, 'stacks_per_step': 1,
               'api': 'Engine.analyze_host -> pb_analyze_host, one frame per call'}
    if rank == 0:
        cfg = {'workload': f'1PF time series (BASELINE configs[4]): {n_frames} frames of {Hf}x{Hf}x{N} uint16 (generator G1 on the device), dark {DARK:.0f}, '
                           f'bin {a.bin}x{a.bin}, disk-cone LUT, G1 threshold, rho/psi/intensity maps f32 + mask + per-frame rho/psi histograms (60 bins)',
               'ilow': ilow, 'bin': [a.bin, a.bin], 'l2': 'inputs larger than L2 (no flush needed)', 'precision_mode': 'exact (fp64 replay, bit-exact bins)'}
        print(json.dumps({'metric': 'Mpixel-angles/s', 'value': value, 'unit': 'Mpx-angles/s', 'n_gpus': world, 'steps': a.steps, 'warmup': W3,
                          'ms_per_step': total_ms / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
                          'data': 'synthetic', 'config': cfg, 'frames_per_gpu': F, 'frames_total': n_frames, 'valid_pixel_fraction': round(valid_frac, 4),
                          'plan': o['plan'], 'roofline': roofline, 'cpu_baseline': None, 'e2e': e2e, 'gpu_launches': int(launches_per_step * a.steps),
                          'collective': 'one all_reduce(SUM) of the [frames, groups, vars, nbins] int64 buffer per step' if world > 1 else None,
                          'cpus_per_rank': ncpu, 'clocks': clocks}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--bin', type=int, default=int(os.environ.get('PB_BENCH_BIN', '3')))
    ap.add_argument('--batch', type=int, default=int(os.environ.get('PB_BENCH_BATCH', '32')), help='stacks resident per GPU')
    ap.add_argument('--e2e-batch', type=int, default=16, help='stacks per host-buffer call of the e2e leg (8: 0.90, 16: 0.92, 32: 0.94 of the pinned-H2D ceiling)')
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--ilow', type=float, default=None, help='intensity threshold (1pf workloads); default = SURVEY G1 (20th percentile of the '
                                                             'intensity image, ~80 %% of the pixels valid); 0 keeps every pixel')
    ap.add_argument('--workload', default='1pf', choices=sorted(WORKLOADS) + ['timeseries'],
                    help='1pf = the headline (BASELINE configs[3]); shg / 4polar3d = configs[1] / [2]; timeseries = configs[4] (use --bin 5)')
    ap.add_argument('--frames-per-gpu', type=int, default=8, help='timeseries: resident frames per GPU (8 = one GPU\'s share of the 64 frames at 8 GPUs)')
    ap.add_argument('--field', type=int, default=8192, help='timeseries: frame height = width')
    ap.add_argument('--no-e2e', action='store_true', help='skip the host-buffer leg')
    ap.add_argument('--rho-span', type=float, default=180.0, help='1pf: orientation range of the synthetic sample in degrees (180 = generator G1; '
                                                                  '6 = an aligned fibre whose rho falls into 2-3 histogram bins: contention variant)')
    ap.add_argument('--strip', action='store_true', help='1pf bin 3: the strip-march kernel + redo pass (PB_FLAG_STRIP_MARCH; experimental, slower)')
    ap.add_argument('--hist-match', action='store_true', help='warp-aggregate the histogram updates with a match on the bin (PB_FLAG_HIST_MATCH)')
    a = ap.parse_args()
    if a.impl == 'reference':
        return run_reference_arm(a)
    if a.workload == 'timeseries':
        return run_timeseries(a)

    import torch
    import torch.distributed as dist
    from polarimetry_b200 import AnalysisParams, Calibration, Engine, synth
    from polarimetry_b200.batch import BatchAnalyzer

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    ncpu = pin_rank_cpus(local, world)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)
    eng = Engine(local)
    global N
    method, N = WORKLOADS[a.workload]
    d = np.load(DISK)
    if method == '1PF':
        cal = Calibration(d['RoTest'], d['PsiTest'], int(d['NbMapValues']))
        gen = lambda seed: synth.stack_1pf(H, W, N, seed=seed, rho_lo=40.0 if a.rho_span < 180.0 else 0.0, rho_span=a.rho_span)
        ilow = a.ilow
        if ilow is None:      # G1 threshold from the threshold image of the seed-0 stack, through the product's own seam
            ilow = g1_ilow(eng.get_intensity(gen(0), DARK, (a.bin, a.bin)).cpu().numpy())
        p = AnalysisParams(method='1PF', dark=DARK, bins=(a.bin, a.bin), ilow=ilow, out_dtype='float32', hist_match=a.hist_match, strip_march=a.strip)
    elif method == 'SHG':
        cal = None
        ilow = 40000.0
        p = AnalysisParams(method='SHG', dark=0.0, bins=(a.bin, a.bin), ilow=ilow, out_dtype='float32')
        gen = lambda seed: synth.stack_4harm(H, W, N, seed=seed)
    else:
        cal = Calibration.from_K(synth.K_TH_3D)
        ilow = 1000.0
        p = AnalysisParams(method='4POLAR 3D', dark=0.0, bins=(a.bin, a.bin), ilow=ilow, out_dtype='float32')
        gen = lambda seed: synth.stack_4polar(H, W, seed=seed, three_d=True)
    out_bpp = {'1PF': 13, 'SHG': 21, '4POLAR 3D': 17}[method]

    # synthetic stacks: a few distinct seeded G1 stacks per rank, tiled to the resident batch (19 MB/stack of host RNG work each)
    B = a.batch
    ndistinct = min(B, 4)
    base = [torch.from_numpy(gen(1000 * rank + k).view(np.int16)) for k in range(ndistinct)]
    stacks = torch.empty((B, N, H, W), dtype=torch.int16, device=dev)
    for b in range(B):
        stacks[b].copy_(base[b % ndistinct])
    stacks = stacks.view(torch.uint16)
    runner = BatchAnalyzer(eng, p, cal)
    out = runner._out
    launches0 = eng.launch_count()

    def step():
        # this rank's shard of the batch; per-ROI histograms summed over ranks with one NCCL all-reduce (no-op at N=1)
        return runner.run(stacks, reduce=world > 1)

    for _ in range(max(a.warmup, 3)):
        step()
    torch.cuda.synchronize()
    launches_per_step = (eng.launch_count() - launches0) // max(a.warmup, 3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    ev[0].record()
    for s in range(a.steps):
        step()
        ev[s + 1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[s].elapsed_time(ev[s + 1]) for s in range(a.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    clocks = sampler.stop() if sampler else None
    px_angles_step = B * H * W * N * world
    value = px_angles_step * a.steps / (total_ms * 1e-3) / 1e6
    plan = out['plan']
    valid_frac = float(out['mask'].float().mean().item())

    # roofline of the dominant kernel: algorithmic bytes of one launch / its CUDA-event duration (median step)
    peak, peak_src = measured_peak_gbs()
    alg_bytes = B * (H * W * N * 2 + H * W * out_bpp)
    kern_ms = float(np.median(step_ms))
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': None,
                'peak_source': peak_src, 'kernel': plan, 'algorithmic_bytes_per_launch': alg_bytes,
                'note': 'one launch = the whole per-GPU batch; duration = median CUDA-event step time'
                        + (' (unfused plan: the step is several launches; achieved is step-level)' if plan == 'unfused' else '')}
    tfile = ROOT / 'profiles' / 'traffic.json'
    if tfile.exists():
        try:
            tj = json.loads(tfile.read_text())
            key = f'{plan}_bin{a.bin}'
            if key in tj:
                roofline['traffic'] = tj[key]['dram_bytes_per_stack'] * B
                roofline['traffic_source'] = tj[key].get('source')
        except Exception:
            pass

    # e2e: the same metric through the host-buffer API (pinned host stacks in, maps + mask + histograms back to the host)
    Be = a.e2e_batch
    host = torch.empty((Be, N, H, W), dtype=torch.int16).pin_memory()
    for b in range(Be):
        host[b].copy_(base[b % ndistinct])
    host_u16 = host.view(torch.uint16)
    hout = {}
    for _ in range(2):
        eng.analyze_host(host_u16, p, cal, out=hout)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_steps = max(3, min(a.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ho = eng.analyze_host(host_u16, p, cal, out=hout)      # synchronous: returns when the results are in host memory
        _ = int(ho['hist'].sum())
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = Be * H * W * N * world * e2e_steps / float(te.item()) / 1e6
    # PCIe ceiling of this box for the same buffers: pinned H2D of the input batch alone (what bounds e2e)
    dprobe = torch.empty_like(host, device=dev)
    evp = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    dprobe.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    evp[0].record()
    for _ in range(3):
        dprobe.copy_(host, non_blocking=True)
    evp[1].record()
    torch.cuda.synchronize()
    h2d_gbs = 3 * host.numel() * 2 / (evp[0].elapsed_time(evp[1]) * 1e-3) / 1e9
    # ... and the same with every rank of the box copying at the same time (the host side is shared: memory, PCIe root, IOMMU)
    h2d_all = h2d_gbs
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        evp[0].record()
        for _ in range(3):
            dprobe.copy_(host, non_blocking=True)
        evp[1].record()
        torch.cuda.synchronize()
        tl = torch.tensor([evp[0].elapsed_time(evp[1])], dtype=torch.float64, device=dev)
        dist.all_reduce(tl, op=dist.ReduceOp.MAX)
        h2d_all = 3 * host.numel() * 2 / (float(tl.item()) * 1e-3) / 1e9
    del dprobe
    # batch mode of configs[3] wants the reduced histograms only: the same call without the map D2H
    hout2 = {}
    eng.analyze_host(host_u16, p, cal, out=hout2, want_maps=False)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ho = eng.analyze_host(host_u16, p, cal, out=hout2, want_maps=False)
        _ = int(ho['hist'].sum())
    th = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(th, op=dist.ReduceOp.MAX)
    e2e_hist = Be * H * W * N * world * e2e_steps / float(th.item()) / 1e6
    e2e = {'value': e2e_val, 'unit': 'Mpx-angles/s', 'h2d_bytes_per_step': Be * H * W * N * 2,
           'd2h_bytes_per_step': Be * (H * W * out_bpp) + 2 * 60 * 8, 'stacks_per_step': Be, 'h2d_pinned_gbs': round(h2d_gbs, 2),
           'h2d_pinned_gbs_per_gpu_all_ranks_copying': round(h2d_all, 2),
           'frac_of_h2d_ceiling': round(e2e_val / world * 1e6 * 2 / 1e9 / h2d_all, 3),
           'histograms_only': {'value': e2e_hist, 'd2h_bytes_per_step': 2 * 60 * 8, 'frac_of_h2d_ceiling': round(e2e_hist / world * 1e6 * 2 / 1e9 / h2d_all, 3)},
           'cpus_per_rank': ncpu,
           'api': 'Engine.analyze_host -> pb_analyze_host (pinned host buffers, 2-stream copy/compute overlap)'}

    # size-independent sanity of the timed outputs: every valid pixel has a rho inside [0, 180] -> one histogram count each
    # (the histograms were summed over the ranks by the step's all-reduce: compare with the valid pixels of all ranks)
    hist_np = out['hist'].cpu().numpy()
    nv = out['mask'].sum().to(torch.int64).reshape(1)
    if world > 1:
        dist.all_reduce(nv)
    n_valid = int(nv.item())
    assert int(hist_np[0, 0, 0].sum()) == n_valid, (int(hist_np[0, 0, 0].sum()), n_valid)

    cpu = parity = None
    if rank == 0 and world == 1 and not a.no_cpu and a.workload == '1pf':
        # cpu_baseline leg: the oracle on the GPU box's host cores, on a bounded sample of the same workload (same generator,
        # seeds 0.., same threshold), and -- the oracle acting as the checker -- the GPU results of the first distinct stacks of
        # the timed batch compared with it: histograms and mask / map checksums must be identical.
        workers = host_workers()
        arm = CpuArm(workers, 1)
        arm.step((a.bin, a.bin), ilow)                                   # warm-up pass
        v, dt, recs = arm.step((a.bin, a.bin), ilow, want_digest=True)
        arm.close()
        cpu = {'value': v, 'unit': 'Mpx-angles/s', 'cores': workers, 'kind': 'port',
               'sample': f'{len(recs)} stacks {H}x{W}x18 (one per worker process, {workers} processes) in {dt:.1f} s, oracle/restate.py (the reference\'s NumPy/SciPy statements: threshold image, maps, mask, rho/psi histograms)'}
        by_seed = {r['seed']: r for r in recs}
        ncheck = [k for k in range(ndistinct) if k in by_seed]
        ok = True
        for k in ncheck:
            g = digest([out['mask'][k].cpu().numpy(), out['vars']['Rho'][k].cpu().numpy(), out['vars']['Psi'][k].cpu().numpy(), out['intmap'][k].cpu().numpy()])
            ok = ok and g == by_seed[k]['crc'] and int(out['mask'][k].sum().item()) == by_seed[k]['valid']
        hist_ok = None
        if len(ncheck) == ndistinct and B % ndistinct == 0:
            want = (B // ndistinct) * sum(np.asarray(by_seed[k]['hist'], dtype=np.int64) for k in ncheck)
            hist_ok = bool(np.array_equal(hist_np[0, 0], want))
            ok = ok and hist_ok
        parity = {'checked_stacks': len(ncheck), 'maps_mask_crc_equal': bool(ok), 'batch_histograms_equal': hist_ok}
        assert ok, f'GPU results differ from the oracle: {parity}'
    if rank == 0:
        line = {'metric': 'Mpixel-angles/s', 'value': value, 'unit': 'Mpx-angles/s', 'n_gpus': world, 'steps': a.steps, 'warmup': max(a.warmup, 3),
                'ms_per_step': total_ms / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
                'data': 'synthetic', 'config': workload_config(a, ilow), 'stacks_per_gpu': B, 'valid_pixel_fraction': round(valid_frac, 4),
                'plan': plan, 'roofline': roofline, 'cpu_baseline': cpu, 'parity_check': parity, 'e2e': e2e,
                'gpu_launches': int(launches_per_step * a.steps), 'clocks': clocks}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
