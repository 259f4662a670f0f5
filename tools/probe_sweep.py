#!/usr/bin/env python
"""Diagnostic: where does the time of one analyze_batch call go (bin 3, 8 stacks) for different thresholds / disk handling."""
import sys, time, json
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from polarimetry_b200 import AnalysisParams, Calibration, Engine, synth
H = W = 2048; N = 18
eng = Engine(0)
cdir = ROOT / 'tests' / 'golden' / 'calib'
files = sorted(cdir.glob('Disk_*.npz'))
disks = [Calibration(np.load(f)['RoTest'], np.load(f)['PsiTest'], 401, name=f.stem) for f in files]
base = [torch.from_numpy(synth.stack_1pf(H, W, N, seed=500 + k).view(np.int16)) for k in range(4)]
stacks = torch.empty((8, N, H, W), dtype=torch.int16, device='cuda')
for b in range(8):
    stacks[b].copy_(base[b % 4])
stacks = stacks.view(torch.uint16)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps, 3), round((time.perf_counter() - t0) / reps * 1e3, 3)
out = {}
res = {}
for bins in (3, 1):
    for ilow in (40000.0, 23220.0):
        for d, cal in enumerate(disks):
            p = AnalysisParams(method='1PF', dark=480.0, bins=(bins, bins), ilow=ilow, out_dtype='float32')
            res[f'bin{bins}_ilow{int(ilow)}_disk{d}'] = timed(lambda: eng.analyze_batch(stacks, p, cal, want_intensity=False, out=out))
    p = AnalysisParams(method='1PF', dark=480.0, bins=(bins, bins), ilow=40000.0, out_dtype='float32')
    def alt():
        for cal in disks: eng.analyze_batch(stacks, p, cal, want_intensity=False, out=out)
    res[f'bin{bins}_alternating_4_disks_upload_each'] = timed(alt, 3)
    eng.set_diskcones(disks)
    res[f'bin{bins}_alternating_4_disks_resident'] = timed(alt, 3)
    eng._resident = {}; eng._key_calib = None
print(json.dumps(res, indent=1))
