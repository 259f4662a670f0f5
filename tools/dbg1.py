import sys, json, numpy as np, torch
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT))
from polarimetry_b200 import AnalysisParams, Calibration, Engine, synth
eng = Engine(0)
files = sorted((ROOT/'tests/golden/calib').glob('Disk_*.npz'))
H=W=2048; N=18
st = torch.from_numpy(synth.stack_1pf(H, W, N, seed=500).view(np.int16)).cuda().view(torch.uint16)[None].repeat(4,1,1,1).contiguous()
for bins in ((1,1),(3,3)):
    p = AnalysisParams(method='1PF', dark=480.0, bins=bins, ilow=40000.0, out_dtype='float32')
    for f in files:
        d = np.load(f); cal = Calibration(d['RoTest'], d['PsiTest'], int(d['NbMapValues']), name=f.stem)
        o = {}
        eng.analyze_batch(st, p, cal, want_intensity=False, out=o); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): eng.analyze_batch(st, p, cal, want_intensity=False, out=o)
        e1.record(); torch.cuda.synchronize()
        print(bins, f.stem[:30], 'ms per 4 stacks', round(e0.elapsed_time(e1)/3, 3), 'valid', float(o['mask'].float().mean()))
