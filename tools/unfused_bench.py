"""Time the unfused plan (field -> box H -> box W -> per-pixel kernel) on the configurations that have no fused kernel:
CARS with binning, fractional dark, float stacks.   python tools/unfused_bench.py"""
import json, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from polarimetry_b200 import AnalysisParams, Calibration, Engine, synth

eng = Engine()
d = np.load(Path(__file__).resolve().parent.parent / 'tests/golden/calib/Disk_Ga0_Pa0_Ta0_Gb0_Pb0_Tb0_Gc0_Pc0_Tc0.npz')
cal = Calibration(d['RoTest'], d['PsiTest'], int(d['NbMapValues']))
H = W = 2048
cases = [('CARS 2048x2048x36 bin 3x3', synth.stack_4harm(H, W, 36, seed=1), AnalysisParams(method='CARS', dark=0.0, bins=(3, 3), ilow=40000.0, out_dtype='float32'), None),
         ('SHG 2048x2048x36 bin 3x3 (fused march, for comparison)', synth.stack_4harm(H, W, 36, seed=1), AnalysisParams(method='SHG', dark=0.0, bins=(3, 3), ilow=40000.0, out_dtype='float32'), None),
         ('1PF 2048x2048x18 bin 3x3, dark 480.5 (fractional)', synth.stack_1pf(H, W, 18, seed=1), AnalysisParams(method='1PF', dark=480.5, bins=(3, 3), ilow=23000.0, out_dtype='float32'), cal),
         ('1PF 2048x2048x18 float32 stack bin 3x3', synth.stack_1pf(H, W, 18, seed=1).astype(np.float32), AnalysisParams(method='1PF', dark=480.0, bins=(3, 3), ilow=23000.0, out_dtype='float32'), cal)]
for name, v, p, c in cases:
    t = torch.from_numpy(v.view(np.int16) if v.dtype == np.uint16 else v).cuda()
    if v.dtype == np.uint16:
        t = t.view(torch.uint16)
    out = {}
    r = eng.analyze_batch(t[None], p, c, want_intensity=False, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        eng.analyze_batch(t[None], p, c, want_intensity=False, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps({'case': name, 'plan': r['plan'], 'ms_per_stack': round(ms, 3), 'mpx_angles_s': round(v.size / ms / 1e3)}))
