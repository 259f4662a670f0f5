"""Debug driver: one small 1PF 18-angle 3x3 analysis through the strip-march kernel (run under compute-sanitizer)."""
import sys
import numpy as np
import torch
from polarimetry_b200 import AnalysisParams, Calibration, Engine, synth

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (70, 88)
d = np.load('tests/golden/calib/Disk_Ga0_Pa0_Ta0_Gb0_Pb0_Tb0_Gc0_Pc0_Tc0.npz')
cal = Calibration(d['RoTest'], d['PsiTest'], int(d['NbMapValues']))
eng = Engine()
v = synth.stack_1pf(H, W, 18, seed=3)
kw = dict(method='1PF', dark=480.0, bins=(3, 3), ilow=21000.0)
a = eng.analyze(v, AnalysisParams(**kw), cal)
torch.cuda.synchronize()
b = eng.analyze(v, AnalysisParams(force_unfused=True, **kw), cal)
torch.cuda.synchronize()
for va, vb in zip(a.vars, b.vars):
    x, y = torch.nan_to_num(va.values, nan=-1.0), torch.nan_to_num(vb.values, nan=-1.0)
    print(va.name, 'equal', bool(torch.equal(x, y)), 'mismatches', int((x != y).sum()))
    if not torch.equal(x, y):
        idx = (x != y).nonzero()[:10]
        print(idx.tolist(), x[x != y][:5].tolist(), y[x != y][:5].tolist())
print('mask equal', bool(torch.equal(a.mask, b.mask)), int(a.mask.sum()), int(b.mask.sum()))
print('hist equal', all((a.hist[k] == b.hist[k]).all() for k in a.hist))
o = eng.analyze_batch(v[None], AnalysisParams(out_dtype='float32', **kw), cal, want_intensity=False)
torch.cuda.synchronize()
r = eng.analyze(v, AnalysisParams(force_unfused=True, out_dtype='float32', **kw), cal)
for var in r.vars:
    x, y = torch.nan_to_num(o['vars'][var.name][0], nan=-1.0), torch.nan_to_num(var.values, nan=-1.0)
    print('fast', var.name, 'equal', bool(torch.equal(x, y)), 'mismatches', int((x != y).sum()))
print('fast mask equal', bool(torch.equal(o['mask'][0].bool(), r.mask.bool())), 'hist', bool((o['hist'][0, 0, 0].cpu().numpy() == r.hist[('Rho', None)]).all()))
