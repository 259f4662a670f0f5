import numpy as np, torch
from polarimetry_b200 import AnalysisParams, Calibration, Engine, synth
d = np.load('tests/golden/calib/Disk_Ga0_Pa0_Ta0_Gb0_Pb0_Tb0_Gc0_Pc0_Tc0.npz')
cal = Calibration(d['RoTest'], d['PsiTest'], int(d['NbMapValues']))
eng = Engine()
v = synth.stack_1pf(70, 88, 18, seed=3)
a = eng.analyze(v, AnalysisParams(method='1PF', dark=480.0, bins=(3, 3), ilow=21000.0), cal)
torch.cuda.synchronize()
print('ok')
