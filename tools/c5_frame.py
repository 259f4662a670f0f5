#!/usr/bin/env python
"""BASELINE configs[4] at FULL frame size: one 8192x8192x90 uint16 frame (12.1 GB), bin 5x5, rho/psi/intensity maps + mask +
per-frame histograms, on one B200. Checks (size-independent properties, the oracle cannot run this size):
  * rows [y0, y0+R) of the full-frame maps are bit-identical to the analysis of the row crop [y0-8, y0+R+8) (full width: the W pass of
    the box filter carries its history from column 0, the H pass reaches 2 rows) -- at the top edge, in the middle (beyond the
    2^32-element offset) and at the bottom edge;
  * the per-frame histogram equals the histogram of the rho/psi values recomputed from the written maps' mask count.
Prints one JSON line.   python tools/c5_frame.py [--field 8192] [--angles 90] [--bin 5]"""
import argparse, json, math, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from polarimetry_b200 import AnalysisParams, Calibration, Engine

def run(field: int = 8192, angles: int = 90, bin: int = 5, reps: int = 3) -> dict:
    H = W = field; N = angles
    dev = torch.device('cuda', 0)
    eng = Engine(0)
    d = np.load(ROOT / 'tests' / 'golden' / 'calib' / 'Disk_Ga0_Pa0_Ta0_Gb0_Pb0_Tb0_Gc0_Pc0_Tc0.npz')
    cal = Calibration(d['RoTest'], d['PsiTest'], int(d['NbMapValues']))

    # generator G1 on the device, plane by plane
    torch.manual_seed(7)
    x = torch.arange(W, device=dev, dtype=torch.float32)[None, :]
    y = torch.arange(H, device=dev, dtype=torch.float32)[:, None]
    rho0 = torch.remainder(180.0 * x / W, 180.0) * (math.pi / 180.0)
    mod = 0.05 + 0.9 * y / H
    I0 = 2000.0 + 1500.0 * torch.sin(x / 37.0) * torch.cos(y / 53.0)
    frame = torch.empty((N, H, W), dtype=torch.int16, device=dev)
    for k in range(N):
        al = -math.pi * k / N
        lam = I0 * (1.0 + mod * torch.cos(2.0 * (al - rho0)))
        v = torch.poisson(lam.clamp_(min=0.0)) + 480.0
        frame[k] = v.clamp_(0, 65535).to(torch.int32).to(torch.int16)       # the uint16 bit pattern
        del lam, v
    frame = frame.view(torch.uint16)
    p = AnalysisParams(method='1PF', dark=480.0, bins=(bin, bin), ilow=40000.0 * N / 18.0, out_dtype='float32')
    out = {}
    eng.analyze_batch(frame[None], p, cal, want_intensity=False, out=out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for r in range(reps):
        eng.analyze_batch(frame[None], p, cal, want_intensity=False, out=out)
        ev[r + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[r].elapsed_time(ev[r + 1]) for r in range(reps))[reps // 2]
    full = {k: out['vars'][k][0].clone() for k in ('Rho', 'Psi')}
    full['intmap'] = out['intmap'][0].clone(); full['mask'] = out['mask'][0].clone()
    hist = out['hist'].clone()
    valid = float(full['mask'].float().mean())

    R, M = min(256, H // 4), 8
    checks = {}
    for name, y0 in (('top', 0), ('middle', H // 2 + 37), ('bottom', H - R)):
        lo, hi = max(0, y0 - M), min(H, y0 + R + M)
        crop = frame[:, lo:hi, :].contiguous()
        o2 = {}
        eng.analyze_batch(crop[None], p, cal, want_intensity=False, out=o2)
        torch.cuda.synchronize()
        ok = True
        for k in ('Rho', 'Psi'):
            ok &= torch.equal(o2['vars'][k][0][y0 - lo:y0 - lo + R].view(torch.int32), full[k][y0:y0 + R].view(torch.int32))
        ok &= torch.equal(o2['intmap'][0][y0 - lo:y0 - lo + R].view(torch.int32), full['intmap'][y0:y0 + R].view(torch.int32))
        ok &= torch.equal(o2['mask'][0][y0 - lo:y0 - lo + R], full['mask'][y0:y0 + R])
        checks[name] = bool(ok)
        del crop, o2
    # histogram self-consistency: every valid pixel has a finite rho in [0,180): the rho counts sum to the valid pixels
    rho = full['Rho'][full['mask'].bool()]
    h_rho = int(hist.reshape(-1, 2, p.nbins)[0, 0].sum())
    checks['hist_rho_total_equals_valid'] = h_rho == int(rho.numel())
    alg_bytes = H * W * N * 2 + H * W * 13
    res = {'workload': f'1PF time-series frame {H}x{W}x{N} uint16 (BASELINE configs[4] frame), bin {bin}x{bin}, f32 maps + mask + per-frame histograms',
           'plan': out['plan'], 'ms_per_frame': ms, 'mpx_angles_per_s': H * W * N / (ms * 1e-3) / 1e6, 'algorithmic_gbs': alg_bytes / (ms * 1e-3) / 1e9,
           'valid_pixel_fraction': round(valid, 4), 'input_gb': H * W * N * 2 / 1e9, 'checks': checks,
           'all_ok': all(checks.values()), 'frames64_s': 64 * ms * 1e-3}
    eng.close()
    del frame, full, out
    torch.cuda.empty_cache()
    return res


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--field', type=int, default=8192); ap.add_argument('--angles', type=int, default=90); ap.add_argument('--bin', type=int, default=5)
    ap.add_argument('--reps', type=int, default=3)
    a = ap.parse_args()
    print(json.dumps(run(a.field, a.angles, a.bin, a.reps)))
