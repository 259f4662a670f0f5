#!/usr/bin/env python
"""Folder mode: a folder of multi-page uint16 TIFF stacks analysed file by file, (a) serially as the reference's
loop does (PP:2124-2131: decode -> copy -> dark -> analysis, one after the other) and (b) with FolderAnalyzer's pipeline
(page-wise decode into a pinned ring, per-page asynchronous copies, analysis of file k overlapping decode / copy of file k+1).
Prints one JSON line with both rates and the pipeline's stage timeline.   python tools/folder_bench.py [--files 12] [--size 2048]"""
import argparse, json, sys, tempfile, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--files', type=int, default=12)
    ap.add_argument('--size', type=int, default=2048)
    ap.add_argument('--bin', type=int, default=3)
    ap.add_argument('--threads', type=int, default=4)
    ap.add_argument('--slots', type=int, default=4)
    a = ap.parse_args()
    import cv2, torch
    from polarimetry_b200 import AnalysisParams, Calibration, Engine, synth
    from polarimetry_b200.ingest import FolderAnalyzer
    d = np.load(ROOT / 'tests' / 'golden' / 'calib' / 'Disk_Ga0_Pa0_Ta0_Gb0_Pb0_Tb0_Gc0_Pc0_Tc0.npz')
    cal = Calibration(d['RoTest'], d['PsiTest'], int(d['NbMapValues']))
    eng = Engine(0)
    H = W = a.size
    tmp = Path(tempfile.mkdtemp(prefix='pb_folder_', dir='/dev/shm' if Path('/dev/shm').exists() else None))
    files = []
    for k in range(a.files):
        v = synth.stack_1pf(H, W, 18, seed=k % 3)
        f = tmp / f'stack_{k:03d}.tif'
        assert cv2.imwritemulti(str(f), list(v))
        files.append(f)
    nbytes = sum(f.stat().st_size for f in files)
    p = AnalysisParams(method='1PF', bins=(a.bin, a.bin), ilow=23000.0, out_dtype='float32')
    px = a.files * H * W * 18

    def serial():
        t0 = time.perf_counter()
        tot = 0
        for f in files:
            st = eng.define_stack(f)
            dark = float('{:.0f}'.format(st['dark']))
            res = eng.analyze(st['values'], AnalysisParams(**{**p.__dict__, 'dark': dark}), cal)
            tot += int(res.mask.sum().item())
        torch.cuda.synchronize()
        return time.perf_counter() - t0, tot

    serial()
    ts, valid_s = serial()
    fa = FolderAnalyzer(eng, p, cal, slots=a.slots, decode_threads=a.threads)
    fa.run(files[:2])
    t0 = time.perf_counter()
    recs = fa.run(files)
    torch.cuda.synchronize()
    tp = time.perf_counter() - t0
    valid_p = sum(r['n_valid'] for r in recs)
    assert valid_p == valid_s, (valid_p, valid_s)
    rep = fa.overlap_report()
    print(json.dumps({'workload': f'{a.files} TIFF stacks {H}x{W}x18 uint16 ({nbytes / 1e6:.0f} MB on {tmp.parent}), 1PF bin {a.bin}x{a.bin}, dark per file',
                      'serial_s': ts, 'pipelined_s': tp, 'speedup': ts / tp, 'serial_mpx_angles_s': px / ts / 1e6, 'pipelined_mpx_angles_s': px / tp / 1e6,
                      'decode_threads': a.threads, 'slots': a.slots, 'overlap': rep,
                      'timeline_first_files': [{k: t[k] for k in ('index', 'slot', 'decode_ms', 'h2d_ms', 'compute_ms')} for t in fa.timeline[:6]]}))
    for f in files:
        f.unlink()
    tmp.rmdir()


if __name__ == '__main__':
    main()
