"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/ from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):  python -m oracle.make_golden
Writes
  tests/golden/calib/<disk>.npz      the reference's shipped disk-cone tables (RoTest, PsiTest, NbMapValues)
  tests/golden/calib/<K>.txt.npy     the reference's shipped 4POLAR K matrices
  tests/golden/cases/<case>.npz      per parity case (oracle/cases.py): the input stack and every output of
                                     the reference's compute_intensity -> analyze_stack / compute_fields ->
                                     np.histogram path, executed through oracle/ref_loader.py
The fixtures are data, not source; they let the CPU and GPU parity tests run where the reference is absent.
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
from scipy.io import loadmat

from . import cases, ref_loader as rl

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / 'tests' / 'golden'


def main(only=None):
    """only: iterable of case names to (re)generate; the other fixtures and their index entries are left as they are."""
    ref = rl.reference_dir()
    (GOLD / 'calib').mkdir(parents=True, exist_ok=True)
    (GOLD / 'cases').mkdir(parents=True, exist_ok=True)
    for disk in cases.DISKS:
        d = loadmat(str(ref / 'diskcones' / (disk + '.mat')), squeeze_me=True)
        np.savez_compressed(GOLD / 'calib' / (disk + '.npz'), RoTest=np.asarray(d['RoTest'], dtype=np.float64),
                            PsiTest=np.asarray(d['PsiTest'], dtype=np.float64), NbMapValues=int(d['NbMapValues']))
    for k in ('Calib_th_2D', 'Calib_th_3D', 'Calib_20231009_2D'):
        np.save(GOLD / 'calib' / (k + '.npy'), np.genfromtxt(str(ref / 'calibration' / (k + '.txt')), dtype=np.float64))
    idx_file = GOLD / 'cases' / 'index.json'
    index = json.loads(idx_file.read_text()) if (only and idx_file.exists()) else {}
    for c in cases.all_cases():
        if only and c['name'] not in only:
            continue
        v = cases.make_stack(c['stack'])
        edge = cases.make_edge(v.shape[1:]) if c['edge'] else None
        out = rl.run_reference(v, c['method'], c['dark'], c['ilow'], c['offset_angle'], c['polar_dir'], c['bins'],
                               c['rotation'], c['removed_intensity'],
                               disk=str(ref / 'diskcones' / (c['disk'] + '.mat')), calib_label=c['calib_label'],
                               rois=c['rois'], per_roi=c['per_roi'], for_calib=c['for_calib'], edge_contours=edge,
                               nbins=c['nbins'], mask_image=cases.make_mask_image(v.shape[1:]) if c['mask'] else None)
        arrs = dict(values=v, intensity=out['intensity'], intmap=out['intmap'], roi_map=out['roi_map'])
        for name, val in out['vars'].items():
            arrs['var_' + name] = val
        for (name, s), h in out['hist'].items():
            arrs[f'hist_{name}_{"all" if s is None else s}'] = h
        if not c['for_calib']:
            for k in ('xmap', 'ymap', 'chi2'):
                arrs[k] = out[k]
            if np.ndim(out['field_fit']) == 3:
                arrs['field_fit'] = out['field_fit'].astype(np.float64)
        np.savez_compressed(GOLD / 'cases' / (c['name'] + '.npz'), **arrs)
        index[c['name']] = dict(var_order=list(out['vars']), rows=out.get('rows'),
                                n_valid=int(np.isfinite(out['vars']['Rho']).sum()))
        print(c['name'], index[c['name']]['n_valid'])
    (GOLD / 'cases' / 'index.json').write_text(json.dumps(index, indent=1, default=str))


if __name__ == '__main__':
    import sys
    main(set(sys.argv[1:]) or None)
