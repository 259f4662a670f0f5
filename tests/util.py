"""Shared helpers for the parity tests: golden fixtures, calibration tables, case -> Params."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from oracle import cases, restate as rs

GOLD = Path(__file__).resolve().parent / 'golden'
CASES = cases.all_cases()
CASE_BY_NAME = {c['name']: c for c in CASES}
INDEX = json.loads((GOLD / 'cases' / 'index.json').read_text())


def load_disk(name: str) -> rs.Calib:
    d = np.load(GOLD / 'calib' / (name + '.npz'))
    return rs.Calib.from_disk(d['RoTest'], d['PsiTest'], int(d['NbMapValues']), name)


def load_K(name: str) -> np.ndarray:
    return np.load(GOLD / 'calib' / (name + '.npy'))


def calib_for(c) -> rs.Calib | None:
    if c['method'] == '1PF':
        return load_disk(c['disk'])
    if c['method'].startswith('4POLAR'):
        nm = 'Calib_th_' + c['method'][-2] + 'D' if c['calib_label'] == 'no distortions' else c['calib_label']
        return rs.Calib.from_K(load_K(nm), nm)
    return None


def params_for(c) -> rs.Params:
    return rs.Params(method=c['method'], dark=c['dark'], removed_intensity=c['removed_intensity'], bins=c['bins'],
                     polar_dir=c['polar_dir'], offset_angle=c['offset_angle'], rotation=c['rotation'], ilow=c['ilow'],
                     nbins=c['nbins'])


def golden(name: str):
    return np.load(GOLD / 'cases' / (name + '.npz'))


def hist_key(name, s):
    return f'hist_{name}_{"all" if s is None else s}'


def eq(a, b) -> bool:
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def circ_diff(a, b, period=180.0):
    """Circular distance for rho-like maps (NaN positions must agree, checked separately)."""
    d = np.abs(a - b)
    return np.minimum(d, period - d)


def base_mask_for(c, shape):
    """The mask image of a 'Mask'-option case (PP:2082-2094), or None."""
    return cases.make_mask_image(shape) if c.get('mask') else None
