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


def hist_explained_by_ties(h_got, h_ref, ref_values, edges, tol, is_rho=False, rotation_fig=0.0):
    """Tie-aware histogram comparison for variables that go through a transcendental (atan2 / acos / hypot / cos).
    Where the reference's value lies within `tol` of a bin edge its bin is decided by the last bits of the host's libm / SVML
    (numpy's AVX-512 arctan2 differs from glibc's by 1 ulp on 7 % of random inputs), so it is not a property of the algorithm.
    Exact integer data produce such values structurally (e.g. 4POLAR: equal counts in two channels -> rho = 45 deg up to rounding).
    Requirement: the net number of values that crossed each edge (cumulative count difference) is at most the number of
    reference values within `tol` of that edge; everything else must match exactly. Returns (ok, n_ties, n_moved)."""
    d = np.asarray(ref_values, dtype=np.float64)
    d = d[np.isfinite(d)]
    if is_rho:
        d = np.mod(2 * (d + rotation_fig), 360) / 2
    edges = np.asarray(edges, dtype=np.float64)
    ties = np.array([np.count_nonzero(np.abs(d - e) <= tol) for e in edges])
    diff = np.asarray(h_got, dtype=np.int64) - np.asarray(h_ref, dtype=np.int64)
    # crossed[k] = values counted below edge k+1 by `got` but not by `ref` (k = 0 .. nbins-1; the last one concerns the closed top edge)
    crossed = np.cumsum(diff)
    lower = np.concatenate(([0], crossed[:-1]))          # net inflow through the bottom edge of the histogram is bounded by ties[0]
    ok = True
    for k in range(len(diff)):
        # values entering / leaving bin k can only come through its two edges
        if abs(diff[k]) > ties[k] + ties[k + 1]:
            ok = False
    if abs(int(diff.sum())) > ties[0] + ties[-1]:
        ok = False
    return ok, int(ties.sum()), int(np.abs(diff).sum() // 2 + abs(int(diff.sum())))
