"""TEST INFRASTRUCTURE ONLY -- ctypes driver for oracle/contract.c (the op-for-op fp64 contract).

`build()` compiles contract.c with gcc (-ffp-contract=off) into oracle/_build/libcontract.so;
`analyze_stack()` chains the C routines in the order of PyPOLAR.py:2948-2968 and returns the same dict
layout as `oracle.restate.analyze_stack`. Never imported by the product.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

from . import restate as rs

HERE = Path(__file__).resolve().parent
SO = HERE / '_build' / 'libcontract.so'
METHOD_ID = {'1PF': 0, 'CARS': 1, 'SRS': 2, 'SHG': 3, '2PF': 4, '4POLAR 2D': 5, '4POLAR 3D': 6}
DTYPE_ID = {np.dtype('uint8'): 0, np.dtype('uint16'): 1, np.dtype('float32'): 2, np.dtype('float64'): 3}
_lib = None


def build(force: bool = False) -> Path:
    src = HERE / 'contract.c'
    if force or not SO.exists() or SO.stat().st_mtime < src.stat().st_mtime:
        SO.parent.mkdir(exist_ok=True)
        subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-fno-fast-math', '-shared', '-fPIC',
                               '-o', str(SO), str(src), '-lm'])
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _d(x):
    return C.c_double(float(x))


def _l(x):
    return C.c_long(int(x))


def get_intensity(values, dark, bins):
    v = np.ascontiguousarray(values)
    N, H, W = v.shape
    out = np.empty((H, W))
    lib().pc_intensity(_p(v), DTYPE_ID[v.dtype], _l(N), _l(H), _l(W), _d(dark), int(bins[0]), int(bins[1]), _p(out))
    return out


def field(values, p: rs.Params):
    v = np.ascontiguousarray(values)
    N, H, W = v.shape
    out = np.empty((N, H, W))
    lib().pc_field(_p(v), DTYPE_ID[v.dtype], _l(N), _l(H), _l(W), _d(float(p.dark) + float(p.removed_intensity)),
                   int(p.method == 'CARS'), int(p.method.startswith('4POLAR')), _p(out))
    bh, bw = int(p.bins[0]), int(p.bins[1])
    if bh > 1 or bw > 1:
        lib().pc_box(_p(out), _l(N), _l(H), _l(W), bh, bw)
    return out


def compute_fields(f, mask_in, p: rs.Params, calib: Optional[rs.Calib], edge=None, want_fit=False):
    N, H, W = f.shape
    P = H * W
    mask = np.ascontiguousarray(mask_in, dtype=np.uint8).copy()
    out = {'vars': {}}
    nan = lambda: np.full((H, W), np.nan)
    a0 = nan()
    rot = float(p.rotation[0])
    if not p.method.startswith('4POLAR'):
        _, e2, e4 = rs.basis(N, p.polar_dir, p.offset_angle)
        c2, s2 = np.ascontiguousarray(e2.real), np.ascontiguousarray(e2.imag)
        c4, s4 = np.ascontiguousarray(e4.real), np.ascontiguousarray(e4.imag)
        nh = 1 if p.method == '1PF' else 2
        A2, B2, A4, B4, chi2 = nan(), nan(), nan(), nan(), nan()
        fit = np.empty_like(f) if want_fit else None
        lib().pc_harmonics(_p(f), _l(N), _l(H), _l(W), nh, _p(c2), _p(s2), _p(c4), _p(s4), _d(p.chi2threshold),
                           _p(mask), _p(a0), _p(A2), _p(B2), _p(A4), _p(B4), _p(chi2), _p(fit))
        out['chi2'], out['A2'], out['B2'], out['A4'], out['B4'] = chi2, A2, B2, A4, B4
        if want_fit:
            out['field_fit'] = fit
        rho = nan()
        if p.method == '1PF':
            psi = nan()
            g = np.ascontiguousarray(calib.grid)
            lib().pc_lut_1pf(_p(A2), _p(B2), _l(P), _p(np.ascontiguousarray(calib.rho_table)),
                             _p(np.ascontiguousarray(calib.psi_table)), _p(g), _l(g.size), _d(rot), _p(mask), _p(rho), _p(psi))
            out['vars'] = {'Rho': rho, 'Psi': psi}
        else:
            s2_, s4_ = nan(), nan()
            sshg = nan() if p.method == 'SHG' else None
            lib().pc_out_4harm(_p(A2), _p(B2), _p(A4), _p(B4), _l(P), _d(rot), _p(mask), _p(rho), _p(s2_), _p(s4_), _p(sshg))
            out['vars'] = {'Rho': rho, 'S2': s2_, 'S4': s4_}
            if sshg is not None:
                out['vars']['S_SHG'] = sshg
    else:
        three_d = '3D' in p.method
        rho, psi = nan(), nan()
        eta = nan() if three_d else None
        lib().pc_4polar(_p(f), _l(P), _p(np.ascontiguousarray(calib.invK)), int(three_d), _d(rot), _p(mask), _p(a0),
                        _p(rho), _p(psi), _p(eta))
        out['vars'] = {'Rho': rho, 'Psi': psi}
        if three_d:
            out['vars']['Eta'] = eta
    ref_angle = float(p.rotation[2])
    rho_a = nan() if ref_angle else None
    rho_c = nan() if edge is not None else None
    lib().pc_tail(_l(P), _p(mask), _p(a0), _p(out['vars']['Rho']), _d(ref_angle), _p(rho_a),
                  _p(None if edge is None else np.ascontiguousarray(edge)), _p(rho_c))
    if rho_c is not None:
        both = np.isfinite(out['vars']['Rho']) & np.isfinite(edge)
        names = list(out['vars'])[1:]
        out['vars']['Rho_contour'] = rho_c
        for nm in names:
            out['vars'][nm + '_contour'] = np.where(both, out['vars'][nm], np.nan)
    if rho_a is not None:
        out['vars']['Rho_angle'] = rho_a
    out['intmap'] = a0
    out['mask'] = mask.astype(bool)
    return out


def histo(values, sel, name, vmin, vmax, rotation_fig, nbins, fold90=False):
    edges = np.linspace(vmin, 90 if fold90 else vmax, nbins + 1)
    counts = np.zeros(nbins, dtype=np.int64)
    v = np.ascontiguousarray(values, dtype=np.float64)
    s = None if sel is None else np.ascontiguousarray(sel, dtype=np.uint8)
    lib().pc_hist(_p(v), _p(s), _l(v.size), int(name in ('Rho', 'Rho_angle')), _d(rotation_fig), int(fold90),
                  _p(edges), int(nbins), _p(counts))
    return counts


def analyze_stack(values, p: rs.Params, calib=None, rois: Optional[Sequence[dict]] = None, per_roi=False,
                  base_mask=None, edge=None, want_fit=False):
    intensity = get_intensity(values, p.dark, p.bins)
    roi_map, rmask = rs.roi_masks(intensity, p, rois, per_roi, base_mask)   # host raster + compare (PP:2869-2892)
    gmask = np.any(rmask, axis=0)
    f = field(values, p)
    res = compute_fields(f, gmask, p, calib, edge, want_fit)
    res['intensity'], res['roi_map'], res['field'] = intensity, roi_map, f
    slices = [None] if not (per_roi and rois) else [r['indx'] for r in rois if r.get('select')]
    res['hist'] = {}
    for s in slices:
        sel = rs.slice_roi(roi_map, s)
        for name, v in res['vars'].items():
            if name in p.ranges:
                vmin, vmax = p.ranges[name]
                res['hist'][(name, s)] = histo(v, sel, name, vmin, vmax, float(p.rotation[1]), p.nbins)
    return res


# ---- SURVEY 8f N3 / N4 ------------------------------------------------------------------------------------------
def rescale(a: np.ndarray) -> np.ndarray:
    """define_stack's rescale (PP:1929-1930) through the C contract (float32 or float64 stacks)."""
    a = np.ascontiguousarray(a)
    out = np.empty(a.shape, dtype=np.uint16)
    fn = lib().pc_rescale_f32 if a.dtype == np.float32 else lib().pc_rescale_f64
    if a.dtype not in (np.float32, np.float64):
        raise TypeError(a.dtype)
    fn(_p(a), _l(a.size), _p(out))
    return out


def f64_to_f16(x: np.ndarray) -> np.ndarray:
    """binary64 -> binary16 with one rounding to nearest even, element by element (C contract)."""
    x = np.ascontiguousarray(x, dtype=np.float64).ravel()
    f = lib().pc_f64_to_f16
    f.restype, f.argtypes = C.c_uint16, [C.c_double]
    return np.array([f(float(v)) for v in x], dtype=np.uint16).view(np.float16)


def compact_f16(values: np.ndarray, sel: Optional[np.ndarray] = None) -> np.ndarray:
    """save_mat's gather + cast (PP:2838-2840) through the C contract."""
    v = np.ascontiguousarray(values, dtype=np.float64).ravel()
    s = None if sel is None else np.ascontiguousarray(sel, dtype=np.uint8).ravel()
    out = np.empty(v.size, dtype=np.uint16)
    f = lib().pc_compact_f16
    f.restype = C.c_long
    n = f(_p(v), _p(s), _l(v.size), _p(out))
    return out[:n].view(np.float16)
