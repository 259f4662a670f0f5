"""TEST INFRASTRUCTURE ONLY -- the parity case list shared by `make_golden.py` and the tests.

Each case = a seeded synthetic stack (polarimetry_b200.synth) + the analysis settings the reference
GUI would hold (PyPOLAR.py:2950-2968). Sizes are small so that the reference, the NumPy restatement and
the C contract all finish in well under a second per case.
"""
from __future__ import annotations

import numpy as np

from polarimetry_b200 import synth

DISK_NODIST = 'Disk_Ga0_Pa0_Ta0_Gb0_Pb0_Tb0_Gc0_Pc0_Tc0'
DISK_561 = 'Disk_Ga0.1_Pa20_Ta0_Gb0_Pb10_Tb45_Gc0.2_Pc0_Tc0'
DISK_488 = 'Disk_Ga-0.1_Pa0_Ta45_Gb0.1_Pb10_Tb45_Gc-0.2_Pc0_Tc0'
DISK_640 = 'Disk_Ga0.2_Pa10_Ta0_Gb-0.1_Pb20_Tb0_Gc0.2_Pc0_Tc0'
DISKS = (DISK_NODIST, DISK_561, DISK_488, DISK_640)

H, W = 40, 56


def _rois():
    # vertices are float [2, nv] arrays (x row, y row) as stored by the reference at PP:1458 / PP:2061
    r1 = dict(indx=1, vertices=np.array([[4.6, 30.2, 33.9, 8.1], [3.3, 5.7, 28.4, 25.0]]), ILow='20000', select=True)
    r2 = dict(indx=2, vertices=np.array([[20.5, 50.0, 44.2], [10.0, 12.5, 36.9]]), ILow='36000', select=True)
    r3 = dict(indx=3, vertices=np.array([[1.0, 10.0, 10.0, 1.0], [30.0, 30.0, 38.0, 38.0]]), ILow='0', select=False)
    return [r1, r2, r3]


def _c(name, stack, **kw):
    d = dict(name=name, stack=stack, method='1PF', dark=480.0, ilow=0.0, offset_angle=0.0, polar_dir='clockwise',
             bins=(1, 1), rotation=(0.0, 0.0, 0.0), removed_intensity=0, disk=DISK_NODIST,
             calib_label='no distortions', rois=None, per_roi=False, for_calib=True, edge=False, nbins=60, mask=False)
    d.update(kw)
    return d


def all_cases():
    c = []
    g1 = ('1pf', dict(H=H, W=W, N=18, seed=0))
    for b in [(1, 1), (2, 2), (3, 3), (5, 5), (3, 1), (1, 4), (2, 5), (7, 3)]:
        c.append(_c(f'1pf_bin{b[0]}x{b[1]}', g1, bins=b, ilow=45000.0))
    c.append(_c('1pf_disk561_bin3', g1, bins=(3, 3), disk=DISK_561, offset_angle=60.0, ilow=42000.0))
    c.append(_c('1pf_disk488_bin1', ('1pf', dict(H=H, W=W, N=18, seed=1)), disk=DISK_488, offset_angle=60.0))
    c.append(_c('1pf_disk640_bin2', ('1pf', dict(H=H, W=W, N=18, seed=2)), bins=(2, 2), disk=DISK_640, offset_angle=60.0))
    c.append(_c('1pf_ccw_rot', g1, polar_dir='counterclockwise', offset_angle=17.5, rotation=(33.0, 12.0, 0.0), bins=(3, 3)))
    c.append(_c('1pf_removed', g1, removed_intensity=150, bins=(3, 3), ilow=30000.0, nbins=90))
    c.append(_c('1pf_removed_bin1', g1, removed_intensity=150, ilow=30000.0))
    c.append(_c('shg_removed_bin1', ('4harm', dict(H=H, W=W, N=36, seed=2, dark=100)), method='SHG', dark=100.0, removed_intensity=40))
    c.append(_c('1pf_n12', ('1pf', dict(H=H, W=W, N=12, seed=3)), bins=(3, 3), nbins=10))
    c.append(_c('1pf_n90_bin5', ('1pf', dict(H=24, W=40, N=90, seed=4)), bins=(5, 5)))
    c.append(_c('1pf_u8', ('1pf', dict(H=H, W=W, N=18, seed=5, dark=0, dtype='uint8', level=100.0, amp=60.0)), dark=0.0, bins=(3, 3)))
    c.append(_c('1pf_adversarial_bin1', ('adv', dict(H=32, W=48, N=18, seed=0))))
    c.append(_c('1pf_adversarial_bin3', ('adv', dict(H=32, W=48, N=18, seed=0)), bins=(3, 3)))
    c.append(_c('1pf_fracdark', g1, dark=480.5, bins=(3, 3)))
    c.append(_c('1pf_rois', g1, rois=_rois(), per_roi=True, bins=(3, 3)))
    c.append(_c('1pf_rois_union', g1, rois=_rois(), per_roi=False, bins=(2, 2)))
    c.append(_c('1pf_mask', g1, bins=(3, 3), ilow=30000.0, mask=True))
    c.append(_c('1pf_mask_rois', g1, rois=_rois(), per_roi=True, mask=True))
    c.append(_c('1pf_noncalib', g1, bins=(3, 3), for_calib=False, rotation=(10.0, 5.0, 40.0), edge=True))
    g2 = ('4harm', dict(H=H, W=W, N=36, seed=0))
    g2d = ('4harm', dict(H=H, W=W, N=36, seed=1, dark=100))
    for m in ('SHG', 'CARS', 'SRS', '2PF'):
        c.append(_c(f'{m.lower()}_bin1', g2, method=m, dark=0.0))
    c.append(_c('shg_bin3', g2d, method='SHG', dark=100.0, bins=(3, 3), rotation=(25.0, 0.0, 0.0)))
    c.append(_c('cars_bin3', g2d, method='CARS', dark=100.0, bins=(3, 3)))
    c.append(_c('cars_bin2x5', g2d, method='CARS', dark=100.0, bins=(2, 5)))
    c.append(_c('shg_noncalib', g2, method='SHG', dark=0.0, for_calib=False, rotation=(0.0, 0.0, 30.0)))
    g3 = ('4polar', dict(H=H, W=W, seed=0, three_d=True))
    g3b = ('4polar', dict(H=H, W=W, seed=1, three_d=False))
    c.append(_c('4polar3d_bin1', g3, method='4POLAR 3D', dark=0.0))
    c.append(_c('4polar3d_bin3', g3, method='4POLAR 3D', dark=0.0, bins=(3, 3), rotation=(15.0, 0.0, 0.0)))
    c.append(_c('4polar2d_bin1', g3b, method='4POLAR 2D', dark=0.0))
    c.append(_c('4polar2d_bin3', g3b, method='4POLAR 2D', dark=0.0, bins=(3, 3)))
    c.append(_c('4polar2d_calib2023', g3b, method='4POLAR 2D', dark=0.0, calib_label='Calib_20231009_2D'))
    return c


def make_stack(spec):
    kind, kw = spec
    kw = dict(kw)
    if 'dtype' in kw:
        kw['dtype'] = np.dtype(kw['dtype']).type
    if kind == '1pf':
        return synth.stack_1pf(**kw)
    if kind == '4harm':
        return synth.stack_4harm(**kw)
    if kind == '4polar':
        return synth.stack_4polar(**kw)
    if kind == 'adv':
        return synth.adversarial_1pf(**kw)
    raise ValueError(kind)


def make_edge(shape):
    """A synthetic contour-angle map (what define_rho_ct, PP:1109-1125, would hand to compute_fields):
    finite angles on a band, NaN elsewhere."""
    yy, xx = np.indices(shape, dtype=float)
    e = np.full(shape, np.nan)
    band = (np.abs(yy - shape[0] / 2) < 8)
    e[band] = (37.0 + 3.0 * xx[band]) % 180
    return e


def make_mask_image(shape):
    """A grey-level mask image as the reference reads it from '<stack name>.png' (PP:2082-2094): 0 = excluded, anything else
    kept (the reference divides by the maximum and tests != 0)."""
    yy, xx = np.indices(shape)
    im = np.full(shape, 255, dtype=np.uint8)
    im[(yy - shape[0] * 0.4) ** 2 + (xx - shape[1] * 0.55) ** 2 < (0.22 * min(shape)) ** 2] = 0      # a hole
    im[:, : shape[1] // 7] = 0                                                                          # a dark band
    im[shape[0] // 2:, shape[1] // 2:] = 90                                                             # grey, still kept
    return im
