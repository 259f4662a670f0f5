"""the result writers' gathers (save_mat / save_csv, PP:2794-2842). Golden = the files the UNMODIFIED reference
wrote (oracle/make_golden_writers.py -> tests/golden/writers.npz); the oracle restatement and the device compaction
(pb_compact through Engine.mat_arrays / csv_columns / save_mat / save_csv) must reproduce them exactly."""
import csv

import numpy as np
import pytest

from oracle import restate as rs
from tests import util

CASE = dict(dark=480.0, ilow=38000.0, bins=(3, 3), rotation=(17.0, 5.0, 0.0))
ROIS = [dict(indx=1, select=True, ILow=30000.0, vertices=np.array([[5.2, 90.7, 90.1, 6.0], [4.0, 6.5, 70.2, 60.9]])),
        dict(indx=3, select=True, ILow=42000.0, vertices=np.array([[40.0, 110.0, 75.5], [10.0, 40.0, 78.0]]))]
VARIANTS = (('whole', None, False, None, 0), ('roi3', ROIS, True, 3, 1), ('union', ROIS, False, None, 0),
            ('contour', None, False, None, 0), ('contour_roi1', ROIS, True, 1, 0))


def edge_for(tag, shape):
    from oracle import cases
    return cases.make_edge(shape) if tag.startswith('contour') else None


def mat_keys(tag):
    return ('X', 'Y', 'Rho', 'Psi') + (('X_contour', 'Y_contour', 'Rho_contour', 'Psi_contour') if tag.startswith('contour') else ())


def golden():
    return np.load(util.GOLD / 'writers.npz')


def table(cols):
    rows = [[rs.fmt3(x) if isinstance(x, float) else str(x) for x in (c.tolist() if c.dtype.kind == 'f' else list(c))] for c in cols]
    return np.array(list(zip(*rows)), dtype=str).reshape(-1, len(cols))


@pytest.mark.parametrize('tag,rois,per_roi,roi,group', VARIANTS)
def test_oracle_writers_match_reference_files(tag, rois, per_roi, roi, group):
    g = golden()
    p = rs.Params(method='1PF', dark=CASE['dark'], bins=CASE['bins'], ilow=CASE['ilow'], rotation=CASE['rotation'])
    res = rs.analyze_stack(g['values'], p, util.load_disk(util.cases.DISK_NODIST), rois, per_roi, edge_contours=edge_for(tag, g['values'].shape[1:]),
                           for_calib=False)
    m = rs.mat_arrays(res, res['roi_map'], roi)
    for k in mat_keys(tag):
        assert np.array_equal(g[f'{tag}__mat_{k}'].astype(np.float64), m[k].astype(np.float64)), (tag, k)
    assert m['Rho'].dtype == np.float16
    names, cols = rs.csv_columns(res, res['roi_map'], roi)
    assert list(names) == list(g[f'{tag}__csv_names'])
    assert np.array_equal(table(cols), g[f'{tag}__csv_rows'])
    if tag.startswith('contour'):
        names, cols = rs.csv_contour_columns(res, res['roi_map'], roi)
        assert list(names) == list(g[f'{tag}__csvct_names'])
        assert np.array_equal(table(cols), g[f'{tag}__csvct_rows'])


@pytest.mark.gpu
@pytest.mark.parametrize('tag,rois,per_roi,roi,group', VARIANTS)
def test_device_writers_match_reference_files(tag, rois, per_roi, roi, group, tmp_path):
    from scipy.io import loadmat
    from polarimetry_b200 import AnalysisParams, Calibration, Engine
    g = golden()
    d = util.load_disk(util.cases.DISK_NODIST)
    eng = Engine(0)
    p = AnalysisParams(method='1PF', dark=CASE['dark'], bins=CASE['bins'], ilow=CASE['ilow'], rotation=CASE['rotation'])
    res = eng.analyze(g['values'], p, Calibration(d.rho_table, d.psi_table, 401), rois, per_roi, None, edge_for(tag, g['values'].shape[1:]))
    m = eng.mat_arrays(res, roi, group)
    for k in mat_keys(tag):
        assert m[k].dtype == (np.uint16 if k[0] in 'XY' else np.float16)
        assert np.array_equal(g[f'{tag}__mat_{k}'].astype(np.float64), m[k].astype(np.float64)), (tag, k)
    names, cols = eng.csv_columns(res, roi, group)
    assert list(names) == list(g[f'{tag}__csv_names'])
    assert np.array_equal(table(cols), g[f'{tag}__csv_rows'])
    if tag.startswith('contour'):
        names, cols = eng.csv_contour_columns(res, roi, group)
        assert list(names) == list(g[f'{tag}__csvct_names'])
        assert np.array_equal(table(cols), g[f'{tag}__csvct_rows'])
    # and the files themselves
    eng.save_mat(res, tmp_path / 'o.mat', '1PF', 'synthetic.tif', roi, group)
    eng.save_csv(res, tmp_path / 'o.csv', '1PF', 'synthetic.tif', roi, group)
    mm = loadmat(str(tmp_path / 'o.mat'))
    assert mm['polarimetry'][0] == '1PF'
    for k in ('Rho', 'Psi'):
        assert np.array_equal(np.asarray(mm[k]).squeeze(), g[f'{tag}__mat_{k}'])
    with open(tmp_path / 'o.csv', newline='') as f:
        rows = list(csv.reader(f))
    assert rows[1] == list(g[f'{tag}__csv_names'])
    assert np.array_equal(np.array(rows[2:], dtype=str).reshape(-1, 4), g[f'{tag}__csv_rows'])
    if tag.startswith('contour'):
        with open(tmp_path / 'o_contour.csv', newline='') as f:
            rows = list(csv.reader(f))
        assert rows[1] == list(g[f'{tag}__csvct_names'])
        assert np.array_equal(np.array(rows[2:], dtype=str).reshape(-1, 4), g[f'{tag}__csvct_rows'])
    eng.close()


@pytest.mark.gpu
def test_compact_properties_large():
    """Stable compaction on a 2048x2048 map (2048 tiles, multi-chunk scan path at 3000x3000): equals torch boolean indexing, with a
    selection mask, ROI bits, a separate filter map, float32 and float64 maps, empty and full selections."""
    import torch
    from polarimetry_b200 import Engine
    eng = Engine(0)
    gen = torch.Generator(device='cuda').manual_seed(5)
    for n_side, dt in ((2048, torch.float64), (3000, torch.float32), (37, torch.float64)):
        n = n_side * n_side
        v = torch.rand(n, generator=gen, device='cuda', dtype=torch.float64) * 180.0
        v[torch.rand(n, generator=gen, device='cuda') < 0.4] = float('nan')
        v = v.to(dt)
        f = torch.rand(n, generator=gen, device='cuda', dtype=torch.float64).to(dt)
        f[torch.rand(n, generator=gen, device='cuda') < 0.5] = float('inf')
        sel = (torch.rand(n, generator=gen, device='cuda') < 0.7).to(torch.uint8)
        bits = torch.randint(0, 8, (n,), generator=gen, device='cuda', dtype=torch.int32)
        for use_f, use_s, use_b in ((False, False, False), (True, True, True), (False, True, False), (True, False, True)):
            keep = torch.isfinite(f if use_f else v)
            if use_s:
                keep &= sel.bool()
            if use_b:
                keep &= (bits & 5) != 0
            o = eng.compact(v, f if use_f else None, sel if use_s else None, bits.cpu().numpy().astype(np.uint32) if use_b else None, 5,
                            want=('f16', 'val', 'index'))
            assert o['count'] == int(keep.sum())
            assert torch.equal(o['index'].long(), torch.nonzero(keep).squeeze(1))
            ref = v[keep]
            assert torch.equal(o['val'].nan_to_num(-1.0), ref.nan_to_num(-1.0))
            assert np.array_equal(o['f16'].cpu().numpy(), ref.double().cpu().numpy().astype(np.float16), equal_nan=True)
    z = torch.full((5000,), float('nan'), device='cuda', dtype=torch.float64)
    assert eng.compact(z)['count'] == 0
    assert eng.compact(torch.zeros(5000, device='cuda', dtype=torch.float64))['count'] == 5000
    eng.close()


def test_c_contract_half_rounding_and_gather():
    """binary64 -> binary16 in ONE rounding (numpy's astype, the device's cvt.rn.f16.f64): the C contract against numpy on
    ties, subnormals, overflow, NaN and random values; and the gather against the reference's .mat arrays."""
    from oracle import contract
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-200, 200, 4000), rng.uniform(-1e-6, 1e-6, 500), 10.0 ** rng.uniform(-9, 6, 500),
                        [0.0, -0.0, np.inf, -np.inf, np.nan, 65504.0, 65519.99, 65520.0, 1e9, 5.96e-8, 2.98e-8, 2.9802322387695312e-08,
                         6.1e-5, 6.097555160522461e-05, 1.0 + 2.0 ** -11, 1.0 + 3 * 2.0 ** -11, 2049.0, 2051.0, 179.96875, 179.984375]])
    with np.errstate(over='ignore'):
        ref = x.astype(np.float16)
    got = contract.f64_to_f16(x)
    assert np.array_equal(got.view(np.uint16)[~np.isnan(x)], ref.view(np.uint16)[~np.isnan(x)])
    assert np.isnan(got[np.isnan(x)]).all()
    g = golden()
    p = rs.Params(method='1PF', dark=CASE['dark'], bins=CASE['bins'], ilow=CASE['ilow'], rotation=CASE['rotation'])
    res = rs.analyze_stack(g['values'], p, util.load_disk(util.cases.DISK_NODIST), None, False, for_calib=False)
    for k in ('Rho', 'Psi'):
        assert np.array_equal(contract.compact_f16(res['vars'][k]).astype(np.float64), g[f'whole__mat_{k}'])
