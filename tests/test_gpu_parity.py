"""GPU parity: the CUDA path (through the C ABI) against the golden vectors of the unmodified reference
and against the oracle on the same seeded inputs. Bars: masks, intensity/intmap, histogram counts and all
1PF maps BIT-EXACT (float64 outputs); rho/psi/eta within 1e-9 deg and S2/S4 within 1e-12 where a
transcendental is involved (north_star asks 1e-3 deg); float32 output mode within 1e-3 deg."""
import numpy as np
import pytest
import torch

from oracle import cases, restate as rs
from polarimetry_b200 import AnalysisParams, Calibration, Engine
from tests import util

pytestmark = pytest.mark.gpu
NAMES = [c['name'] for c in util.CASES]
ANGLE_VARS = ('Rho', 'Psi', 'Eta', 'Rho_angle', 'Rho_contour', 'Psi_contour', 'Eta_contour')


@pytest.fixture(scope='module')
def eng():
    e = Engine()
    yield e
    e.close()


_calibs = {}


def calib_for(c):
    key = (c['method'], c['disk'], c['calib_label'])
    if key not in _calibs:
        o = util.calib_for(c)
        if o is None:
            _calibs[key] = None
        elif o.invK is not None:
            _calibs[key] = Calibration(invK=o.invK, name=o.name)
        else:
            _calibs[key] = Calibration(o.rho_table, o.psi_table, o.grid.size, name=o.name)
    return _calibs[key]


def params_for(c, **kw):
    return AnalysisParams(method=c['method'], dark=c['dark'], removed_intensity=c['removed_intensity'], bins=c['bins'],
                          polar_dir=c['polar_dir'], offset_angle=c['offset_angle'], rotation=c['rotation'], ilow=c['ilow'],
                          nbins=c['nbins'], **kw)


def check_against_golden(c, res, g, exact_maps, tol_angle=1e-9, tol_other=1e-12, check_hist=True):
    names = util.INDEX[c['name']]['var_order']
    assert [v.name for v in res.vars] == names
    assert util.eq(res.intensity.cpu().numpy(), g['intensity'])
    assert util.eq(res.mask.cpu().numpy().astype(bool), np.isfinite(g['var_Rho']) & np.isfinite(g['intmap']))
    im = res.intmap.cpu().numpy().astype(np.float64)
    if exact_maps:
        assert util.eq(im, g['intmap'])
    else:
        assert util.eq(np.isnan(im), np.isnan(g['intmap']))
    for v in res.vars:
        got = v.values.cpu().numpy().astype(np.float64)
        ref = g['var_' + v.name]
        assert util.eq(np.isnan(got), np.isnan(ref)), v.name
        if exact_maps and c['method'] == '1PF':
            assert util.eq(got, ref), v.name
        else:
            m = np.isfinite(ref)
            d = util.circ_diff(got[m], ref[m]) if v.name.startswith('Rho') else np.abs(got[m] - ref[m])
            assert d.max(initial=0.0) <= (tol_angle if v.name in ANGLE_VARS else tol_other), (v.name, d.max())
    if check_hist:
        for (k, s), h in res.hist.items():
            assert util.eq(h, g[util.hist_key(k, s)]), (k, s)
        assert len(res.hist) == len([k for k in g.files if k.startswith('hist_')])


@pytest.mark.parametrize('name', NAMES)
def test_analyze_matches_reference(eng, name):
    c = util.CASE_BY_NAME[name]
    g = util.golden(name)
    edge = cases.make_edge(g['values'].shape[1:]) if c['edge'] else None
    res = eng.analyze(g['values'], params_for(c), calib_for(c), c['rois'], c['per_roi'], util.base_mask_for(c, g['values'].shape[1:]), edge)
    check_against_golden(c, res, g, exact_maps=True)


@pytest.mark.parametrize('name', NAMES)
def test_unfused_kernels_match_reference(eng, name):
    c = util.CASE_BY_NAME[name]
    g = util.golden(name)
    edge = cases.make_edge(g['values'].shape[1:]) if c['edge'] else None
    res = eng.analyze(g['values'], params_for(c, force_unfused=True), calib_for(c), c['rois'], c['per_roi'], util.base_mask_for(c, g['values'].shape[1:]), edge)
    assert res.plan == 'unfused'
    check_against_golden(c, res, g, exact_maps=True)


@pytest.mark.parametrize('name', [n for n in NAMES if not n.startswith('4polar')])
def test_exact_chi2_flag_gives_same_decisions(eng, name):
    c = util.CASE_BY_NAME[name]
    g = util.golden(name)
    edge = cases.make_edge(g['values'].shape[1:]) if c['edge'] else None
    res = eng.analyze(g['values'], params_for(c, exact_chi2=True), calib_for(c), c['rois'], c['per_roi'], util.base_mask_for(c, g['values'].shape[1:]), edge)
    check_against_golden(c, res, g, exact_maps=True)


@pytest.mark.parametrize('name', ['1pf_bin1x1', '1pf_bin3x3', '1pf_disk561_bin3', 'shg_bin1', 'cars_bin3', '4polar3d_bin1', '4polar2d_bin3'])
def test_float32_output_mode(eng, name):
    """Throughput mode: float32 maps (<= 1e-3 deg, north_star), histograms still from the fp64 values."""
    c = util.CASE_BY_NAME[name]
    g = util.golden(name)
    res = eng.analyze(g['values'], params_for(c, out_dtype='float32'), calib_for(c), c['rois'], c['per_roi'])
    assert res.vars[0].values.dtype == torch.float32
    check_against_golden(c, res, g, exact_maps=False, tol_angle=1e-3, tol_other=1e-6)


def test_seams_get_intensity_field_binning(eng):
    for name in ('1pf_bin3x3', '1pf_bin2x5', 'cars_bin2x5', '1pf_u8', '1pf_fracdark'):
        c = util.CASE_BY_NAME[name]
        g = util.golden(name)
        v = g['values']
        p = util.params_for(c)
        I = eng.get_intensity(v, c['dark'], c['bins']).cpu().numpy()
        assert util.eq(I, g['intensity']), name
        f_ref = rs.binning(rs.preprocess(v, p), c['bins'])
        f = eng.field(v, params_for(c)).cpu().numpy()
        assert util.eq(f, f_ref), name
        f2 = eng.binning(rs.preprocess(v, p), c['bins']).cpu().numpy()
        assert util.eq(f2, f_ref), name


@pytest.mark.parametrize('shape,bins', [((1, 1, 1), (3, 3)), ((2, 1, 7), (5, 4)), ((1, 2, 3), (7, 7)), ((3, 70, 131), (20, 20)),
                                        ((2, 33, 64), (32, 32)), ((1, 300, 257), (2, 3))])
def test_box_filter_edge_shapes(eng, shape, bins):
    from scipy.ndimage import uniform_filter
    rng = np.random.default_rng(5)
    f = rng.integers(0, 65535, size=shape).astype(np.float64) / 7.0
    ref = uniform_filter(f, size=[0, bins[0], bins[1]], mode='reflect')
    assert util.eq(eng.binning(f, bins).cpu().numpy(), ref)


@pytest.mark.parametrize('name', ['1pf_bin3x3', '1pf_noncalib', 'shg_noncalib', '4polar3d_bin3', '1pf_adversarial_bin3'])
def test_compute_fields_seam(eng, name):
    """compute_fields on the oracle's double field + incoming mask (the S1 seam), incl. fit/chi2 maps."""
    c = util.CASE_BY_NAME[name]
    g = util.golden(name)
    v = g['values']
    p = util.params_for(c)
    edge = cases.make_edge(v.shape[1:]) if c['edge'] else None
    intensity = rs.get_intensity(v, p.dark, p.bins)
    _, rmask = rs.roi_masks(intensity, p)
    field = rs.binning(rs.preprocess(v, p), p.bins)
    res = eng.compute_fields(field, rmask[0], params_for(c), calib_for(c), edge, want_fit=not name.startswith('4polar'))
    res.intensity = torch.from_numpy(g['intensity'])
    check_against_golden(c, res, g, exact_maps=True, check_hist=False)
    if 'chi2' in g.files:
        m = res.mask.cpu().numpy().astype(bool)
        chi2 = res.chi2.cpu().numpy()
        assert util.eq(np.isnan(chi2), np.isnan(g['chi2']))
        assert np.allclose(chi2[m], g['chi2'][m], rtol=1e-12, atol=0)
        if 'field_fit' in g.files:
            assert np.allclose(res.field_fit.cpu().numpy()[:, m], g['field_fit'][:, m], rtol=1e-13, atol=0)


def test_histo_and_stats_seams(eng):
    c = util.CASE_BY_NAME['1pf_rois']
    g = util.golden('1pf_rois')
    p = util.params_for(c)
    ref = rs.analyze_stack(g['values'], p, util.calib_for(c), c['rois'], True)
    res = eng.analyze(g['values'], params_for(c), calib_for(c), c['rois'], True)
    for gi, idx in enumerate([1, 2]):
        sel = rs.slice_roi(ref['roi_map'], idx)
        for nm, rot in (('Rho', 12.0), ('Psi', 0.0)):
            vmin, vmax = p.ranges[nm]
            want = rs.histo(ref['vars'][nm], sel, nm, vmin, vmax, rot, 45)[0]
            got = eng.histo(res.get_var(nm).values, sel.astype(np.uint8), nm, vmin, vmax, rot, 45)
            assert util.eq(got, want)
        want = rs.histo(ref['vars']['Rho'], sel, 'Rho_contour', 0, 180, 0.0, 30, 'polar3')[0]
        got = eng.histo(res.get_var('Rho').values, sel.astype(np.uint8), 'Rho_contour', 0, 180, 0.0, 30, 'polar3')
        assert util.eq(got, want)
        st_ref = rs.vecstats(ref, ref['roi_map'], p, idx)
        st = eng.roi_stats(res, params_for(c), idx, gi)
        assert st['N'] == st_ref['N']
        for k in ('MeanRho', 'StdRho', 'MeanDeltaRho', 'MeanPsi', 'StdPsi', 'MeanInt', 'StdInt'):
            assert abs(st[k] - st_ref[k]) <= 1e-9 * max(1.0, abs(st_ref[k])), (k, st[k], st_ref[k])
            assert rs.fmt3(float(st[k])) == rs.fmt3(float(st_ref[k]))


def test_batch_and_host_paths(eng):
    from polarimetry_b200 import synth
    B, N, H, W = 5, 18, 64, 96
    stacks = np.stack([synth.stack_1pf(H, W, N, seed=100 + b) for b in range(B)])
    c = util.CASE_BY_NAME['1pf_bin1x1']
    cal = calib_for(c)
    for bins in ((1, 1), (3, 3)):
        p = AnalysisParams(method='1PF', dark=480.0, bins=bins, ilow=45000.0)
        o = eng.analyze_batch(stacks, p, cal)
        oh = eng.analyze_host(stacks, p, cal)
        hsum = np.zeros((2, 60), dtype=np.int64)
        for b in range(B):
            ref = rs.analyze_stack(stacks[b], rs.Params(method='1PF', dark=480.0, bins=bins, ilow=45000.0), util.calib_for(c))
            for nm in ('Rho', 'Psi'):
                assert util.eq(o['vars'][nm][b].cpu().numpy(), ref['vars'][nm])
                assert util.eq(oh['vars'][nm][b].numpy(), ref['vars'][nm])
            assert util.eq(o['intmap'][b].cpu().numpy(), ref['intmap'])
            assert util.eq(oh['mask'][b].numpy().astype(bool), ref['mask'])
            hsum[0] += ref['hist'][('Rho', None)]
            hsum[1] += ref['hist'][('Psi', None)]
        assert util.eq(o['hist'][0, 0].cpu().numpy(), hsum)
        assert util.eq(oh['hist'][0, 0].numpy(), hsum)


def test_larger_stack_properties(eng):
    """Full-size-ish (1024x1024x18) checks that need no CPU oracle run: fused == unfused == exact-chi2 bitwise,
    histogram total == number of finite values, idempotence of a second run."""
    from polarimetry_b200 import synth
    v = synth.stack_1pf(1024, 1024, 18, seed=7)
    cal = calib_for(util.CASE_BY_NAME['1pf_bin1x1'])
    for bins in ((1, 1), (3, 3)):
        p = AnalysisParams(method='1PF', dark=480.0, bins=bins, ilow=50000.0)
        a = eng.analyze(v, p, cal)
        b = eng.analyze(v, AnalysisParams(method='1PF', dark=480.0, bins=bins, ilow=50000.0, force_unfused=True), cal)
        c2 = eng.analyze(v, AnalysisParams(method='1PF', dark=480.0, bins=bins, ilow=50000.0, exact_chi2=True), cal)
        for x in (b, c2):
            for va, vb in zip(a.vars, x.vars):
                assert torch.equal(torch.nan_to_num(va.values, nan=-1.0), torch.nan_to_num(vb.values, nan=-1.0))
            assert torch.equal(a.mask, x.mask)
            for k in a.hist:
                assert util.eq(a.hist[k], x.hist[k])
        rho = a.vars[0].values
        assert int(a.hist[('Rho', None)].sum()) == int(torch.isfinite(rho).sum())
        assert 0 < int(a.mask.sum()) < v.shape[1] * v.shape[2]


def test_compute_dark_on_device(eng):
    """"next" row N2: Stack.compute_dark (PC:266-278), exact integer sums on the device."""
    from polarimetry_b200 import synth
    for seed, shape, dt in ((1, (18, 100, 130), np.uint16), (2, (4, 64, 64), np.uint16), (3, (6, 41, 59), np.uint8), (4, (18, 512, 512), np.uint16)):
        v = synth.stack_1pf(shape[1], shape[2], shape[0], seed, dtype=dt, level=2000.0 if dt == np.uint16 else 100.0,
                            amp=1500.0 if dt == np.uint16 else 60.0, dark=480 if dt == np.uint16 else 0)
        v[:, :25, :25] = np.minimum(v[:, :25, :25], 60)
        v[0, 3:9, 4:11] = 0
        assert eng.compute_dark(v) == rs.compute_dark(v), shape
    assert eng.compute_dark(np.zeros((3, 10, 10), dtype=np.uint16)) == 0.0     # smaller than one cell -> 0.0 like an empty stack


def test_calibration_sweep(eng):
    """"next" row N1: disks x stacks sweep (PP:883-898), statistics rows against the oracle at the reference's 3 decimals."""
    from polarimetry_b200 import synth
    stacks = [synth.stack_1pf(48, 64, 18, seed=200 + k) for k in range(2)]
    names = [cases.DISK_NODIST, cases.DISK_561]
    disks = [Calibration(util.load_disk(n).rho_table, util.load_disk(n).psi_table, 401, name=n) for n in names]
    p = AnalysisParams(method='1PF', dark=480.0, bins=(2, 2), ilow=30000.0, offset_angle=60.0)
    rows = eng.calibration_sweep(stacks, disks, p)
    assert len(rows) == 4
    for row in rows:
        d = names.index(row['disk'])
        ref = rs.analyze_stack(stacks[row['stack']], rs.Params(method='1PF', dark=480.0, bins=(2, 2), ilow=30000.0, offset_angle=60.0),
                               util.load_disk(names[d]))
        st = ref['stats'][None]
        assert row['N'] == st['N']
        for k in ('MeanRho', 'StdRho', 'MeanPsi', 'StdPsi', 'MeanInt', 'StdInt'):
            assert rs.fmt3(float(row[k])) == rs.fmt3(float(st[k])), (k, row[k], st[k])


def test_project_then_lut_apply_equals_fused(eng):
    """N1: project once + disk-dependent tail == the fused analysis, bit for bit (maps, mask, histograms), for the fused pixel,
    fused march and unfused plans, with ROIs and rotations; and the sweep rows agree with the per-(disk, stack) analyses."""
    from polarimetry_b200 import synth
    v = np.stack([synth.stack_1pf(80, 112, 18, seed=300 + k) for k in range(3)])
    v[0, :, 10:14, 20:30] = 0                    # a0 == 0 pixels
    v[1, 0, 5, 5] = 65535                         # high-modulation pixel: exact chi2 path
    rois = [dict(indx=1, select=True, ILow=30000.0, vertices=np.array([[5.2, 90.7, 90.1, 6.0], [4.0, 6.5, 70.2, 60.9]])),
            dict(indx=3, select=True, ILow=42000.0, vertices=np.array([[40.0, 110.0, 75.5], [10.0, 40.0, 78.0]]))]
    names = [cases.DISK_NODIST, cases.DISK_561]
    disks = [Calibration(util.load_disk(n).rho_table, util.load_disk(n).psi_table, 401, name=n) for n in names]
    for bins, kw, plan in (((1, 1), {}, 'fused_pixel'), ((3, 3), {}, 'fused_march'), ((3, 2), dict(force_unfused=True), 'unfused')):
        for use_rois, per_roi in ((False, False), (True, True)):
            p = AnalysisParams(method='1PF', dark=480.0, bins=bins, ilow=35000.0, rotation=(17.0, 5.0, 0.0), **kw)
            proj = eng.project(v, p, rois if use_rois else None, per_roi)
            assert proj['plan'] == plan
            for cal in disks:
                a = eng.lut_apply(proj, p, cal)
                b = eng.analyze_batch(v, p, cal, rois if use_rois else None, per_roi, want_intensity=False)
                for n in ('Rho', 'Psi'):
                    assert util.eq(a['vars'][n].cpu().numpy(), b['vars'][n].cpu().numpy()), (bins, use_rois, cal.name, n)
                assert util.eq(a['intmap'].cpu().numpy(), b['intmap'].cpu().numpy())
                assert util.eq(a['mask'].cpu().numpy(), b['mask'].cpu().numpy())
                assert util.eq(a['hist'].cpu().numpy(), b['hist'].cpu().numpy())
                assert int(a['mask'].sum()) > 0
    p = AnalysisParams(method='1PF', dark=480.0, bins=(3, 3), ilow=35000.0)
    r1 = eng.calibration_sweep(list(v), disks, p, rois, per_roi=True)
    r2 = eng.calibration_sweep(list(v), disks, p, rois, per_roi=True, project_once=False)
    assert len(r1) == len(r2) == 2 * 3 * 2
    for x, y in zip(r1, r2):
        assert x.keys() == y.keys()
        for k in x:
            # the maps are bit-identical (above); the moment sums of pb_stats accumulate with atomics, so the statistics are
            # compared the way the reference prints them (3 decimals, PP:2791-2792)
            fx, fy = (rs.fmt3(float(z)) if isinstance(z, (float, np.floating)) else z for z in (x[k], y[k]))
            assert fx == fy, (k, x[k], y[k])
    with pytest.raises(ValueError):
        eng.project(v, AnalysisParams(method='SHG'))


def test_full_width_stack_against_oracle(eng):
    """1024x1024x18 (BASELINE configs[0] shape), bin 3x3 and 1x1, straight against the oracle: every map and bin exact."""
    from polarimetry_b200 import synth
    v = synth.stack_1pf(1024, 1024, 18, seed=11)
    cal = calib_for(util.CASE_BY_NAME['1pf_bin1x1'])
    ocal = util.load_disk(cases.DISK_NODIST)
    for bins in ((3, 3), (1, 1)):
        res = eng.analyze(v, AnalysisParams(method='1PF', dark=480.0, bins=bins, ilow=45000.0), cal)
        ref = rs.analyze_stack(v, rs.Params(method='1PF', dark=480.0, bins=bins, ilow=45000.0), ocal)
        assert res.plan == ('fused_march' if bins[0] > 1 else 'fused_pixel')
        for var in res.vars:
            assert util.eq(var.values.cpu().numpy(), ref['vars'][var.name]), (bins, var.name)
        assert util.eq(res.mask.cpu().numpy().astype(bool), ref['mask'])
        assert util.eq(res.intmap.cpu().numpy(), ref['intmap'])
        for k, h in res.hist.items():
            assert util.eq(h, ref['hist'][k]), (bins, k)


def test_many_rois_and_bins(eng):
    """32 ROIs, per-ROI histograms with 128 bins: the per-CTA histogram area exceeds 48 KB of shared memory."""
    c = util.CASE_BY_NAME['shg_bin1']
    g = util.golden('shg_bin1')
    v = g['values']
    H, W = v.shape[1:]
    rois = [dict(indx=k + 1, vertices=np.array([[2.0 + k, 50.0, 50.0, 2.0 + k], [1.0, 1.0, 38.0 - (k % 7), 38.0]]), ILow='0', select=True)
            for k in range(32)]
    p = util.params_for(c)
    p.nbins = 128
    ref = rs.analyze_stack(v, p, None, rois, True)
    res = eng.analyze(v, params_for(c, ), None, rois, True) if False else eng.analyze(v, AnalysisParams(method='SHG', dark=0.0, nbins=128), None, rois, True)
    assert len(res.hist) == 32 * 4
    for k, h in res.hist.items():
        assert util.eq(h, ref['hist'][k]), k


def test_error_paths_return_codes(eng):
    """Bad calls fail with a negative status and a message; nothing falls back."""
    import ctypes as C
    from polarimetry_b200 import _cabi
    lib, ctx = eng.lib, eng._ctx
    # establish the context state (basis for N = 18, a disk cone) with one valid call
    eng.analyze(np.full((18, 8, 8), 1000, dtype=np.uint16), AnalysisParams(method='1PF'), calib_for(util.CASE_BY_NAME['1pf_bin1x1']))
    p = _cabi.PbParams(0, 1, 18, 8, 8, 1, 1, 0, 0.0, 0.0, 0.0, 0.0, 0.0, 500.0, 0.0)
    o = _cabi.PbOutputs()
    assert lib.pb_analyze(ctx, C.byref(p), None, 1, None, None, C.byref(o), None) == -1          # NULL stack
    assert b'NULL' in lib.pb_last_error(ctx)
    p.method = 9
    assert lib.pb_analyze(ctx, C.byref(p), C.c_void_p(8), 1, None, None, C.byref(o), None) == -1   # unknown method
    p.method, p.N = 5, 18
    assert lib.pb_analyze(ctx, C.byref(p), C.c_void_p(8), 1, None, None, C.byref(o), None) == -1   # 4POLAR needs 4 planes
    p.method, p.N = 0, 12
    assert lib.pb_analyze(ctx, C.byref(p), C.c_void_p(8), 1, None, None, C.byref(o), None) == -3   # basis not set for this N
    p.method, p.N, p.bin_h = 0, 18, 40
    assert lib.pb_analyze(ctx, C.byref(p), C.c_void_p(8), 1, None, None, C.byref(o), None) == -4   # unsupported window
    with pytest.raises(ValueError):
        eng.analyze(np.zeros((18, 8, 8), dtype=np.uint16), AnalysisParams(method='1PF'), None)     # 1PF without a disk cone
    with pytest.raises(TypeError):
        eng.analyze(np.zeros((18, 8, 8), dtype=np.int64), AnalysisParams(method='SHG'))            # unsupported dtype


def test_drop_in_functions_with_reference_signatures(eng):
    """The S1-S3 seams exactly as the reference calls them (PP:2968, PP:2962, PP:2245): module-level compute_fields with
    the reference's positional/keyword arguments and its side effects (mask updated in place, datastack filled),
    Polarimetry.binning(self, field) reading the spin boxes, Stack.get_intensity(self, dark, bin)."""
    import types
    import polarimetry_b200 as pb
    from polarimetry_b200 import api
    api._default_engine = eng
    c = util.CASE_BY_NAME['1pf_noncalib']
    g = util.golden('1pf_noncalib')
    v = g['values']
    p = util.params_for(c)
    ocal = util.calib_for(c)
    edge = cases.make_edge(v.shape[1:])

    class Spin:
        def __init__(self, x): self.x = x
        def get(self): return self.x

    app = types.SimpleNamespace(bin_spinboxes=[Spin(str(c['bins'][0])), Spin(str(c['bins'][1]))])
    stack = types.SimpleNamespace(values=v)
    intensity = pb.get_intensity(stack, dark=c['dark'], bin=[c['bins'][0], c['bins'][1]])
    assert util.eq(intensity, g['intensity'])
    assert pb.compute_dark(stack) == rs.compute_dark(v) and pb.compute_dark(stack, 16) == rs.compute_dark(v, 16)   # PC:266-278
    field = np.maximum(v - (c['dark'] + 0.0), 0)
    field = pb.binning(app, field)
    assert util.eq(field, rs.binning(rs.preprocess(v, p), p.bins))
    mask = intensity >= c['ilow']
    calibration = types.SimpleNamespace(RhoPsi=np.moveaxis(np.stack((ocal.rho_table, ocal.psi_table)), 0, -1), xy=(ocal.grid, ocal.grid), name='disk')
    H, W = v.shape[1:]
    yy, xx = np.indices((H, W), dtype=float)
    ds = types.SimpleNamespace(nangle=v.shape[0], height=H, width=W, vars=[], xct=xx.copy(), yct=yy.copy())
    out = pb.compute_fields(field, ds, mask, method='1PF', polar_dir=c['polar_dir'], offset_angle=c['offset_angle'], calibration=calibration,
                            edge_contours=edge, reference_angle=c['rotation'][2], rotation=c['rotation'][0], for_calib=False)
    assert out is ds
    assert [x.name for x in ds.vars] == util.INDEX['1pf_noncalib']['var_order']
    for x in ds.vars:
        assert util.eq(x.values, g['var_' + x.name]), x.name
    assert util.eq(ds.intmap, g['intmap'])
    assert util.eq(mask, np.isfinite(g['intmap']))                       # updated in place, as PP:3000 / PP:3025 do
    assert util.eq(ds.xmap, g['xmap']) and util.eq(ds.ymap, g['ymap'])
    assert util.eq(np.isnan(ds.chi2), np.isnan(g['chi2']))
    m = np.isfinite(g['chi2'])
    assert np.allclose(ds.chi2[m], g['chi2'][m], rtol=1e-12, atol=0)
    assert np.isnan(ds.field[:, ~mask]).all() and util.eq(ds.field[:, mask], rs.binning(rs.preprocess(v, p), p.bins)[:, mask])
    api._default_engine = None


def test_time_series_per_frame_histograms(eng):
    """Time-series mode (BASELINE configs[4] semantics): per-frame histograms [frames, groups, vars, nbins] from one launch,
    each equal to the oracle's histogram of that frame; their sum equals the batch histogram."""
    from polarimetry_b200 import synth
    frames = np.stack([synth.stack_1pf(40, 72, 18, seed=300 + k) for k in range(4)])
    cal = calib_for(util.CASE_BY_NAME['1pf_bin1x1'])
    ocal = util.load_disk(cases.DISK_NODIST)
    for bins in ((1, 1), (3, 3)):
        p = AnalysisParams(method='1PF', dark=480.0, bins=bins, ilow=38000.0)
        o = eng.analyze_batch(frames, p, cal, hist_per_stack=True)
        per = o['hist'].cpu().numpy()
        assert per.shape == (4, 1, 2, 60)
        tot = eng.analyze_batch(frames, p, cal)['hist'].cpu().numpy()
        assert util.eq(per.sum(axis=0), tot[0])
        for k in range(4):
            ref = rs.analyze_stack(frames[k], rs.Params(method='1PF', dark=480.0, bins=bins, ilow=38000.0), ocal)
            assert util.eq(per[k, 0, 0], ref['hist'][('Rho', None)]) and util.eq(per[k, 0, 1], ref['hist'][('Psi', None)])


def test_time_series_analyzer_places_frames(eng):
    """TimeSeriesAnalyzer (configs[4]): a rank that owns frames [2, 4) of a 6-frame series returns a [6, ...] buffer holding
    its two frames' histograms (equal to the per-frame analysis) and zeros elsewhere."""
    from polarimetry_b200 import synth
    from polarimetry_b200.batch import TimeSeriesAnalyzer, shard_range
    frames = np.stack([synth.stack_1pf(40, 72, 18, seed=320 + k) for k in range(6)])
    cal = calib_for(util.CASE_BY_NAME['1pf_bin1x1'])
    p = AnalysisParams(method='1PF', dark=480.0, bins=(3, 3), ilow=38000.0)
    own = shard_range(6, 1, 3)
    assert (own.start, own.stop) == (2, 4)
    o = TimeSeriesAnalyzer(eng, p, cal).run(eng._to_device(frames[own.start:own.stop]), 6, 1, 3, reduce=False)
    hf = o['hist_frames'].cpu().numpy()
    assert hf.shape == (6, 1, 2, 60)
    assert not hf[:2].any() and not hf[4:].any()
    ref = eng.analyze_batch(frames, p, cal, hist_per_stack=True)['hist'].cpu().numpy()
    assert util.eq(hf[2:4], ref[2:4])


def test_full_size_time_series_frame_properties():
    """BASELINE configs[4] at its FULL frame size (8192x8192x90 uint16, 12.1 GB, bin 5x5) -- the oracle cannot run this size, so the
    checks are size-independent properties: rows of the full-frame maps are bit-identical to the analysis of full-width row crops
    at the top edge, beyond the 2^32-element offset and at the bottom edge; the rho histogram total equals the valid-pixel count."""
    import importlib.util
    from pathlib import Path
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip('needs 40 GB of free device memory')
    spec = importlib.util.spec_from_file_location('c5_frame', Path(__file__).resolve().parent.parent / 'tools' / 'c5_frame.py')
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    r = mod.run(8192, 90, 5, reps=1)
    assert r['plan'] == 'fused_march' and r['all_ok'], r['checks']
    assert 0.2 < r['valid_pixel_fraction'] < 0.6


@pytest.mark.parametrize('method', ['1PF', 'SHG'])
def test_batch_statistics_over_two_shards(eng, method):
    """ROI statistics of a batch split into two shards (what two ranks hold) through pb_stats_batch -- moment sums on the
    device, the two accumulators summed between / after the passes as the all-reduce does -- equal the reference's statistics on the
    concatenated values of all stacks at its 3 decimals (and to 1e-9), per ROI group."""
    from polarimetry_b200 import batch, synth
    from polarimetry_b200.api import VAR_NAMES
    B, H, W = 6, 48, 64
    c = util.CASE_BY_NAME['1pf_rois']
    if method == '1PF':
        stacks = np.stack([synth.stack_1pf(H, W, 18, seed=300 + b) for b in range(B)])
        p = AnalysisParams(method='1PF', dark=480.0, bins=(3, 3), rotation=(0.0, 25.0, 0.0))
        op = rs.Params(method='1PF', dark=480.0, bins=(3, 3), rotation=(0.0, 25.0, 0.0))
        cal, ocal = calib_for(util.CASE_BY_NAME['1pf_bin1x1']), util.load_disk(cases.DISK_NODIST)
    else:
        stacks = np.stack([synth.stack_4harm(H, W, 36, seed=300 + b) for b in range(B)])
        p, op, cal, ocal = AnalysisParams(method='SHG'), rs.Params(method='SHG'), None, None
    rois = c['rois']
    shards = [stacks[:4], stacks[4:]]
    outs = [eng.analyze_batch(s, p, cal, rois, True, want_hist=False) for s in shards]
    keep = []                                                    # what the all-reduce would carry

    def fake_allreduce(t):
        keep.append(t)

    accs = [eng.batch_stats_accumulate(o, p, None) for o in outs]          # single-shard sanity: runs both passes
    assert all(torch.isfinite(a1[:, 0, 0]).all() for a1, _ in accs)
    # two-rank emulation: pass 1 on both shards, sum, pass 2 on both shards with the summed acc1, sum
    import ctypes as C
    from polarimetry_b200._cabi import PbOutputs
    names = VAR_NAMES[method]

    def run_pass(o, npass, acc1, acc2):
        Bs = o['intmap'].shape[0]
        po = (PbOutputs * Bs)()
        for b in range(Bs):
            for k, n in enumerate(names):
                po[b].var[k] = o['vars'][n][b].data_ptr()
            po[b].intmap = o['intmap'][b].data_ptr()
        pp = eng._params(p, 'uint16', stacks.shape[1], H, W)
        eng._ck(eng.lib.pb_stats_batch(eng._ctx, C.byref(pp), npass, 0, po, Bs, o['_keep'][1].data_ptr(), acc1.data_ptr(), acc2.data_ptr(),
                                       eng._stream()), 'pb_stats_batch')

    ng = len(outs[0]['roi_indices'])
    a1 = [torch.zeros((ng, 5, 4), dtype=torch.float64, device=eng.device) for _ in shards]
    a2 = [torch.zeros((ng, 5, 2), dtype=torch.float64, device=eng.device) for _ in shards]
    for o, x1, x2 in zip(outs, a1, a2):
        run_pass(o, 1, x1, x2)
    tot1 = a1[0] + a1[1]
    for o, x2 in zip(outs, a2):
        run_pass(o, 2, tot1, x2)
    rows = batch.finalize_stats(tot1, a2[0] + a2[1], names)
    # reference: the statistics of the concatenated values (every stack analysed by the oracle, maps stacked along the rows)
    refs = [rs.analyze_stack(s, op, ocal, rois, True, want_stats=False) for s in stacks]
    cat = {'vars': {n: np.concatenate([r['vars'][n] for r in refs], axis=0) for n in names},
           'intmap': np.concatenate([r['intmap'] for r in refs], axis=0)}
    roi_map = np.concatenate([refs[0]['roi_map']] * B, axis=1)
    for g, idx in enumerate(outs[0]['roi_indices']):
        want = rs.vecstats(cat, roi_map, op, idx)
        assert rows[g]['N'] == want['N'] and want['N'] > 50
        for k, v in want.items():
            if k == 'N':
                continue
            assert rs.fmt3(float(rows[g][k])) == rs.fmt3(float(v)), (g, k, rows[g][k], v)
            assert abs(rows[g][k] - v) <= 1e-9 * max(1.0, abs(v)), (g, k)


def test_histogram_match_aggregation_on_an_aligned_sample(eng):
    """PB_FLAG_HIST_MATCH (one shared-memory atomic per distinct bin of a warp) on a sample whose rho falls into two or three bins:
    identical histograms and maps, with and without the flag, fused pixel and fused march plans; and equal to the oracle."""
    from polarimetry_b200 import synth
    v = synth.stack_1pf(96, 128, 18, seed=9, rho_lo=40.0, rho_span=5.0)
    cal = calib_for(util.CASE_BY_NAME['1pf_bin1x1'])
    for bins in ((1, 1), (3, 3)):
        a = eng.analyze(v, AnalysisParams(method='1PF', dark=480.0, bins=bins, ilow=30000.0), cal)
        b = eng.analyze(v, AnalysisParams(method='1PF', dark=480.0, bins=bins, ilow=30000.0, hist_match=True), cal)
        ref = rs.analyze_stack(v, rs.Params(method='1PF', dark=480.0, bins=bins, ilow=30000.0), util.load_disk(cases.DISK_NODIST), want_stats=False)
        for k in a.hist:
            assert util.eq(a.hist[k], b.hist[k]) and util.eq(a.hist[k], ref['hist'][k]), (bins, k)
        hr = np.sort(ref['hist'][('Rho', None)])[::-1]
        assert hr[:3].sum() > 0.9 * hr.sum()                     # an aligned sample: three bins hold nearly every pixel
        for va, vb in zip(a.vars, b.vars):
            assert util.eq(va.values.cpu().numpy(), vb.values.cpu().numpy())


def test_4polar_binned_takes_the_fused_march(eng):
    """4POLAR with spatial binning runs in ONE launch of the row-march kernel (box filter of the four channel planes + the
    polarization-matrix inversion of PP:3001-3014 / 3041-3051 per pixel); a stack whose width the tensor map cannot describe falls
    back to the unfused kernels with the same results."""
    from polarimetry_b200 import synth
    for name in ('4polar3d_bin3', '4polar2d_bin3'):
        c = util.CASE_BY_NAME[name]
        g = util.golden(name)
        n0 = eng.launch_count()
        res = eng.analyze(g['values'], params_for(c), calib_for(c), c['rois'], c['per_roi'])
        assert res.plan == 'fused_march' and eng.launch_count() - n0 == 1
        check_against_golden(c, res, g, exact_maps=True)
    cal = calib_for(util.CASE_BY_NAME['4polar3d_bin3'])
    for shape in ((40, 52), (64, 96), (17, 16)):
        v = synth.stack_4polar(*shape, seed=3, three_d=True)
        kw = dict(method='4POLAR 3D', dark=0.0, bins=(3, 3), ilow=1000.0)
        a = eng.analyze(v, AnalysisParams(**kw), cal)
        b = eng.analyze(v, AnalysisParams(force_unfused=True, **kw), cal)
        for va, vb in zip(a.vars, b.vars):
            assert torch.equal(torch.nan_to_num(va.values, nan=-1.0), torch.nan_to_num(vb.values, nan=-1.0)), (shape, va.name)
        assert torch.equal(a.mask, b.mask) and torch.equal(a.intensity, b.intensity)
        for k in a.hist:
            assert util.eq(a.hist[k], b.hist[k]), (shape, k)
