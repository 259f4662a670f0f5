"""install(): the drop-in boundary on the reference's own modules (seams S1-S5).

CPU part (runs where the reference tree is present): install() on the modules loaded by oracle/ref_loader.py rebinds every
seam, each rebound callable presents the reference's own signature, and uninstall() restores the originals.
GPU part: a stand-in application object with exactly the state analyze_stack reads (PP:2949-2968) is driven through the
rebound fused `analyze_stack` and compared with the golden outputs of the unmodified reference."""
import inspect
import types

import numpy as np
import pytest

from oracle import cases, ref_loader as rl, restate as rs
from polarimetry_b200 import api
from tests import util


@pytest.mark.skipif(not rl.available(), reason='reference tree not present')
def test_install_rebinds_every_seam_with_the_reference_signatures():
    pp, pc = rl.load()
    originals = {'S1': pp.compute_fields, 'S2': pp.Polarimetry.binning, 'S4': pp.Polarimetry.analyze_stack,
                 'S3': pc.Stack.get_intensity, 'N2': pc.Stack.compute_dark, 'S5': pc.np}
    saved = api.install(pp, pc)
    try:
        assert set(saved['attrs']) == set(originals)
        rebound = {'S1': pp.compute_fields, 'S2': pp.Polarimetry.binning, 'S4': pp.Polarimetry.analyze_stack,
                   'S3': pc.Stack.get_intensity, 'N2': pc.Stack.compute_dark}
        for key, fn in rebound.items():
            assert fn is not originals[key], key
            assert inspect.signature(fn) == inspect.signature(originals[key]), key
            assert fn.__name__ == originals[key].__name__, key
        # the only call sites of the fused seam keep working by name (PP:898, PP:2128, PP:2133 call self.analyze_stack)
        assert list(inspect.signature(pp.Polarimetry.analyze_stack).parameters) == ['self', 'datastack', 'for_calib']
        # S5: the classes module's numpy is numpy, except histogram
        assert pc.np.histogram is api.histogram
        assert pc.np.mean is np.mean and pc.np.pi == np.pi
        h, e = pc.np.histogram(np.array([1, 2, 3]), bins=3)             # non-float / implicit bins: numpy's own
        assert h.tolist() == [1, 1, 1]
    finally:
        api.uninstall(saved)
    assert pp.compute_fields is originals['S1'] and pp.Polarimetry.analyze_stack is originals['S4']
    assert pc.Stack.get_intensity is originals['S3'] and pc.np is np


def test_install_needs_no_gpu_until_called():
    mod = types.ModuleType('PyPOLAR')

    class Polarimetry:
        def binning(self, field): return field
        def analyze_stack(self, datastack, for_calib=False): return None

    mod.Polarimetry = Polarimetry
    mod.compute_fields = lambda *a, **k: None
    saved = api.install(mod)
    assert set(saved['attrs']) == {'S1', 'S2', 'S4'}
    api.uninstall(saved)
    assert mod.Polarimetry.binning(None, 5) == 5


# ---------------------------------------------------------------------------------------------------------------------
class _Var:
    def __init__(self, v): self.v = v
    def get(self): return self.v
    def set(self, v): self.v = v
    def select(self): self.v = True
    def deselect(self): self.v = False


def _fake_reference():
    """Modules with the attributes install() touches and an application class carrying the state of PP:2949-2979."""
    pm, cm = types.ModuleType('PyPOLAR'), types.ModuleType('pypolar_classes')

    class Variable:                                   # constructor signature of PC:323
        def __init__(self, name='', values=None, datastack=None):
            self.name, self.values = name, values

    class Stack:
        def get_intensity(self, dark=0.0, bin=[1, 1]): raise AssertionError('not rebound')
        def compute_dark(self, size_cell=20): raise AssertionError('not rebound')

    class Polarimetry:
        def binning(self, field): raise AssertionError('not rebound')
        def analyze_stack(self, datastack, for_calib=False): raise AssertionError('not rebound')

        def get_mask(self, datastack, display=True):                      # PP:2082-2094
            return np.ones((datastack.height, datastack.width)) if self.mask_image is None else self.mask_image / np.amax(self.mask_image)

        def compute_roi_map(self, datastack):                             # PP:2869-2892 (restated by the oracle)
            p = rs.Params(ilow=float(self.ilow.get()))
            return rs.roi_masks(datastack.intensity, p, datastack.rois, bool(self.per_roi.get()), self.get_mask(datastack))

        def define_rho_ct(self, edge):                                    # PP:1109-1125: hands the contour-angle map over
            yy, xx = np.indices(edge.shape, dtype=float)
            return edge, xx, yy

        def return_vecexcel(self, datastack, roi_map, roi=[], simplify=False):
            self.calls.append(('return_vecexcel', roi['indx'] if roi else None))
            return [('row', roi['indx'] if roi else None)]

        def plot_data(self, datastack, roi_map=None): self.calls.append(('plot_data',))
        def save_data(self, datastack, roi_map=None): self.calls.append(('save_data',))

    pm.Polarimetry, pm.Variable, pm.compute_fields = Polarimetry, Variable, (lambda *a, **k: (_ for _ in ()).throw(AssertionError('not rebound')))
    cm.Stack, cm.Variable, cm.np = Stack, Variable, np
    return pm, cm


def _app(pm, cm, c, g):
    app = pm.Polarimetry()
    app.method, app.dark, app.ilow = _Var(c['method']), _Var(repr(c['dark'])), _Var(repr(c['ilow']))
    app.offset_angle, app.polar_dir, app.per_roi = _Var(repr(c['offset_angle'])), _Var(c['polar_dir']), _Var(bool(c['per_roi']))
    app.bin_spinboxes = [_Var(c['bins'][0]), _Var(c['bins'][1])]
    app.rotation = [_Var(repr(r)) for r in c['rotation']]
    app.removed_intensity = c['removed_intensity']
    app.calls = []
    v = g['values']
    app.mask_image = cases.make_mask_image(v.shape[1:]).astype(np.float64) if c['mask'] else None
    o = util.calib_for(c)
    if c['method'] == '1PF':
        app.CD = types.SimpleNamespace(RhoPsi=np.moveaxis(np.stack((o.rho_table, o.psi_table)), 0, -1), xy=(o.grid, o.grid), name=o.name)
    elif c['method'].startswith('4POLAR'):
        app.CD = types.SimpleNamespace(invKmat=o.invK)
    else:
        app.CD = None
    if c['edge']:
        app.edge_contours = cases.make_edge(v.shape[1:])
    app.stack = cm.Stack()
    app.stack.values = v
    ds = types.SimpleNamespace(height=v.shape[1], width=v.shape[2], nangle=v.shape[0], rois=c['rois'] or [], intensity=g['intensity'],
                               vars=[], xct=[], yct=[], intmap=[], dark=0, method='')
    return app, ds


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['1pf_bin3x3', '1pf_bin1x1', '1pf_rois', '1pf_mask', '1pf_mask_rois', '1pf_removed', 'cars_bin3', 'shg_bin1',
                                  '4polar3d_bin3', '1pf_noncalib', 'shg_noncalib'])
def test_rebound_analyze_stack_on_a_stand_in_application(name):
    c, g = util.CASE_BY_NAME[name], util.golden(name)
    pm, cm = _fake_reference()
    saved = api.install(pm, cm)
    try:
        app, ds = _app(pm, cm, c, g)
        # S3 first, as the GUI does at load (PP:2242-2245): the threshold image
        assert util.eq(app.stack.get_intensity(c['dark'], list(c['bins'])), g['intensity'])
        out = app.analyze_stack(ds, for_calib=c['for_calib'])
        assert [v.name for v in ds.vars] == util.INDEX[name]['var_order']
        assert all(isinstance(v, pm.Variable) for v in ds.vars)
        exact = c['method'] == '1PF'
        for v in ds.vars:
            ref = g['var_' + v.name]
            assert util.eq(np.isnan(v.values), np.isnan(ref)), v.name
            if exact:
                assert util.eq(v.values, ref), v.name
            else:
                m = np.isfinite(ref)
                d = util.circ_diff(v.values[m], ref[m]) if v.name.startswith('Rho') else np.abs(v.values[m] - ref[m])
                assert d.max(initial=0.0) <= 1e-9, v.name
        assert util.eq(ds.intmap, g['intmap'])
        assert ds.method == c['method'] and ds.dark == c['dark']
        if c['for_calib']:
            want = [r['indx'] for r in (c['rois'] or []) if r['select']] if c['per_roi'] else [None]
            assert out == [('row', k) for k in want]
            assert app.calls == [('return_vecexcel', k) for k in want]
        else:
            assert app.calls == [('plot_data',), ('save_data',)]
            assert util.eq(ds.xmap, g['xmap']) and util.eq(ds.ymap, g['ymap'])
            # the individual-fit viewer's arrays (PP:2485-2494), produced on first use
            assert np.allclose(np.asarray(ds.chi2), g['chi2'], rtol=1e-9, atol=0, equal_nan=True)
            assert np.allclose(ds.field_fit[:, 20, 30], g['field_fit'][:, 20, 30], rtol=1e-12, atol=0, equal_nan=True)
            assert ds.field.shape == g['values'].shape
    finally:
        api.uninstall(saved)


@pytest.mark.gpu
def test_histogram_seam_equals_numpy():
    rng = np.random.default_rng(3)
    d = rng.uniform(-10, 200, 100003)
    d[:7] = [0.0, 180.0, 3.0, 177.0, 179.99999999999997, -0.0, 90.0]
    for lo, hi, nb in ((0, 180, 60), (40, 180, 60), (0, 90, 10), (-1, 1, 90)):
        bins = np.linspace(lo, hi, nb + 1)
        h, e = api.histogram(d, bins=bins, range=(lo, hi))
        h0, e0 = np.histogram(d, bins=bins, range=(lo, hi))
        assert util.eq(h, h0) and util.eq(e, e0) and h.dtype == h0.dtype
