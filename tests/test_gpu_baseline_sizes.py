"""GPU parity at the BASELINE.json sizes, straight against the oracle (oracle/restate.py, bit-identical to the
reference): the shape bench.py times (1PF 2048x2048x18, bin 3x3 and 1x1, G1 threshold), configs[1] (SHG and CARS
2048x2048x36) and configs[2] (4POLAR 3D 4x2048x2048). Bars as in test_gpu_parity.py: masks, intensity/intmap,
every histogram count and every 1PF map bit-exact; rho/psi/eta <= 1e-9 deg and S2/S4 <= 1e-12 where a
transcendental is involved."""
import numpy as np
import pytest

from oracle import cases, restate as rs
from polarimetry_b200 import AnalysisParams, Calibration, Engine, synth
from tests import util

pytestmark = pytest.mark.gpu
H = W = 2048
ANGLE_VARS = ('Rho', 'Psi', 'Eta')


@pytest.fixture(scope='module')
def eng():
    e = Engine()
    yield e
    e.close()


def g1_ilow(intensity):
    """20th percentile of the threshold image, '{:.0f}'-rounded (what bench.py uses)."""
    return float('{:.0f}'.format(np.percentile(intensity, 20.0)))


def compare(res, ref, exact_maps, p=None):
    assert util.eq(res.mask.cpu().numpy().astype(bool), ref['mask'])
    assert util.eq(res.intensity.cpu().numpy(), ref['intensity'])
    assert util.eq(res.intmap.cpu().numpy(), ref['intmap'])
    assert [v.name for v in res.vars] == list(ref['vars'])
    for v in res.vars:
        got, want = v.values.cpu().numpy(), ref['vars'][v.name]
        if exact_maps:
            assert util.eq(got, want), v.name
        else:
            assert util.eq(np.isnan(got), np.isnan(want)), v.name
            m = np.isfinite(want)
            d = util.circ_diff(got[m], want[m]) if v.name == 'Rho' else np.abs(got[m] - want[m])
            assert d.max(initial=0.0) <= (1e-9 if v.name in ANGLE_VARS else 1e-12), (v.name, d.max())
    assert set(res.hist) == set(ref['hist'])
    for k, h in res.hist.items():
        if exact_maps:
            assert util.eq(h, ref['hist'][k]), k
            continue
        # transcendental outputs: exact wherever the reference value is farther than the map tolerance from a bin edge
        # (util.hist_explained_by_ties), and exactly the histogram of the device's own fp64 map
        name = k[0]
        vmin, vmax = p.ranges[name]
        edges = np.linspace(vmin, vmax, p.nbins + 1)
        tol = 1e-9 if name in ANGLE_VARS else 1e-12
        ok, n_ties, n_moved = util.hist_explained_by_ties(h, ref['hist'][k], ref['vars'][name], edges, tol, name == 'Rho', float(p.rotation[1]))
        assert ok, (k, n_ties, n_moved, h - ref['hist'][k])
        own = rs.histo(res.get_var(name).values.cpu().numpy(), None, name, vmin, vmax, float(p.rotation[1]), p.nbins)[0]
        assert util.eq(h, own), (k, 'device histogram vs histogram of the device map')


@pytest.mark.parametrize('bins', [(3, 3), (1, 1)])
def test_1pf_2048x2048x18_headline_shape(eng, bins):
    """The workload bench.py times: G1 stack, dark 480, G1 threshold (~80 % of the pixels valid), no-distortion disk."""
    v = synth.stack_1pf(H, W, 18, seed=0)
    o = util.load_disk(cases.DISK_NODIST)
    cal = Calibration(o.rho_table, o.psi_table, o.grid.size)
    ilow = g1_ilow(rs.get_intensity(v, 480.0, bins))
    res = eng.analyze(v, AnalysisParams(method='1PF', dark=480.0, bins=bins, ilow=ilow), cal)
    assert res.plan == ('fused_march' if bins[0] > 1 else 'fused_pixel')
    ref = rs.analyze_stack(v, rs.Params(method='1PF', dark=480.0, bins=bins, ilow=ilow), o, want_stats=False)
    assert 0.79 < ref['mask'].mean() < 0.81
    compare(res, ref, exact_maps=True)
    # throughput mode: float32 maps are the rounded fp64 values, histograms unchanged
    r32 = eng.analyze(v, AnalysisParams(method='1PF', dark=480.0, bins=bins, ilow=ilow, out_dtype='float32'), cal)
    for var in r32.vars:
        assert util.eq(var.values.cpu().numpy(), ref['vars'][var.name].astype(np.float32)), var.name
    for k, h in r32.hist.items():
        assert util.eq(h, ref['hist'][k]), k


@pytest.mark.parametrize('method,bins', [('SHG', (1, 1)), ('CARS', (1, 1)), ('SHG', (3, 3))])
def test_4harmonic_2048x2048x36_configs1(eng, method, bins):
    v = synth.stack_4harm(H, W, 36, seed=1, dark=100)
    p = dict(method=method, dark=100.0, bins=bins, ilow=30000.0 if method == 'SHG' else 0.0)
    res = eng.analyze(v, AnalysisParams(**p))
    ref = rs.analyze_stack(v, rs.Params(**p), None, want_stats=False)
    assert ref['mask'].mean() > 0.5
    compare(res, ref, exact_maps=False, p=rs.Params(**p))


def test_4polar3d_4x2048x2048_configs2(eng):
    v = synth.stack_4polar(H, W, seed=2, three_d=True)
    K = util.load_K('Calib_th_3D')
    for bins in ((1, 1), (3, 3)):
        p = dict(method='4POLAR 3D', dark=0.0, bins=bins, ilow=1000.0)
        res = eng.analyze(v, AnalysisParams(**p), Calibration.from_K(K))
        ref = rs.analyze_stack(v, rs.Params(**p), rs.Calib.from_K(K), want_stats=False)
        assert ref['mask'].mean() > 0.3
        compare(res, ref, exact_maps=False, p=rs.Params(**p))
        same_decisions_in_float32_mode(eng, v, p, Calibration.from_K(K), res)


def same_decisions_in_float32_mode(eng, v, p, cal, res64):
    """float32-output mode runs the certified fast evaluation of the 4POLAR angles: the mask and EVERY histogram count must be
    those of the double-precision mode (decisions are only taken from a fast value when its error bound proves them), the maps
    within 1e-3 degrees (north_star)."""
    r32 = eng.analyze(v, AnalysisParams(out_dtype='float32', **p), cal)
    assert util.eq(r32.mask.cpu().numpy(), res64.mask.cpu().numpy())
    for k, h in r32.hist.items():
        assert util.eq(h, res64.hist[k]), k
    for a, b in zip(r32.vars, res64.vars):
        x, y = a.values.cpu().numpy().astype(np.float64), b.values.cpu().numpy()
        assert util.eq(np.isnan(x), np.isnan(y)), a.name
        m = np.isfinite(y)
        d = util.circ_diff(x[m], y[m]) if a.name == 'Rho' else np.abs(x[m] - y[m])
        assert d.max(initial=0.0) <= 1e-3, (a.name, d.max())


def test_4polar2d_fast_path_decisions(eng):
    v = synth.stack_4polar(1024, 1024, seed=5, three_d=False)
    K = util.load_K('Calib_th_2D')
    for kw in (dict(bins=(1, 1)), dict(bins=(2, 2), rotation=(33.0, 12.0, 0.0))):
        p = dict(method='4POLAR 2D', dark=0.0, ilow=1000.0, **kw)
        res = eng.analyze(v, AnalysisParams(**p), Calibration.from_K(K))
        ref = rs.analyze_stack(v, rs.Params(**p), rs.Calib.from_K(K), want_stats=False)
        compare(res, ref, exact_maps=False, p=rs.Params(**p))
        same_decisions_in_float32_mode(eng, v, p, Calibration.from_K(K), res)


def test_two_temporary_calibrations_back_to_back(eng):
    """The disk-cone upload cache must not confuse two short-lived Calibration objects (CPython recycles id())."""
    v = synth.stack_1pf(64, 96, 18, seed=4)
    p = AnalysisParams(method='1PF', dark=480.0, ilow=30000.0, offset_angle=60.0)
    outs = []
    for name in (cases.DISK_NODIST, cases.DISK_561, cases.DISK_NODIST, cases.DISK_640):
        o = util.load_disk(name)
        res = eng.analyze(v, p, Calibration(o.rho_table, o.psi_table, o.grid.size))        # temporary: freed right after the call
        ref = rs.analyze_stack(v, rs.Params(method='1PF', dark=480.0, ilow=30000.0, offset_angle=60.0), o, want_stats=False)
        for var in res.vars:
            assert util.eq(var.values.cpu().numpy(), ref['vars'][var.name]), (name, var.name)
        outs.append(res.vars[1].values.cpu().numpy())
    assert not util.eq(outs[0], outs[1])


def test_rejected_hist_setup_keeps_engine_usable(eng):
    """A refused pb_set_hist (vmin >= vmax) must leave the context consistent: the next analysis with the previous settings works."""
    v = synth.stack_1pf(40, 56, 18, seed=6)
    o = util.load_disk(cases.DISK_NODIST)
    cal = Calibration(o.rho_table, o.psi_table, o.grid.size)
    good = AnalysisParams(method='1PF', dark=480.0, ilow=30000.0)
    a = eng.analyze(v, good, cal)
    bad = AnalysisParams(method='1PF', dark=480.0, ilow=30000.0)
    bad.ranges = dict(bad.ranges, Rho=(10.0, 10.0))
    with pytest.raises(Exception):
        eng.analyze(v, bad, cal)
    b = eng.analyze(v, good, cal)
    for k in a.hist:
        assert util.eq(a.hist[k], b.hist[k])
    assert int(b.hist[('Rho', None)].sum()) == int(b.mask.sum())
