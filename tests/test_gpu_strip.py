"""GPU parity of the strip-march kernel (pb_strip2.cuh: two lanes per image row carry the box-filter state of all 18 planes in
registers; opt-in with AnalysisParams(strip_march=True) = PB_FLAG_STRIP_MARCH) and of its redo pass: 1PF, 18 angles, 3x3 window. Bars as everywhere: masks, histograms and every 1PF map bit-exact against the
unfused kernels and the exact-chi2 path (which never takes the strip kernel), in both output modes, on shapes that exercise the
strip geometry: a partial last strip, the narrowest legal width, widths that are / are not a multiple of the 8-column tile pair,
rows whose 3-row window reflects at the top and at the bottom edge."""
import numpy as np
import pytest
import torch

from polarimetry_b200 import AnalysisParams, Engine, synth
from tests import util
from tests.test_gpu_parity import calib_for

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng():
    e = Engine()
    yield e
    e.close()


def same(a, b):
    for va, vb in zip(a.vars, b.vars):
        assert torch.equal(torch.nan_to_num(va.values, nan=-1.0), torch.nan_to_num(vb.values, nan=-1.0)), va.name
    assert torch.equal(torch.nan_to_num(a.intmap, nan=-1.0), torch.nan_to_num(b.intmap, nan=-1.0))
    assert torch.equal(a.mask, b.mask)
    assert a.hist.keys() == b.hist.keys()
    for k in a.hist:
        assert util.eq(a.hist[k], b.hist[k]), k


def fast_equals(eng, v, kw, cal, ref):
    """The FAST variant (float32 maps + mask only, whole-image threshold: 128-bit stores, unrolled epilogue) against a Result."""
    kw = dict(kw, out_dtype='float32')
    o = eng.analyze_batch(v[None], AnalysisParams(strip_march=True, **kw), cal, want_intensity=False)
    kw.pop('strip_march', None)
    r = ref if ref.vars[0].values.dtype == torch.float32 else None
    if r is None:
        r = eng.analyze(v, AnalysisParams(force_unfused=True, **kw), cal)
    for var in r.vars:
        assert torch.equal(torch.nan_to_num(o['vars'][var.name][0], nan=-1.0), torch.nan_to_num(var.values, nan=-1.0)), var.name
    assert torch.equal(torch.nan_to_num(o['intmap'][0], nan=-1.0), torch.nan_to_num(r.intmap, nan=-1.0))
    assert torch.equal(o['mask'][0].bool(), r.mask.bool())
    assert util.eq(o['hist'][0, 0, 0].cpu().numpy(), r.hist[('Rho', None)])
    assert util.eq(o['hist'][0, 0, 1].cpu().numpy(), r.hist[('Psi', None)])


@pytest.mark.parametrize('shape', [(2, 16), (31, 24), (32, 16), (33, 40), (77, 136), (96, 128), (200, 264)])
@pytest.mark.parametrize('out_dtype', ['float64', 'float32'])
def test_strip_equals_unfused_and_exact(eng, shape, out_dtype):
    H, W = shape
    v = synth.stack_1pf(H, W, 18, seed=H * 1000 + W)
    cal = calib_for(util.CASE_BY_NAME['1pf_bin1x1'])
    kw = dict(method='1PF', dark=480.0, bins=(3, 3), ilow=21000.0, out_dtype=out_dtype)
    a = eng.analyze(v, AnalysisParams(strip_march=True, **kw), cal)
    assert a.plan == 'fused_march'
    n0 = eng.launch_count()
    eng.analyze(v, AnalysisParams(strip_march=True, **kw), cal)
    assert eng.launch_count() - n0 == 2                    # k_strip2 + the redo pass of k_march_tma
    b = eng.analyze(v, AnalysisParams(force_unfused=True, **kw), cal)
    c = eng.analyze(v, AnalysisParams(exact_chi2=True, **kw), cal)
    same(a, b)
    same(a, c)
    kw.pop('out_dtype')
    fast_equals(eng, v, kw, cal, b)
    assert int(a.hist[('Rho', None)].sum()) == int(a.mask.sum())


@pytest.mark.parametrize('out_dtype', ['float64', 'float32'])
def test_strip_redo_pass_on_undecidable_pixels(eng, out_dtype):
    """Pixels the certified screening cannot decide (constant, saturated, pure cosine, alternating) scattered over several strips
    and 16-row blocks: the strip kernel leaves them to k_march_tma's redo pass."""
    H, W = 120, 160
    v = synth.stack_1pf(H, W, 18, seed=5)
    adv = synth.adversarial_1pf(32, 48, 18, seed=1)
    for (y, x) in ((0, 0), (40, 56), (88, 112), (3, 100)):
        hh, ww = min(32, H - y), min(48, W - x)
        v[:, y:y + hh, x:x + ww] = adv[:, :hh, :ww]
    cal = calib_for(util.CASE_BY_NAME['1pf_bin1x1'])
    kw = dict(method='1PF', dark=480.0, bins=(3, 3), ilow=0.0, out_dtype=out_dtype)
    a = eng.analyze(v, AnalysisParams(strip_march=True, **kw), cal)
    b = eng.analyze(v, AnalysisParams(force_unfused=True, **kw), cal)
    c = eng.analyze(v, AnalysisParams(exact_chi2=True, **kw), cal)
    same(a, b)
    same(a, c)
    kw.pop('out_dtype')
    fast_equals(eng, v, kw, cal, b)


def test_strip_batch_and_removed_intensity(eng):
    """A batch (every stack its own strips and flags) and removed_intensity != 0 (threshold image clamps at the dark value alone)."""
    cal = calib_for(util.CASE_BY_NAME['1pf_bin1x1'])
    stacks = np.stack([synth.stack_1pf(70, 88, 18, seed=s) for s in range(5)])
    for removed in (0.0, 37.0):
        kw = dict(method='1PF', dark=480.0, removed_intensity=removed, bins=(3, 3), ilow=21000.0, out_dtype='float32')
        o = eng.analyze_batch(stacks, AnalysisParams(strip_march=True, **kw), cal, want_intensity=False)       # float32 maps + mask only: the FAST variant
        hsum = 0
        for s in range(5):
            r = eng.analyze(stacks[s], AnalysisParams(force_unfused=True, **kw), cal)
            assert torch.equal(o['mask'][s].bool(), r.mask.bool())
            for var in r.vars:
                assert torch.equal(torch.nan_to_num(o['vars'][var.name][s], nan=-1.0), torch.nan_to_num(var.values, nan=-1.0)), var.name
            assert torch.equal(torch.nan_to_num(o['intmap'][s], nan=-1.0), torch.nan_to_num(r.intmap, nan=-1.0))
            hsum = hsum + r.hist[('Rho', None)]
        assert util.eq(o['hist'][0, 0, 0].cpu().numpy(), hsum)
