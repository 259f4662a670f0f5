"""The oracle pinned against outputs of the unmodified reference (tests/golden, made by oracle/make_golden.py).

restate.py (same NumPy/SciPy calls as the reference) must be bit-identical on every array, histogram and
statistic. contract.c (op-for-op scalar fp64) must be bit-identical wherever no transcendental is involved
(all of 1PF, every mask, intensity/intmap, every histogram) and within 1e-9 deg / 1e-12 elsewhere.
"""
import numpy as np
import pytest

from oracle import cases, contract as ct, restate as rs
from tests import util

NAMES = [c['name'] for c in util.CASES]


def _run(mod, c, v, edge):
    p = util.params_for(c)
    if mod is rs:
        return rs.analyze_stack(v, p, util.calib_for(c), c['rois'], c['per_roi'], util.base_mask_for(c, v.shape[1:]), edge, c['for_calib'])
    return ct.analyze_stack(v, p, util.calib_for(c), c['rois'], c['per_roi'], util.base_mask_for(c, v.shape[1:]), edge)


@pytest.mark.parametrize('name', NAMES)
def test_restate_bit_identical_to_reference(name):
    c = util.CASE_BY_NAME[name]
    g = util.golden(name)
    v = g['values']
    assert util.eq(v, cases.make_stack(c['stack'])), 'synthetic generator drifted from the committed input'
    edge = cases.make_edge(v.shape[1:]) if c['edge'] else None
    r = _run(rs, c, v, edge)
    assert list(r['vars']) == util.INDEX[name]['var_order']
    assert util.eq(r['intensity'], g['intensity'])
    assert util.eq(r['intmap'], g['intmap'])
    assert util.eq(r['roi_map'], g['roi_map'])
    for k, val in r['vars'].items():
        assert util.eq(val, g['var_' + k]), k
    for (k, s), h in r['hist'].items():
        assert util.eq(h, g[util.hist_key(k, s)]), (k, s)
    if not c['for_calib']:
        for k in ('xmap', 'ymap', 'chi2'):
            assert util.eq(r[k], g[k]), k
    rows = util.INDEX[name]['rows']
    if rows:
        sl = [None] if not (c['per_roi'] and c['rois']) else [q['indx'] for q in c['rois'] if q['select']]
        for row, s in zip(rows, sl):
            st = r['stats'][s]
            assert row[2] == rs.fmt3(float(st['MeanRho'])) and row[3] == rs.fmt3(float(st['StdRho']))
            names = [k for k in r['vars'] if k not in ('Rho', 'Rho_contour', 'Rho_angle')]
            col = 4
            for nm in names:
                assert row[col] == rs.fmt3(float(st['Mean' + nm])) and row[col + 1] == rs.fmt3(float(st['Std' + nm]))
                col += 2
            assert row[col] == rs.fmt3(float(st['MeanInt'])) and row[col + 1] == rs.fmt3(float(st['StdInt']))
            assert row[col + 3] == st['N']


@pytest.mark.parametrize('name', NAMES)
def test_contract_matches_reference(name):
    c = util.CASE_BY_NAME[name]
    g = util.golden(name)
    v = g['values']
    edge = cases.make_edge(v.shape[1:]) if c['edge'] else None
    r = _run(ct, c, v, edge)
    exact = c['method'] == '1PF'
    assert util.eq(r['intensity'], g['intensity'])
    assert util.eq(r['intmap'], g['intmap'])
    assert util.eq(r['mask'], np.isfinite(g['var_Rho']))
    for k, val in r['vars'].items():
        ref = g['var_' + k]
        assert util.eq(np.isnan(val), np.isnan(ref)), k
        if exact:
            assert util.eq(val, ref), k
        else:
            m = np.isfinite(ref)
            d = util.circ_diff(val[m], ref[m]) if k.startswith('Rho') else np.abs(val[m] - ref[m])
            assert d.max(initial=0.0) <= (1e-9 if k in ('Rho', 'Psi', 'Eta', 'Rho_angle', 'Rho_contour') else 1e-12), k
    for (k, s), h in r['hist'].items():
        assert util.eq(h, g[util.hist_key(k, s)]), (k, s)


def test_contract_chi2_and_fit_bits_on_fma_hosts():
    """fit/chi2 bits are host dependent (numpy's fused complex multiply); on FMA hosts the contract's
    fma() spelling reproduces them exactly. Only their comparisons are parity relevant."""
    c = util.CASE_BY_NAME['1pf_noncalib']
    g = util.golden('1pf_noncalib')
    v = g['values']
    edge = cases.make_edge(v.shape[1:])
    r = ct.analyze_stack(v, util.params_for(c), util.calib_for(c), None, False, None, edge, want_fit=True)
    m = r['mask']
    assert np.allclose(r['chi2'][m], g['chi2'][m], rtol=1e-12, atol=0)
    assert np.allclose(r['field_fit'][:, m], g['field_fit'][:, m], rtol=1e-13, atol=0)
