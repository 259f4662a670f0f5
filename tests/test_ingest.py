"""raw-stack ingest (Polarimetry.define_stack, PP:1923-1942). Golden = what the UNMODIFIED reference's define_stack
returned for multi-page TIFFs written here (oracle/make_golden_ingest.py -> tests/golden/ingest.npz)."""
import numpy as np
import pytest

from oracle import restate as rs
from tests import util

KINDS = ('f32', 'f32_small', 'f32_flatmax', 'u16')


def golden():
    return np.load(util.GOLD / 'ingest.npz')


@pytest.mark.parametrize('kind', KINDS)
def test_oracle_define_stack_values(kind):
    g = golden()
    out = rs.define_stack_values(g[f'{kind}__in'])
    assert out.dtype == np.uint16 and np.array_equal(out, g[f'{kind}__out'])


def test_oracle_float64_pages():
    a = golden()['f32__in'].astype(np.float64) * 1.000001
    out = rs.define_stack_values(a)
    assert out.dtype == np.uint16 and out.max() == 65535 and out.min() == 0


@pytest.mark.gpu
@pytest.mark.parametrize('kind', KINDS)
def test_device_define_stack_matches_reference(kind, tmp_path):
    import cv2
    from polarimetry_b200 import Engine
    g = golden()
    a = g[f'{kind}__in']
    eng = Engine(0)
    st = eng.define_stack(a, default_dark=0)
    assert (st['nangle'], st['height'], st['width']) == a.shape and st['dark'] == 0.0
    assert np.array_equal(st['values'].cpu().view(dtype=__import__('torch').int16).numpy().view(np.uint16), g[f'{kind}__out'])
    # the TIFF path: same reader call as the reference (cv2.imreadmulti, IMREAD_ANYDEPTH), dark from the device compute_dark
    f = tmp_path / f'{kind}.tif'
    assert cv2.imwritemulti(str(f), list(a))
    st2 = eng.define_stack(f)
    vals = st2['values'].cpu().view(dtype=__import__('torch').int16).numpy().view(np.uint16)
    assert np.array_equal(vals, g[f'{kind}__out'])
    assert st2['dark'] == rs.compute_dark(g[f'{kind}__out'])
    eng.close()


@pytest.mark.gpu
def test_device_rescale_large_and_float64():
    """2048x2048x6 float32 and float64 stacks against the NumPy statement (bit-exact), NaN refused."""
    import torch
    from polarimetry_b200 import Engine
    eng = Engine(0)
    rng = np.random.default_rng(3)
    for dt in (np.float32, np.float64):
        a = (rng.gamma(2.0, 700.0, size=(6, 1024, 2048)) - 250.0).astype(dt)
        st = eng.define_stack(a, default_dark=0)
        got = st['values'].cpu().view(dtype=torch.int16).numpy().view(np.uint16)
        assert np.array_equal(got, rs.define_stack_values(a))
    a = np.ones((2, 8, 8), dtype=np.float32)
    a[1, 3, 3] = np.nan
    with pytest.raises(RuntimeError):
        eng.define_stack(a, default_dark=0)
    with pytest.raises(ValueError):
        eng.define_stack(np.ones((8, 8), dtype=np.float32))
    eng.close()


@pytest.mark.gpu
def test_hwn_layout_conversion():
    """angle-minor [H,W,N] stacks are converted to the reference's [N,H,W] on the device (bit copy), ragged sizes, every
    dtype, batches; the analysis of the converted stack is the analysis of the original."""
    import torch
    from polarimetry_b200 import AnalysisParams, Calibration, Engine, synth
    eng = Engine(0)
    rng = np.random.default_rng(9)
    for dt, N, H, W in ((np.uint16, 18, 37, 53), (np.uint8, 12, 64, 64), (np.float32, 36, 19, 130), (np.float64, 128, 9, 21), (np.uint16, 4, 1, 1)):
        a = (rng.random((2, H, W, N)) * 250).astype(dt)
        got = eng.from_hwn(a)
        assert tuple(got.shape) == (2, N, H, W)
        g = got.cpu()
        g = g.view(torch.int16).numpy().view(np.uint16) if dt == np.uint16 else g.numpy()
        assert np.array_equal(g, np.ascontiguousarray(np.moveaxis(a, -1, 1)))
    with pytest.raises(RuntimeError):
        eng.from_hwn(np.zeros((4, 4, 129), dtype=np.uint16))
    v = synth.stack_1pf(72, 96, 18, seed=77)
    d = util.load_disk(util.cases.DISK_NODIST)
    cal = Calibration(d.rho_table, d.psi_table, 401)
    p = AnalysisParams(method='1PF', dark=480.0, bins=(3, 3), ilow=40000.0)
    r1 = eng.analyze(v, p, cal)
    r2 = eng.analyze(eng.from_hwn(np.ascontiguousarray(np.moveaxis(v, 0, -1))), p, cal)
    for x, y in zip(r1.vars, r2.vars):
        assert util.eq(x.values.cpu().numpy(), y.values.cpu().numpy())
    assert util.eq(r1.mask.cpu().numpy(), r2.mask.cpu().numpy())
    eng.close()


def test_c_contract_rescale_matches_reference():
    """oracle/contract.c restates the rescale one rounded operation at a time: identical to what the reference returned."""
    from oracle import contract
    g = golden()
    for kind in ('f32', 'f32_small', 'f32_flatmax'):
        assert np.array_equal(contract.rescale(g[f'{kind}__in']), g[f'{kind}__out'])
    a = g['f32__in'].astype(np.float64) * 1.000001
    assert np.array_equal(contract.rescale(a), rs.define_stack_values(a))


@pytest.mark.gpu
def test_folder_pipeline_equals_file_by_file(tmp_path):
    """Folder mode with overlap (FolderAnalyzer): a mix of uint16 and float32 multi-page TIFFs of two sizes; every file's maps and
    histograms equal the serial define_stack -> set_dark -> analyze of the same file, and (uint16 files) the oracle."""
    import cv2
    import torch
    from polarimetry_b200 import AnalysisParams, Calibration, Engine, synth
    from polarimetry_b200.ingest import FolderAnalyzer
    from oracle import cases
    eng = Engine(0)
    o = util.load_disk(cases.DISK_NODIST)
    cal = Calibration(o.rho_table, o.psi_table, o.grid.size)
    files, stacks = [], []
    for k in range(7):
        h, w = ((64, 96), (48, 80))[k % 2]
        v = synth.stack_1pf(h, w, 18, seed=40 + k)
        v[:, :22, :22] = np.minimum(v[:, :22, :22], 520 + k)            # a dark corner: compute_dark finds it
        a = v.astype(np.float32) * 0.37 + 3.0 if k == 3 else v           # one float TIFF (rescaled on the device, PP:1929-1930)
        f = tmp_path / f'stack_{k}.tif'
        assert cv2.imwritemulti(str(f), list(a))
        files.append(f)
        stacks.append(a)
    p = AnalysisParams(method='1PF', bins=(3, 3), ilow=30000.0)
    got = {}

    def keep(i, f, res, dark):
        got[i] = (dark, {v.name: v.values.cpu().numpy() for v in res.vars}, res.mask.cpu().numpy().astype(bool), dict(res.hist))

    fa = FolderAnalyzer(eng, p, cal, slots=3, decode_threads=2)
    recs = fa.run(files, keep)
    assert [r['file'] for r in recs] == [str(f) for f in files] and len(fa.timeline) == len(files)
    rep = fa.overlap_report()
    assert rep['files'] == 7 and rep['wall_ms'] > 0
    for k, f in enumerate(files):
        st = eng.define_stack(f)
        dark = float('{:.0f}'.format(st['dark']))
        assert got[k][0] == dark
        res = eng.analyze(st['values'], AnalysisParams(method='1PF', bins=(3, 3), ilow=30000.0, dark=dark), cal)
        for v in res.vars:
            assert util.eq(got[k][1][v.name], v.values.cpu().numpy()), (k, v.name)
        assert util.eq(got[k][2], res.mask.cpu().numpy().astype(bool))
        for key, h in res.hist.items():
            assert util.eq(got[k][3][key], h)
        vals = rs.define_stack_values(stacks[k])
        assert dark == float('{:.0f}'.format(rs.compute_dark(vals)))
        ref = rs.analyze_stack(vals, rs.Params(method='1PF', bins=(3, 3), ilow=30000.0, dark=dark), o, want_stats=False)
        for name, arr in ref['vars'].items():
            assert util.eq(got[k][1][name], arr), (k, name)
        assert util.eq(got[k][2], ref['mask'])
    eng.close()
