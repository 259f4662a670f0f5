"""Live check of the NumPy restatement against the UNMODIFIED reference, executed through the stub
loader. Runs only where the reference tree is present (this build container); skipped elsewhere -- the
committed golden vectors (test_oracle_golden.py) carry the same pin to machines without the reference."""
import numpy as np
import pytest

from oracle import ref_loader as rl, restate as rs
from polarimetry_b200 import synth
from tests import util

pytestmark = pytest.mark.skipif(not rl.available(), reason='reference tree not present')


def _disk(name):
    return str(rl.reference_dir() / 'diskcones' / (name + '.mat'))


@pytest.mark.parametrize('bins,seed,disk', [((1, 1), 11, util.cases.DISK_NODIST), ((3, 3), 12, util.cases.DISK_561),
                                            ((4, 2), 13, util.cases.DISK_640)])
def test_1pf_live(bins, seed, disk):
    v = synth.stack_1pf(96, 128, 18, seed)
    ref = rl.run_reference(v, '1PF', dark=480, ilow=40000, bins=bins, disk=_disk(disk), offset_angle=60.0)
    p = rs.Params(method='1PF', dark=480.0, bins=bins, ilow=40000.0, offset_angle=60.0)
    mine = rs.analyze_stack(v, p, util.load_disk(disk))
    for k in ref['vars']:
        assert util.eq(ref['vars'][k], mine['vars'][k])
    assert util.eq(ref['intmap'], mine['intmap']) and util.eq(ref['intensity'], mine['intensity'])
    for k in ref['hist']:
        assert util.eq(ref['hist'][k], mine['hist'][k])


@pytest.mark.parametrize('method', ['SHG', 'CARS'])
def test_4harm_live(method):
    v = synth.stack_4harm(64, 96, 36, 21)
    ref = rl.run_reference(v, method, dark=0.0, bins=(3, 3))
    mine = rs.analyze_stack(v, rs.Params(method=method, dark=0.0, bins=(3, 3)), None)
    for k in ref['vars']:
        assert util.eq(ref['vars'][k], mine['vars'][k])


@pytest.mark.parametrize('method', ['4POLAR 2D', '4POLAR 3D'])
def test_4polar_live(method):
    v = synth.stack_4polar(64, 96, 5, three_d='3D' in method)
    ref = rl.run_reference(v, method, dark=0.0, bins=(2, 2))
    nm = 'Calib_th_' + method[-2] + 'D'
    mine = rs.analyze_stack(v, rs.Params(method=method, dark=0.0, bins=(2, 2)), rs.Calib.from_K(util.load_K(nm)))
    for k in ref['vars']:
        assert util.eq(ref['vars'][k], mine['vars'][k])


def test_compute_dark_live():
    pp, pc = rl.load()
    from pathlib import Path
    for seed, shape in ((1, (18, 100, 130)), (2, (4, 64, 64)), (3, (6, 41, 59))):
        v = synth.stack_1pf(shape[1], shape[2], shape[0], seed)
        v[:, :25, :25] = np.minimum(v[:, :25, :25], 600)      # a dark corner
        v[0, 3:9, 4:11] = 0                                    # zeros are masked out
        st = pc.Stack(Path('/tmp/x.tif'))
        st.values, (st.nangle, st.height, st.width) = v, v.shape
        assert rs.compute_dark(v) == st.compute_dark()
