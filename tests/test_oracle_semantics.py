"""Micro-semantics of the third-party routines on the path, pinned on the C contract against
the live SciPy/NumPy routines: box filter window/reflect rule, LUT cell rule on exact grid hits and the
NaN rim, python-sign modulo, histogram edge rule."""
import ctypes as C

import numpy as np
import pytest
from scipy.interpolate import interpn
from scipy.ndimage import uniform_filter

from oracle import contract as ct
from tests import util


@pytest.mark.parametrize('shape', [(1, 1, 1), (2, 1, 7), (1, 2, 3), (3, 17, 23), (2, 4, 5)])
@pytest.mark.parametrize('bins', [(2, 2), (3, 3), (4, 1), (1, 5), (7, 6), (20, 20)])
def test_box_filter_matches_scipy_bitwise(shape, bins):
    rng = np.random.default_rng(hash((shape, bins)) % 2**32)
    f = rng.integers(0, 65535, size=shape).astype(np.float64) / 7.0
    ref = uniform_filter(f, size=[0, bins[0], bins[1]], mode='reflect')
    out = f.copy()
    ct.lib().pc_box(ct._p(out), ct._l(shape[0]), ct._l(shape[1]), ct._l(shape[2]), bins[0], bins[1])
    assert util.eq(out, ref)


def _lut(A2, B2, cal, rot=0.0):
    P = A2.size
    mask = np.ones(P, dtype=np.uint8)
    rho, psi = np.empty(P), np.empty(P)
    g = np.ascontiguousarray(cal.grid)
    ct.lib().pc_lut_1pf(ct._p(A2), ct._p(B2), ct._l(P), ct._p(np.ascontiguousarray(cal.rho_table)),
                        ct._p(np.ascontiguousarray(cal.psi_table)), ct._p(g), ct._l(g.size), ct._d(rot), ct._p(mask),
                        ct._p(rho), ct._p(psi))
    return rho, psi, mask


@pytest.mark.parametrize('disk', ['Disk_Ga0_Pa0_Ta0_Gb0_Pb0_Tb0_Gc0_Pc0_Tc0', 'Disk_Ga0.2_Pa10_Ta0_Gb-0.1_Pb20_Tb0_Gc0.2_Pc0_Tc0'])
def test_lut_cell_rule_on_grid_hits_rim_and_outside(disk):
    cal = util.load_disk(disk)
    g = cal.grid
    rng = np.random.default_rng(1)
    pts = [rng.uniform(-1.05, 1.05, size=(4000, 2))]
    gi = rng.integers(0, g.size, size=(2000, 2))
    pts.append(np.stack([g[gi[:, 0]], g[gi[:, 1]]], 1))                                # exact grid hits
    pts.append(np.stack([g[gi[:, 0]], rng.uniform(-1, 1, 2000)], 1))                   # one coordinate on a line
    pts.append(np.array([[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0], [0.0, -1.0], [1.0, 1.0], [0.0, 0.0], [np.nan, 0.0],
                         [0.0, np.nan], [1.0000001, 0.0], [-1.0000001, 0.2], [np.nextafter(1.0, 0), 0.0]]))
    th = rng.uniform(0, 2 * np.pi, 3000)
    rr = rng.uniform(0.97, 1.0, 3000)
    pts.append(np.stack([rr * np.cos(th), rr * np.sin(th)], 1))                        # NaN rim
    xy = np.ascontiguousarray(np.concatenate(pts))
    table = np.moveaxis(np.stack((cal.rho_table, cal.psi_table)), 0, -1)
    ref = interpn((g, g), table, xy, bounds_error=False, fill_value=np.nan)
    rho, psi, mask = _lut(np.ascontiguousarray(xy[:, 0]), np.ascontiguousarray(xy[:, 1]), cal)
    ref_rho = np.where(np.isfinite(ref[:, 0]), ref[:, 0] % 180, ref[:, 0])
    assert util.eq(rho, ref_rho)
    assert util.eq(psi, ref[:, 1])
    assert util.eq(mask.astype(bool), np.isfinite(ref[:, 0]) & np.isfinite(ref[:, 1]))
    assert np.isnan(rho).sum() > 100 and np.isfinite(rho).sum() > 1000


def test_constant_pixel_hits_lut_centre():
    cal = util.load_disk('Disk_Ga0_Pa0_Ta0_Gb0_Pb0_Tb0_Gc0_Pc0_Tc0')
    rho, psi, _ = _lut(np.zeros(1), np.zeros(1), cal)
    assert rho[0] == cal.rho_table[200, 200] % 180 and psi[0] == cal.psi_table[200, 200]


def test_pymod_tiny_negative_is_exactly_180():
    lib = ct.lib()
    X = np.array([-1e-14, 0.0, 179.99999999999997, 180.0, -0.0, 359.5, -180.0, np.nan])
    out = np.zeros(8, dtype=np.int64)
    # route through pc_tail's Rho_angle = (rho - 0) % 180 with reference_angle = 0 (computed regardless)
    ra = np.empty(8)
    a0 = np.zeros(8)
    lib.pc_tail(ct._l(8), ct._p(np.ones(8, dtype=np.uint8)), ct._p(a0), ct._p(X), ct._d(0.0), ct._p(ra), None, None)
    with np.errstate(invalid='ignore'):
        ref = X % 180
    assert util.eq(ra, ref) and ra[0] == 180.0


@pytest.mark.parametrize('nbins', [10, 60, 90])
@pytest.mark.parametrize('rng_', [(0.0, 180.0), (40.0, 180.0), (-1.0, 1.0), (0.0, 1.0)])
def test_histogram_edge_rule(nbins, rng_):
    vmin, vmax = rng_
    rng = np.random.default_rng(nbins)
    edges = np.linspace(vmin, vmax, nbins + 1)
    d = np.concatenate([rng.uniform(vmin - 1, vmax + 1, 5000), edges, np.nextafter(edges, np.inf),
                        np.nextafter(edges, -np.inf), [np.nan, np.inf, -np.inf]])
    sel = rng.random(d.size) < 0.9
    ref, _ = np.histogram(d[sel & np.isfinite(d)], bins=edges, range=(vmin, vmax))
    got = ct.histo(d, sel, 'Psi', vmin, vmax, 0.0, nbins)
    assert util.eq(got, ref)


def test_histogram_rho_rewrap():
    rng = np.random.default_rng(7)
    d = np.concatenate([rng.uniform(0, 180, 4000), [0.0, 180.0, 3.0, 177.0, 90.0]])
    for rot in (0.0, 12.0, -33.5, 180.0):
        dd = np.mod(2 * (d + rot), 360) / 2
        ref, _ = np.histogram(dd, bins=np.linspace(0, 180, 61), range=(0, 180))
        assert util.eq(ct.histo(d, None, 'Rho', 0, 180, rot, 60), ref)
