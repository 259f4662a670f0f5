"""The exact-arithmetic shortcuts of the CUDA kernels, proven on the CPU against plain IEEE operations:
  * x / n  ==  fma(x, hi, x*lo)   with (hi, lo) the double-double reciprocal of the integer n
  * RN(f*c) ==  fma(2^52 + f, c, -(2^52 c))   for integer f
(the product path relies on them to replay the reference's divisions / multiplications with fewer fp64 ops)."""
import ctypes as C

import numpy as np
import pytest

from oracle import contract as ct


@pytest.mark.parametrize('n', sorted(set(range(1, 33)) | {36, 45, 60, 90, 180}))      # box sizes 1..32 (unfused box filter), angle counts
def test_division_by_small_integer_via_double_double_reciprocal(n):
    lib = ct.lib()
    lib.pc_check_div_by_const.restype = C.c_long
    assert lib.pc_check_div_by_const(C.c_int(n), C.c_long(2_000_000), C.c_uint64(n * 7919 + 1)) == 0


def test_magic_product_equals_rounded_product():
    lib = ct.lib()
    lib.pc_check_magic_product.restype = C.c_long
    rng = np.random.default_rng(0)
    e2 = np.exp(2j * np.deg2rad(-np.linspace(0, 180, 18, endpoint=False) + 17.5))
    for c in list(e2.real) + list(e2.imag) + list(rng.uniform(-1, 1, 20)) + [1.0, -1.0, 0.0, 6.123233995736766e-17]:
        assert lib.pc_check_magic_product(C.c_double(float(c)), C.c_long(200_000), C.c_uint64(int(abs(c) * 1e9) + 3)) == 0


def test_markstein_division_by_grid_spacing():
    """t = (x - g[i]) / (g[i+1] - g[i]) via the reciprocal table: exact for every spacing of linspace(-1, 1, n)."""
    lib = ct.lib()
    lib.pc_check_markstein_div.restype = C.c_long
    for n in (401, 201, 1001):
        g = np.linspace(-1, 1, n)
        for h in np.unique(np.diff(g)):
            assert lib.pc_check_markstein_div(C.c_double(float(h)), C.c_long(400_000), C.c_uint64(int(h * 1e12) % 2**31 + 5)) == 0
