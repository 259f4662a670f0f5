"""The C-ABI library builds for sm_100a, loads, and exports every entry point include/polar_b200.h declares
(no compute call is made: this runs without a GPU). Also: the product refuses to run without a CUDA device
instead of falling back to anything."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_functions():
    h = (ROOT / 'include' / 'polar_b200.h').read_text()
    h = re.sub(r'/\*.*?\*/', '', h, flags=re.S)
    return sorted(set(re.findall(r'\b(pb_[a-z_0-9]+)\s*\(', h)))


def test_library_exports_every_declared_symbol():
    from polarimetry_b200 import _cabi, build
    build.build()
    lib = _cabi.load()
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_cabi.EXPORTS), set(names) ^ set(_cabi.EXPORTS)
    assert lib.pb_abi_version() == 1


def test_struct_layouts_match_header():
    from polarimetry_b200 import _cabi
    # pb_params: 8 int32 then 7 doubles ; pb_outputs: 4 + 6 pointers
    assert C.sizeof(_cabi.PbParams) == 8 * 4 + 7 * 8
    assert C.sizeof(_cabi.PbOutputs) == 10 * C.sizeof(C.c_void_p)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    from polarimetry_b200 import Engine, _cabi
    lib = _cabi.load()
    ctx = C.c_void_p()
    assert lib.pb_create(C.byref(ctx), 0) != 0          # fails loudly: no device, no CPU path
    assert b'no CPU path' in lib.pb_last_error(None)
    with pytest.raises(RuntimeError):
        Engine()


def test_product_never_imports_the_oracle():
    for f in (ROOT / 'polarimetry_b200').rglob('*.py'):
        src = f.read_text()
        assert 'oracle' not in src, f
    for f in (ROOT / 'polarimetry_b200' / 'csrc').iterdir():
        assert 'oracle/' not in f.read_text().replace('oracle/contract.c', ''), f


def test_flag_constants_match_header():
    """The ctypes mirror carries the header's flag / enum values (a drifted constant would silently select another code path)."""
    from polarimetry_b200 import _cabi
    h = (ROOT / 'include' / 'polar_b200.h').read_text()
    flags = {m.group(1): int(m.group(2)) for m in re.finditer(r'#define\s+(PB_FLAG_[A-Z0-9_]+)\s+(\d+)u', h)}
    assert flags == {'PB_FLAG_OUT_F64': _cabi.FLAG_OUT_F64, 'PB_FLAG_NO_HIST': _cabi.FLAG_NO_HIST, 'PB_FLAG_EXACT_CHI2': _cabi.FLAG_EXACT_CHI2,
                     'PB_FLAG_FORCE_UNFUSED': _cabi.FLAG_FORCE_UNFUSED, 'PB_FLAG_PROJECT_ONLY': _cabi.FLAG_PROJECT_ONLY,
                     'PB_FLAG_HIST_MATCH': _cabi.FLAG_HIST_MATCH, 'PB_FLAG_STRIP_MARCH': _cabi.FLAG_STRIP_MARCH}
    m = re.search(r'typedef enum pb_dtype \{([^}]*)\}', h)
    dt = {k.strip(): int(v) for k, v in (x.split('=') for x in m.group(1).split(','))}
    assert dt == {'PB_U8': _cabi.DTYPES['uint8'], 'PB_U16': _cabi.DTYPES['uint16'], 'PB_F32': _cabi.DTYPES['float32'], 'PB_F64': _cabi.DTYPES['float64']}
    for name, val in (('PB_MAX_ROIS', _cabi.PB_MAX_ROIS), ('PB_MAX_VARS', _cabi.PB_MAX_VARS), ('PB_MAX_BINS', _cabi.PB_MAX_BINS)):
        mm = re.search(r'#define\s+' + name + r'\s+(\d+)', h)
        assert mm and int(mm.group(1)) == val, name
