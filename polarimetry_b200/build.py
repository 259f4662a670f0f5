"""Build libpolarb200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc.

    python -m polarimetry_b200.build [--force]

The library is compiled for B200 only (-gencode arch=compute_100a,code=sm_100a), with -fmad=false so that
no fp64 multiply-add is contracted behind the arithmetic contract's back, and -lineinfo so ncu's source
page maps to these files. nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / 'csrc'
TAG = os.environ.get('PB_BUILD_TAG', '')          # developer A/B builds: libpolarb200_<tag>.so (select with PB_LIB=...)
LIB = PKG / (f'libpolarb200_{TAG}.so' if TAG else 'libpolarb200.so')
SOURCES = ['pb_strip_1.cu', 'pb_march_tma_1.cu', 'pb_march_tma_2.cu', 'pb_march_tma_4p.cu', 'pb_march_u16_1.cu', 'pb_march_u16_2.cu', 'pb_pixel.cu', 'pb_march_u8_1.cu', 'pb_march_u8_2.cu', 'pb_api.cu', 'pb_box.cu', 'pb_march.cu',
           'pb_reduce.cu', 'pb_compact.cu', 'pb_ingest.cu']
NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-fmad=false',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + list(CSRC.glob('*.cuh')) + [PKG.parent / 'include' / 'polar_b200.h']
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not _stale():
        return LIB
    objs = []
    bdir = PKG / 'build' / TAG if TAG else PKG / 'build'
    bdir.mkdir(parents=True, exist_ok=True)
    procs = []
    for s in SOURCES:
        o = bdir / (s + '.o')
        cmd = ['nvcc', *NVCC_FLAGS, *os.environ.get('PB_EXTRA_NVCC', '').split(), '-c', str(CSRC / s), '-o', str(o)] + (['-Xptxas', '-v'] if verbose else [])
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(o))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f'--- {s}\n{out}\n')
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    subprocess.check_call(['nvcc', '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', str(LIB), *objs])
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
